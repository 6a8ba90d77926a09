#!/bin/sh
# Disassembles the pass-kernel instantiations that launch_pass() actually dispatches for the
# default arithmetic (grouped walk, taps in the parameter block, 32-bit cell indices) out of the
# built library, one file per kernel under profiles/, so that per-step instruction claims can
# be checked against the shipped code (tools/sass_steps.py counts the same SASS).
#   sh tools/dump_sass.sh r2
set -e
cd "$(dirname "$0")/.."
TAG=${1:-r2}
LIB=rlic_b200/librlic_b200.so
cuobjdump -sass "$LIB" > /tmp/rlic_all.sass
for spec in "f32_velocity:IfLb0ENS_9ParamTapsIfLi768EEEiLi16ELi16ELi8ELi8ELi2ELi4ELb1ELi7E" \
            "f32_polarization:IfLb1ENS_9ParamTapsIfLi768EEEiLi16ELi16ELi4ELi8ELi0ELi3ELb1ELi1E" \
            "f64_velocity:IdLb0ENS_9ParamTapsIdLi384EEEiLi16ELi16ELi4ELi5ELi0ELi2ELb1ELi9E" \
            "f64_polarization:IdLb1ENS_9ParamTapsIdLi384EEEiLi16ELi16ELi4ELi4ELi0ELi2ELb1ELi9E"; do
    name=${spec%%:*}; key=${spec#*:}
    out=profiles/${TAG}_sass_lic_pass_kernel_${name}.txt
    awk -v key="lic_pass_kernel$key" '
        /Function :/ { on = index($0, key) > 0 }
        on { print }' /tmp/rlic_all.sass | sed 's#/\* 0x[0-9a-f]* \*/##' | grep -v "^\s*$" > "$out"
    echo "$out: $(grep -c ';' "$out") instructions"
done
