#!/bin/sh
# Disassembles the kernels launch_pass() actually dispatches for the default options (grouped
# walk, default arithmetic, taps in the parameter block, 32-bit cell indices) out of the built
# library, one file per kernel under profiles/, so that per-step instruction claims can be
# checked against the shipped code (tools/sass_steps.py counts the same SASS):
#   the four walking pass kernels, the two recording instantiations of the f32 walk (pass 1 of
#   a call of two or more iterations) and the replay kernels of passes 2.. (one and two groups
#   of 32 steps per half: kernels of up to 65 / 129 taps).
#   sh tools/dump_sass.sh r2
set -e
cd "$(dirname "$0")/.."
TAG=${1:-r2}
LIB=rlic_b200/librlic_b200.so
ALL=$(mktemp)
trap 'rm -f "$ALL"' EXIT
cuobjdump -sass "$LIB" > "$ALL"
# name : kernel : the template arguments as they are mangled (a substring of the symbol)
for spec in \
    "lic_pass_kernel_f32_velocity:lic_pass_kernel:IfLb0ENS_9ParamTapsIfLi768EEEiLi16ELi16ELi8ELi8ELi2ELi4ELb1ELi7ELb0EE" \
    "lic_pass_kernel_f32_polarization:lic_pass_kernel:IfLb1ENS_9ParamTapsIfLi768EEEiLi16ELi16ELi4ELi8ELi0ELi3ELb1ELi1ELb0EE" \
    "lic_pass_kernel_f64_velocity:lic_pass_kernel:IdLb0ENS_9ParamTapsIdLi384EEEiLi16ELi16ELi4ELi5ELi0ELi2ELb1ELi9ELb0EE" \
    "lic_pass_kernel_f64_polarization:lic_pass_kernel:IdLb1ENS_9ParamTapsIdLi384EEEiLi16ELi16ELi4ELi4ELi0ELi2ELb1ELi9ELb0EE" \
    "lic_pass_kernel_f32_velocity_recording:lic_pass_kernel:IfLb0ENS_9ParamTapsIfLi768EEEiLi16ELi16ELi8ELi6ELi2ELi4ELb1ELi7ELb1EE" \
    "lic_pass_kernel_f64_velocity_recording:lic_pass_kernel:IdLb0ENS_9ParamTapsIdLi384EEEiLi16ELi16ELi4ELi5ELi0ELi2ELb1ELi9ELb1EE" \
    "lic_replay_kernel_f32_one_group:lic_replay_kernel:IfNS_8StepTapsIfLi384EEEiLi1ELb0ELi16ELi16ELi8EE" \
    "lic_replay_kernel_f32_two_groups:lic_replay_kernel:IfNS_8StepTapsIfLi384EEEiLi2ELb0ELi16ELi16ELi8EE" \
    "lic_replay_kernel_f64_two_groups:lic_replay_kernel:IdNS_8StepTapsIdLi192EEEiLi2ELb0ELi16ELi16ELi8EE"; do
    name=${spec%%:*}; rest=${spec#*:}; kernel=${rest%%:*}; key=${rest#*:}
    out=profiles/${TAG}_sass_${name}.txt
    awk -v key="$kernel$key" '
        /Function :/ { on = index($0, key) > 0 }
        on { print }' "$ALL" | sed 's#/\* 0x[0-9a-f]* \*/##' | grep -v "^\s*$" > "$out"
    n=$(grep -c ';' "$out" || true)
    if [ "$n" -eq 0 ]; then
        echo "$out: NOT FOUND in $LIB (the template arguments changed: update the key)" >&2
        exit 1
    fi
    echo "$out: $n instructions"
done
