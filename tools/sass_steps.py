#!/usr/bin/env python
"""Static instruction counts of the pass-kernel formulations (no GPU needed).

    python tools/sass_steps.py [> profiles/rN_sass_steps_*.txt]

Compiles tools/sass_variants.cu for sm_100a with the library's flags, disassembles it and,
for every kernel, finds the walk's loops (backward branches), removes the rarely-taken
blocks (whatever a forward conditional branch inside the loop jumps over: the generic
step, the wall crossing, the NaN stop) and reports what is left per step: the instructions
every step issues, of which FP64-pipe ones, plus registers and spills from ptxas.

The pass kernels are bound by instruction issue (f32) and by issue + the FP64 pipe (f64)
-- DESIGN.md section 5.1 -- so these counts rank candidate formulations before GPU time is
spent on them; tools/kernel_lab.cu then times the short list.
"""
from __future__ import annotations

import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rlic_b200._build import NVCC_FLAGS, find_nvcc  # noqa: E402

BRANCH = re.compile(r"(@!?U?P\d\s+)?BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)")


def compile_and_disassemble(src: Path):
    with tempfile.TemporaryDirectory() as tmp:
        cubin = Path(tmp) / "variants.cubin"
        flags = [f for f in NVCC_FLAGS if f not in ("-shared", "-Xcompiler", "-fPIC", "-cudart", "static")]
        r = subprocess.run([find_nvcc(), *flags, "-Xptxas", "-v", "-cubin", "-o", str(cubin), str(src)],
                           capture_output=True, text=True)
        if r.returncode:
            sys.exit(r.stderr)
        sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True, check=True).stdout
    usage, cur = {}, None
    for line in r.stderr.splitlines():
        if m := re.search(r"Compiling entry function '(\S+)'", line):
            cur = m.group(1)
        if (m := re.search(r"Used (\d+) registers", line)) and cur:
            usage.setdefault(cur, {})["regs"] = int(m.group(1))
        # the first such line after the entry is the kernel's own; later ones belong to the
        # out-of-line functions it calls
        if (m := re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)) and cur:
            usage.setdefault(cur, {}).setdefault("spill", int(m.group(1)) + int(m.group(2)))
    functions, cur = {}, None
    for line in sass.splitlines():
        if m := re.search(r"Function : (\S+)", line):
            cur = m.group(1)
            functions[cur] = []
        elif (m := re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)) and cur:
            functions[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return usage, functions


def demangle(name: str) -> str:
    full = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.split("(")[0]
    return full.replace("void rlic::lic_pass_kernel", "").replace("rlic::", "")


def describe(label: str) -> str:
    """<T, POL, Taps, Idx, TW, TH, UNROLL, MINB, FLAVOR, ADMIT, BRANCHLESS, WALK, REC> -> short text"""
    args = [a.strip() for a in re.sub(r"ParamTaps<(\w+), (\d+)>", r"ParamTaps", label.strip("<>")).split(",")]
    t, pol = args[0], args[1] == "true"
    unroll, minb, flavor, admit, walk = args[6], args[7], args[8], args[9], args[11]
    rec = " recording" if len(args) > 12 and args[12] == "true" else ""
    return (f"{'f32' if t == 'float' else 'f64'} {'pol' if pol else 'vel'} unroll {unroll} blocks {minb} "
            f"flavor {flavor} admit {admit} walk {walk:>2}{rec}")


def pipe_of(op: str) -> str:
    if op in ("FFMA", "FADD", "FMUL", "IMAD", "HFMA2", "HADD2", "HMUL2"):
        return "fma"
    if op[0] == "D":
        return "fp64"
    if op in ("LDG", "STG", "LDC", "LDCU", "LDL", "STL", "LDS", "STS"):
        return "mem"
    if op in ("BRA", "BSSY", "BSYNC", "BREAK", "CALL", "RET", "EXIT", "WARPSYNC"):
        return "ctl"
    return "alu"


def loops_of(body):
    found = []
    for addr, ins in body:
        m = BRANCH.match(ins)
        if m and int(m.group(2), 16) < addr:
            found.append((int(m.group(2), 16), addr))
    return found


def fast_path(body, lo, hi):
    inside = [(a, i) for a, i in body if lo <= a <= hi]
    rare = set()
    for a, ins in inside:
        m = BRANCH.match(ins)
        if m and m.group(1) and a < int(m.group(2), 16) <= hi:
            rare.update(x for x, _ in inside if a < x < int(m.group(2), 16))
    return [(a, i) for a, i in inside if a not in rare]


def main() -> None:
    usage, functions = compile_and_disassemble(ROOT / "tools" / "sass_variants.cu")
    rows = []
    for name, body in functions.items():
        label = demangle(name)
        if not label.startswith("<"):
            continue
        # <T, POL, Taps, Idx, TW, TH, UNROLL, ...>: counted from the front (the list grows at the back)
        steps_per_group = int([a.strip() for a in re.sub(r"ParamTaps<(\w+), (\d+)>", "ParamTaps",
                                                          label.strip("<>")).split(",")][6])
        loops = sorted(loops_of(body), key=lambda lh: lh[0] - lh[1])[:2]    # the two main loops
        per_step, fp64, local = [], [], 0
        mix = collections.Counter()
        for lo, hi in loops:
            fast = fast_path(body, lo, hi)
            ops = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", i).split()[0].split(".")[0] for _, i in fast)
            per_step.append(len(fast) / steps_per_group)
            fp64.append(sum(n for op, n in ops.items() if op[0] == "D") / steps_per_group)
            local += ops.get("LDL", 0) + ops.get("STL", 0)
            for op, n in ops.items():
                mix[pipe_of(op)] += n / steps_per_group / len(loops)
        u = usage.get(name, {})
        rows.append((describe(label), sum(per_step) / len(per_step), per_step, sum(fp64) / len(fp64),
                     u.get("regs"), u.get("spill", 0), local, len(body), mix))
    print("fast-path instructions per step in the walk's two main loops (forward, backward), sm_100a SASS")
    print("pipes: fma = FFMA/FADD/FMUL/IMAD/HFMA2, alu = integer, logic, compare, select, min/max, move, "
          "fp64 = D*, mem = loads/stores, ctl = branches and reconvergence")
    print(f"{'kernel':62s} {'instr/step':>10s} {'fwd':>6s} {'bwd':>6s} {'fma':>6s} {'alu':>6s} {'fp64':>6s} "
          f"{'mem':>5s} {'ctl':>5s} {'regs':>4s} {'spill B':>7s} {'LDL+STL in loops':>16s} {'SASS total':>10s}")
    for d, mean, per, f64, regs, spill, local, total, mix in sorted(rows, key=lambda r: r[:2]):
        print(f"{d:62s} {mean:10.2f} {per[0]:6.2f} {per[-1]:6.2f} {mix['fma']:6.2f} {mix['alu']:6.2f} "
              f"{mix['fp64']:6.2f} {mix['mem']:5.2f} {mix['ctl']:5.2f} {regs!s:>4s} {spill:7d} {local:16d} {total:10d}")


if __name__ == "__main__":
    main()
