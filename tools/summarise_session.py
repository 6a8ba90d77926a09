#!/usr/bin/env python
"""Turns the output of tools/sessions/gpu_session.sh / gpu_session_2gpu.sh (gpurun_out/session*/) into
the artefacts that get committed under profiles/ and prints the decisions they support.

    python tools/summarise_session.py [--round 2] [--session gpurun_out/session]

Copies: the bench lines (one JSON per variant), the lab tables, the pytest summary (with the
XPASS / XFAIL list of the first-run tests), the ncu launch list.  Prints, for every opt-in
switch, its number next to the default's -- flipping a default is a decision for a human (or
the next session) to take with these numbers in hand, not something this script does.
"""
from __future__ import annotations

import argparse
import json
import re
import shutil
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def bench_line(path: Path) -> dict | None:
    if not path.exists():
        return None
    for line in path.read_text().splitlines():
        if line.startswith("{"):
            try:
                return json.loads(line)
            except json.JSONDecodeError:
                pass
    return None


def lab_table(path: Path) -> list[tuple[str, float, str, int]]:
    """(variant, ms, same-as-shipped, registers) rows of a kernel_lab log."""
    rows = []
    if not path.exists():
        return rows
    for line in path.read_text().splitlines():
        m = re.match(r"(.{44})\s+([\d.]+)\s+[\d.]+\s+\d+\s+(yes|NO)\s+(\d+)\s*$", line)
        if m:
            rows.append((m.group(1).strip(), float(m.group(2)), m.group(3), int(m.group(4))))
    return rows


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", type=int, default=2)
    ap.add_argument("--session", default="gpurun_out/session")
    args = ap.parse_args()
    src = ROOT / args.session
    if not src.is_dir():
        raise SystemExit(f"{src} does not exist: run tools/sessions/gpu_session.sh through gpurun first")
    dst = ROOT / "profiles"
    tag = f"r{args.round}_{src.name}"

    copied = []
    for name in ("summary.txt", "launches.csv", "bench_lines.jsonl", "lines.jsonl", "gpu.csv"):
        if (src / name).exists():
            shutil.copy(src / name, dst / f"{tag}_{name}")
            copied.append(name)
    for log in sorted(src.glob("lab_*.log")) + sorted(src.glob("pytest_*.log")):
        shutil.copy(log, dst / f"{tag}_{log.stem}.txt")
        copied.append(log.name)
    print(f"copied into profiles/ as {tag}_*: {', '.join(copied) or 'nothing'}\n")

    # bench variants against the default
    base = bench_line(src / "bench_default.log") or bench_line(src / "bench_nccl.log")
    if base:
        print(f"{'bench variant':28s} {'Mpix/s':>10s} {'ms/step':>9s} {'pass ms':>8s} {'roofline':>9s} "
              f"{'e2e Mpix/s':>11s} {'e2e ms':>8s}")
        for log in sorted(src.glob("bench_*.log")):
            line = bench_line(log)
            if not line:
                print(f"{log.stem:28s} no JSON line (see {log})")
                continue
            roof, e2e = line.get("roofline") or {}, line.get("e2e") or {}
            print(f"{log.stem:28s} {line['value']:10.0f} {line['ms_per_step']:9.3f} "
                  f"{roof.get('launch_ms', float('nan')):8.3f} {roof.get('frac', float('nan')):9.3f} "
                  f"{e2e.get('value', float('nan')):11.0f} {e2e.get('ms_per_step', float('nan')):8.3f}"
                  f"   vs default x{line['value'] / base['value']:.3f}, e2e x"
                  f"{e2e.get('value', 0) / max(base.get('e2e', {}).get('value', 1), 1):.3f}")
        parity = base.get("parity")
        if parity:
            print(f"\nparity of the default run: {json.dumps(parity)}")

    # lab: best grouped formulation per type against the shipped kernel of the same run
    for log in sorted(src.glob("lab_*.log")):
        rows = lab_table(log)
        if not rows:
            continue
        print(f"\n{log.stem}:")
        shipped = [r for r in rows if r[0] == "shipped"]
        for k, (name, ms, same, regs) in enumerate(rows):
            ref = max((s for s in shipped if rows.index(s) <= k), key=rows.index, default=None)
            rel = f"x{ref[1] / ms:.3f} vs shipped" if ref and name != "shipped" else ""
            flag = "" if same == "yes" else "   <-- RESULT DIFFERS FROM THE SHIPPED KERNEL"
            print(f"  {name:44s} {ms:8.3f} ms  {regs:3d} regs  {rel}{flag}")

    # first-run tests
    for log in sorted(src.glob("pytest_*.log")):
        text = log.read_text()
        xpass = re.findall(r"^XPASS (\S+)", text, re.M)
        xfail = re.findall(r"^XFAIL (\S+)", text, re.M)
        tail = [ln for ln in text.splitlines() if re.search(r"\d+ (passed|failed)", ln)]
        print(f"\n{log.stem}: {tail[-1] if tail else 'no summary line'}")
        print(f"  first-run tests green (drop their marker): {len(xpass)}")
        for t in xfail:
            print(f"  first-run test FAILED: {t}")


if __name__ == "__main__":
    main()
