#!/usr/bin/env python3
"""Dynamic code footprint from an ncu report with source data (--set full --import-source on):
how many distinct SASS instructions carry the executed instruction stream, and where they sit.

    python tools/hot_code.py <report.ncu-rep> [bucket_bytes]
    NCU_ARGS="--launch-count 1" ...   extra `ncu -i` selection arguments for reports that hold several launches
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
import os  # noqa: E402
out = subprocess.run(["ncu", "-i", rep, *os.environ.get("NCU_ARGS", "").split(), "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(lines[start:]))
ins = []
for r in rows:
    try:
        a = int(r["Address"], 16) if r["Address"].startswith("0x") else int(r["Address"])
        ins.append((a, int(r["Instructions Executed"] or 0), int(r["# Samples"] or 0), int(r.get("stall_no_inst") or 0), r["Source"]))
    except (ValueError, KeyError):
        continue
base = min(a for a, *_ in ins)
tot = sum(e for _, e, *_ in ins)
tots = sum(s for _, _, s, *_ in ins)
print(f"{len(ins)} SASS instructions, {tot:.3e} warp-instructions executed, {tots} samples")
srt = sorted(ins, key=lambda x: -x[1])
acc = 0
marks = [0.5, 0.8, 0.9, 0.95, 0.99]
mi = 0
for k, (a, e, *_r) in enumerate(srt):
    acc += e
    while mi < len(marks) and acc >= marks[mi] * tot:
        print(f"  {marks[mi]*100:4.0f}% of executed instructions come from {k+1:5d} SASS instructions ({(k+1)*16/1024:.1f} KB)")
        mi += 1
print(f"\nper {bucket}-byte bucket of code: executed share, sample share, no_inst samples")
b = {}
for a, e, s, ni, _ in ins:
    k = (a - base) // bucket
    x = b.setdefault(k, [0, 0, 0])
    x[0] += e; x[1] += s; x[2] += ni
for k in sorted(b):
    e, s, ni = b[k]
    if e / tot > 0.004 or s / tots > 0.004:
        print(f"  +{k*bucket/1024:6.1f} KB  exec {100*e/tot:5.1f}%  samples {100*s/tots:5.1f}%  no_inst {ni}")
