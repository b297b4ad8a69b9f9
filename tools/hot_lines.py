#!/usr/bin/env python3
"""Join the dynamic per-SASS-instruction counts of an ncu report (source page) with the static
address -> file:line map of nvdisasm -g: static size and executed share of the HOT code per source
function region.   python tools/hot_lines.py <report.ncu-rep> <kernel-fragment> [min_exec]
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep, frag = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
dyn = {}
for r in csv.DictReader(lines[start:]):
    try:
        a = int(r["Address"], 16) if r["Address"].startswith("0x") else int(r["Address"])
        dyn[a] = (int(r["Instructions Executed"] or 0), int(r["# Samples"] or 0))
    except (ValueError, KeyError):
        pass
base = min(dyn)
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "build" / "kernels.cu.o")], cwd=td, check=True, stdout=subprocess.DEVNULL)
    cubin = next(Path(td).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-g", str(cubin)], check=True, stdout=subprocess.PIPE, text=True).stdout
cur_fn, cur_src, stat = None, None, {}
for ln in dis.splitlines():
    m = re.match(r"\.text\.(\S+):", ln)
    if m:
        cur_fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_src = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
    if m and cur_fn and frag in cur_fn:
        stat[int(m.group(1), 16)] = cur_src
tot = sum(e for e, _ in dyn.values())
n_calls = max(e for e, _ in dyn.values())
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0002
by = collections.defaultdict(lambda: [0, 0, 0])   # static hot instr, executed, samples
for a, (e, s) in dyn.items():
    src = stat.get(a - base)
    if src is None:
        continue
    key = src
    x = by[key]
    if e > thr * n_calls:
        x[0] += 1
    x[1] += e; x[2] += s
# aggregate per file in blocks of source lines that are contiguous
agg = collections.defaultdict(lambda: [0, 0, 0])
def region(f, l):
    table = {
        "mcc_band.h": [(0,150,"band:setup/weights"),(151,204,"band:segs/ldg"),(205,225,"band:ring_ld"),(226,335,"band:interior_item"),(336,405,"band:specials"),(406,454,"band:interior_A"),(455,472,"band:interior_sum"),(473,532,"band:cfac"),(533,657,"band:finish_in"),(658,720,"band:finish_out"),(721,760,"band:collect")],
        "mcc_core.h": [(0,130,"core:ctx/off/tb"),(131,215,"core:misc"),(216,250,"core:pair_type/ss/stems"),(251,275,"core:special_loop"),(276,345,"core:hairpin/row_sum"),(346,440,"core:dots/multi_dot"),(441,570,"core:prologue"),(571,800,"core:inside generic"),(801,880,"core:nick"),(881,1130,"core:outside generic"),(1131,1335,"core:wide"),(1336,1635,"core:unstru"),(1636,1700,"core:outputs")],
    }
    for lo, hi, name in table.get(f, []):
        if lo <= l <= hi:
            return name
    return f
for (f, l), (h, e, s) in by.items():
    x = agg[region(f, l)]
    x[0] += h; x[1] += e; x[2] += s
print(f"{'region':28s} {'hot instr':>9s} {'KB':>6s} {'exec %':>7s} {'samp %':>7s}")
tots = sum(s for _, s in dyn.values())
for k, (h, e, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} {h:9d} {h*16/1024:6.1f} {100*e/tot:7.2f} {100*s/tots:7.2f}")
print("total hot static instr:", sum(v[0] for v in agg.values()))
