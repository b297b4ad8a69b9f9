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
import os  # noqa: E402
out = subprocess.run(["ncu", "-i", rep, *os.environ.get("NCU_ARGS", "").split(), "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
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
# aggregate per enclosing source function (found by scanning the source files for definitions at column 0)
_defs = {}
def _load(f):
    out = []
    for path in list((ROOT / "ractip_b200" / "csrc").glob(f)) :
        for n, line in enumerate(path.read_text().splitlines(), 1):
            m = re.match(r"^(?:RP_HD|__device__|__global__|inline|static)\b[^;]*?\b(\w+)\s*\(", line)
            if m and not line.rstrip().endswith(";"):
                out.append((n, m.group(1)))
    return out
def region(f, l):
    if f not in _defs:
        _defs[f] = _load(f)
    name = None
    for n, fn in _defs[f]:
        if n <= l:
            name = fn
        else:
            break
    return f"{f.split('.')[0]}:{name}" if name else f
agg = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), (h, e, s) in by.items():
    x = agg[region(f, l)]
    x[0] += h; x[1] += e; x[2] += s
print(f"{'region':28s} {'hot instr':>9s} {'KB':>6s} {'exec %':>7s} {'samp %':>7s}")
tots = sum(s for _, s in dyn.values())
for k, (h, e, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} {h:9d} {h*16/1024:6.1f} {100*e/tot:7.2f} {100*s/tots:7.2f}")
print("total hot static instr:", sum(v[0] for v in agg.values()))
