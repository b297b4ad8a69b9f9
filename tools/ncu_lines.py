#!/usr/bin/env python3
"""Aggregate an ncu report's per-instruction samples by CUDA source line.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [kernel_substring] [top]
    NCU_KERNEL=<name> ...   picks that kernel's launch out of a report that holds several (ncu --kernel-name)

kernel_substring selects the SASS section by its MANGLED name: give enough of it to single out one template
instance (mcc_band_kernelILi512ELi1, mcc_persistentILi1ELi10), or the line map of another instance is used.

Uses `ncu --page source --csv` for the per-SASS-instruction counters and
`nvdisasm -g` on the cubin extracted from the built library for the SASS->line map.
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "mcc_persistent"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30

tmp = Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "ractip_b200" / "libractip_prob.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
addr2line = {}
for cubin in tmp.glob("*.cubin"):
    sass = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                          text=True).stdout.split("\n")
    infunc, cur = False, None
    for l in sass:
        if l.startswith(".text.") or l.startswith("\t.section"):
            infunc = kern in l
        if not infunc:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
        if m:
            addr2line[int(m.group(1), 16)] = cur
import os  # noqa: E402
sel = ["--kernel-name", os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []
out = subprocess.run(["ncu", "-i", rep, *sel, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stallcols = [(k, c) for k, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
agg = collections.defaultdict(lambda: [0, 0])
stall = collections.defaultdict(collections.Counter)
base = None
ti = ts = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    if r[ia] == "Address":   # a second kernel's section starts: only the first captured launch is summarised
        break
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    ln = addr2line.get(a - base)
    inst, s = int(r[ii] or 0), int(r[isamp] or 0)
    agg[ln][0] += inst
    agg[ln][1] += s
    ti += inst
    ts += s
    for k, c in stallcols:
        if r[k] and r[k] != "0":
            stall[ln][c] += int(r[k])
print(f"total warp-instructions {ti}  samples {ts}")
src_cache = {}
for ln, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = ""
    if ln:
        f = ROOT / "ractip_b200" / "csrc" / ln[0]
        if f.exists():
            src_cache.setdefault(f, f.read_text().split("\n"))
            text = src_cache[f][ln[1] - 1].strip()[:90]
    st = ",".join(f"{k[6:]}:{v * 100 // max(s, 1)}" for k, v in stall[ln].most_common(3))
    print(f"{str(ln):28s} inst {100 * i / ti:5.1f}%  samp {100 * s / ts:5.1f}%  [{st}]  {text}")

# --- per-function totals (line ranges from the source)
import bisect
funcs = []
for fname in ("mcc_core.h", "mcc_band.h", "mcc_band_shfl.cuh", "mcc_wide_shfl.cuh", "mcc_driver.h", "kernels.cu"):
    f = ROOT / "ractip_b200" / "csrc" / fname
    for no, l in enumerate(f.read_text().split("\n"), start=1):
        m = re.match(r"(?:template.*\n)?(?:RP_HD|__global__|__device__|inline).*?\b(\w+)\(", l)
        if m and not l.startswith(" "):
            funcs.append((fname, no, m.group(1)))
tot = collections.defaultdict(lambda: [0, 0])
for ln, (i, s) in agg.items():
    if not ln:
        tot["?"][0] += i; tot["?"][1] += s
        continue
    cand = [(no, name) for fn, no, name in funcs if fn == ln[0] and no <= ln[1]]
    name = cand[-1][1] if cand else ln[0]
    tot[name][0] += i; tot[name][1] += s
print("\nby function:")
for name, (i, s) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"  {name:20s} inst {100 * i / ti:5.1f}%  samp {100 * s / ts:5.1f}%")
