#!/usr/bin/env python3
"""Static code size of the kernels in build/kernels.cu.o: SASS instructions per kernel and, for one
kernel, per source file / line range (from the -lineinfo annotations of nvdisasm -g).

    python tools/sass_size.py [kernel-name-fragment]

The per-diagonal loops of the wavefront kernels have to fit the SM's 32 KB instruction cache
(2048 instructions of 16 bytes): misses go to the GPC-level cache, which all SMs of a GPC share.
"""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
obj = ROOT / "build" / "kernels.cu.o"
frag = sys.argv[1] if len(sys.argv) > 1 else "mcc_band_kernelILi512"
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(obj)], cwd=td, check=True, stdout=subprocess.DEVNULL)
    cubin = next(Path(td).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-g", str(cubin)], check=True, stdout=subprocess.PIPE, text=True).stdout
cur_fn, cur_src = None, None
per_fn = collections.Counter()
per_src = collections.Counter()
for ln in dis.splitlines():
    m = re.match(r"\.text\.(\S+):", ln)
    if m:
        cur_fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_src = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]+\*/", ln) and cur_fn:
        per_fn[cur_fn] += 1
        if frag in cur_fn and cur_src:
            per_src[cur_src] += 1
for fn, c in per_fn.most_common():
    print(f"{c:7d} instr {c * 16 / 1024:7.1f} KB  {fn}")
print(f"\nby source file for *{frag}*:")
by_file = collections.Counter()
for (f, l), c in per_src.items():
    by_file[f] += c
for f, c in by_file.most_common():
    print(f"{c:7d}  {f}")
print("\ntop source lines:")
for (f, l), c in per_src.most_common(40):
    print(f"{c:6d}  {f}:{l}")
