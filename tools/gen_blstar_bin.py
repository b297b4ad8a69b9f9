#!/usr/bin/env python3
"""Pack the Andronescu BL* integer energy tables into params/blstar.bin.

The numbers are *data* (Andronescu et al., RNA 2010) that the reference keeps as
C initialisers in src/boltzmann_param.c:21-5906.  This script reads them where
they lie under /root/reference (read-only, only available in the build
container), and writes them in this repo's own packed binary layout; no
reference source text is copied.  The committed output params/blstar.bin is what
travels; re-run this only to regenerate it.

Layout (little endian):
    8 bytes  magic  b"RPBLSTR1"
    int32    n_arrays
    per array: 24-byte NUL-padded name, int32 count, count x int32 values
    int32    n_tetra
    per tetraloop: 8-byte NUL-padded 6-mer, int32 bonus
Macro values: INF=10000000 (ViennaRNA energy_const.h), NST=0, DEF=-50
(boltzmann_param.c:17-18).
"""
import re
import struct
import sys
from pathlib import Path

SRC = Path("/root/reference/src/boltzmann_param.c")
OUT = Path(__file__).resolve().parent.parent / "params" / "blstar.bin"
MACROS = {"INF": 10000000, "NST": 0, "DEF": -50}
EXPECT = {  # name -> count
    "stack37a": 49, "mismatchH37a": 175, "mismatchI37a": 175,
    "dangle5_37a": 40, "dangle3_37a": 40, "int11_37a": 1225,
    "int21_37a": 6125, "int22_37a": 12544, "hairpin37a": 31,
    "bulge37a": 31, "internal_loop37a": 31, "MLparams_a": 4, "ninio_a": 2,
}


def main():
    text = SRC.read_text()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    arrays = {}
    for m in re.finditer(r"static\s+int\s+(\w+)\s*\[\]\s*=\s*\{(.*?)\};", text, flags=re.S):
        name, body = m.group(1), m.group(2)
        vals = []
        for tok in body.replace("\n", " ").split(","):
            tok = tok.strip()
            if not tok:
                continue
            vals.append(MACROS[tok] if tok in MACROS else int(tok))
        arrays[name] = vals
    for name, cnt in EXPECT.items():
        assert name in arrays, name
        assert len(arrays[name]) == cnt, (name, len(arrays[name]), cnt)
    tetra = re.findall(r'\{\s*"([ACGU]{6})"\s*,\s*(-?\d+)\s*\}', text)
    assert len(tetra) == 30, len(tetra)

    blob = bytearray(b"RPBLSTR1")
    blob += struct.pack("<i", len(EXPECT))
    for name in EXPECT:
        blob += name.encode().ljust(24, b"\0")
        blob += struct.pack("<i", len(arrays[name]))
        blob += struct.pack("<%di" % len(arrays[name]), *arrays[name])
    blob += struct.pack("<i", len(tetra))
    for s, e in tetra:
        blob += s.encode().ljust(8, b"\0") + struct.pack("<i", int(e))
    OUT.parent.mkdir(exist_ok=True)
    OUT.write_bytes(bytes(blob))
    print("wrote", OUT, len(blob), "bytes;", sum(EXPECT.values()), "ints +", len(tetra), "tetraloops")


if __name__ == "__main__":
    sys.exit(main())
