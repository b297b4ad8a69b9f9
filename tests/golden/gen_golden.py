#!/usr/bin/env python3
"""Generate tests/golden/*.json from the reference tree (build container only).

* bundled_pairs.json : the 15 input sequences of /root/reference/data/*.fa and the
  8 pairings implied by examples/RNA-RNAdata.zip, plus the
  only known-answer in the reference, README.md:91-97 (DIS x DIS dot-brackets).
* ushuffle.json : outputs of the REFERENCE's own src/ushuffle.c (compiled where it
  lies into oracle/_ref by oracle/Makefile) driven exactly like
  src/ractip.cpp:1636-1643: srandom(seed); alternate shuffle(s1,k=2), shuffle(s2,k=2).
* default_model.bin : the bytes of the rp_model struct rp_model_default(use_bl=1) fills (BL* tables from
  src/boltzmann_param.c over the residual Turner-2004 tables), so that bench.py's reference arm and other
  checker-side code can run the oracle without loading the product library.
"""
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))

PAIRS = [("CopA", "CopT"), ("DIS", "DIS"), ("IncRNA54", "RepZ"), ("MicA", "ompA"),
         ("OxyS", "fhlA"), ("R1inv", "R2inv"), ("RyhB", "SodB"), ("Tar", "Tarstar")]


def read_fa(p):
    name, seq = None, ""
    for line in p.read_text().splitlines():
        if line.startswith(">"):
            if name is not None:
                break
            name = line[1:].strip()
        elif line and line[0].isalpha():
            seq += "".join(ch for ch in line.strip() if ch.isalpha())
    return name, seq


def write_default_model():
    import ctypes
    from ractip_b200 import default_model
    m = default_model()
    (ROOT / "tests" / "golden" / "default_model.bin").write_bytes(ctypes.string_at(ctypes.addressof(m), ctypes.sizeof(m)))


def main():
    write_default_model()
    seqs = {}
    for fa in sorted((REF / "data").glob("*.fa")):
        _, s = read_fa(fa)
        seqs[fa.stem] = s
    out = {"sequences": seqs, "pairs": PAIRS,
           "readme_dis": {"s1": "((((.(((((((..[[[[[[.)))))))...))))", "s2": "((((.(((((((..]]]]]].)))))))...))))",
                          "cite": "README.md:91-97"}}
    (ROOT / "tests" / "golden" / "bundled_pairs.json").write_text(json.dumps(out, indent=1))

    from oracle import oracle as orc
    orc.build()
    ref = C.CDLL(str(orc.REF_USHUFFLE))
    libc = C.CDLL(None)
    cases = []
    for a, b, seed, num in [("MicA", "ompA", 1, 6), ("DIS", "DIS", 7, 4), ("CopA", "CopT", 12345, 3),
                            ("Tar", "Tarstar", 2, 4)]:
        s1, s2 = seqs[a], seqs[b]
        libc.srandom(C.c_uint(seed))
        ref.set_randfunc(C.cast(libc.random, C.c_void_p))
        b1 = C.create_string_buffer(len(s1) + 1)
        b2 = C.create_string_buffer(len(s2) + 1)
        r1, r2 = [], []
        for _ in range(num):
            ref.shuffle(s1.encode(), b1, len(s1), 2)
            ref.shuffle(s2.encode(), b2, len(s2), 2)
            r1.append(b1.raw[:len(s1)].decode())
            r2.append(b2.raw[:len(s2)].decode())
        cases.append({"s1": a, "s2": b, "seed": seed, "k": 2, "mode": 12, "out1": r1, "out2": r2})
    # k=3 and k=1 single-sequence streams
    for name, seed, k, num in [("ompA", 5, 3, 3), ("MicA", 9, 1, 3)]:
        s = seqs[name]
        libc.srandom(C.c_uint(seed))
        ref.set_randfunc(C.cast(libc.random, C.c_void_p))
        b1 = C.create_string_buffer(len(s) + 1)
        r = []
        for _ in range(num):
            ref.shuffle(s.encode(), b1, len(s), k)
            r.append(b1.raw[:len(s)].decode())
        cases.append({"s1": name, "s2": name, "seed": seed, "k": k, "mode": 1, "out1": r, "out2": [s] * num})
    (ROOT / "tests" / "golden" / "ushuffle.json").write_text(json.dumps({"cases": cases}, indent=1))
    print("wrote golden fixtures")


if __name__ == "__main__":
    main()
