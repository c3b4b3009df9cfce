#!/usr/bin/env python
"""Constants of csrc/g1.cuh (BLS12-381 Fq, Montgomery R = 2^384, 12 x u32 little-endian): prints them and, with
--check, verifies that the header carries exactly these values."""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import kzg_ref as K

Q = K.Q
Rm = 1 << 384


def limbs(v):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(12)]


vals = {
    "P": limbs(Q),
    "ONE": limbs(Rm % Q),
    "R2": limbs(Rm * Rm % Q),
    "gx": limbs(K.G1[0] * Rm % Q),
    "gy": limbs(K.G1[1] * Rm % Q),
}
inv = (-pow(Q, -1, 1 << 32)) % (1 << 32)


def fmt(v):
    return ", ".join("0x%08xu" % x for x in v)


if __name__ == "__main__":
    if "--check" in sys.argv:
        src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zk-research-implementations_b200", "csrc", "g1.cuh")).read()
        flat = re.sub(r"\s+", " ", src)
        ok = ("INV = 0x%08xu" % inv) in flat
        for name, v in vals.items():
            ok &= fmt(v) in flat
            if fmt(v) not in flat:
                print("MISMATCH", name, fmt(v))
        print("g1.cuh constants", "ok" if ok else "WRONG")
        sys.exit(0 if ok else 1)
    print("INV = 0x%08xu" % inv)
    for name, v in vals.items():
        print(name, "=", fmt(v))
