#!/usr/bin/env python3
"""sha256 (first 16 hex digits) over the sources libzkb200.so is built from, in a fixed order.  The Makefile bakes it
into the library (zkb_version()); engine.tree_src_hash() recomputes it from the tree, so a stale prebuilt .so is noticed."""
import glob
import hashlib
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def src_hash() -> str:
    files = sorted(glob.glob(os.path.join(HERE, "*.cu")) + glob.glob(os.path.join(HERE, "*.cuh")) + glob.glob(os.path.join(HERE, "*.hpp"))
                   + glob.glob(os.path.join(HERE, "*.h")))
    files.append(os.path.join(HERE, "..", "..", "include", "zkb200.h"))
    h = hashlib.sha256()
    for f in files:
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(src_hash())
