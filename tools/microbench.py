#!/usr/bin/env python
"""Multiplier-pipe microbenchmarks on the B200: raw IMAD variants and the register-resident Montgomery
product.  Fills the IMAD-roofline denominator that MEASURED_PEAKS.json does not carry."""
import ctypes as C
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
z = importlib.import_module("zk-research-implementations_b200")
out = {}
for fid, name in ((0, "bn254_fr"), (2, "bls12_381_fr")):
    with z.Context(fid, 0, 0) as ctx:
        L = z.engine.lib()
        v = C.c_double()
        if fid == 0:
            for mode, nm in enumerate(["imad_lo", "imad_hi", "imad_wide", "imad_wide_x_carry"]):
                z.engine._ck(ctx, L.zkb_bench_imad(ctx.handle, mode, 4096, C.byref(v)))
                out[nm + "_per_s"] = v.value
        for var, nm in enumerate(["wide_ilp1", "wide_ilp2", "split_ilp1", "split_ilp2"]):
            z.engine._ck(ctx, L.zkb_bench_modmul(ctx.handle, var, 2048, C.byref(v)))
            out[f"modmul_{name}_{nm}_per_s"] = v.value
print(json.dumps(out, indent=1))
