#!/usr/bin/env python
"""Cycle breakdown of the device-side transcript step of k_sc_small:
    ZKB200_TRACE=2 python tools/dt_trace.py 2>&1 | grep "device transcript"
prints, per round, the clock64 cycles of interpolate+serialise / absorb+Keccak-f / challenge on warp 0 and of
claim+record / system fence on the last warp."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
z = importlib.import_module("zk-research-implementations_b200")
ctx = z.Context(0, 0, 1)
ctx.set_device_transcript(True)  # off by default
tabs = [z.MultilinearPoly.generate(ctx, 5, t, 10) for t in range(2)]
sp = z.SumPoly(ctx, [z.ProductPoly.from_polys(ctx, tabs)])
S, T = z.sum_check_protocol, z.fiat_shamir.Transcript
for _ in range(3):
    S.gkr_prove(0, sp, T(0))
print("device transcript launches / challenges checked:", ctx.device_transcript_stats())
