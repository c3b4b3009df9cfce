import importlib, os, sys, random
sys.path.insert(0, "/root/repo")
z = importlib.import_module("zk-research-implementations_b200")
ctx = z.Context(0, 0, 1)
p = z.engine.MODULI[0]
tabs = [z.MultilinearPoly.generate(ctx, 5, t, 10) for t in range(2)]
sp = z.SumPoly(ctx, [z.ProductPoly.from_polys(ctx, tabs)])
S, T = z.sum_check_protocol, z.fiat_shamir.Transcript
for _ in range(3):
    S.gkr_prove(0, sp, T(0))
