#!/usr/bin/env python
"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck), checked against the oracle:
exercises every round-driver regime, the GKR path and the MLE kernels at sizes a sanitizer finishes in seconds.
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import importlib, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
z = importlib.import_module("zk-research-implementations_b200")
from oracle import c_oracle as O

fid, p = 0, z.engine.MODULI[0]
rng = random.Random(11)
ok = True
for mode, P, D, n in ((1, 1, 2, 13), (0, 2, 2, 10), (1, 2, 3, 9)):
    ctx = z.Context(fid, 0, mode)
    tabs = [[rng.randrange(p) for _ in range(1 << n)] for _ in range(P * D)]
    ref = O.gkr_sumcheck_prove(O.Transcript(fid), mode, P, D, [O.ints_to_arr(t) for t in tabs])
    sp = z.SumPoly(ctx, [z.ProductPoly(ctx, tabs[q * D:(q + 1) * D]) for q in range(P)])
    for tail, small in ((40, 200 * 1024), (0, 0), (40, 4096)):
        ctx.set_tail_threshold(tail)
        ctx.set_small_threshold(small)
        pr = z.sum_check_protocol.gkr_prove(0, sp, z.fiat_shamir.Transcript(fid))
        ok &= [q.coefficients for q in pr.proof_polynomials] == ref["coeffs"] and pr.final_values == ref["final_vals"]
    m = z.MultilinearPoly(ctx, tabs[0])
    rs = [rng.randrange(p) for _ in range(n)]
    ok &= m.evaluate(rs) == O.mle_evaluate(fid, O.ints_to_arr(tabs[0]), rs)
    ok &= m.partial_evaluate(1, rs[0]).evaluation == O.arr_to_ints(O.mle_partial_evaluate(fid, O.ints_to_arr(tabs[0]), 1, rs[0]))
    pl = z.sum_check_protocol.prove(m)
    claimed, msgs, _ = O.sumcheck_prove(fid, O.ints_to_arr(tabs[0]))
    ok &= (pl.claimed_sum, pl.proof_polynomials) == (claimed, msgs)
    ctx.close()
ctx = z.Context(2, 0, 0)
p2 = z.engine.MODULI[2]
L = 7
gates = [1 << (L - 1 - l) for l in range(L)]
ops = [[rng.randrange(2) for _ in range(g)] for g in gates]
inputs = [rng.randrange(p2) for _ in range(2 * gates[0])]
c = z.gkr_circuit.Circuit(ctx, [[z.Operation(o) for o in layer] for layer in ops])
pr = z.gkr_protocol.prove(c, inputs)
ref = O.gkr_prove(2, gates, np.array([o for l in ops for o in l], dtype=np.uint8), O.ints_to_arr(inputs))
ok &= [[q.coefficients for q in layer] for layer in pr.proof_polynomials] == ref["proof_polynomials"]
ok &= z.gkr_protocol.verify(pr, c, inputs)
ctx.close()
print("SANITIZE_TARGET", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
