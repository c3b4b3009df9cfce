#!/usr/bin/env python
"""Stress the host<->device handshakes: thousands of proofs back to back over a mix of shapes and regimes; every
repetition must reproduce the oracle's proof bit for bit (a torn mailbox read, a lost challenge or a stale table
would change the bytes)."""
import importlib, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
z = importlib.import_module("zk-research-implementations_b200")
from oracle import c_oracle as O

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
fid, p = 0, z.engine.MODULI[0]
rng = random.Random(5)
ctx = z.Context(fid, 0, z.MODE_FULL)
S, T = z.sum_check_protocol, z.fiat_shamir.Transcript
cases = []
# (the shapes from 2^17 entries up run their large rounds on the tensor cores: TMA / MMA / TMEM pipelines, Gram drains)
for P, D, n in ((1, 2, 4), (1, 2, 11), (1, 2, 14), (2, 2, 12), (1, 3, 10), (1, 2, 17), (2, 3, 13), (1, 2, 19), (1, 3, 18), (2, 2, 18), (2, 3, 17)):
    tabs = [O.synth_table(fid, 1000 + n, t, n) for t in range(P * D)]
    ref = O.gkr_sumcheck_prove(O.Transcript(fid), 1, P, D, tabs)
    mont = [z.engine.to_mont(fid, t) for t in tabs]
    polys = [z.MultilinearPoly.from_montgomery(ctx, m) for m in mont]
    sp = z.SumPoly(ctx, [z.ProductPoly.from_polys(ctx, polys[q * D:(q + 1) * D]) for q in range(P)])
    raw = S.RawGkrProver(sp)
    want_c = np.zeros_like(raw.coeffs)
    for k, c in enumerate(ref["coeffs"]):
        if c:
            want_c[k, :len(c)] = z.engine.to_mont(fid, O.ints_to_arr(c))
    want_f = z.engine.to_mont(fid, O.ints_to_arr(ref["final_vals"]))
    cases.append((raw, want_c, want_f, (P, D, n)))
# GKR circuit too
L = 9
gates = [1 << (L - 1 - l) for l in range(L)]
ops = [[rng.randrange(2) for _ in range(g)] for g in gates]
gin = O.synth_table(fid, 77, 0, L)
circ = z.gkr_circuit.Circuit(ctx, [[z.Operation(o) for o in layer] for layer in ops])
gref = O.gkr_prove(fid, gates, np.array([o for l in ops for o in l], dtype=np.uint8), gin)
gp = z.gkr_protocol.RawGkrProver(circ, z.engine.to_mont(fid, gin))
gp.prove()
g_want = gp.coeffs.copy()
flat = [c for layer in gref["proof_polynomials"] for c in layer]
for k, c in enumerate(flat):
    assert ctx.unmont(g_want[k, :len(c)]) == c, "gkr reference mismatch"
t0 = time.time()
n_proofs = bad = 0
regimes = [(40, 200 * 1024), (40, 0), (0, 0), (40, 4096), (14, 200 * 1024)]
while time.time() - t0 < secs:
    tail, small = regimes[n_proofs % len(regimes)]
    ctx.set_tail_threshold(tail)
    ctx.set_small_threshold(small)
    raw, want_c, want_f, shape = cases[rng.randrange(len(cases))]
    raw.coeffs[:] = 0
    raw.prove(T(fid))
    if not (np.array_equal(raw.coeffs, want_c) and np.array_equal(raw.fin, want_f)):
        bad += 1
        print("MISMATCH", shape, tail, small, flush=True)
    if n_proofs % 7 == 0:
        gp.coeffs[:] = 0
        gp.prove()
        if not np.array_equal(gp.coeffs, g_want):
            bad += 1
            print("GKR MISMATCH", tail, small, flush=True)
    n_proofs += 1
print(f"STRESS proofs={n_proofs} mismatches={bad} in {time.time() - t0:.1f}s")
sys.exit(1 if bad else 0)
