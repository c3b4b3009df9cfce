#!/usr/bin/env python
"""BASELINE configs[4]: MultilinearPoly partial_evaluate / evaluate sweep over 2^16 .. 2^30 entries on one GPU,
against the HBM roofline (algorithmic bytes: fold 48 N, evaluate 32 N) and the multiplier
roofline (fold: 80 wide MACs per output entry; evaluate: 80 per input entry)."""
import ctypes as C, importlib, json, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
z = importlib.import_module("zk-research-implementations_b200")
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 30
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
ctx = z.Context(0, 0, 0)
L, p = z.engine.lib(), z.engine.MODULI[0]
rng = random.Random(3)
rows = []
for n in range(lo, hi + 1, 2):
    m = z.MultilinearPoly.generate(ctx, 11, 0, n)
    rs = [rng.randrange(p) for _ in range(n)]
    arr = ctx.mont(rs)
    out = (C.c_uint64 * 4)()
    h = C.c_uint64()
    reps = 5 if n <= 26 else 2
    def fold():
        z.engine._ck(ctx, L.zkb_mle_partial_evaluate(ctx.handle, m.handle, 0, z.engine._p(arr[:1].copy()), C.byref(h)))
        L.zkb_mle_free(ctx.handle, h.value)
    def evaluate():
        z.engine._ck(ctx, L.zkb_mle_evaluate(ctx.handle, m.handle, z.engine._p(arr), n, out))
    res = {"n_vars": n, "entries": 1 << n}
    for name, fn, alg in (("partial_evaluate", fold, 48.0 * (1 << n)), ("evaluate", evaluate, 32.0 * (1 << n))):
        fn(); ctx.sync()
        ctx.profile(True)
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        ctx.sync()
        wall = (time.perf_counter() - t0) / reps
        prof = ctx.profile_read(); ctx.profile(False)
        kms = sum(v[1] for v in prof.values()) / reps
        res[name] = {"kernel_ms": round(kms, 4), "api_ms": round(wall * 1e3, 4), "GBps_kernel": round(alg / (kms * 1e-3) / 1e9, 1),
                     "frac_hbm": round(alg / (kms * 1e-3) / 1e9 / peaks, 3), "launches": sum(v[0] for v in prof.values()) // reps}
    rows.append(res)
    print(json.dumps(res), flush=True)
    m.free()
