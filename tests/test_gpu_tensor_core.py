"""Parity of the tensor-core paths (csrc/tcfold.cuh: folds as u8 x u8 -> s32 tcgen05.mma, sums of products as Gram
matrices, evaluate as eight accumulated products) -- against the CPU oracle and against the CUDA-core kernels of the
same library (ZKB200_NO_TC=1), bit for bit, for all three fields, every instantiated shape, and byte patterns that drive
the column sums to their maximum.  Sizes are the smallest the tensor-core paths take (>= 2^17 entries per table)."""
import os

import pytest

from oracle import pyref as R

pytestmark = pytest.mark.gpu

FIELDS = [(0, R.BN254_FR), (1, R.BN254_FQ), (2, R.BLS12_381_FR)]


@pytest.fixture(scope="module")
def ctx_pairs(zkb):
    """(tensor-core ctx, CUDA-core ctx) per (field, mode); the switch is read when the context is created."""
    cache = {}

    def get(fid, mode=1):
        if (fid, mode) not in cache:
            tc = zkb.Context(fid, 0, mode)
            os.environ["ZKB200_NO_TC"] = "1"
            try:
                cc = zkb.Context(fid, 0, mode)
            finally:
                del os.environ["ZKB200_NO_TC"]
            cache[(fid, mode)] = (tc, cc)
        return cache[(fid, mode)]

    yield get
    for a, b in cache.values():
        a.close()
        b.close()


def prove(zkb, ctx, fid, tabs, P, D):
    sp = zkb.SumPoly(ctx, [zkb.ProductPoly.from_polys(ctx, tabs[q * D:(q + 1) * D]) for q in range(P)])
    pr = zkb.sum_check_protocol.gkr_prove(0, sp, zkb.fiat_shamir.Transcript(fid))
    sp.free()
    return [q.coefficients for q in pr.proof_polynomials], pr.random_challenges, pr.final_values


@pytest.mark.parametrize("fid,p", FIELDS)
@pytest.mark.parametrize("P,D,n", [(1, 2, 18), (2, 2, 18), (1, 3, 18), (2, 3, 18), (3, 2, 17), (5, 3, 17)])
def test_tc_proof_equals_oracle_and_cuda_cores(zkb, ctx_pairs, oracle, fid, p, P, D, n):
    tc, cc = ctx_pairs(fid)
    seed = 0xB2007C00 + 16 * P + D
    got = []
    for ctx in (tc, cc):
        tabs = [zkb.MultilinearPoly.generate(ctx, seed, t, n) for t in range(P * D)]
        got.append(prove(zkb, ctx, fid, tabs, P, D))
        for t in tabs:
            t.free()
    oracle.set_threads(max(1, len(os.sched_getaffinity(0))))
    ref = oracle.gkr_sumcheck_prove(oracle.Transcript(fid), 1, P, D, [oracle.synth_table(fid, seed, t, n) for t in range(P * D)])
    assert got[0] == got[1], "tensor-core proof differs from the CUDA-core proof"
    assert got[0] == (ref["coeffs"], ref["challenges"], ref["final_vals"])
    assert tc.launch_count > 0


@pytest.mark.parametrize("fid,p", FIELDS)
def test_tc_compat_shape(zkb, ctx_pairs, oracle, fid, p):
    """compat mode with 3 declared factors: only factors 0, 1 of products 0, 1 enter the round polynomial, evaluated at 4
    points (composed_polynomial.rs:52-54,88-99) -- the (D = 2, 4 points) instantiation of the tensor-core kernels."""
    tc, cc = ctx_pairs(fid, 0)
    n, P, D, seed = 17, 2, 3, 0xB2007D00
    got = []
    for ctx in (tc, cc):
        tabs = [zkb.MultilinearPoly.generate(ctx, seed, t, n) for t in range(P * D)]
        got.append(prove(zkb, ctx, fid, tabs, P, D))
        for t in tabs:
            t.free()
    ref = oracle.gkr_sumcheck_prove(oracle.Transcript(fid), 0, P, D, [oracle.synth_table(fid, seed, t, n) for t in range(P * D)])
    assert got[0] == got[1] == (ref["coeffs"], ref["challenges"], ref["final_vals"])


@pytest.mark.parametrize("fid,p", FIELDS)
def test_tc_extreme_bytes(zkb, ctx_pairs, fid, p):
    """Constant tables whose bytes drive the s32 column sums of the byte-matrix products to their maximum: p - 1 (the largest
    residue), 2^248 - 1 (31 bytes of 0xff) and 0: folds, sums of products and evaluate against plain integer arithmetic."""
    tc, _ = ctx_pairs(fid)
    n = 17
    N = 1 << n
    big = (1 << 248) - 1
    for va, vb, vc in [(p - 1, p - 1, p - 1), (big, p - 1, big), (0, big, p - 1)]:
        tabs = [zkb.MultilinearPoly(tc, [v] * N) for v in (va, vb, vc)]
        for D in (2, 3):
            coeffs, chal, fin = prove(zkb, tc, fid, tabs[:D], 1, D)
            # a constant table stays constant under every fold; round k sums 2^(n-1-k) copies of prod(v) at every point
            prod = 1
            for v in (va, vb, vc)[:D]:
                prod = prod * v % p
            assert fin == [v % p for v in (va, vb, vc)[:D]]
            for k, c in enumerate(coeffs):
                want = prod * (1 << (n - 1 - k)) % p
                assert (c + [0])[0] == want and all(x == 0 for x in c[1:]), (k, D)
        r = [(7 * i + 3) % p for i in range(n)]
        assert tabs[1].evaluate(r) == vb % p
        for t in tabs:
            t.free()


@pytest.mark.parametrize("fid,p", FIELDS)
@pytest.mark.parametrize("n", [17, 19, 20])
def test_tc_evaluate_equals_oracle(zkb, ctx_pairs, oracle, fid, p, n):
    """evaluate (three variables per pass on the tensor cores + the remainder on the CUDA cores) == the CUDA-core chain ==
    the oracle's evaluate of the same synthetic table (multilinear_polynomial_evaluation.rs:79-91)."""
    import random

    tc, cc = ctx_pairs(fid)
    rng = random.Random(n * 31 + fid)
    r = [rng.randrange(p) for _ in range(n)]
    r[0], r[1], r[2] = 0, 1, p - 1
    vals = []
    for ctx in (tc, cc):
        m = zkb.MultilinearPoly.generate(ctx, 0xB2007E00, 0, n)
        vals.append(m.evaluate(r))
        k = 5
        part = m.multi_partial_evaluate(r[:k])
        vals.append(part.evaluate(r[k:]))
        part.free()
        m.free()
    want = oracle.mle_evaluate(fid, oracle.synth_table(fid, 0xB2007E00, 0, n), r)
    assert vals[0] == vals[1] == vals[2] == vals[3] == want
