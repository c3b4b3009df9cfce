"""k_sc_small as ONE THREAD-BLOCK CLUSTER (kernels.cuh: tables sharded over the CTAs' shared memory on the low index bits,
partial sums over distributed shared memory, collection into CTA 0 for the last rounds) against the CPU oracle and against
the same library restricted to a single CTA (ZKB200_CLUSTER_MAX=1), bit for bit; sizes from the first one a cluster takes
(2 CTAs) to the largest (16 CTAs), products of 2..4 factors, sums of products, round 0 inside the kernel, all three fields."""
import os

import pytest

from oracle import pyref as R

pytestmark = pytest.mark.gpu

FIELDS = [(0, R.BN254_FR), (1, R.BN254_FQ), (2, R.BLS12_381_FR)]


@pytest.fixture(scope="module")
def ctx_pairs(zkb):
    cache = {}

    def get(fid, mode=1):
        if (fid, mode) not in cache:
            cl = zkb.Context(fid, 0, mode)
            os.environ["ZKB200_CLUSTER_MAX"] = "1"
            try:
                one = zkb.Context(fid, 0, mode)
            finally:
                del os.environ["ZKB200_CLUSTER_MAX"]
            cache[(fid, mode)] = (cl, one)
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
@pytest.mark.parametrize("P,D", [(1, 2), (2, 3), (1, 4), (5, 2), (1, 3)])
@pytest.mark.parametrize("n", [11, 12, 13, 14, 15, 16])
def test_cluster_proof_equals_oracle_and_single_cta(zkb, ctx_pairs, oracle, fid, p, P, D, n):
    cl, one = ctx_pairs(fid)
    seed = 0xB200C100 + 16 * P + D
    got = []
    for ctx in (cl, one):
        tabs = [zkb.MultilinearPoly.generate(ctx, seed, t, n) for t in range(P * D)]
        got.append(prove(zkb, ctx, fid, tabs, P, D))
        for t in tabs:
            t.free()
    oracle.set_threads(max(1, len(os.sched_getaffinity(0))))
    ref = oracle.gkr_sumcheck_prove(oracle.Transcript(fid), 1, P, D, [oracle.synth_table(fid, seed, t, n) for t in range(P * D)])
    assert got[0] == got[1], "cluster proof differs from the single-CTA proof"
    assert got[0] == (ref["coeffs"], ref["challenges"], ref["final_vals"])


def test_cluster_is_used(zkb):
    """The device places a cluster of k_sc_small (otherwise the tests above compare the single-CTA kernel with itself)."""
    ctx = zkb.Context(0, 0, 1)
    try:
        assert ctx.small_cluster_max() >= 8
    finally:
        ctx.close()
