"""The CUDA path against the committed golden proofs, in every round-driver regime (one launch per round,
persistent cooperative kernel, shared-memory kernel) -- the stored bytes do not depend on the regime."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proofs.json")))
REGIMES = [(0, 0), (40, 0), (0, 200 * 1024), (40, 200 * 1024), (40, 2048)]  # (tail_log2, small_bytes)


def ints(v):
    return [int(x, 16) for x in v]


@pytest.fixture(params=REGIMES, ids=lambda r: f"tail{r[0]}_small{r[1]}")
def regime(request, ctxs):
    tail, small = request.param
    used = []

    def get(field, mode):
        c = ctxs(field, mode)
        c.set_tail_threshold(tail)
        c.set_small_threshold(small)
        used.append(c)
        return c

    yield get
    for c in used:
        c.set_tail_threshold(40)
        c.set_small_threshold(200 * 1024)


def test_plain_golden(zkb, regime):
    for g in GOLD["plain"]:
        ctx = regime(g["field"], 0)
        pr = zkb.sum_check_protocol.prove(zkb.MultilinearPoly(ctx, ints(g["table"])))
        assert pr.claimed_sum == int(g["claimed_sum"], 16)
        assert pr.proof_polynomials == [ints(m) for m in g["msgs"]] and pr.random_challenges == ints(g["challenges"])


def test_composed_golden(zkb, regime):
    for g in GOLD["composed"]:
        ctx = regime(g["field"], 0 if g["mode"] == "compat" else 1)
        D = g["D"]
        tabs = [ints(t) for t in g["tables"]]
        sp = zkb.SumPoly(ctx, [zkb.ProductPoly(ctx, tabs[q * D:(q + 1) * D]) for q in range(g["P"])])
        pr = zkb.sum_check_protocol.gkr_prove(0, sp, zkb.fiat_shamir.Transcript(g["field"]))
        assert [q.coefficients for q in pr.proof_polynomials] == [ints(c) for c in g["coeffs"]]
        assert pr.random_challenges == ints(g["challenges"])
        sp.free()


def test_gkr_golden(zkb, regime):
    for g in GOLD["gkr"]:
        ctx = regime(g["field"], 0)
        c = zkb.gkr_circuit.Circuit(ctx, [[zkb.Operation(o) for o in layer] for layer in g["ops"]])
        pr = zkb.gkr_protocol.prove(c, ints(g["inputs"]))
        assert pr.output_poly == ints(g["output_poly"])
        assert [[q.coefficients for q in layer] for layer in pr.proof_polynomials] == [[ints(x) for x in layer] for layer in g["proof_polynomials"]]
        assert [list(x) for x in pr.claimed_evaluations] == [ints(x) for x in g["claimed_evaluations"]]
        assert list(pr.final_openings) == ints(g["final_openings"])
        assert zkb.gkr_protocol.verify(pr, c, ints(g["inputs"]))
        c.free()


def test_wired_gkr_golden(zkb, regime):
    """zkb_gkr_prove_wired against the committed general-wiring proofs, in every round-driver regime."""
    W = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wired_proofs.json")))["wired"]
    for g in W:
        ctx = regime(g["field"], 0)
        c = zkb.gkr_circuit.WiredCircuit(ctx, g["n_inputs"], [(l["ops"], l["in1"], l["in2"]) for l in g["layers"]])
        pr = zkb.gkr_protocol.prove_wired(c, ints(g["inputs"]))
        assert pr.output_poly == ints(g["output_poly"])
        assert [[q.coefficients for q in layer] for layer in pr.proof_polynomials] == [[ints(x) for x in layer] for layer in g["proof_polynomials"]]
        assert [list(x) for x in pr.claimed_evaluations] == [ints(x) for x in g["claimed_evaluations"]]
        assert list(pr.final_openings) == ints(g["final_openings"])
        assert zkb.gkr_protocol.verify_wired(pr, c, ints(g["inputs"]))
        c.free()


def test_regimes_agree_on_larger_tables(zkb, ctxs, oracle):
    """n = 15, 2 x 2 and 1 x 3 shapes: every regime gives the oracle's proof."""
    import random

    from oracle import pyref as R
    from oracle.c_oracle import ints_to_arr

    fid, p = 0, R.BN254_FR
    rng = random.Random(99)
    for mode, P, D in ((0, 2, 2), (1, 1, 3), (1, 1, 2)):
        ctx = ctxs(fid, mode)
        n = 15
        tabs = [[rng.randrange(p) for _ in range(1 << n)] for _ in range(P * D)]
        ref = oracle.gkr_sumcheck_prove(oracle.Transcript(fid), mode, P, D, [ints_to_arr(t) for t in tabs])
        sp = zkb.SumPoly(ctx, [zkb.ProductPoly(ctx, tabs[q * D:(q + 1) * D]) for q in range(P)])
        try:
            for tail, small in REGIMES + [(12, 200 * 1024), (40, 65536)]:
                ctx.set_tail_threshold(tail)
                ctx.set_small_threshold(small)
                pr = zkb.sum_check_protocol.gkr_prove(0, sp, zkb.fiat_shamir.Transcript(fid))
                assert [q.coefficients for q in pr.proof_polynomials] == ref["coeffs"], (mode, P, D, tail, small)
                assert pr.final_values == ref["final_vals"]
        finally:
            ctx.set_tail_threshold(40)
            ctx.set_small_threshold(200 * 1024)
            sp.free()
