"""The committed golden proofs (tests/golden/proofs.json, made by tests/golden/make_golden.py from the
pure-Python restatement, GKR by the reference's DENSE construction) against the C oracle and against the
host-resident code of libzkb200 (verifier, transcript).  No GPU needed."""
import json
import os

import numpy as np
import pytest

from oracle.c_oracle import ints_to_arr

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proofs.json")))


def ints(v):
    return [int(x, 16) for x in v]


def test_plain_sumcheck_golden(oracle):
    for g in GOLD["plain"]:
        claimed, msgs, ch = oracle.sumcheck_prove(g["field"], ints_to_arr(ints(g["table"])))
        assert claimed == int(g["claimed_sum"], 16)
        assert msgs == [ints(m) for m in g["msgs"]] and ch == ints(g["challenges"])


def test_composed_sumcheck_golden(oracle, zkb):
    U = zkb.univariate_polynomial.UnivariatePoly
    for g in GOLD["composed"]:
        fid = g["field"]
        mode = 0 if g["mode"] == "compat" else 1
        tabs = [ints_to_arr(ints(t)) for t in g["tables"]]
        pr = oracle.gkr_sumcheck_prove(oracle.Transcript(fid), mode, g["P"], g["D"], tabs)
        assert pr["coeffs"] == [ints(c) for c in g["coeffs"]] and pr["challenges"] == ints(g["challenges"])
        # the library's host verifier replays the same transcript: challenges must match the stored ones
        p = zkb.engine.MODULI[fid]
        polys = [U(ints(c), fid) for c in g["coeffs"]]
        claim = (polys[0].evaluate(0) + polys[0].evaluate(1)) % p
        v = zkb.sum_check_protocol.gkr_verify(polys, claim, zkb.fiat_shamir.Transcript(fid))
        assert v.verified and v.random_challenges == ints(g["challenges"])


def test_gkr_golden(oracle):
    for g in GOLD["gkr"]:
        flat = np.array([o for layer in g["ops"] for o in layer], dtype=np.uint8)
        pr = oracle.gkr_prove(g["field"], g["gates"], flat, ints_to_arr(ints(g["inputs"])))
        assert pr["output_poly"] == ints(g["output_poly"])
        assert pr["proof_polynomials"] == [[ints(c) for c in layer] for layer in g["proof_polynomials"]]
        assert [list(x) for x in pr["claimed_evaluations"]] == [ints(c) for c in g["claimed_evaluations"]]
        assert list(pr["final_openings"]) == ints(g["final_openings"])


def test_wired_gkr_golden():
    """General-wiring proofs (tests/golden/wired_proofs.json, made from the dense general-index construction):
    the oracle's two-phase form reproduces the stored bytes and its verifier accepts them."""
    from oracle import pyref as R

    W = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wired_proofs.json")))["wired"]
    assert len(W) == 12
    for g in W:
        p = R.MODULI_BY_ID[g["field"]]
        layers, w = [], g["n_inputs"]
        for l in g["layers"]:
            layers.append(R.WiredLayer(l["ops"], l["in1"], l["in2"], w))
            w = len(l["ops"])
        circ, inputs = R.WiredCircuit(layers), ints(g["inputs"])
        pr = R.wired_prove_sparse(circ, inputs, p)
        assert pr.output_poly == ints(g["output_poly"])
        assert pr.proof_polynomials == [[ints(c) for c in layer] for layer in g["proof_polynomials"]]
        assert [list(x) for x in pr.claimed_evaluations] == [ints(c) for c in g["claimed_evaluations"]]
        assert list(pr.final_openings) == ints(g["final_openings"])
        assert R.wired_verify_sparse(pr, circ, inputs, p)
