"""Pins BOTH oracles (oracle/pyref.py and oracle/zk_oracle.c) against every
known answer the reference's own unit tests hold for the hot path
(SURVEY.md section 4).  Citations are paths under /root/reference/."""
import numpy as np
import pytest

from oracle import pyref as R
from oracle.c_oracle import arr_to_ints, ints_to_arr

FQ, FR, BLS = R.BN254_FQ, R.BN254_FR, R.BLS12_381_FR
FQ_ID, FR_ID, BLS_ID = 1, 0, 2


def test_keccak_kats(oracle):
    for msg, hx in [(b"", "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"),
                    (b"abc", "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45")]:
        assert R.keccak256(msg).hex() == hx
        assert oracle.keccak256(msg).hex() == hx
    # multi-block + boundary lengths: C (incremental sponge) vs Python (one-shot)
    for n in (1, 31, 32, 135, 136, 137, 271, 272, 273, 1000):
        msg = bytes((i * 7 + 3) & 0xFF for i in range(n))
        assert oracle.keccak256(msg) == R.keccak256(msg)


def test_partial_evaluate_known_answer(oracle):
    # multilinear_polynomial_evaluation.rs:174-186
    assert R.MultilinearPoly([0, 0, 3, 10], FQ).partial_evaluate(0, 5).evaluation == [15, 50]
    out = oracle.mle_partial_evaluate(FQ_ID, ints_to_arr([0, 0, 3, 10]), 0, 5)
    assert arr_to_ints(out) == [15, 50]


def test_evaluate_known_answer(oracle):
    # multilinear_polynomial_evaluation.rs:189-198
    assert R.MultilinearPoly([0, 0, 3, 10], FQ).evaluate([5, 1]) == 50
    assert oracle.mle_evaluate(FQ_ID, ints_to_arr([0, 0, 3, 10]), [5, 1]) == 50


def test_invalid_evaluations_panics():
    # multilinear_polynomial_evaluation.rs:30
    with pytest.raises(ValueError, match="Invalid evaluations"):
        R.MultilinearPoly([1, 2, 3], FQ)
    with pytest.raises(ValueError, match="Invalid number of values"):
        R.MultilinearPoly([1, 2], FQ).evaluate([1, 2])
    with pytest.raises(ValueError, match="Invalid number of values"):
        R.MultilinearPoly([1, 2], FQ).multi_partial_evaluate([1, 2])


def test_product_poly_known_answers():
    # composed_polynomial.rs:113-155
    pp = R.ProductPoly([[0, 0, 0, 3], [0, 0, 0, 2]], FQ)
    assert pp.evaluate([2, 3]) == 216
    assert [m.evaluation for m in pp.partial_evaluate(2).evaluation] == [[0, 6], [0, 4]]
    # :157-175 should_panic
    with pytest.raises(ValueError, match="all evaluations must have same length"):
        R.ProductPoly([[0, 0, 0, 3], [0, 0, 0, 4, 0, 0, 0, 4]], FQ)


def test_sum_poly_known_answers():
    # composed_polynomial.rs:184-256
    sp = R.SumPoly([R.ProductPoly([[0, 0, 0, 3], [0, 0, 0, 2]], FQ), R.ProductPoly([[0, 0, 0, 4], [0, 0, 0, 5]], FQ)])
    assert sp.evaluate([2, 3]) == 936
    got = [[m.evaluation for m in q.evaluation] for q in sp.partial_evaluate(2).polys]
    assert got == [[[0, 6], [0, 4]], [[0, 8], [0, 10]]]
    with pytest.raises(ValueError, match="all product polys must have same degree"):
        R.SumPoly([R.ProductPoly([[0, 1], [0, 1]], FQ), R.ProductPoly([[0, 1]], FQ)])


def test_univariate_known_answers(oracle):
    # univariate_polynomial_dense.rs:117-196
    assert R.uni_evaluate([3, 4, 3], 3, FQ) == 42
    assert R.uni_interpolate([(0, 2), (1, 4), (2, 6)], FQ) == [2, 2]
    assert oracle.uni_interpolate(FQ_ID, [0, 1, 2], [2, 4, 6]) == [2, 2]
    assert R.uni_interpolate([(0, 0), (1, 0), (2, 0)], FQ) == []
    assert oracle.uni_interpolate(FQ_ID, [0, 1, 2], [0, 0, 0]) == []


def test_get_gkr_round_poly(oracle):
    # sum_check_protocol.rs:225-245: coefficients of interpolate{(0,20),(1,68),(2,156)}
    e = [[0, 3, 2, 5], [0, 6, 4, 10], [0, 1, 1, 2], [0, 2, 2, 4]]
    sp = R.SumPoly([R.ProductPoly(e[:2], FQ), R.ProductPoly(e[2:], FQ)])
    expect = R.uni_interpolate([(0, 20), (1, 68), (2, 156)], FQ)
    assert expect == [20, 28, 20]
    assert R.get_round_partial_polynomial_proof_gkr(sp) == expect
    tr = oracle.Transcript(FQ_ID)
    res = oracle.gkr_sumcheck_prove(tr, 0, 2, 2, [ints_to_arr(x) for x in e])
    assert res["coeffs"][0] == expect and res["evals"][0] == [20, 68, 156]


def test_sumcheck_valid_and_invalid(oracle):
    # sum_check_protocol.rs:194-204 -- constant-10 table (2^20 in the reference; 2^12 in the
    # Python twin to stay fast, the full 2^20 in the C oracle)
    poly = R.MultilinearPoly([10] * (1 << 12), FQ)
    assert R.verify(poly, R.prove(poly))
    tab = ints_to_arr([10]) * np.ones((1 << 20, 1), dtype=np.uint64)
    claimed, msgs, _ = oracle.sumcheck_prove(FQ_ID, tab)
    assert claimed == (10 << 20) % FQ
    assert oracle.sumcheck_verify(FQ_ID, tab, claimed, msgs, redundant_fold=True)
    # sum_check_protocol.rs:207-222 -- forged proof rejected
    bad = R.Proof([[3, 9], [1, 2]], 20)
    assert not R.verify(R.MultilinearPoly([0, 3, 2, 5], FQ), bad)
    assert not oracle.sumcheck_verify(FQ_ID, ints_to_arr([0, 3, 2, 5]), 20, [[3, 9], [1, 2]])


def test_gkr_prover_and_verifier_roundtrip(oracle):
    # sum_check_protocol.rs:247-269
    sp = R.SumPoly([R.ProductPoly([[0, 0, 0, 2], [0, 0, 0, 3]], FQ), R.ProductPoly([[0, 0, 0, 2], [0, 0, 0, 3]], FQ)])
    res = R.gkr_prove(12, sp, R.Transcript(FQ))
    assert R.gkr_verify(res.proof_polynomials, res.claimed_sum, R.Transcript(FQ)).verified
    tabs = [ints_to_arr(x) for x in ([0, 0, 0, 2], [0, 0, 0, 3], [0, 0, 0, 2], [0, 0, 0, 3])]
    c = oracle.gkr_sumcheck_prove(oracle.Transcript(FQ_ID), 0, 2, 2, tabs)
    assert c["coeffs"] == res.proof_polynomials and c["challenges"] == res.random_challenges
    ok, _, _ = oracle.gkr_sumcheck_verify(oracle.Transcript(FQ_ID), c["coeffs"], 12)
    assert ok
    ok, fin, ch = oracle.gkr_sumcheck_verify(oracle.Transcript(FQ_ID), c["coeffs"], 13)
    assert (ok, fin, ch) == (False, 0, [0])  # :129-133


def test_circuit_known_answers(oracle):
    # gkr_circuit.rs:152-186
    c = R.Circuit([[R.MUL] * 4, [R.ADD] * 2, [R.ADD]])
    inp = [5, 2, 2, 4, 10, 0, 3, 3]
    assert c.evaluate(inp, FQ) == [[10, 8, 0, 9], [18, 9], [27]]
    outs = oracle.circuit_evaluate(FQ_ID, [4, 2, 1], np.array([1, 1, 1, 1, 0, 0, 0], dtype=np.uint8), ints_to_arr(inp))
    assert [arr_to_ints(o) for o in outs] == [[10, 8, 0, 9], [18, 9], [27]]
    # :189-202 -- gate outputs [3,12,11,56] = Add(1,2) Mul(3,4) Add(5,6) Mul(7,8)
    c2 = R.Circuit([[R.ADD, R.MUL, R.ADD, R.MUL]])
    assert c2.evaluate([1, 2, 3, 4, 5, 6, 7, 8], FQ) == [[3, 12, 11, 56]]
    # :205-256 -- 1-gate layer: 8-entry indicator with a 1 at index 1 (a=0,b=0,c=1)
    for op, other in ((R.ADD, R.MUL), (R.MUL, R.ADD)):
        lay = R.Layer([op])
        assert lay.get_add_mul_i(op, FQ).evaluation == [0, 1, 0, 0, 0, 0, 0, 0]
        assert lay.get_add_mul_i(other, FQ).evaluation == [0] * 8


def test_tensor_tables():
    # gkr_protocol.rs:363-420
    T = R.MultilinearPoly.tensor_add_mul_polynomials
    assert T([0, 2], [0, 3], R.ADD, BLS).evaluation == [0, 3, 2, 5]
    assert T([0, 3], [0, 0, 0, 2], R.ADD, BLS).evaluation == [0, 0, 0, 2, 3, 3, 3, 5]
    assert T([0, 2], [0, 3], R.MUL, BLS).evaluation == [0, 0, 0, 6]
    assert T([0, 3], [0, 0, 0, 2], R.MUL, BLS).evaluation == [0, 0, 0, 0, 0, 0, 0, 6]


def test_get_fbc_poly():
    # gkr_protocol.rs:423-452: 1 add gate, r=5, w=[2,12]
    fbc = R.get_fbc_poly(5, R.Layer([R.ADD]), [2, 12], [2, 12], BLS)
    got = [[m.evaluation for m in q.evaluation] for q in fbc.polys]
    assert got == [[[0, (-4) % BLS, 0, 0], [4, 14, 14, 24]], [[0, 0, 0, 0], [4, 24, 24, 144]]]


def test_gkr_protocol_roundtrip_and_invalid(oracle):
    # gkr_protocol.rs:474-506 (KZG omitted)
    struct = [[R.ADD] * 4, [R.MUL, R.ADD], [R.ADD]]
    inp = [5, 2, 2, 4, 10, 0, 3, 3]
    proof = R.gkr_protocol_prove_dense(R.Circuit(struct), inp, BLS)
    assert proof.output_poly == [58, 0]
    assert R.gkr_protocol_verify_dense(proof, R.Circuit(struct), inp, BLS)
    assert R.gkr_protocol_prove_sparse(R.Circuit(struct), inp, BLS) == proof
    c = oracle.gkr_prove(BLS_ID, [4, 2, 1], np.array([0, 0, 0, 0, 1, 0, 0], dtype=np.uint8), ints_to_arr(inp))
    assert c["output_poly"] == proof.output_poly
    assert c["proof_polynomials"] == proof.proof_polynomials
    assert c["claimed_evaluations"] == proof.claimed_evaluations
    assert c["final_openings"] == proof.final_openings
    # :509-570 -- dummy proof rejected (every round poly = interpolate{(0,10),(1,5)})
    dummy = R.uni_interpolate([(0, 10), (1, 5)], BLS)
    bad = R.GkrProof([10, 0], [[dummy] * 2, [dummy] * 4], [(10, 5)], (1, 2))
    assert not R.gkr_protocol_verify_dense(bad, R.Circuit([[R.MUL, R.MUL], [R.ADD]]), [1, 1, 1, 1], BLS)


def test_survey_appendix_c_vectors(oracle):
    """Cross-check vectors recorded in SURVEY.md App. C (a Python model written by
    the surveyor, independent of this repo's two oracles)."""
    t = R.Transcript(FQ)
    t.append(b"zero knowledge")
    assert t.get_random_challenge() == 0x020d8026e5dccbca38647e1d8c0b3173d413b6557454aa37235c10ec78521249
    assert t.get_random_challenge() == 0x0d7d17fcf24c3a1c7464e02991aecd8adaafbe08b0a0764432af2ddbab9a173d
    for fid, c1, c2 in ((FR_ID, 0x020d8026e5dccbca38647e1d8c0b31759149bf792f36122704566af81a460761,
                         0x0d7d17fcf24c3a1c7464e02991aecd8c97e5c72c6b81de3413a987e74d8e0c55),
                        (BLS_ID, 0x4fb1129f4105cf28e66bbcef886ebae4de5bbc98161d786d13de4148da460764,
                         0x5b20aa754d753d7b226c1efb8e1256fbe4f7c44b5269447a23315e380d8e0c58)):
        ct = oracle.Transcript(fid)
        ct.append(b"zero knowledge")
        assert ct.challenge() == c1 and ct.challenge() == c2
    claimed, msgs, ch = oracle.sumcheck_prove(FR_ID, ints_to_arr([0, 0, 0, 2, 0, 10, 0, 17]))
    assert claimed == 29 and msgs[0] == [2, 0x1b]
    assert msgs[1] == [0x2812d3d4e2c3bd8fe80078f7ac88d430ae0678f9d7f740e414d0b4a342b7ce5f,
                       0x23ea1685e38ccc42ffd89298420c921a70efc1528716290d7d48142aec13b590]
    assert msgs[2] == [0, 0x0d6102457ee32228e46bd71a5183791b796090a166f7b168f7941207728f806d]
    assert ch == [0x1c34093520df963cd9c1c88d9ece5b003f1a99d6d27571f90a6c400d7eabfb0a,
                  0x1d086f8eb6b4f8183c7a9d1e89c589bc65ae594e2756ac66c17f7bb1acffea4b,
                  0x1e5a288cd215226ea4ebef9fb908c8b6f68ac31cb0f04a16cb53fecc37b338e7]
    assert oracle.mle_evaluate(FR_ID, ints_to_arr([0, 0, 0, 2, 0, 10, 0, 17]), ch) == \
        0x12e654089c1043a57f3674eba041c4d32c8c79e64f5f44adf5ca87d86aeb4a80
    proof = R.gkr_protocol_prove_dense(R.Circuit([[R.ADD] * 4, [R.MUL, R.ADD], [R.ADD]]), [5, 2, 2, 4, 10, 0, 3, 3], BLS)
    assert [[len(x) for x in l] for l in proof.proof_polynomials] == [[3, 3], [3, 3, 2, 3], [3, 3, 3, 3, 3, 3]]
    blob = R.fq_vec_to_bytes(proof.output_poly) + b"".join(R.fq_vec_to_bytes(c) for l in proof.proof_polynomials for c in l) \
        + b"".join(R.fq_vec_to_bytes(list(e)) for e in proof.claimed_evaluations) + R.fq_vec_to_bytes(list(proof.final_openings))
    assert len(blob) == 1376
    assert R.keccak256(blob).hex() == "f36dd78dc8541b9e70d074028d90a5f35623682be6076b5a6e3ca2e7d9526e5d"
