"""The two independent restatements (oracle/pyref.py, Python ints; and
oracle/zk_oracle.c, 4x64 Montgomery) must agree everywhere."""
import random

import numpy as np
import pytest

from oracle import pyref as R
from oracle.c_oracle import arr_to_ints, ints_to_arr

FIELDS = [(0, R.BN254_FR), (1, R.BN254_FQ), (2, R.BLS12_381_FR)]


def edge_values(p):
    return [0, 1, 2, p - 1, p - 2, (1 << 256) % p, ((1 << 256) % p) - 1, p >> 1, (p >> 1) + 1, (1 << 64) - 1, 1 << 64,
            (1 << 128) - 1, (1 << 192) + 12345]


@pytest.mark.parametrize("fid,p", FIELDS)
def test_field_ops(oracle, fid, p):
    rng = random.Random(100 + fid)
    a = edge_values(p) + [rng.randrange(p) for _ in range(200)]
    b = list(reversed(edge_values(p))) + [rng.randrange(p) for _ in range(200)]
    A, B = ints_to_arr(a), ints_to_arr(b)
    assert arr_to_ints(oracle.vec_op(fid, 0, A, B)) == [(x + y) % p for x, y in zip(a, b)]
    assert arr_to_ints(oracle.vec_op(fid, 1, A, B)) == [(x - y) % p for x, y in zip(a, b)]
    assert arr_to_ints(oracle.vec_op(fid, 2, A, B)) == [(x * y) % p for x, y in zip(a, b)]
    M = oracle.to_mont(fid, A)
    assert arr_to_ints(M) == [R.to_mont(x, p) for x in a]
    assert arr_to_ints(oracle.from_mont(fid, M)) == a


@pytest.mark.parametrize("fid,p", FIELDS)
def test_fold_all_bits_and_evaluate(oracle, fid, p):
    rng = random.Random(7 + fid)
    for n in (1, 2, 3, 5, 8):
        tab = [rng.randrange(p) for _ in range(1 << n)]
        m = R.MultilinearPoly(tab, p)
        for bit in range(n):
            r = rng.randrange(p)
            got = arr_to_ints(oracle.mle_partial_evaluate(fid, ints_to_arr(tab), bit, r))
            assert got == m.partial_evaluate(bit, r).evaluation
        rs = [rng.randrange(p) for _ in range(n)]
        assert oracle.mle_evaluate(fid, ints_to_arr(tab), rs) == m.evaluate(rs)
    assert oracle.mle_evaluate(fid, ints_to_arr([7]), []) == 7  # 0 variables


@pytest.mark.parametrize("fid,p", FIELDS)
def test_plain_sumcheck(oracle, fid, p):
    rng = random.Random(21 + fid)
    for n in (1, 2, 4, 9):
        tab = [rng.randrange(p) for _ in range(1 << n)]
        ref = R.prove(R.MultilinearPoly(tab, p))
        claimed, msgs, ch = oracle.sumcheck_prove(fid, ints_to_arr(tab))
        assert (claimed, msgs, ch) == (ref.claimed_sum, ref.proof_polynomials, ref.challenges)
        assert oracle.sumcheck_verify(fid, ints_to_arr(tab), claimed, msgs, redundant_fold=True)
        msgs[-1][0] = (msgs[-1][0] + 1) % p
        assert not oracle.sumcheck_verify(fid, ints_to_arr(tab), claimed, msgs)


@pytest.mark.parametrize("fid,p", FIELDS)
@pytest.mark.parametrize("mode,P,d", [(0, 2, 2), (0, 3, 3), (1, 1, 1), (1, 1, 2), (1, 2, 2), (1, 1, 3), (1, 2, 3), (1, 3, 4)])
def test_composed_sumcheck(oracle, fid, p, mode, P, d):
    rng = random.Random(1000 * fid + 100 * mode + 10 * P + d)
    for n in (1, 3, 6):
        tabs = [[rng.randrange(p) for _ in range(1 << n)] for _ in range(P * d)]
        sp = R.SumPoly([R.ProductPoly(tabs[q * d:(q + 1) * d], p) for q in range(P)])
        m = "compat" if mode == 0 else "full"
        ref = R.gkr_prove(5, sp, R.Transcript(p), m)
        got = oracle.gkr_sumcheck_prove(oracle.Transcript(fid), mode, P, d, [ints_to_arr(t) for t in tabs])
        assert got["coeffs"] == ref.proof_polynomials
        assert got["challenges"] == ref.random_challenges
        if mode == 1:
            assert got["evals"][0] == R.round_evals_full([tabs[q * d:(q + 1) * d] for q in range(P)], p)
            claim = (got["evals"][0][0] + got["evals"][0][1]) % p
            ok, fin, ch = oracle.gkr_sumcheck_verify(oracle.Transcript(fid), got["coeffs"], claim)
            assert ok and ch == ref.random_challenges
            prod = 0
            for q in range(P):
                m_ = 1
                for f in range(d):
                    m_ = m_ * got["final_vals"][q * d + f] % p
                prod = (prod + m_) % p
            assert fin == prod == sp.evaluate(ref.random_challenges)


def test_compat_needs_two_products(oracle):
    # SumPoly::reduce indexes polys[1] (composed_polynomial.rs:90) -> panic with one product
    sp = R.SumPoly([R.ProductPoly([[1, 2], [3, 4]], R.BN254_FR)])
    with pytest.raises(IndexError):
        R.gkr_prove(0, sp, R.Transcript(R.BN254_FR), "compat")
    with pytest.raises(ValueError):
        oracle.gkr_sumcheck_prove(oracle.Transcript(0), 0, 1, 2, [ints_to_arr([1, 2]), ints_to_arr([3, 4])])


def random_circuit(rng, depth, out_gates):
    struct, G = [], out_gates
    for _ in range(depth):
        struct.insert(0, [rng.choice([R.ADD, R.MUL]) for _ in range(G)])
        G *= 2
    return struct, G


@pytest.mark.parametrize("fid,p", FIELDS)
def test_gkr_sparse_equals_dense(oracle, fid, p):
    rng = random.Random(55 + fid)
    for depth, outg in [(1, 1), (1, 2), (2, 1), (2, 2), (3, 1), (3, 2), (4, 1)]:
        struct, nin = random_circuit(rng, depth, outg)
        inp = [rng.randrange(p) for _ in range(nin)]
        dense = R.gkr_protocol_prove_dense(R.Circuit(struct), inp, p)
        sparse = R.gkr_protocol_prove_sparse(R.Circuit(struct), inp, p)
        assert dense == sparse
        assert R.gkr_protocol_verify_dense(dense, R.Circuit(struct), inp, p)
        assert R.gkr_protocol_verify_sparse(dense, R.Circuit(struct), inp, p)
        ops = np.array([o for l in struct for o in l], dtype=np.uint8)
        c = oracle.gkr_prove(fid, [len(l) for l in struct], ops, ints_to_arr(inp))
        assert c["output_poly"] == dense.output_poly
        assert c["proof_polynomials"] == dense.proof_polynomials
        assert c["claimed_evaluations"] == dense.claimed_evaluations
        assert c["final_openings"] == dense.final_openings
        # tampering is rejected
        bad = R.GkrProof(list(dense.output_poly), [[list(c_) for c_ in l] for l in dense.proof_polynomials],
                         list(dense.claimed_evaluations), dense.final_openings)
        bad.proof_polynomials[-1][0] = [(x + 1) % p for x in bad.proof_polynomials[-1][0]] or [1]
        assert not R.gkr_protocol_verify_sparse(bad, R.Circuit(struct), inp, p)


def test_gkr_larger_sparse_c_vs_python(oracle):
    rng = random.Random(99)
    p = R.BN254_FR
    struct, nin = random_circuit(rng, 7, 2)  # 256 inputs; dense tables would need 2^23 entries
    inp = [rng.randrange(p) for _ in range(nin)]
    sparse = R.gkr_protocol_prove_sparse(R.Circuit(struct), inp, p)
    assert R.gkr_protocol_verify_sparse(sparse, R.Circuit(struct), inp, p)
    ops = np.array([o for l in struct for o in l], dtype=np.uint8)
    c = oracle.gkr_prove(0, [len(l) for l in struct], ops, ints_to_arr(inp))
    assert c["proof_polynomials"] == sparse.proof_polynomials
    assert c["final_openings"] == sparse.final_openings


@pytest.mark.parametrize("fid", [0, 1, 2])
def test_synthetic_tables(oracle, fid):
    p = R.MODULI_BY_ID[fid]
    t = arr_to_ints(oracle.synth_table(fid, 0xB2000002, 3, 6))
    assert t == R.synth_table(0xB2000002, 3, 6, fid)
    assert all(v < p for v in t)
    # strided access = what a low-bit shard of rank 1 of 4 holds
    s = arr_to_ints(oracle.synth_table(fid, 0xB2000002, 3, 6, first=1, stride=4, count=16))
    assert s == t[1::4]


def test_openmp_matches_single_thread(oracle):
    p = R.BN254_FR
    tabs = [oracle.synth_table(0, 5, t, 13) for t in range(4)]
    oracle.set_threads(1)
    a = oracle.gkr_sumcheck_prove(oracle.Transcript(0), 1, 2, 2, tabs)
    a1 = oracle.sumcheck_prove(0, tabs[0])
    oracle.set_threads(4)
    b = oracle.gkr_sumcheck_prove(oracle.Transcript(0), 1, 2, 2, tabs)
    b1 = oracle.sumcheck_prove(0, tabs[0])
    oracle.set_threads(1)
    assert a == b and a1 == b1


# ------------------------------------------------------------- general wiring (extension, SURVEY 8d C3 ii)
def random_wired(rng, n_inputs, gates_per_layer):
    layers, w = [], n_inputs
    for G in gates_per_layer:
        layers.append(R.WiredLayer([rng.choice([R.ADD, R.MUL]) for _ in range(G)], [rng.randrange(w) for _ in range(G)],
                                   [rng.randrange(w) for _ in range(G)], w))
        w = G
    return R.WiredCircuit(layers)


def proof_fields(pr):
    return (pr.output_poly, pr.proof_polynomials, pr.claimed_evaluations, pr.final_openings)


@pytest.mark.parametrize("fid,p", FIELDS)
def test_wired_reduces_to_reference_wiring(fid, p):
    """With in1 = 2g, in2 = 2g+1 and <= 2 outputs the general-wiring prover IS the reference's."""
    rng = random.Random(77 + fid)
    for depth, outg in [(1, 1), (1, 2), (2, 1), (3, 2), (4, 1)]:
        struct, nin = random_circuit(rng, depth, outg)
        inp = [rng.randrange(p) for _ in range(nin)]
        ref = R.gkr_protocol_prove_dense(R.Circuit(struct), inp, p)
        wc = R.WiredCircuit.binary_tree(struct)
        assert proof_fields(R.wired_prove_dense(wc, inp, p)) == proof_fields(ref)
        assert proof_fields(R.wired_prove_sparse(wc, inp, p)) == proof_fields(ref)
        assert R.wired_verify_sparse(ref, wc, inp, p)


@pytest.mark.parametrize("fid,p", FIELDS)
def test_wired_sparse_equals_dense(fid, p):
    rng = random.Random(88 + fid)
    for nin, gates in [(2, [1]), (2, [4]), (4, [4, 4]), (8, [4, 8, 2]), (4, [8, 4, 4, 1]), (8, [8, 8])]:
        wc = random_wired(rng, nin, gates)
        inp = [rng.randrange(p) for _ in range(nin)]
        dense = R.wired_prove_dense(wc, inp, p)
        sparse = R.wired_prove_sparse(wc, inp, p)
        assert proof_fields(dense) == proof_fields(sparse), (nin, gates)
        assert len(dense.output_poly) == max(gates[-1], 2)
        assert R.wired_verify_sparse(sparse, wc, inp, p)
        bad = R.GkrProof(list(sparse.output_poly), [[list(c_) for c_ in l] for l in sparse.proof_polynomials],
                         list(sparse.claimed_evaluations), sparse.final_openings)
        bad.output_poly[0] = (bad.output_poly[0] + 1) % p
        assert not R.wired_verify_sparse(bad, wc, inp, p)
        bad_inp = list(inp)
        bad_inp[-1] = (bad_inp[-1] + 1) % p
        assert not R.wired_verify_sparse(sparse, wc, bad_inp, p)


@pytest.mark.parametrize("fid,p", FIELDS)
def test_wired_c_oracle_matches_python(oracle, fid, p):
    """The two restatements of the general-wiring prover (C two-phase, Python two-phase == Python dense) agree."""
    rng = random.Random(99 + fid)
    for nin, gates in [(2, [1]), (2, [4]), (8, [4, 8, 2]), (4, [8, 4, 4, 1]), (64, [128, 32, 64]), (256, [256, 256])]:
        wc = random_wired(rng, nin, gates)
        inp = [rng.randrange(p) for _ in range(nin)]
        ref = R.wired_prove_sparse(wc, inp, p)
        c = oracle.gkr_prove_wired(fid, nin, [(l.ops, l.in1, l.in2) for l in wc.layers], ints_to_arr(inp))
        assert c["output_poly"] == ref.output_poly and c["proof_polynomials"] == ref.proof_polynomials
        assert c["claimed_evaluations"] == ref.claimed_evaluations and c["final_openings"] == ref.final_openings
    # and on the reference's own wiring the C wired prover is the C tree prover
    struct, nin = random_circuit(rng, 5, 2)
    inp = [rng.randrange(p) for _ in range(nin)]
    ops = np.array([o for l in struct for o in l], dtype=np.uint8)
    a = oracle.gkr_prove(fid, [len(l) for l in struct], ops, ints_to_arr(inp))
    b = oracle.gkr_prove_wired(fid, nin, [(l, [2 * g for g in range(len(l))], [2 * g + 1 for g in range(len(l))]) for l in struct], ints_to_arr(inp))
    assert a == b
