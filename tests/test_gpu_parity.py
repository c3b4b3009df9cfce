"""Parity tests proper: the CUDA path, called through the C ABI (via the Python
mirror of the reference's crate API), against the CPU oracle on identical
seeded inputs.  Integer/byte work: the bar is BIT-EXACT equality."""
import os
import random

import numpy as np
import pytest

from oracle import pyref as R
from oracle.c_oracle import arr_to_ints, ints_to_arr

pytestmark = pytest.mark.gpu

FIELDS = [(0, R.BN254_FR), (1, R.BN254_FQ), (2, R.BLS12_381_FR)]


def edge_values(p):
    return [0, 1, 2, p - 1, p - 2, (1 << 256) % p, ((1 << 256) % p) - 1, p >> 1, (p >> 1) + 1, (1 << 64) - 1, 1 << 64,
            (1 << 128) - 1, (1 << 192) + 12345, (1 << 32) - 1, 1 << 32, p - (1 << 32)]


def rand_table(rng, p, n, edges=True):
    t = [rng.randrange(p) for _ in range(1 << n)]
    if edges:
        ev = edge_values(p)
        for i in range(min(len(t), len(ev))):
            t[rng.randrange(len(t))] = ev[i]
    return t


# ------------------------------------------------------------ reference known answers on the device
@pytest.mark.parametrize("fid,p", FIELDS)
def test_reference_known_answers_on_device(zkb, ctxs, fid, p):
    ctx = ctxs(fid, zkb.MODE_COMPAT)
    M, PP, SP = zkb.MultilinearPoly, zkb.ProductPoly, zkb.SumPoly
    # multilinear_polynomial_evaluation.rs:174-198
    m = M(ctx, [0, 0, 3, 10])
    assert m.partial_evaluate(0, 5).evaluation == [15, 50]
    assert m.evaluate([5, 1]) == 50
    # composed_polynomial.rs:113-155
    pp = PP(ctx, [[0, 0, 0, 3], [0, 0, 0, 2]])
    assert pp.evaluate([2, 3]) == 216
    assert [q.evaluation for q in pp.partial_evaluate(2).evaluation] == [[0, 6], [0, 4]]
    with pytest.raises(ValueError, match="all evaluations must have same length"):
        PP(ctx, [[0, 0, 0, 3], [0, 2]])
    with pytest.raises(ValueError, match="Invalid evaluations"):
        M(ctx, [1, 2, 3])
    # :184-256
    sp = SP(ctx, [PP(ctx, [[0, 0, 0, 3], [0, 0, 0, 2]]), PP(ctx, [[0, 0, 0, 4], [0, 0, 0, 5]])])
    assert sp.evaluate([2, 3]) == 936
    assert [[q.evaluation for q in pr.evaluation] for pr in sp.partial_evaluate(2).polys] == [[[0, 6], [0, 4]], [[0, 8], [0, 10]]]
    with pytest.raises(ValueError, match="all product polys must have same degree"):
        SP(ctx, [PP(ctx, [[0, 1], [0, 1]]), PP(ctx, [[0, 1]])])
    # sum_check_protocol.rs:225-245: round polynomial [20, 28, 20]
    sp = SP(ctx, [PP(ctx, [[0, 3, 2, 5], [0, 6, 4, 10]]), PP(ctx, [[0, 1, 1, 2], [0, 2, 2, 4]])])
    assert sp.round_evals() == [20, 68, 156]
    pr = zkb.sum_check_protocol.gkr_prove(0, sp, zkb.fiat_shamir.Transcript(fid))
    assert pr.proof_polynomials[0].coefficients == [20, 28, 20]
    # kzg.rs:344-366 open of [0,4,0,4,0,4,3,7] at (6,4,0) -> 72
    assert M(ctx, [0, 4, 0, 4, 0, 4, 3, 7]).evaluate([6, 4, 0]) == 72
    # gkr_protocol.rs:363-452 tensor tables
    w = M(ctx, [2, 12])
    assert M.tensor_add_mul_polynomials(w, w, zkb.Operation.Add).evaluation == [4, 14, 14, 24]
    assert M.tensor_add_mul_polynomials(w, w, zkb.Operation.Mul).evaluation == [4, 24, 24, 144]
    # gkr_circuit.rs:152-186
    Op = zkb.Operation
    c = zkb.gkr_circuit.Circuit(ctx, [[Op.Mul] * 4, [Op.Add] * 2, [Op.Add]])
    assert c.evaluate([5, 2, 2, 4, 10, 0, 3, 3]) == [[10, 8, 0, 9], [18, 9], [27]]


def test_survey_appendix_c_vectors(zkb, ctxs):
    """SURVEY App. C (derived from the Python restatement): plain sumcheck on the bench table, BN254 Fr and Fq."""
    S = zkb.sum_check_protocol
    ctx = ctxs(0, 0)
    pr = S.prove(zkb.MultilinearPoly(ctx, [0, 0, 0, 2, 0, 10, 0, 17]))
    assert pr.claimed_sum == 29
    assert pr.proof_polynomials == [
        [0x2, 0x1B],
        [0x2812D3D4E2C3BD8FE80078F7AC88D430AE0678F9D7F740E414D0B4A342B7CE5F, 0x23EA1685E38CCC42FFD89298420C921A70EFC1528716290D7D48142AEC13B590],
        [0x0, 0x0D6102457EE32228E46BD71A5183791B796090A166F7B168F7941207728F806D]]
    assert pr.random_challenges == [0x1C34093520DF963CD9C1C88D9ECE5B003F1A99D6D27571F90A6C400D7EABFB0A,
                                    0x1D086F8EB6B4F8183C7A9D1E89C589BC65AE594E2756AC66C17F7BB1ACFFEA4B,
                                    0x1E5A288CD215226EA4EBEF9FB908C8B6F68AC31CB0F04A16CB53FECC37B338E7]
    ctx = ctxs(1, 0)
    pr = S.prove(zkb.MultilinearPoly(ctx, [0, 0, 0, 2, 0, 10, 0, 17]))
    assert pr.random_challenges[0] == 0x1C34093520DF963CD9C1C88D9ECE5AFF607F9544F504BE0119EF1307ADB2007E
    # gkr_prove(12, 2 x ([0,0,0,2]*[0,0,0,3])), BN254 Fq (sum_check_protocol.rs:247-261)
    PP = zkb.ProductPoly
    sp = zkb.SumPoly(ctx, [PP(ctx, [[0, 0, 0, 2], [0, 0, 0, 3]]), PP(ctx, [[0, 0, 0, 2], [0, 0, 0, 3]])])
    g = S.gkr_prove(12, sp, zkb.fiat_shamir.Transcript(1))
    assert [q.coefficients for q in g.proof_polynomials] == [
        [0, 0, 0xC], [0, 0, 0x1AC63CC2631E670019DEB1F8E0850F806B71CE2FDEDFE58CB0601B2DA90A1185]]
    assert g.random_challenges == [0x13ACBBB2EF729BC9A1B076B70C7355CD2D6519B11AC85C7F4A20DDD551646F10,
                                   0x044E514C81BDFE4D1706E73834AD93C751FBF397E5B233361DC342722243A751]
    v = S.gkr_verify(g.proof_polynomials, 12, zkb.fiat_shamir.Transcript(1))
    assert v.verified


# ------------------------------------------------------------------------- MultilinearPoly kernels
@pytest.mark.parametrize("fid,p", FIELDS)
def test_mle_kernels(zkb, ctxs, oracle, fid, p):
    ctx = ctxs(fid, 0)
    rng = random.Random(1000 + fid)
    M = zkb.MultilinearPoly
    for n in (1, 2, 3, 5, 9, 12):
        tab = rand_table(rng, p, n)
        A = ints_to_arr(tab)
        m = M(ctx, tab)
        assert m.num_of_vars == n
        assert m.evaluation == tab  # upload -> planar -> canonical download round trip
        assert arr_to_ints(zkb.engine.from_mont(fid, m.montgomery())) == tab
        for bit in sorted(set([0, n - 1, n // 2])):
            r = rng.choice(edge_values(p) + [rng.randrange(p)] * 4)
            assert m.partial_evaluate(bit, r).evaluation == arr_to_ints(oracle.mle_partial_evaluate(fid, A, bit, r))
        rs = [rng.randrange(p) for _ in range(n)]
        assert m.evaluate(rs) == oracle.mle_evaluate(fid, A, rs)
        k = n // 2
        ref = A
        for i in range(k):
            ref = oracle.mle_partial_evaluate(fid, ref, 0, rs[i])
        assert m.multi_partial_evaluate(rs[:k]).evaluation == arr_to_ints(ref)
        s0 = sum(tab[: len(tab) // 2]) % p
        s1 = sum(tab[len(tab) // 2:]) % p
        assert m.sum_halves() == [s0, s1]
        other = rand_table(rng, p, n)
        o = M(ctx, other)
        B = ints_to_arr(other)
        assert (m + o).evaluation == arr_to_ints(oracle.vec_op(fid, 0, A, B))
        assert (m - o).evaluation == arr_to_ints(oracle.vec_op(fid, 1, A, B))
        assert (m * o).evaluation == arr_to_ints(oracle.vec_op(fid, 2, A, B))
        s = rng.randrange(p)
        assert m.scale(s).evaluation == [(x * s) % p for x in tab]
        assert m.clone().evaluation == tab
    with pytest.raises(ValueError, match="Invalid number of values"):
        M(ctx, [1, 2, 3, 4]).evaluate([1])
    with pytest.raises(ValueError, match="Invalid number of values"):
        M(ctx, [1, 2, 3, 4]).multi_partial_evaluate([1, 2, 3])
    assert M(ctx, [7]).evaluate([]) == 7
    a, b = rand_table(rng, p, 3), rand_table(rng, p, 2)
    for op in (zkb.Operation.Add, zkb.Operation.Mul):
        got = M.tensor_add_mul_polynomials(M(ctx, a), M(ctx, b), op).evaluation
        assert got == [op.apply(x, y, p) for x in a for y in b]


@pytest.mark.parametrize("fid,p", FIELDS)
def test_generate_matches_oracle_synth(zkb, ctxs, oracle, fid, p):
    ctx = ctxs(fid, 0)
    for n, seed, tid in ((1, 1, 0), (10, 0xB2000002, 3), (14, 99, 7)):
        m = zkb.MultilinearPoly.generate(ctx, seed, tid, n)
        assert m.evaluation == arr_to_ints(oracle.synth_table(fid, seed, tid, n))


# ---------------------------------------------------------------------------- plain sumcheck
@pytest.mark.parametrize("fid,p", FIELDS)
def test_plain_sumcheck(zkb, ctxs, oracle, fid, p):
    ctx = ctxs(fid, 0)
    S = zkb.sum_check_protocol
    rng = random.Random(2000 + fid)
    for n in (1, 2, 3, 7, 11, 14):
        tab = rand_table(rng, p, n)
        for absorb in (True, False):
            claimed, msgs, ch = oracle.sumcheck_prove(fid, ints_to_arr(tab), absorb_table=absorb)
            m = zkb.MultilinearPoly(ctx, tab)
            pr = S.prove(m, absorb_table=absorb)
            assert (pr.claimed_sum, pr.proof_polynomials, pr.random_challenges) == (claimed, msgs, ch)
            assert S.verify(m, pr, absorb_table=absorb)
            assert oracle.sumcheck_verify(fid, ints_to_arr(tab), pr.claimed_sum, pr.proof_polynomials, absorb_table=absorb)
            bad = S.Proof([list(x) for x in pr.proof_polynomials], pr.claimed_sum)
            bad.proof_polynomials[-1][1] = (bad.proof_polynomials[-1][1] + 1) % p
            assert not S.verify(m, bad, absorb_table=absorb)
    # sum_check_protocol.rs:207-222: forged proof on [0,3,2,5]
    m = zkb.MultilinearPoly(ctx, [0, 3, 2, 5])
    assert not S.verify(m, S.Proof([[3, 7], [1, 2]], 10))
    assert not S.verify(m, S.Proof([[3, 8], [1, 2]], 10))


def test_reference_2pow20_constant_table(zkb, ctxs):
    """sum_check_protocol.rs:194-204: 2^20 entries of 10 over BN254 Fq, prove -> verify == true."""
    ctx = ctxs(1, 0)
    ten = zkb.engine.to_mont(1, zkb.engine.ints_to_limbs([10]))
    m = zkb.MultilinearPoly.from_montgomery(ctx, np.repeat(ten, 1 << 20, axis=0))
    pr = zkb.sum_check_protocol.prove(m)
    assert pr.claimed_sum == 10 << 20
    assert zkb.sum_check_protocol.verify(m, pr)


# --------------------------------------------------------------------- composed sumcheck (gkr_prove)
SHAPES_COMPAT = [(2, 2), (2, 3), (3, 2), (2, 4), (3, 3)]
SHAPES_FULL = [(1, 2), (2, 2), (1, 3), (2, 3), (1, 4), (2, 4), (3, 2), (5, 3)]


@pytest.mark.parametrize("fid,p", FIELDS)
@pytest.mark.parametrize("mode", [0, 1])
def test_composed_sumcheck(zkb, ctxs, oracle, fid, p, mode):
    ctx = ctxs(fid, mode)
    S = zkb.sum_check_protocol
    rng = random.Random(3000 + 10 * fid + mode)
    for P, D in (SHAPES_FULL if mode else SHAPES_COMPAT):
        for n in (1, 2, 3, 6, 10, 13):
            tabs = [rand_table(rng, p, n, edges=(n > 3)) for _ in range(P * D)]
            ref = oracle.gkr_sumcheck_prove(oracle.Transcript(fid), mode, P, D, [ints_to_arr(t) for t in tabs])
            sp = zkb.SumPoly(ctx, [zkb.ProductPoly(ctx, tabs[q * D: (q + 1) * D]) for q in range(P)])
            assert sp.round_evals() == ref["evals"][0]
            pr = S.gkr_prove(0, sp, zkb.fiat_shamir.Transcript(fid))
            assert [q.coefficients for q in pr.proof_polynomials] == ref["coeffs"], (P, D, n)
            assert pr.random_challenges == ref["challenges"]
            assert pr.final_values == ref["final_vals"]
            # the handle is reusable (the reference clones): a second prove gives the same proof
            pr2 = S.gkr_prove(0, sp, zkb.fiat_shamir.Transcript(fid))
            assert [q.coefficients for q in pr2.proof_polynomials] == ref["coeffs"]
            # source tables untouched
            assert sp.polys[0].evaluation[0].evaluation == tabs[0]
            if mode == 1 or (P, D) == (2, 2):
                claim = 0
                for q in range(P):
                    for i in range(1 << n):
                        t = 1
                        for f in range(D):
                            t = t * tabs[q * D + f][i] % p
                        claim += t
                claim %= p
                v = S.gkr_verify(pr.proof_polynomials, claim, zkb.fiat_shamir.Transcript(fid))
                assert v.verified
                fv = pr.final_values
                want = 0
                for q in range(P):
                    t = 1
                    for f in range(D):
                        t = t * fv[q * D + f] % p
                    want += t
                assert v.final_claimed_sum == want % p
            sp.free()


def test_composed_step_api(zkb, ctxs, oracle):
    """round_evals / bind_and_next / final_values == the oracle's per-round evaluations."""
    fid, p = 0, R.BN254_FR
    ctx = ctxs(fid, 1)
    rng = random.Random(4)
    P, D, n = 2, 3, 9
    tabs = [rand_table(rng, p, n) for _ in range(P * D)]
    ref = oracle.gkr_sumcheck_prove(oracle.Transcript(fid), 1, P, D, [ints_to_arr(t) for t in tabs])
    sp = zkb.SumPoly(ctx, [zkb.ProductPoly(ctx, tabs[q * D: (q + 1) * D]) for q in range(P)])
    ev = sp.round_evals()
    for k in range(n):
        assert ev == ref["evals"][k]
        ev = sp.bind_and_next(ref["challenges"][k], last=(k == n - 1))
    assert sp.final_values() == ref["final_vals"]


def test_compat_shape_errors(zkb, ctxs):
    ctx = ctxs(0, 0)
    PP = zkb.ProductPoly
    with pytest.raises(zkb.ZkbError) as ei:  # reference indexes polys[1] and panics (composed_polynomial.rs:90)
        zkb.SumPoly(ctx, [PP(ctx, [[1, 2], [3, 4]])]).handle()
    assert ei.value.status == -10


# ------------------------------------------------------------------------------------------- GKR
def tree_circuit(rng, n_layers, out_gates=1):
    gates = [out_gates << (n_layers - 1 - l) for l in range(n_layers)]
    ops = [[rng.randrange(2) for _ in range(g)] for g in gates]
    return gates, ops


@pytest.mark.parametrize("fid,p", FIELDS)
def test_gkr_prove_matches_oracle(zkb, ctxs, oracle, fid, p):
    ctx = ctxs(fid, 0)
    G = zkb.gkr_protocol
    rng = random.Random(5000 + fid)
    for n_layers, out_gates in ((1, 1), (1, 2), (2, 1), (3, 1), (3, 2), (5, 1), (8, 2), (10, 1)):
        gates, ops = tree_circuit(rng, n_layers, out_gates)
        inputs = [rng.randrange(p) for _ in range(2 * gates[0])]
        flat = np.array([o for layer in ops for o in layer], dtype=np.uint8)
        ref = oracle.gkr_prove(fid, gates, flat, ints_to_arr(inputs))
        c = zkb.gkr_circuit.Circuit(ctx, [[zkb.Operation(o) for o in layer] for layer in ops])
        layers = c.evaluate(inputs)
        want = oracle.circuit_evaluate(fid, gates, flat, ints_to_arr(inputs))
        assert layers == [arr_to_ints(x) for x in want]
        pr = G.prove(c, inputs)
        assert pr.output_poly == ref["output_poly"]
        assert [[q.coefficients for q in layer] for layer in pr.proof_polynomials] == ref["proof_polynomials"], (n_layers, out_gates)
        assert pr.claimed_evaluations == ref["claimed_evaluations"]
        assert pr.final_openings == ref["final_openings"]
        assert pr.challenges == ref["challenges"]
        assert G.verify(pr, c, inputs)
        # tampering: a coefficient, a claimed evaluation, an opening, the inputs
        if pr.proof_polynomials[0][0].coefficients:
            saved = list(pr.proof_polynomials[0][0].coefficients)
            pr.proof_polynomials[0][0].coefficients[0] = (saved[0] + 1) % p
            assert not G.verify(pr, c, inputs)
            pr.proof_polynomials[0][0].coefficients = saved
        fo = pr.final_openings
        pr.final_openings = ((fo[0] + 1) % p, fo[1])
        assert not G.verify(pr, c, inputs)
        pr.final_openings = fo
        bad_inputs = list(inputs)
        bad_inputs[0] = (bad_inputs[0] + 1) % p
        assert not G.verify(pr, c, bad_inputs)
        assert G.verify(pr, c, inputs)
        c.free()


@pytest.mark.parametrize("fid,p", FIELDS)
def test_dense_reference_construction_on_device(zkb, ctxs, fid, p):
    """Layer::get_add_mul_i / get_fbc_poly as the reference builds them (dense), on the device:
    the known answers of gkr_circuit.rs:205-256 and gkr_protocol.rs:423-452, and the dense prover
    == the two-phase prover on small circuits (compat mode: 2 products x 2 factors)."""
    ctx = ctxs(fid, zkb.MODE_COMPAT)
    Op = zkb.Operation
    G = zkb.gkr_protocol
    # one Add gate -> add_i has its single one at index 1 (a=0,b=0,c=1); mul_i is all zero
    c1 = zkb.gkr_circuit.Circuit(ctx, [[Op.Add]])
    assert c1.layers[0].get_add_mul_i(Op.Add).evaluation == [0, 1, 0, 0, 0, 0, 0, 0]
    assert c1.layers[0].get_add_mul_i(Op.Mul).evaluation == [0] * 8
    c2 = zkb.gkr_circuit.Circuit(ctx, [[Op.Mul]])
    assert c2.layers[0].get_add_mul_i(Op.Mul).evaluation == [0, 1, 0, 0, 0, 0, 0, 0]
    # test_get_fbc_poly: 1 add gate, r = 5, w = [2, 12]
    fbc = G.get_fbc_poly(5, c1.layers[0], [2, 12], [2, 12])
    assert fbc.polys[0].evaluation[0].evaluation == [0, (p - 4) % p, 0, 0]
    assert fbc.polys[0].evaluation[1].evaluation == [4, 14, 14, 24]
    assert fbc.polys[1].evaluation[1].evaluation == [4, 24, 24, 144]
    # 2-gate layer: widths (1, 2, 2): gate 1 sits at a=1, b=2, c=3 -> index 0b1_10_11 = 27
    c3 = zkb.gkr_circuit.Circuit(ctx, [[Op.Add, Op.Mul], [Op.Add]])
    tab = c3.layers[0].get_add_mul_i(Op.Mul).evaluation
    assert len(tab) == 32 and [i for i, v in enumerate(tab) if v] == [27] and tab[27] == 1
    rng = random.Random(6000 + fid)
    for n_layers, out_gates in ((1, 1), (2, 1), (3, 1), (3, 2), (4, 1)):
        gates, ops = tree_circuit(rng, n_layers, out_gates)
        inputs = [rng.randrange(p) for _ in range(2 * gates[0])]
        c = zkb.gkr_circuit.Circuit(ctx, [[Op(o) for o in layer] for layer in ops])
        dense = G.prove_dense(c, inputs)
        sparse = G.prove(c, inputs)
        assert [[q.coefficients for q in layer] for layer in dense.proof_polynomials] == \
               [[q.coefficients for q in layer] for layer in sparse.proof_polynomials], (n_layers, out_gates)
        assert dense.claimed_evaluations == sparse.claimed_evaluations and dense.final_openings == sparse.final_openings
        assert dense.challenges == sparse.challenges
        c.free()


def test_gkr_reference_test_circuit(zkb, ctxs):
    """gkr_protocol.rs:474-506 over BLS12-381 Fr + SURVEY App. C digest of the whole proof."""
    ctx = ctxs(2, 0)
    Op = zkb.Operation
    c = zkb.gkr_circuit.Circuit(ctx, [[Op.Add] * 4, [Op.Mul, Op.Add], [Op.Add]])
    inputs = [5, 2, 2, 4, 10, 0, 3, 3]
    pr = zkb.gkr_protocol.prove(c, inputs)
    assert pr.output_poly == [58, 0]
    assert [[len(q.coefficients) for q in layer] for layer in pr.proof_polynomials] == [[3, 3], [3, 3, 2, 3], [3, 3, 3, 3, 3, 3]]
    assert pr.claimed_evaluations[0] == (0x5F68E1D7B90A6A0C7F6949557492B69D68670D9FF4349F1A9DE701BA63781FFB,
                                         0x267EC4515C3CA65CB4FDB7CB98343C8B1CFCBA5831B22140F40571A63DB775B9)
    blob = b"".join(int(v).to_bytes(32, "little") for v in pr.output_poly)
    for layer in pr.proof_polynomials:
        for q in layer:
            blob += b"".join(int(v).to_bytes(32, "little") for v in q.coefficients)
    for a, b in pr.claimed_evaluations:
        blob += int(a).to_bytes(32, "little") + int(b).to_bytes(32, "little")
    blob += int(pr.final_openings[0]).to_bytes(32, "little") + int(pr.final_openings[1]).to_bytes(32, "little")
    assert len(blob) == 1376
    assert zkb.engine.keccak256(blob).hex() == "f36dd78dc8541b9e70d074028d90a5f35623682be6076b5a6e3ca2e7d9526e5d"
    assert zkb.gkr_protocol.verify(pr, c, inputs)


def wired_pair(zkb, ctx, rng, n_inputs, gates_per_layer):
    """The same random general-wiring circuit for the oracle and for the device."""
    layers, spec, w = [], [], n_inputs
    for G in gates_per_layer:
        ops = [rng.randrange(2) for _ in range(G)]
        in1 = [rng.randrange(w) for _ in range(G)]
        in2 = [rng.randrange(w) for _ in range(G)]
        layers.append(R.WiredLayer(ops, in1, in2, w))
        spec.append((ops, in1, in2))
        w = G
    return R.WiredCircuit(layers), zkb.gkr_circuit.WiredCircuit(ctx, n_inputs, spec)


@pytest.mark.parametrize("fid,p", FIELDS)
def test_wired_gkr_matches_oracle(zkb, ctxs, fid, p):
    """General wiring + wide output layer (extension for BASELINE configs[2]): device == oracle, bit for bit."""
    ctx = ctxs(fid, 0)
    G = zkb.gkr_protocol
    rng = random.Random(7000 + fid)
    shapes = [(2, [1]), (2, [4]), (4, [4, 4]), (8, [4, 8, 2]), (4, [8, 4, 4, 1]), (16, [16, 16, 16]), (64, [128, 32, 64]),
              (256, [256, 256, 256])]
    if fid == 0:
        shapes += [(1 << 12, [1 << 13, 1 << 12]), (1 << 14, [1 << 14])]  # persistent-kernel and multi-launch regimes
    for nin, gates in shapes:
        oc, dc = wired_pair(zkb, ctx, rng, nin, gates)
        inputs = [rng.randrange(p) for _ in range(nin)]
        assert dc.evaluate(inputs) == oc.evaluate(inputs, p)
        ref = R.wired_prove_sparse(oc, inputs, p)
        pr = G.prove_wired(dc, inputs)
        assert pr.output_poly == ref.output_poly
        assert [[q.coefficients for q in layer] for layer in pr.proof_polynomials] == ref.proof_polynomials, (nin, gates)
        assert pr.claimed_evaluations == ref.claimed_evaluations
        assert pr.final_openings == ref.final_openings
        assert G.verify_wired(pr, dc, inputs)
        assert R.wired_verify_sparse(ref, oc, inputs, p)
        if pr.proof_polynomials[-1][0].coefficients:
            saved = list(pr.proof_polynomials[-1][0].coefficients)
            pr.proof_polynomials[-1][0].coefficients[0] = (saved[0] + 1) % p
            assert not G.verify_wired(pr, dc, inputs)
            pr.proof_polynomials[-1][0].coefficients = saved
        out = list(pr.output_poly)
        pr.output_poly[-1] = (out[-1] + 1) % p
        assert not G.verify_wired(pr, dc, inputs)
        pr.output_poly = out
        bad_inputs = list(inputs)
        bad_inputs[-1] = (bad_inputs[-1] + 1) % p
        assert not G.verify_wired(pr, dc, bad_inputs)
        assert G.verify_wired(pr, dc, inputs)
        dc.free()


def test_wired_gkr_large_vs_c_oracle(zkb, ctxs, oracle):
    """2^16-wide uniform layers with random wiring (all three round-driver regimes per phase): device == C oracle."""
    ctx = ctxs(0, 0)
    p = R.BN254_FR
    nrng = np.random.default_rng(17)
    G = 1 << 16
    spec = [(nrng.integers(0, 2, size=G, dtype=np.uint8), nrng.integers(0, G, size=G, dtype=np.uint32),
             nrng.integers(0, G, size=G, dtype=np.uint32)) for _ in range(3)]
    inputs = oracle.synth_table(0, 41, 0, 16)
    ref = oracle.gkr_prove_wired(0, G, spec, inputs)
    dc = zkb.gkr_circuit.WiredCircuit(ctx, G, spec)
    pr = zkb.gkr_protocol.prove_wired(dc, arr_to_ints(inputs))
    assert pr.output_poly == ref["output_poly"]
    assert [[q.coefficients for q in layer] for layer in pr.proof_polynomials] == ref["proof_polynomials"]
    assert pr.claimed_evaluations == ref["claimed_evaluations"] and pr.final_openings == ref["final_openings"]
    assert pr.challenges == ref["challenges"]
    assert zkb.gkr_protocol.verify_wired(pr, dc, arr_to_ints(inputs))
    dc.free()


@pytest.mark.parametrize("fid,p", FIELDS)
def test_wired_reduces_to_reference_wiring_on_device(zkb, ctxs, fid, p):
    """in1 = 2g, in2 = 2g+1, <= 2 outputs: zkb_gkr_prove_wired produces the bytes of zkb_gkr_prove."""
    ctx = ctxs(fid, 0)
    G = zkb.gkr_protocol
    rng = random.Random(7100 + fid)
    for n_layers, out_gates in ((1, 1), (1, 2), (3, 1), (6, 2), (9, 1)):
        gates, ops = tree_circuit(rng, n_layers, out_gates)
        inputs = [rng.randrange(p) for _ in range(2 * gates[0])]
        struct = [[zkb.Operation(o) for o in layer] for layer in ops]
        c = zkb.gkr_circuit.Circuit(ctx, struct)
        w = zkb.gkr_circuit.WiredCircuit.binary_tree(ctx, struct)
        a, b = G.prove(c, inputs), G.prove_wired(w, inputs)
        assert a.output_poly == b.output_poly and a.claimed_evaluations == b.claimed_evaluations
        assert [[q.coefficients for q in l] for l in a.proof_polynomials] == [[q.coefficients for q in l] for l in b.proof_polynomials]
        assert a.final_openings == b.final_openings and a.challenges == b.challenges
        assert G.verify_wired(a, w, inputs) and G.verify(b, c, inputs)
        c.free()
        w.free()


def test_wired_shape_errors(zkb, ctxs):
    ctx = ctxs(0, 0)
    W = zkb.gkr_circuit.WiredCircuit
    for nin, spec in [(3, [([0], [0], [1])]),                      # inputs not a power of two
                      (4, [([0, 1, 0], [0, 1, 2], [1, 2, 3])]),    # 3 gates
                      (4, [([0, 1], [0, 4], [1, 2])]),             # wire out of range
                      (4, [([0], [0], [1]), ([1], [0], [0])])]:    # a layer above a single gate
        with pytest.raises(zkb.ZkbError) as ei:
            W(ctx, nin, spec)
        assert ei.value.status == -11


@pytest.mark.parametrize("fid,p", FIELDS)
def test_device_transcript(zkb, ctxs, oracle, fid, p):
    """in the on-chip kernel the GPU runs Keccak / interpolation / trimming itself and the host only
    replays.  Same bytes as the oracle and as the host-transcript path; the counters prove the device path ran."""
    ctx = ctxs(fid, zkb.MODE_FULL)
    rng = random.Random(8000 + fid)
    S, T = zkb.sum_check_protocol, zkb.fiat_shamir.Transcript
    try:
        for n in (1, 2, 5, 10, 13):
            tabs = [rand_table(rng, p, n) for _ in range(2)]
            if n == 5:  # a round polynomial that trims: second factor zero -> every message is empty
                tabs[1] = [0] * (1 << n)
            if n == 2:  # degree drops: constant second factor
                tabs[1] = [7] * (1 << n)
            ref = oracle.gkr_sumcheck_prove(oracle.Transcript(fid), 1, 1, 2, [ints_to_arr(t) for t in tabs])
            sp = zkb.SumPoly(ctx, [zkb.ProductPoly(ctx, tabs)])
            out = []
            for on in (True, False):
                ctx.set_device_transcript(on)
                l0, r0 = ctx.device_transcript_stats()
                t = T(fid)
                t.append(b"seed")  # a partly filled sponge at entry (4 bytes: not whole words -> host path) ...
                pr = S.gkr_prove(0, sp, t)
                t2 = T(fid)
                t2.append(b"12345678" * 5)  # ... and one with whole words pending -> device path continues it
                pr2 = S.gkr_prove(0, sp, t2)
                pr3 = S.gkr_prove(0, sp, T(fid))
                l1, r1 = ctx.device_transcript_stats()
                if on:
                    assert l1 > l0 and r1 - r0 >= n, (n, l0, l1, r0, r1)
                else:
                    assert (l1, r1) == (l0, r0)
                out.append([([q.coefficients for q in x.proof_polynomials], x.random_challenges, x.final_values) for x in (pr, pr2, pr3)])
                # the transcripts continue identically after the proof
                assert t2.get_random_challenge() == _replay(zkb, fid, b"12345678" * 5, pr2)
            assert out[0] == out[1]
            assert out[0][2][0] == ref["coeffs"] and out[0][2][1] == ref["challenges"] and out[0][2][2] == ref["final_vals"]
            sp.free()
    finally:
        ctx.set_device_transcript(False)  # the default


def _replay(zkb, fid, seed, pr):
    """Challenge that follows a proof on a fresh host transcript seeded like the prover's."""
    t = zkb.fiat_shamir.Transcript(fid)
    t.append(seed)
    for q in pr.proof_polynomials:
        t.append(zkb.fiat_shamir.fq_vec_to_bytes(q.coefficients))
        t.get_random_challenge()
    return t.get_random_challenge()


def test_circuit_shape_errors(zkb, ctxs):
    ctx = ctxs(0, 0)
    Op = zkb.Operation
    with pytest.raises(zkb.ZkbError) as ei:
        zkb.gkr_circuit.Circuit(ctx, [[Op.Add] * 4, [Op.Add] * 4])
    assert ei.value.status == -11


# --------------------------------------------------------------- full-size, size-independent properties
def test_fullsize_composed_sumcheck_matches_oracle(zkb, ctxs, oracle):
    """BASELINE configs[1] at its full size (n = 24, one ProductPoly of 2 factors, full mode): the proof -- every
    trimmed coefficient vector, every challenge, the bound values -- equals the CPU oracle's proof of the same
    synthetic tables bit for bit (sum_check_protocol.rs:86-115), plus the independent-kernel properties."""
    fid, p = 0, R.BN254_FR
    ctx = ctxs(fid, 1)
    n = 24
    a = zkb.MultilinearPoly.generate(ctx, 0xB2000002, 0, n)
    b = zkb.MultilinearPoly.generate(ctx, 0xB2000002, 1, n)
    # spot-check generated entries against the oracle's generator
    head = arr_to_ints(zkb.engine.from_mont(fid, a.montgomery()[:4]))
    assert head == arr_to_ints(oracle.synth_table(fid, 0xB2000002, 0, n, count=4))
    sp = zkb.SumPoly(ctx, [zkb.ProductPoly.from_polys(ctx, [a, b])])
    S = zkb.sum_check_protocol
    pr = S.gkr_prove(0, sp, zkb.fiat_shamir.Transcript(fid))
    oracle.set_threads(max(1, len(os.sched_getaffinity(0))))
    ref = oracle.gkr_sumcheck_prove(oracle.Transcript(fid), 1, 1, 2, [oracle.synth_table(fid, 0xB2000002, t, n) for t in range(2)])
    assert [q.coefficients for q in pr.proof_polynomials] == ref["coeffs"]
    assert pr.random_challenges == ref["challenges"] and pr.final_values == ref["final_vals"]
    prod = a * b
    claim = sum(prod.sum_halves()) % p
    ev0 = pr.proof_polynomials[0]
    assert (ev0.evaluate(0) + ev0.evaluate(1)) % p == claim
    v = S.gkr_verify(pr.proof_polynomials, claim, zkb.fiat_shamir.Transcript(fid))
    assert v.verified and v.random_challenges == pr.random_challenges
    assert v.final_claimed_sum == pr.final_values[0] * pr.final_values[1] % p
    assert a.evaluate(pr.random_challenges) == pr.final_values[0]
    assert b.evaluate(pr.random_challenges) == pr.final_values[1]
    sp.free()
    for t in (a, b, prod):
        t.free()


@pytest.mark.parametrize("P,D,n", [(2, 3, 20), (1, 3, 21), (2, 2, 20)])
def test_large_composed_shapes_match_oracle(zkb, ctxs, oracle, P, D, n):
    """The shapes of BASELINE configs[3] (2 products x 3 factors) and of the north-star target (1 x 3) at 2^20-2^21
    entries per table -- large enough for every regime (one launch per round, persistent kernel, on-chip kernel) --
    against the oracle, bit for bit."""
    fid = 0
    ctx = ctxs(fid, 1)
    tabs = [zkb.MultilinearPoly.generate(ctx, 0xB2000010 + P, t, n) for t in range(P * D)]
    sp = zkb.SumPoly(ctx, [zkb.ProductPoly.from_polys(ctx, tabs[q * D:(q + 1) * D]) for q in range(P)])
    pr = zkb.sum_check_protocol.gkr_prove(0, sp, zkb.fiat_shamir.Transcript(fid))
    oracle.set_threads(max(1, len(os.sched_getaffinity(0))))
    ref = oracle.gkr_sumcheck_prove(oracle.Transcript(fid), 1, P, D, [oracle.synth_table(fid, 0xB2000010 + P, t, n) for t in range(P * D)])
    assert [q.coefficients for q in pr.proof_polynomials] == ref["coeffs"]
    assert pr.random_challenges == ref["challenges"] and pr.final_values == ref["final_vals"]
    sp.free()
    for t in tabs:
        t.free()


def test_multi_gpu_parity_two_ranks():
    """tests/multi_gpu_check.py under torchrun with 2 ranks (real NCCL, one process per GPU): sharded == single-GPU ==
    oracle for plain and composed sumchecks, both exchange paths.  Needs two visible GPUs."""
    import subprocess
    import sys
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run on a multi-GPU box: gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ZKB_REPEAT="5")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", os.path.join(root, "tests", "multi_gpu_check.py")], cwd=root, env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MULTI_GPU_PARITY OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_stress_short(zkb):
    """tools/stress.py for a few seconds: thousands of small proofs through every regime of the round driver (the
    persistent kernels publish round messages without a system fence and rely on checksums), each compared with the oracle."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "stress.py"), "5"], cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


def test_fullsize_evaluate_linearity(zkb, ctxs):
    fid, p = 0, R.BN254_FR
    ctx = ctxs(fid, 0)
    n = 22
    rng = random.Random(8)
    a = zkb.MultilinearPoly.generate(ctx, 5, 0, n)
    b = zkb.MultilinearPoly.generate(ctx, 5, 1, n)
    rs = [rng.randrange(p) for _ in range(n)]
    s, t = rng.randrange(p), rng.randrange(p)
    lhs = (a.scale(s) + b.scale(t)).evaluate(rs)
    assert lhs == (s * a.evaluate(rs) + t * b.evaluate(rs)) % p
    # a boolean point reads the table entry (variable 0 = MSB)
    idx = rng.randrange(1 << n)
    bits = [(idx >> (n - 1 - k)) & 1 for k in range(n)]
    from oracle import c_oracle
    assert a.evaluate(bits) == arr_to_ints(c_oracle.synth_table(fid, 5, 0, n, first=idx, count=1))[0]


# ------------------------------------------------------------------------ error behaviour through the C ABI
def test_abi_error_codes_on_device(zkb, ctxs):
    """The reference's panics arrive as status codes (and as the same strings through the mirror); bad handles and
    null pointers are rejected instead of crashing."""
    import ctypes as C

    E = zkb.engine
    L = E.lib()
    ctx = ctxs(0, 1)
    h = C.c_uint64()
    three = ctx.mont([1, 2, 3])
    assert L.zkb_mle_upload(ctx.handle, three.ctypes.data, 3, C.byref(h)) == -2          # "Invalid evaluations"
    assert L.zkb_mle_upload(ctx.handle, None, 4, C.byref(h)) == -1                        # null pointer
    assert L.zkb_mle_free(ctx.handle, 0xDEADBEEF) == -1                                   # unknown handle
    m = zkb.MultilinearPoly(ctx, [1, 2, 3, 4])
    out = (C.c_uint64 * 4)()
    one = ctx.mont([5])
    assert L.zkb_mle_evaluate(ctx.handle, m.handle, E._p(one), 1, out) == -3              # "Invalid number of values"
    assert L.zkb_mle_partial_evaluate(ctx.handle, m.handle, 2, E._p(one), C.byref(h)) == -3
    m8 = zkb.MultilinearPoly(ctx, list(range(8)))
    arr = (C.c_uint64 * 2)(m.handle, m8.handle)
    assert L.zkb_sumpoly_create(ctx.handle, arr, 1, 2, C.byref(h)) == -4                  # "all evaluations must have same length"
    assert L.zkb_mle_binary(ctx.handle, m.handle, m8.handle, 0, C.byref(h)) == -4
    big = (C.c_uint64 * 20)(*([m.handle] * 20))
    assert L.zkb_sumpoly_create(ctx.handle, big, 10, 2, C.byref(h)) == -9                 # more than 16 tables
    assert L.zkb_sumpoly_create(ctx.handle, big, 1, 5, C.byref(h)) == -9                  # degree > 4
    assert b"16 tables" in L.zkb_ctx_last_error(ctx.handle) or b"degree" in L.zkb_ctx_last_error(ctx.handle)
    gates = (C.c_uint32 * 2)(4, 4)
    ops = (C.c_uint8 * 8)(*([0] * 8))
    assert L.zkb_circuit_create(ctx.handle, 2, gates, ops, C.byref(h)) == -11             # not expressible by the reference wiring
    # a failed call leaves the context usable
    assert m.evaluate([0, 1]) == 2
    raw = C.c_void_p()
    assert L.zkb_ctx_create(7, 0, 0, C.byref(raw)) == -1                                  # unknown field id
    assert L.zkb_ctx_create(0, 99, 0, C.byref(raw)) == -6                                 # no such device
