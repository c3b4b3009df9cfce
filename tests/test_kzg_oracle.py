"""The KZG oracle (oracle/kzg_ref.py) against every known answer the reference's own tests hold for the prover side
of pcs/src/kzg_pcs/kzg.rs (tests at :224-389) and against the curve's defining properties."""
from oracle import kzg_ref as K

R = K.R


def fr(x):
    return x % R


POLY = [0, 4, 0, 4, 0, 4, 3, 7]
TAUS = [5, 2, 3]


def test_curve_constants():
    assert K.on_curve(K.G1)
    assert K.g1_mul(K.G1, R - 1) == K.g1_neg(K.G1)            # r * G = O
    assert K.g1_add(K.g1_mul(K.G1, R - 1), K.G1) is None
    assert K.g1_add(K.g1_mul(K.G1, 5), K.g1_mul(K.G1, 7)) == K.g1_mul(K.G1, 12)
    assert K.on_curve(K.g1_mul(K.G1, 0xDEADBEEF))


def test_blow_up_poly():  # kzg.rs:224-233
    assert K.blow_up_poly([0, 4], 4) == [0, 4, 0, 4]


def test_get_lagrange_basis():  # kzg.rs:236-259
    want = [fr(v) for v in (-8, 12, 16, -24, 10, -15, -20, 30)]
    assert K.lagrange_scalars(3, TAUS) == want
    assert K.get_lagrange_basis(3, TAUS) == [K.g1_mul(K.G1, s) for s in want]


def test_evaluate_poly_with_l_basis_and_commit():  # kzg.rs:262-286, :318-341
    k = K.KZG(3, TAUS)
    assert K.evaluate_poly_with_l_basis_in_g1(POLY, k.g1_lagrange_basis) == K.g1_mul(K.G1, 42)
    assert k.commit(POLY) == K.g1_mul(K.G1, 42)


def test_get_remainder_and_quotient():  # kzg.rs:289-315
    main = [fr(v) for v in (-72, -68, -54, -50)]
    assert K.get_remainder(main, 4) == [0, 4]
    assert K.get_quotient(main)[0] == 18


def test_open_and_get_proof():  # kzg.rs:344-389
    k = K.KZG(3, TAUS)
    z = [6, 4, 0]
    v = k.open(z, POLY)
    assert v == 72
    assert k.get_proof(v, z, POLY) == [K.g1_mul(K.G1, s) for s in (6, 18, 4)]


def test_opening_identity_in_the_exponent():
    """What KZG::verify (:97-129) checks with pairings, on scalars: f(tau) - v = sum_i q_i(tau_{i+1..}) (tau_i - z_i)."""
    import random

    rng = random.Random(3)
    n = 4
    poly = [rng.randrange(R) for _ in range(1 << n)]
    taus = [rng.randrange(R) for _ in range(n)]
    z = [rng.randrange(R) for _ in range(n)]
    v = K.evaluate(poly, z)
    pmv = [(e - v) % R for e in poly]
    rhs = 0
    for i in range(n):
        q = K.get_quotient(pmv)
        rhs = (rhs + K.evaluate(q, taus[i + 1:]) * (taus[i] - z[i])) % R
        pmv = K.get_remainder(pmv, z[i])
    assert (K.evaluate(poly, taus) - v) % R == rhs
    k = K.KZG(n, taus)
    assert k.commit(poly) == K.g1_mul(K.G1, K.evaluate(poly, taus))
