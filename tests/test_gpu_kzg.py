"""The input-layer commitment on the device (zkb_kzg_*, csrc/kzg_impl.cuh) against the KZG oracle
(oracle/kzg_ref.py, pinned to the reference's known answers in tests/test_kzg_oracle.py): affine points bit for bit."""
import random

import pytest

from oracle import kzg_ref as K

pytestmark = pytest.mark.gpu
R = K.R
POLY = [0, 4, 0, 4, 0, 4, 3, 7]
TAUS = [5, 2, 3]


@pytest.fixture(scope="module")
def ctx(zkb):
    c = zkb.Context(zkb.BLS12_381_FR, 0, zkb.MODE_COMPAT)
    yield c
    c.close()


def test_reference_known_answers(zkb, ctx):
    """kzg.rs tests :236-389 on the device."""
    poly = zkb.MultilinearPoly(ctx, POLY)
    k = zkb.kzg.KZG(poly, TAUS)
    want = [v % R for v in (-8, 12, 16, -24, 10, -15, -20, 30)]
    assert k.lagrange_basis() == [K.g1_mul(K.G1, s) for s in want]            # test_get_lagrange_basis
    assert k.commit(poly) == K.g1_mul(K.G1, 42)                               # test_commit / test_evaluate_poly_with_l_basis
    z = [6, 4, 0]
    v = k.open(z, poly)
    assert v == 72                                                            # test_open
    assert k.get_proof(v, z, poly) == [K.g1_mul(K.G1, s) for s in (6, 18, 4)]  # test_get_proof
    with pytest.raises(ValueError):
        zkb.kzg.KZG(poly, [1, 2])                                             # "invalid taus or polynomials"
    k.free()


@pytest.mark.parametrize("n", [1, 2, 5, 9, 11])
def test_random_against_oracle(zkb, ctx, n):
    """Small direct kernel (n <= 8) and the bucket MSM (n >= 9), every folded basis level, random scalars."""
    rng = random.Random(n)
    poly_i = [rng.randrange(R) for _ in range(1 << n)]
    taus = [rng.randrange(R) for _ in range(n)]
    z = [rng.randrange(R) for _ in range(n)]
    poly = zkb.MultilinearPoly(ctx, poly_i)
    k = zkb.kzg.KZG(poly, taus)
    scal = K.lagrange_scalars(n, taus)
    if n <= 5:
        ref = K.KZG(n, taus)
        assert k.lagrange_basis() == ref.g1_lagrange_basis
        assert k.commit(poly) == ref.commit(poly_i)
        v = k.open(z, poly)
        assert v == ref.open(z, poly_i)
        assert k.get_proof(v, z, poly) == ref.get_proof(v, z, poly_i)
    else:
        # the affine oracle is too slow for 2^n scalar multiplications: check in the exponent (the taus are known)
        idx = [0, 1, (1 << n) - 1, rng.randrange(1 << n)]
        for i in idx:
            assert k.lagrange_basis(0, i, 1) == [K.g1_mul(K.G1, scal[i])]
        # folded basis level 3, entry j = sum of the 8 entries with the same low bits
        j = rng.randrange(1 << (n - 3))
        assert k.lagrange_basis(3, j, 1) == [K.g1_mul(K.G1, sum(scal[(i << (n - 3)) + j] for i in range(8)) % R)]
        assert k.commit(poly) == K.g1_mul(K.G1, K.evaluate(poly_i, taus))
        v = k.open(z, poly)
        assert v == K.evaluate(poly_i, z)
        pmv, want = list(poly_i), []
        for i in range(n):
            want.append(K.g1_mul(K.G1, K.evaluate(K.get_quotient(pmv), taus[i + 1:])))
            pmv = K.get_remainder(pmv, z[i])
        assert k.get_proof(v, z, poly) == want
    k.free()


def test_degenerate_scalars(zkb, ctx):
    """Constant, zero and tiny tables: every point lands in a handful of buckets (the whole-CTA bucket path), zero
    digits are skipped, the commitment of the zero polynomial is the point at infinity."""
    n = 11
    taus = [7 + i for i in range(n)]
    for vals in ([0] * (1 << n), [5] * (1 << n), [i & 3 for i in range(1 << n)], [R - 1] * (1 << n)):
        poly = zkb.MultilinearPoly(ctx, vals)
        k = zkb.kzg.KZG(poly, taus)
        assert k.commit(poly) == K.g1_mul(K.G1, K.evaluate(vals, taus))
        k.free()


def test_wrong_field_is_refused(zkb, ctxs):
    ctx0 = ctxs(0, 0)  # BN254 Fr
    poly = zkb.MultilinearPoly(ctx0, [1, 2, 3, 4])
    with pytest.raises(zkb.ZkbError) as ei:
        zkb.kzg.KZG(poly, [1, 2])
    assert ei.value.status == -9
