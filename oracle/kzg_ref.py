"""CPU restatement (test infrastructure -- only tests/, smoke() and bench.py's CPU legs may import this) of the
prover side of the reference's multilinear KZG over BLS12-381 G1, `pcs/src/kzg_pcs/kzg.rs`, as it is used for the
input layer by `gkr/src/gkr_protocol.rs:92-118`:

    get_lagrange_basis   kzg.rs:183-212   (scalars eq(taus, i), then g1 * scalar)
    commit               kzg.rs:51-53 -> evaluate_poly_with_l_basis_in_g1 :131-144  (sum_i basis[i] * poly[i])
    open                 kzg.rs:55-57     (poly.evaluate)
    get_proof            kzg.rs:59-95     (per variable: quotient :152-163, blow_up :165-171, MSM, remainder :146-150)

Plain Python integers; affine short-Weierstrass arithmetic with modular inverses (slow, small cases only).  The
third-party curve arithmetic (ark-bls12-381 0.5.0 / ark-ec 0.5.0, Cargo.lock, not vendored) is restated from the
published curve: y^2 = x^3 + 4 over Fq, the standard generator, prime subgroup order r = the BLS12-381 Fr modulus.

PINNED by the reference's own known answers (kzg.rs tests :239-389): the Lagrange-basis scalars [-8,12,16,-24,10,
-15,-20,30] for taus (5,2,3); commit([0,4,0,4,0,4,3,7]) = g1*42; open at (6,4,0) = 72; get_proof quotients
g1*[6,18,4]; get_remainder / get_quotient / blow_up_poly vectors -- see tests/test_kzg_oracle.py.  Points are
compared as affine coordinates, which is representation-independent (the reference compares projective points for
equality).  NOT restated: KZG::verify (:97-129, pairings and G2) -- the verifier side is out of scope; tests check
openings "in the exponent" with the known taus instead (commitment == g1 * f(taus), etc.).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

Q = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB  # base field
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001  # scalar field = group order
B = 4
G1 = (0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
      0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1)
Point = Optional[Tuple[int, int]]  # None = the point at infinity


def on_curve(p: Point) -> bool:
    return p is None or (p[1] * p[1] - p[0] * p[0] * p[0] - B) % Q == 0


def g1_neg(p: Point) -> Point:
    return None if p is None else (p[0], (-p[1]) % Q)


def g1_add(p: Point, q: Point) -> Point:
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % Q == 0:
            return None
        lam = 3 * p[0] * p[0] * pow(2 * p[1], -1, Q) % Q
    else:
        lam = (q[1] - p[1]) * pow(q[0] - p[0], -1, Q) % Q
    x = (lam * lam - p[0] - q[0]) % Q
    return (x, (lam * (p[0] - x) - p[1]) % Q)


def g1_mul(p: Point, k: int) -> Point:
    """mul_bigint with the canonical integer of the scalar (kzg.rs:143, :211)."""
    k %= R
    acc: Point = None
    while k:
        if k & 1:
            acc = g1_add(acc, p)
        p = g1_add(p, p)
        k >>= 1
    return acc


def g1_sum(points: Sequence[Point]) -> Point:
    acc: Point = None
    for p in points:
        acc = g1_add(acc, p)
    return acc


# ------------------------------------------------------------------ the multilinear pieces (Fr arithmetic)
def fold(table: Sequence[int], r: int) -> List[int]:
    """MultilinearPoly::partial_evaluate(0, r) (multilinear_polynomial_evaluation.rs:52-63): pairs (i, i + N/2)."""
    h = len(table) // 2
    return [(table[i] + r * (table[i + h] - table[i])) % R for i in range(h)]


def evaluate(table: Sequence[int], rs: Sequence[int]) -> int:
    t = list(table)
    for r in rs:
        t = fold(t, r)
    return t[0] % R


def lagrange_scalars(num_of_vars: int, taus: Sequence[int]) -> List[int]:
    """kzg.rs:183-207: for every hypercube point (variable 0 = most significant bit) the product of tau_i or 1 - tau_i."""
    if num_of_vars < 1:
        raise ValueError("Invalid num of vars for lagrange basis")
    out = []
    for i in range(1 << num_of_vars):
        v = 1
        for j in range(num_of_vars):
            bit = (i >> (num_of_vars - 1 - j)) & 1
            v = v * (taus[j] if bit else (1 - taus[j])) % R
        out.append(v)
    return out


def get_lagrange_basis(num_of_vars: int, taus: Sequence[int]) -> List[Point]:
    """kzg.rs:209-212"""
    return [g1_mul(G1, s) for s in lagrange_scalars(num_of_vars, taus)]


def evaluate_poly_with_l_basis_in_g1(evals: Sequence[int], basis: Sequence[Point]) -> Point:
    """kzg.rs:131-144"""
    if len(evals) != len(basis):
        raise ValueError("invalid polynomial or lagrange basis")
    return g1_sum([g1_mul(b, a) for a, b in zip(evals, basis)])


def blow_up_poly(poly: Sequence[int], bigger_len: int) -> List[int]:
    """kzg.rs:165-171: tensor (ones of length factor) x poly, index = i*|poly| + j  ->  the table tiled `factor` times."""
    return [poly[m % len(poly)] % R for m in range(bigger_len)]


def get_remainder(poly: Sequence[int], value: int) -> List[int]:
    """kzg.rs:146-150"""
    return fold(poly, value)


def get_quotient(poly: Sequence[int]) -> List[int]:
    """kzg.rs:152-163: f(1, .) - f(0, .)"""
    h = len(poly) // 2
    return [(poly[i + h] - poly[i]) % R for i in range(h)]


class KZG:
    """kzg.rs:11-34 (the G2 side of the setup is not restated: verifier only)."""

    def __init__(self, n_vars: int, taus: Sequence[int]):
        if len(taus) != n_vars:
            raise ValueError("invalid taus or polynomials")
        self.n_vars = n_vars
        self.taus = [t % R for t in taus]
        self.g1_lagrange_basis = get_lagrange_basis(n_vars, self.taus)

    def commit(self, poly: Sequence[int]) -> Point:  # :51-53
        return evaluate_poly_with_l_basis_in_g1(poly, self.g1_lagrange_basis)

    def open(self, opening_values: Sequence[int], poly: Sequence[int]) -> int:  # :55-57
        return evaluate(poly, opening_values)

    def get_proof(self, opened_value: int, opening_values: Sequence[int], poly: Sequence[int]) -> List[Point]:  # :59-95
        pmv = [(e - opened_value) % R for e in poly]
        n = len(poly)
        out = []
        for value in opening_values:
            quotient = get_quotient(pmv)
            if len(quotient) < n:
                quotient = blow_up_poly(quotient, n)
            out.append(evaluate_poly_with_l_basis_in_g1(quotient, self.g1_lagrange_basis))
            pmv = get_remainder(pmv, value)
        return out


def point_bytes(p: Point) -> bytes:
    """Affine coordinates as 2 x 48 bytes little-endian canonical; infinity = 96 zero bytes (what zkb_kzg_* returns)."""
    if p is None:
        return b"\0" * 96
    return p[0].to_bytes(48, "little") + p[1].to_bytes(48, "little")
