"""Device-backed mirror of `sum_check::sum_check_protocol`
(sum_check_protocol.rs:8-175): prove / verify for one MultilinearPoly and
gkr_prove / gkr_verify for a SumPoly.  The round loop, transcript and
interpolation run inside libzkb200 (one kernel per round, (d+1) field elements
back per round); nothing here computes on the CPU."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from . import engine as E
from .engine import _ck, _p, lib
from .fiat_shamir import Transcript
from .multilinear_polynomial import MultilinearPoly, SumPoly
from .univariate_polynomial import UnivariatePoly

i32p = C.POINTER(C.c_int32)


@dataclass
class Proof:  # :8-12
    proof_polynomials: List[List[int]]
    claimed_sum: int
    random_challenges: List[int] = field(default_factory=list)  # not in the reference struct; kept for tests


@dataclass
class GkrProof:  # :13-17
    proof_polynomials: List[UnivariatePoly]
    claimed_sum: int
    random_challenges: List[int]
    final_values: List[int] = field(default_factory=list)  # bound table values (extra)


@dataclass
class GkrVerify:  # :19-23
    verified: bool
    final_claimed_sum: int
    random_challenges: List[int]


def prove(polynomial: MultilinearPoly, absorb_table: bool = True) -> Proof:  # :25-52
    ctx = polynomial.ctx
    n = polynomial.num_of_vars + (ctx.world.bit_length() - 1)
    claimed = np.zeros((1, 4), dtype=np.uint64)
    msgs = np.zeros((max(n, 1), 2, 4), dtype=np.uint64)
    chals = np.zeros((max(n, 1), 4), dtype=np.uint64)
    _ck(ctx, lib().zkb_sumcheck_prove(ctx.handle, polynomial.handle, 1 if absorb_table else 0, _p(claimed), _p(msgs), _p(chals)))
    m = ctx.unmont(msgs[:n].reshape(-1, 4)) if n else []
    return Proof([m[2 * i: 2 * i + 2] for i in range(n)], ctx.unmont(claimed)[0], ctx.unmont(chals[:n]) if n else [])


def verify(polynomial: MultilinearPoly, proof: Proof, absorb_table: bool = True) -> bool:  # :54-84
    ctx = polynomial.ctx
    for m in proof.proof_polynomials:
        if len(m) == 0 or len(m) & (len(m) - 1):
            raise ValueError("Invalid evaluations")  # MultilinearPoly::new(poly.to_vec()) :63
        if len(m) != 2:
            raise IndexError("round message must have two evaluations")  # poly.evaluation[1] :73
    n = len(proof.proof_polynomials)
    flat = ctx.mont([x for m in proof.proof_polynomials for x in m]) if n else np.zeros((1, 4), dtype=np.uint64)
    ok = C.c_int32()
    _ck(ctx, lib().zkb_sumcheck_verify(ctx.handle, polynomial.handle, 1 if absorb_table else 0, _p(ctx.mont([proof.claimed_sum])),
                                       _p(flat), n, C.byref(ok)))
    return bool(ok.value)


class RawGkrProver:
    """The bare C-ABI call of gkr_prove with preallocated Montgomery-limb outputs -- what a Rust caller pays:
    no per-call allocation and no canonical <-> Montgomery conversion (its `F` IS the Montgomery limbs)."""

    def __init__(self, composed_polynomial: SumPoly):
        self.sp = composed_polynomial
        ctx = self.ctx = composed_polynomial.ctx
        self.d = d = composed_polynomial.get_degree()
        self.n = n = composed_polynomial.polys[0].evaluation[0].num_of_vars + (ctx.world.bit_length() - 1)
        self.coeffs = np.zeros((max(n, 1), d + 1, 4), dtype=np.uint64)
        self.lens = np.zeros(max(n, 1), dtype=np.int32)
        self.chals = np.zeros((max(n, 1), 4), dtype=np.uint64)
        self.fin = np.zeros((len(composed_polynomial.polys) * d, 4), dtype=np.uint64)
        self.claim = np.zeros((1, 4), dtype=np.uint64)
        self._args = (ctx.handle, None, _p(self.claim), composed_polynomial.handle(), _p(self.coeffs),
                      self.lens.ctypes.data_as(i32p), _p(self.chals), _p(self.fin))
        self._fn = lib().zkb_gkr_sumcheck_prove

    def prove(self, transcript: Transcript) -> None:
        a = self._args
        _ck(self.ctx, self._fn(a[0], transcript.handle, a[2], a[3], a[4], a[5], a[6], a[7]))

    def proof(self, claimed_sum: int = 0) -> "GkrProof":
        ctx, n = self.ctx, self.n
        polys = [UnivariatePoly(ctx.unmont(self.coeffs[k, : self.lens[k]]), ctx.field) for k in range(n)]
        return GkrProof(polys, claimed_sum, ctx.unmont(self.chals[:n]) if n else [], ctx.unmont(self.fin))


def gkr_prove(claimed_sum: int, composed_polynomial: SumPoly, transcript: Transcript) -> GkrProof:  # :86-115
    ctx = composed_polynomial.ctx
    d = composed_polynomial.get_degree()
    n = composed_polynomial.polys[0].evaluation[0].num_of_vars + (ctx.world.bit_length() - 1)  # :91
    T = len(composed_polynomial.polys) * d
    coeffs = np.zeros((max(n, 1), d + 1, 4), dtype=np.uint64)
    lens = np.zeros(max(n, 1), dtype=np.int32)
    chals = np.zeros((max(n, 1), 4), dtype=np.uint64)
    fin = np.zeros((T, 4), dtype=np.uint64)
    _ck(ctx, lib().zkb_gkr_sumcheck_prove(ctx.handle, transcript.handle, _p(ctx.mont([claimed_sum])), composed_polynomial.handle(),
                                          _p(coeffs), lens.ctypes.data_as(i32p), _p(chals), _p(fin)))
    polys = [UnivariatePoly(ctx.unmont(coeffs[k, : lens[k]]), ctx.field) for k in range(n)]
    return GkrProof(polys, claimed_sum, ctx.unmont(chals[:n]) if n else [], ctx.unmont(fin))


def gkr_verify(round_polys: Sequence[UnivariatePoly], claimed_sum: int, transcript: Transcript) -> GkrVerify:  # :117-150
    fld = transcript.field
    p = E.MODULI[fld]
    n = len(round_polys)
    slots = max([len(q.coefficients) for q in round_polys] + [1])
    ca = np.zeros((max(n, 1), slots, 4), dtype=np.uint64)
    lens = np.zeros(max(n, 1), dtype=np.int32)
    for k, q in enumerate(round_polys):
        lens[k] = len(q.coefficients)
        if q.coefficients:
            ca[k, : lens[k]] = E.to_mont(fld, E.ints_to_limbs(q.coefficients))
    ok = C.c_int32()
    fin = np.zeros((1, 4), dtype=np.uint64)
    ch = np.zeros((max(n, 1), 4), dtype=np.uint64)
    cs = E.to_mont(fld, E.ints_to_limbs([int(claimed_sum) % p]))
    _ck(None, lib().zkb_gkr_sumcheck_verify(transcript.handle, n, slots, _p(ca), lens.ctypes.data_as(i32p), _p(cs), C.byref(ok),
                                            _p(fin), _p(ch)))
    if not ok.value:
        return GkrVerify(False, 0, [0])  # :129-133
    return GkrVerify(True, E.limbs_to_ints(E.from_mont(fld, fin))[0], E.limbs_to_ints(E.from_mont(fld, ch[:n])) if n else [])


# ------------------------------------------------------------------ proof wire format (zkb_proof_*)
def _encode(field: int, kind: int, msgs: Sequence[Sequence[int]], claimed_sum: int, slots: int) -> bytes:
    n = len(msgs)
    arr = np.zeros((max(n, 1), slots, 4), dtype=np.uint64)
    lens = np.zeros(max(n, 1), dtype=np.int32)
    for k, m in enumerate(msgs):
        if len(m) > slots:
            raise ValueError("round message longer than the slot count")
        lens[k] = len(m)
        if len(m):
            arr[k, : len(m)] = E.to_mont(field, E.ints_to_limbs([int(x) % E.MODULI[field] for x in m]))
    cs = E.to_mont(field, E.ints_to_limbs([int(claimed_sum) % E.MODULI[field]]))
    size = C.c_size_t()
    _ck(None, lib().zkb_proof_encode(field, kind, n, slots, _p(arr), lens.ctypes.data_as(i32p), _p(cs), None, 0, C.byref(size)))
    out = np.zeros(size.value, dtype=np.uint8)
    _ck(None, lib().zkb_proof_encode(field, kind, n, slots, _p(arr), lens.ctypes.data_as(i32p), _p(cs), out.ctypes.data, size.value, C.byref(size)))
    return out.tobytes()


def _decode(data: bytes, slots: int):
    buf = np.frombuffer(data, dtype=np.uint8).copy()
    fld, kind, n = C.c_int32(), C.c_int32(), C.c_uint32()
    _ck(None, lib().zkb_proof_decode(buf.ctypes.data, len(data), C.byref(fld), C.byref(kind), C.byref(n), slots, None, None, None))
    arr = np.zeros((max(n.value, 1), slots, 4), dtype=np.uint64)
    lens = np.zeros(max(n.value, 1), dtype=np.int32)
    cs = np.zeros((1, 4), dtype=np.uint64)
    _ck(None, lib().zkb_proof_decode(buf.ctypes.data, len(data), C.byref(fld), C.byref(kind), C.byref(n), slots, _p(arr), lens.ctypes.data_as(i32p), _p(cs)))
    msgs = [E.limbs_to_ints(E.from_mont(fld.value, arr[k, : lens[k]])) if lens[k] else [] for k in range(n.value)]
    return fld.value, kind.value, msgs, E.limbs_to_ints(E.from_mont(fld.value, cs))[0]


def proof_to_bytes(proof, field: int) -> bytes:
    """Stable byte encoding of a `Proof` (kind 1) or `GkrProof` (kind 2): elements as fq_vec_to_bytes writes them
    (fiat_shamir_transcript.rs:32-37) behind a 12-byte header; see include/zkb200.h."""
    if isinstance(proof, GkrProof):
        msgs = [q.coefficients for q in proof.proof_polynomials]
        return _encode(field, 2, msgs, proof.claimed_sum, max([len(m) for m in msgs] + [1]))
    return _encode(field, 1, proof.proof_polynomials, proof.claimed_sum, 2)


def proof_from_bytes(data: bytes):
    """Inverse of proof_to_bytes; raises ZkbError on malformed bytes or non-canonical elements."""
    kind = data[6] if len(data) > 6 else 0
    fld, kind, msgs, cs = _decode(data, 2 if kind == 1 else 8)
    if kind == 1:
        return Proof(msgs, cs), fld
    return GkrProof([UnivariatePoly(m, fld) for m in msgs], cs, []), fld
