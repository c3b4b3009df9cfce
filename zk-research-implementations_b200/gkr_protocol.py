"""Device-backed mirror of `gkr::gkr_protocol` (gkr_protocol.rs:23-227).

`prove` runs the linear-time two-phase form of the per-layer sumcheck on the
GPU; its round polynomials equal the reference's dense construction
(SURVEY F7).  The KZG commitment/opening of the input layer (:92-118,157-183)
is out of scope: the proof carries the two input-MLE openings and
`verify` recomputes them on the device."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from .engine import _ck, _p, lib
from .gkr_circuit import Circuit
from .univariate_polynomial import UnivariatePoly

i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)


@dataclass
class GkrProof:  # :23-29 (input_proof replaced by final_openings)
    output_poly: List[int]
    proof_polynomials: List[List[UnivariatePoly]]
    claimed_evaluations: List[Tuple[int, int]]
    final_openings: Tuple[int, int]
    challenges: List[List[int]]


def _rounds_per_layer(circuit: Circuit) -> List[int]:
    return [2 * max(1, int(2 * int(g)).bit_length() - 1) for g in circuit.gates[::-1]]


class RawGkrProver:
    """The bare zkb_gkr_prove call on a host array of Montgomery limbs (a Rust `&[F]`), outputs preallocated."""

    def __init__(self, circuit: Circuit, inputs_mont: np.ndarray):
        self.circuit, self.ctx = circuit, circuit.ctx
        self.inputs = np.ascontiguousarray(inputs_mont, dtype=np.uint64)
        rpl = _rounds_per_layer(circuit)
        self.total, L = sum(rpl), len(circuit.gates)
        self.w0 = np.zeros((2, 4), dtype=np.uint64)
        self.coeffs = np.zeros((self.total, 3, 4), dtype=np.uint64)
        self.lens = np.zeros(self.total, dtype=np.int32)
        self.chals = np.zeros((self.total, 4), dtype=np.uint64)
        self.claimed = np.zeros((max(L - 1, 1), 2, 4), dtype=np.uint64)
        self.fin = np.zeros((2, 4), dtype=np.uint64)
        self.nr = C.c_uint32()

    def prove(self) -> None:
        ctx = self.ctx
        _ck(ctx, lib().zkb_gkr_prove(ctx.handle, self.circuit.handle, self.inputs.ctypes.data, self.inputs.shape[0], _p(self.w0),
                                     _p(self.coeffs), self.lens.ctypes.data_as(i32p), _p(self.chals), _p(self.claimed), _p(self.fin),
                                     C.byref(self.nr)))

    def verify(self) -> bool:
        ctx = self.ctx
        ok = C.c_int32()
        _ck(ctx, lib().zkb_gkr_verify(ctx.handle, self.circuit.handle, self.inputs.ctypes.data, self.inputs.shape[0], _p(self.w0),
                                      _p(self.coeffs), self.lens.ctypes.data_as(i32p), _p(self.claimed), _p(self.fin), C.byref(ok)))
        return bool(ok.value)


def prove(circuit: Circuit, inputs: Sequence[int]) -> GkrProof:  # :31-126
    ctx = circuit.ctx
    rpl = _rounds_per_layer(circuit)
    total, L = sum(rpl), len(circuit.gates)
    arr = ctx.mont(inputs)
    w0 = np.zeros((2, 4), dtype=np.uint64)
    coeffs = np.zeros((total, 3, 4), dtype=np.uint64)
    lens = np.zeros(total, dtype=np.int32)
    chals = np.zeros((total, 4), dtype=np.uint64)
    claimed = np.zeros((max(L - 1, 1), 2, 4), dtype=np.uint64)
    fin = np.zeros((2, 4), dtype=np.uint64)
    nr = C.c_uint32()
    _ck(ctx, lib().zkb_gkr_prove(ctx.handle, circuit.handle, arr.ctypes.data, len(inputs), _p(w0), _p(coeffs),
                                 lens.ctypes.data_as(i32p), _p(chals), _p(claimed), _p(fin), C.byref(nr)))
    assert nr.value == total
    polys, ch, off = [], [], 0
    for n in rpl:
        polys.append([UnivariatePoly(ctx.unmont(coeffs[off + k, : lens[off + k]]), ctx.field) for k in range(n)])
        ch.append(ctx.unmont(chals[off: off + n]))
        off += n
    ce = ctx.unmont(claimed[: L - 1].reshape(-1, 4)) if L > 1 else []
    return GkrProof(ctx.unmont(w0), polys, [(ce[2 * i], ce[2 * i + 1]) for i in range(L - 1)], tuple(ctx.unmont(fin)), ch)


def verify(proof: GkrProof, circuit: Circuit, inputs: Sequence[int]) -> bool:  # :128-227
    ctx = circuit.ctx
    rpl = _rounds_per_layer(circuit)
    total, L = sum(rpl), len(circuit.gates)
    if len(proof.proof_polynomials) != L or [len(x) for x in proof.proof_polynomials] != rpl:
        return False
    if len(proof.claimed_evaluations) != L - 1 or len(proof.output_poly) != 2:
        return False
    coeffs = np.zeros((total, 3, 4), dtype=np.uint64)
    lens = np.zeros(total, dtype=np.int32)
    k = 0
    for layer in proof.proof_polynomials:
        for q in layer:
            if len(q.coefficients) > 3:
                return False
            lens[k] = len(q.coefficients)
            if q.coefficients:
                coeffs[k, : lens[k]] = ctx.mont(q.coefficients)
            k += 1
    claimed = np.zeros((max(L - 1, 1), 2, 4), dtype=np.uint64)
    if L > 1:
        claimed[: L - 1] = ctx.mont([x for pr in proof.claimed_evaluations for x in pr]).reshape(L - 1, 2, 4)
    ok = C.c_int32()
    arr = ctx.mont(inputs)
    _ck(ctx, lib().zkb_gkr_verify(ctx.handle, circuit.handle, arr.ctypes.data, len(inputs), _p(ctx.mont(proof.output_poly)),
                                  _p(coeffs), lens.ctypes.data_as(i32p), _p(claimed), _p(ctx.mont(list(proof.final_openings))),
                                  C.byref(ok)))
    return bool(ok.value)
