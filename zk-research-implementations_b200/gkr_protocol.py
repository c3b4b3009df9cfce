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
from .gkr_circuit import Circuit, Layer, WiredCircuit
from .multilinear_polynomial import MultilinearPoly, Operation, ProductPoly, SumPoly
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


def get_fbc_poly(random_challenge: int, layer: Layer, w_b: Sequence[int], w_c: Sequence[int]) -> SumPoly:  # :243-263
    """The reference's DENSE construction of the first layer's composed polynomial, on the device."""
    ctx = layer.ctx
    add_i = layer.get_add_mul_i(Operation.Add).partial_evaluate(0, random_challenge)
    mul_i = layer.get_add_mul_i(Operation.Mul).partial_evaluate(0, random_challenge)
    wb, wc = MultilinearPoly(ctx, w_b), MultilinearPoly(ctx, w_c)
    summed = MultilinearPoly.tensor_add_mul_polynomials(wb, wc, Operation.Add)
    multiplied = MultilinearPoly.tensor_add_mul_polynomials(wb, wc, Operation.Mul)
    return SumPoly(ctx, [ProductPoly.from_polys(ctx, [add_i, summed]), ProductPoly.from_polys(ctx, [mul_i, multiplied])])


def get_folded_fbc_poly(layer: Layer, w_b, w_c, r_b, r_c, alpha: int, beta: int) -> SumPoly:  # :265-292
    ctx = layer.ctx
    add_i, mul_i = layer.get_add_mul_i(Operation.Add), layer.get_add_mul_i(Operation.Mul)
    s_add = add_i.multi_partial_evaluate(r_b).scale(alpha) + add_i.multi_partial_evaluate(r_c).scale(beta)
    s_mul = mul_i.multi_partial_evaluate(r_b).scale(alpha) + mul_i.multi_partial_evaluate(r_c).scale(beta)
    wb, wc = MultilinearPoly(ctx, w_b), MultilinearPoly(ctx, w_c)
    summed = MultilinearPoly.tensor_add_mul_polynomials(wb, wc, Operation.Add)
    multiplied = MultilinearPoly.tensor_add_mul_polynomials(wb, wc, Operation.Mul)
    return SumPoly(ctx, [ProductPoly.from_polys(ctx, [s_add, summed]), ProductPoly.from_polys(ctx, [s_mul, multiplied])])


def prove_dense(circuit: Circuit, inputs: Sequence[int]) -> "GkrProof":
    """gkr_protocol::prove exactly as the reference builds it (:31-91): dense add_i/mul_i tables, tensor tables and
    the composed sumcheck of sum_check_protocol::gkr_prove in `compat` reduce semantics -- every table on the
    device.  Only for small circuits (2^(3g+2) entries per layer); used to check the two-phase prover."""
    from . import sum_check_protocol as S
    from .fiat_shamir import Transcript, fq_vec_to_bytes

    ctx = circuit.ctx
    p = ctx.p
    evals = circuit.evaluate(inputs)
    w0 = list(evals[-1])
    if len(w0) == 1:
        w0.append(0)
    t = Transcript(ctx.field)
    t.append(fq_vec_to_bytes(w0))  # initiate_protocol :229-241
    r0 = t.get_random_challenge()
    claimed = MultilinearPoly(ctx, w0).evaluate([r0])
    t.append(fq_vec_to_bytes([claimed]))
    L = len(circuit.layers)
    polys, claimed_evals, chals = [], [], []
    rb: List[int] = []
    rc: List[int] = []
    alpha = beta = o1 = o2 = 0
    inputs = [int(x) % p for x in inputs]
    for idx, layer in enumerate(reversed(circuit.layers)):
        w_i = inputs if idx == L - 1 else list(reversed(evals))[idx + 1]
        fbc = get_fbc_poly(r0, layer, w_i, w_i) if idx == 0 else get_folded_fbc_poly(layer, w_i, w_i, rb, rc, alpha, beta)
        sc = S.gkr_prove(claimed, fbc, t)
        polys.append(sc.proof_polynomials)
        chals.append(sc.random_challenges)
        mid = len(sc.random_challenges) // 2
        rb, rc = sc.random_challenges[:mid], sc.random_challenges[mid:]
        nxt = MultilinearPoly(ctx, w_i)
        o1, o2 = nxt.evaluate(rb), nxt.evaluate(rc)
        if idx < L - 1:
            t.append(fq_vec_to_bytes([o1]))
            alpha = t.get_random_challenge()
            t.append(fq_vec_to_bytes([o2]))
            beta = t.get_random_challenge()
            claimed = (alpha * o1 + beta * o2) % p
            claimed_evals.append((o1, o2))
        fbc.free()
    return GkrProof(w0, polys, claimed_evals, (o1, o2), chals)


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


# ------------------------------------------------- general wiring (extension beyond the reference, zkb200.h)
class RawWiredGkrProver:
    """The bare zkb_gkr_prove_wired call on a host array of Montgomery limbs, outputs preallocated."""

    def __init__(self, circuit: WiredCircuit, inputs_mont: np.ndarray):
        self.circuit, self.ctx = circuit, circuit.ctx
        self.inputs = np.ascontiguousarray(inputs_mont, dtype=np.uint64)
        self.total, L = circuit.total_rounds, len(circuit.gates)
        self.w0 = np.zeros((circuit.n_w0, 4), dtype=np.uint64)
        self.coeffs = np.zeros((self.total, 3, 4), dtype=np.uint64)
        self.lens = np.zeros(self.total, dtype=np.int32)
        self.chals = np.zeros((self.total, 4), dtype=np.uint64)
        self.claimed = np.zeros((max(L - 1, 1), 2, 4), dtype=np.uint64)
        self.fin = np.zeros((2, 4), dtype=np.uint64)
        self.nr = C.c_uint32()

    def prove(self) -> None:
        ctx = self.ctx
        _ck(ctx, lib().zkb_gkr_prove_wired(ctx.handle, self.circuit.handle, self.inputs.ctypes.data, self.inputs.shape[0], _p(self.w0),
                                           self.w0.shape[0], _p(self.coeffs), self.lens.ctypes.data_as(i32p), _p(self.chals),
                                           _p(self.claimed), _p(self.fin), C.byref(self.nr)))

    def verify(self) -> bool:
        ctx = self.ctx
        ok = C.c_int32()
        _ck(ctx, lib().zkb_gkr_verify_wired(ctx.handle, self.circuit.handle, self.inputs.ctypes.data, self.inputs.shape[0], _p(self.w0),
                                            self.w0.shape[0], _p(self.coeffs), self.lens.ctypes.data_as(i32p), _p(self.claimed),
                                            _p(self.fin), C.byref(ok)))
        return bool(ok.value)


def prove_wired(circuit: WiredCircuit, inputs: Sequence[int]) -> GkrProof:
    ctx = circuit.ctx
    raw = RawWiredGkrProver(circuit, ctx.mont(inputs))
    raw.prove()
    assert raw.nr.value == raw.total
    L = len(circuit.gates)
    polys, ch, off = [], [], 0
    for n in circuit.rounds_per_layer:
        polys.append([UnivariatePoly(ctx.unmont(raw.coeffs[off + k, : raw.lens[off + k]]), ctx.field) for k in range(n)])
        ch.append(ctx.unmont(raw.chals[off: off + n]))
        off += n
    ce = ctx.unmont(raw.claimed[: L - 1].reshape(-1, 4)) if L > 1 else []
    return GkrProof(ctx.unmont(raw.w0), polys, [(ce[2 * i], ce[2 * i + 1]) for i in range(L - 1)], tuple(ctx.unmont(raw.fin)), ch)


def verify_wired(proof: GkrProof, circuit: WiredCircuit, inputs: Sequence[int]) -> bool:
    ctx = circuit.ctx
    rpl, total, L = circuit.rounds_per_layer, circuit.total_rounds, len(circuit.gates)
    if len(proof.proof_polynomials) != L or [len(x) for x in proof.proof_polynomials] != rpl:
        return False
    if len(proof.claimed_evaluations) != L - 1 or len(proof.output_poly) != circuit.n_w0:
        return False
    raw = RawWiredGkrProver(circuit, ctx.mont(inputs))
    k = 0
    for layer in proof.proof_polynomials:
        for q in layer:
            if len(q.coefficients) > 3:
                return False
            raw.lens[k] = len(q.coefficients)
            if q.coefficients:
                raw.coeffs[k, : raw.lens[k]] = ctx.mont(q.coefficients)
            k += 1
    if L > 1:
        raw.claimed[: L - 1] = ctx.mont([x for pr in proof.claimed_evaluations for x in pr]).reshape(L - 1, 2, 4)
    raw.w0[:] = ctx.mont(proof.output_poly)
    raw.fin[:] = ctx.mont(list(proof.final_openings))
    return raw.verify()
