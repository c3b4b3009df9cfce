"""Device-resident mirror of the reference crate `multilinear_polynomial`
(multilinear_polynomial_evaluation.rs + composed_polynomial.rs).

Same names, argument meaning and error behaviour as the Rust items; the dense
tables live in HBM and every method is a kernel launch through the C ABI.
Python `ValueError(msg)` stands for the reference's `panic!(msg)`.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import List, Optional, Sequence

import numpy as np

from . import engine as E
from .engine import Context, _ck, _p, lib


class Operation(enum.IntEnum):
    """multilinear_polynomial_evaluation.rs:4-17"""
    Add = E.OP_ADD
    Mul = E.OP_MUL

    def apply(self, a: int, b: int, p: int) -> int:
        return (a + b) % p if self is Operation.Add else (a * b) % p


class MultilinearPoly:
    """multilinear_polynomial_evaluation.rs:19-164.  `evaluation` downloads the table."""

    def __init__(self, ctx: Context, evaluations: Optional[Sequence[int]] = None, *, _handle: Optional[int] = None):
        self.ctx = ctx
        if _handle is not None:
            self._h = int(_handle)
        else:
            n = len(evaluations)
            if n == 0 or n & (n - 1):
                raise ValueError("Invalid evaluations")  # :30 (len 0 panics in ilog2)
            arr = ctx.mont(evaluations)
            h = C.c_uint64()
            _ck(ctx, lib().zkb_mle_upload(ctx.handle, arr.ctypes.data, n, C.byref(h)))
            self._h = h.value
        nv = C.c_uint32()
        _ck(ctx, lib().zkb_mle_num_vars(ctx.handle, self._h, C.byref(nv)))
        self.num_of_vars = nv.value

    # -- constructors that avoid Python ints for large tables
    @classmethod
    def new(cls, ctx: Context, evaluations: Sequence[int]) -> "MultilinearPoly":
        return cls(ctx, evaluations)

    @classmethod
    def from_montgomery(cls, ctx: Context, aos: np.ndarray, shard: bool = False) -> "MultilinearPoly":
        """aos: (n, 4) uint64 Montgomery limbs -- byte-for-byte a Rust `Vec<F>`."""
        aos = np.ascontiguousarray(aos, dtype=np.uint64)
        n = aos.shape[0]
        if n == 0 or n & (n - 1):
            raise ValueError("Invalid evaluations")
        h = C.c_uint64()
        fn = lib().zkb_mle_upload_shard if shard else lib().zkb_mle_upload
        _ck(ctx, fn(ctx.handle, aos.ctypes.data, n, C.byref(h)))
        return cls(ctx, _handle=h.value)

    @classmethod
    def from_host_pointer(cls, ctx: Context, ptr: int, n: int) -> "MultilinearPoly":
        """Upload n Montgomery elements from a raw host address (e.g. pinned memory)."""
        h = C.c_uint64()
        _ck(ctx, lib().zkb_mle_upload(ctx.handle, C.c_void_p(ptr), n, C.byref(h)))
        return cls(ctx, _handle=h.value)

    @classmethod
    def generate(cls, ctx: Context, seed: int, table_id: int, n_vars: int) -> "MultilinearPoly":
        """Synthetic table made on the device; this rank's shard if a communicator is attached."""
        h = C.c_uint64()
        _ck(ctx, lib().zkb_mle_generate(ctx.handle, seed, table_id, n_vars, C.byref(h)))
        return cls(ctx, _handle=h.value)

    # -- data access
    @property
    def handle(self) -> int:
        return self._h

    def __len__(self) -> int:
        return 1 << self.num_of_vars

    def montgomery(self) -> np.ndarray:
        out = np.empty((len(self), 4), dtype=np.uint64)
        _ck(self.ctx, lib().zkb_mle_download(self.ctx.handle, self._h, out.ctypes.data))
        return out

    def canonical_bytes(self) -> bytes:
        """fq_vec_to_bytes(&self.evaluation) (fiat_shamir_transcript.rs:32-37)."""
        out = np.empty(len(self) * 32, dtype=np.uint8)
        _ck(self.ctx, lib().zkb_mle_download_canonical(self.ctx.handle, self._h, out.ctypes.data))
        return out.tobytes()

    @property
    def evaluation(self) -> List[int]:
        raw = np.frombuffer(self.canonical_bytes(), dtype=np.uint64).reshape(-1, 4)
        return E.limbs_to_ints(raw)

    def free(self) -> None:
        if self._h:
            lib().zkb_mle_free(self.ctx.handle, self._h)
            self._h = 0

    def clone(self) -> "MultilinearPoly":
        h = C.c_uint64()
        _ck(self.ctx, lib().zkb_mle_clone(self.ctx.handle, self._h, C.byref(h)))
        return MultilinearPoly(self.ctx, _handle=h.value)

    # -- the reference's methods
    def partial_evaluate(self, bit: int, value: int) -> "MultilinearPoly":  # :52-63
        h = C.c_uint64()
        _ck(self.ctx, lib().zkb_mle_partial_evaluate(self.ctx.handle, self._h, bit, _p(self.ctx.mont([value])), C.byref(h)))
        return MultilinearPoly(self.ctx, _handle=h.value)

    def multi_partial_evaluate(self, values: Sequence[int]) -> "MultilinearPoly":  # :65-77
        if len(values) > self.num_of_vars:
            raise ValueError("Invalid number of values")
        h = C.c_uint64()
        arr = self.ctx.mont(values) if len(values) else np.zeros((1, 4), dtype=np.uint64)
        _ck(self.ctx, lib().zkb_mle_multi_partial_evaluate(self.ctx.handle, self._h, _p(arr), len(values), C.byref(h)))
        return MultilinearPoly(self.ctx, _handle=h.value)

    def evaluate(self, values: Sequence[int]) -> int:  # :79-91
        out = np.zeros((1, 4), dtype=np.uint64)
        arr = self.ctx.mont(values) if len(values) else np.zeros((1, 4), dtype=np.uint64)
        _ck(self.ctx, lib().zkb_mle_evaluate(self.ctx.handle, self._h, _p(arr), len(values), _p(out)))
        return self.ctx.unmont(out)[0]

    def sum_halves(self) -> List[int]:
        """get_round_partial_polynomial_proof (sum_check_protocol.rs:168-175)."""
        out = np.zeros((2, 4), dtype=np.uint64)
        _ck(self.ctx, lib().zkb_mle_sum_halves(self.ctx.handle, self._h, _p(out)))
        return self.ctx.unmont(out)

    def scale(self, value: int) -> "MultilinearPoly":  # :93-97
        h = C.c_uint64()
        _ck(self.ctx, lib().zkb_mle_scale(self.ctx.handle, self._h, _p(self.ctx.mont([value])), C.byref(h)))
        return MultilinearPoly(self.ctx, _handle=h.value)

    @staticmethod
    def tensor_add_mul_polynomials(a: "MultilinearPoly", b: "MultilinearPoly", op: Operation) -> "MultilinearPoly":  # :99-111
        h = C.c_uint64()
        _ck(a.ctx, lib().zkb_mle_tensor(a.ctx.handle, a._h, b._h, int(op), C.byref(h)))
        return MultilinearPoly(a.ctx, _handle=h.value)

    def _binary(self, other: "MultilinearPoly", op: int) -> "MultilinearPoly":  # :113-156
        h = C.c_uint64()
        _ck(self.ctx, lib().zkb_mle_binary(self.ctx.handle, self._h, other._h, op, C.byref(h)))
        return MultilinearPoly(self.ctx, _handle=h.value)

    def __add__(self, o):
        return self._binary(o, E.OP_ADD)

    def __sub__(self, o):
        return self._binary(o, E.OP_SUB)

    def __mul__(self, o):
        return self._binary(o, E.OP_MUL)


class ProductPoly:
    """composed_polynomial.rs:5-59"""

    def __init__(self, ctx: Context, evaluations: Sequence, *, _polys: Optional[List[MultilinearPoly]] = None):
        self.ctx = ctx
        if _polys is not None:
            self.evaluation = _polys
        else:
            n0 = len(evaluations[0])
            if any(len(e) != n0 for e in evaluations):
                raise ValueError("all evaluations must have same length")  # :20
            self.evaluation = [e if isinstance(e, MultilinearPoly) else MultilinearPoly(ctx, e) for e in evaluations]

    @classmethod
    def from_polys(cls, ctx: Context, polys: List[MultilinearPoly]) -> "ProductPoly":
        if any(len(q) != len(polys[0]) for q in polys):
            raise ValueError("all evaluations must have same length")
        return cls(ctx, [], _polys=list(polys))

    def evaluate(self, values: Sequence[int]) -> int:  # :31-36
        r = 1
        for q in self.evaluation:
            r = r * q.evaluate(values) % self.ctx.p
        return r

    def partial_evaluate(self, value: int) -> "ProductPoly":  # :38-50
        return ProductPoly.from_polys(self.ctx, [q.partial_evaluate(0, value) for q in self.evaluation])

    def reduce(self) -> List[int]:  # :52-54 (factors 0 and 1 only)
        return (self.evaluation[0] * self.evaluation[1]).evaluation

    def get_degree(self) -> int:  # :56-58
        return len(self.evaluation)


class SumPoly:
    """composed_polynomial.rs:10-13,61-104"""

    def __init__(self, ctx: Context, polys: List[ProductPoly]):
        d = polys[0].get_degree()
        if any(q.get_degree() != d for q in polys):
            raise ValueError("all product polys must have same degree")  # :65
        self.ctx = ctx
        self.polys = polys
        self._sp = 0

    def evaluate(self, values: Sequence[int]) -> int:  # :71-76
        return sum(q.evaluate(values) for q in self.polys) % self.ctx.p

    def partial_evaluate(self, value: int) -> "SumPoly":  # :78-86
        return SumPoly(self.ctx, [q.partial_evaluate(value) for q in self.polys])

    def reduce(self) -> List[int]:  # :88-99 (products 0 and 1 only)
        a, b = self.polys[0].reduce(), self.polys[1].reduce()
        return [(x + y) % self.ctx.p for x, y in zip(a, b)]

    def get_degree(self) -> int:  # :101-103
        return self.polys[0].get_degree()

    # -- device handle of the composed sumcheck state
    def handle(self) -> int:
        if not self._sp:
            tabs = [q.handle for pp in self.polys for q in pp.evaluation]
            arr = (C.c_uint64 * len(tabs))(*tabs)
            h = C.c_uint64()
            _ck(self.ctx, lib().zkb_sumpoly_create(self.ctx.handle, arr, len(self.polys), self.get_degree(), C.byref(h)))
            self._sp = h.value
        return self._sp

    def free(self) -> None:
        if self._sp:
            lib().zkb_sumpoly_free(self.ctx.handle, self._sp)
            self._sp = 0

    # step API (the body of gkr_prove's loop)
    def round_evals(self) -> List[int]:
        out = np.zeros((self.get_degree() + 1, 4), dtype=np.uint64)
        _ck(self.ctx, lib().zkb_sc_round_evals(self.ctx.handle, self.handle(), _p(out)))
        return self.ctx.unmont(out)

    def bind_and_next(self, r: int, last: bool = False) -> Optional[List[int]]:
        out = np.zeros((self.get_degree() + 1, 4), dtype=np.uint64)
        _ck(self.ctx, lib().zkb_sc_bind_and_next(self.ctx.handle, self.handle(), _p(self.ctx.mont([r])), None if last else _p(out)))
        return None if last else self.ctx.unmont(out)

    def final_values(self) -> List[int]:
        t = len(self.polys) * self.get_degree()
        out = np.zeros((t, 4), dtype=np.uint64)
        _ck(self.ctx, lib().zkb_sc_final_values(self.ctx.handle, self.handle(), _p(out)))
        return self.ctx.unmont(out)
