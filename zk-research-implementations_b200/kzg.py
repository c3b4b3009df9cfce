"""Device-backed mirror of `pcs::kzg_pcs::kzg::KZG` (prover side, pcs/src/kzg_pcs/kzg.rs:11-95): the multilinear KZG
commitment of the GKR input layer (gkr/src/gkr_protocol.rs:92-118) over BLS12-381 G1.  Everything runs in libzkb200
(fixed-base window multiplication for the Lagrange basis, bucket MSM for commitments and quotient proofs); points come
back as affine coordinate pairs of Python ints, `None` for the point at infinity.  The G2 powers and `verify`
(:97-129, pairings) are verifier-side and not provided."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import engine as E
from .engine import _ck, _p, lib
from .multilinear_polynomial import MultilinearPoly

Point = Optional[Tuple[int, int]]


def _point(b: bytes) -> Point:
    x, y = int.from_bytes(b[:48], "little"), int.from_bytes(b[48:96], "little")
    return None if x == 0 and y == 0 else (x, y)


class KZG:
    def __init__(self, polynomial: MultilinearPoly, taus: Sequence[int]):  # KZG::new, kzg.rs:17-34
        ctx = self.ctx = polynomial.ctx
        if len(taus) != polynomial.num_of_vars:
            raise ValueError("invalid taus or polynomials")  # :19-21
        self.n_vars = polynomial.num_of_vars
        h = C.c_uint64()
        _ck(ctx, lib().zkb_kzg_setup(ctx.handle, self.n_vars, _p(ctx.mont(list(taus))), C.byref(h)))
        self._h = h.value

    def free(self) -> None:
        if self._h:
            lib().zkb_kzg_free(self.ctx.handle, self._h)
            self._h = 0

    def lagrange_basis(self, level: int = 0, first: int = 0, count: Optional[int] = None) -> List[Point]:
        n = 1 << (self.n_vars - level)
        count = n - first if count is None else count
        out = np.zeros(96 * max(count, 1), dtype=np.uint8)
        _ck(self.ctx, lib().zkb_kzg_basis(self.ctx.handle, self._h, level, first, count, out.ctypes.data))
        raw = out.tobytes()
        return [_point(raw[96 * i: 96 * i + 96]) for i in range(count)]

    def commit(self, poly: MultilinearPoly) -> Point:  # :51-53
        out = np.zeros(96, dtype=np.uint8)
        _ck(self.ctx, lib().zkb_kzg_commit(self.ctx.handle, self._h, poly.handle, out.ctypes.data))
        return _point(out.tobytes())

    def open(self, opening_values: Sequence[int], poly: MultilinearPoly) -> int:  # :55-57
        out = np.zeros((1, 4), dtype=np.uint64)
        _ck(self.ctx, lib().zkb_kzg_open(self.ctx.handle, self._h, poly.handle, _p(self.ctx.mont(list(opening_values))),
                                         len(opening_values), _p(out)))
        return self.ctx.unmont(out)[0]

    def get_proof(self, opened_value: int, opening_values: Sequence[int], poly: MultilinearPoly) -> List[Point]:  # :59-95
        n = len(opening_values)
        out = np.zeros(96 * max(n, 1), dtype=np.uint8)
        _ck(self.ctx, lib().zkb_kzg_get_proof(self.ctx.handle, self._h, poly.handle, _p(self.ctx.mont([opened_value])),
                                              _p(self.ctx.mont(list(opening_values))), n, out.ctypes.data))
        raw = out.tobytes()
        return [_point(raw[96 * i: 96 * i + 96]) for i in range(n)]
