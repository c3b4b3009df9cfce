"""Device-backed mirror of `fft/src/fft.rs` (fft_evaluate :31-41, fft_interpolate :43-60): the radix-2 NTT over the ctx's
field with the root of unity ark-ff's `get_root_of_unity(n)` returns, natural order in and out."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from .engine import Context, _ck, _p, lib
from .multilinear_polynomial import MultilinearPoly


def _run(ctx: Context, vals: Sequence[int], fn) -> List[int]:
    n = len(vals)
    if n == 0 or n & (n - 1):
        raise ValueError("Length must be a power of 2")  # fft.rs:34,47 (0 is not a power of two either)
    src = ctx.mont(list(vals))
    out = np.zeros((n, 4), dtype=np.uint64)
    _ck(ctx, fn(ctx.handle, _p(src), n, _p(out)))
    return ctx.unmont(out)


def fft_evaluate(ctx: Context, coefficients: Sequence[int]) -> List[int]:
    return _run(ctx, coefficients, lib().zkb_fft_evaluate)


def fft_interpolate(ctx: Context, evaluations: Sequence[int]) -> List[int]:
    """Coefficients at full length n (UnivariatePoly::new does not trim, fft.rs:59)."""
    return _run(ctx, evaluations, lib().zkb_fft_interpolate)


def ntt(table: MultilinearPoly, inverse: bool = False) -> MultilinearPoly:
    """The same transform on a device-resident table."""
    h = C.c_uint64()
    _ck(table.ctx, lib().zkb_mle_ntt(table.ctx.handle, table.handle, 1 if inverse else 0, C.byref(h)))
    return MultilinearPoly(table.ctx, _handle=h.value)
