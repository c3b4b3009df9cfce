"""Mirror of `univariate_polynomial::univariate_polynomial_dense::UnivariatePoly`
(univariate_polynomial_dense.rs:4-109) -- the part the sumcheck path uses:
interpolate (+ trim) and evaluate, executed by libzkb200's host code."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import engine as E
from .engine import _ck, _p, lib


class UnivariatePoly:
    def __init__(self, coefficients: Sequence[int], field: int = E.BN254_FR):
        self.field = field
        self.p = E.MODULI[field]
        c = [int(x) % self.p for x in coefficients]
        while c and c[-1] == 0:  # trim :14-18
            c.pop()
        self.coefficients = c

    def degree(self) -> int:  # :28-32
        return max(len(self.coefficients) - 1, 0)

    def evaluate(self, x: int) -> int:  # :20-26
        out = np.zeros((1, 4), dtype=np.uint64)
        n = len(self.coefficients)
        co = E.to_mont(self.field, E.ints_to_limbs(self.coefficients)) if n else np.zeros((1, 4), dtype=np.uint64)
        xm = E.to_mont(self.field, E.ints_to_limbs([int(x) % self.p]))
        _ck(None, lib().zkb_uni_evaluate(self.field, _p(co), n, _p(xm), _p(out)))
        return E.limbs_to_ints(E.from_mont(self.field, out))[0]

    @classmethod
    def interpolate(cls, points: Sequence[Tuple[int, int]], field: int = E.BN254_FR) -> "UnivariatePoly":  # :48-74
        p = E.MODULI[field]
        n = len(points)
        xs = E.to_mont(field, E.ints_to_limbs([int(x) % p for x, _ in points]))
        ys = E.to_mont(field, E.ints_to_limbs([int(y) % p for _, y in points]))
        out = np.zeros((max(n, 1), 4), dtype=np.uint64)
        ln = C.c_uint32()
        _ck(None, lib().zkb_uni_interpolate(field, _p(xs), _p(ys), n, _p(out), C.byref(ln)))
        co: List[int] = E.limbs_to_ints(E.from_mont(field, out[: ln.value])) if ln.value else []
        return cls(co, field)
