"""Mirror of the reference crate `fiat_shamir` (fiat_shamir_transcript.rs:5-37).
The Keccak-256 transcript runs on the host inside libzkb200 (host_math.hpp)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import engine as E
from .engine import _ck, _p, lib


def fq_vec_to_bytes(values: Sequence[int]) -> bytes:  # :32-37
    return b"".join(int(v).to_bytes(32, "little") for v in values)


class Transcript:
    def __init__(self, field: int = E.BN254_FR):  # :12-17
        self.field = field
        self.p = E.MODULI[field]
        self._h = C.c_void_p()
        _ck(None, lib().zkb_transcript_new(field, C.byref(self._h)))

    def append(self, incoming_data: bytes) -> None:  # :19-21
        _ck(None, lib().zkb_transcript_append(self._h, incoming_data, len(incoming_data)))

    def get_random_challenge(self) -> int:  # :23-29
        out = np.zeros((1, 4), dtype=np.uint64)
        _ck(None, lib().zkb_transcript_challenge(self._h, _p(out)))
        return E.limbs_to_ints(E.from_mont(self.field, out))[0]

    @property
    def handle(self):
        return self._h

    def __del__(self):
        try:
            if self._h:
                lib().zkb_transcript_free(self._h)
        except Exception:
            pass
