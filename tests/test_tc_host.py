"""The byte matrices of the tensor-core fold (csrc/tcfold.cuh, built on the host by TcMatsBuilder) checked with plain integers,
no GPU: column k is T1_k = (1 - r) 2^(8k+32) mod p resp. T2_k = r 2^(8k+32) mod p, and for any two table entries a, b
  S = sum_k a_k T1_k + b_k T2_k  ==  (a + r (b - a)) 2^32  (mod p),   S < 2^14 p,   every byte-column sum < 2^22
-- the identity and the bounds the kernel's one Montgomery row and one conditional subtraction rely on."""
import ctypes as C
import importlib
import random

import numpy as np
import pytest

from oracle import pyref as R

FIELDS = [(0, R.BN254_FR), (1, R.BN254_FQ), (2, R.BLS12_381_FR)]


@pytest.fixture(scope="module")
def E(zkb):
    return zkb.engine


def matrices(E, fid, p, r):
    r_mont = E.to_mont(fid, E.ints_to_limbs([r]))
    out = np.zeros(2048, dtype=np.uint8)
    assert E.lib().zkb_tc_fold_matrices(fid, E._p(r_mont), out.ctypes.data) == 0
    cols = []
    for half in range(2):
        m = out[1024 * half:1024 * half + 1024]
        cols.append([sum(int(m[(k // 16) * 512 + n * 16 + k % 16]) << (8 * n) for n in range(32)) for k in range(32)])
    return out, cols


@pytest.mark.parametrize("fid,p", FIELDS)
def test_fold_matrices(E, fid, p):
    rng = random.Random(fid)
    for r in [0, 1, p - 1, rng.randrange(p), rng.randrange(p)]:
        raw, (t1, t2) = matrices(E, fid, p, r)
        for k in range(32):
            assert t1[k] == (1 - r) * pow(2, 8 * k + 32, p) % p and t2[k] == r * pow(2, 8 * k + 32, p) % p
        for a, b in [(0, 0), (p - 1, p - 1), ((1 << 248) - 1, p - 1), (rng.randrange(p), rng.randrange(p))]:
            ab, bb = a.to_bytes(32, "little"), b.to_bytes(32, "little")
            S = sum(ab[k] * t1[k] + bb[k] * t2[k] for k in range(32))
            assert S % p == (a + r * (b - a)) * (1 << 32) % p
            assert S < (1 << 14) * p
            # the 32 s32 column sums of the two matrix products (what the tensor core leaves in TMEM)
            colsum = [sum(ab[k] * ((t1[k] >> (8 * n)) & 255) + bb[k] * ((t2[k] >> (8 * n)) & 255) for k in range(32)) for n in range(32)]
            assert max(colsum) < 1 << 22 and sum(c << (8 * n) for n, c in enumerate(colsum)) == S
