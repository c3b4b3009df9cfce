"""fr.cuh compiled for the HOST (carry flag emulated): the limb-level Montgomery
multiplier / add / sub that the CUDA kernels run, checked against Python ints.
Operands are Montgomery residues as raw integers: mul(a,b) = a*b*R^-1 mod p."""
import ctypes as C
import os
import random
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import pyref as R
from oracle.c_oracle import arr_to_ints, ints_to_arr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "zk-research-implementations_b200", "csrc")


@pytest.fixture(scope="module")
def shim():
    d = tempfile.mkdtemp(prefix="zkb_host_fr_")
    so = os.path.join(d, "host_fr.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", CSRC,
                           os.path.join(ROOT, "tests", "host_fr_shim.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.host_fr_op.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    return lib


def run(lib, fid, op, a, b):
    A, B = ints_to_arr(a), ints_to_arr(b)
    out = np.zeros_like(A)
    lib.host_fr_op(fid, op, A.ctypes.data, B.ctypes.data, out.ctypes.data, len(a))
    return arr_to_ints(out)


@pytest.mark.parametrize("fid", [0, 1, 2])
def test_limb_arithmetic(shim, fid):
    p = R.MODULI_BY_ID[fid]
    rinv = pow(R.R256, -1, p)
    rng = random.Random(fid)
    edge = [0, 1, 2, p - 1, p - 2, R.R256 % p, p >> 1, (1 << 32) - 1, 1 << 32, (1 << 224) - 1, p - (1 << 32),
            (1 << 253) - 1, ((1 << 256) - 1) % p]
    a = [x for x in edge for _ in edge] + [rng.randrange(p) for _ in range(3000)]
    b = [y for _ in edge for y in edge] + [rng.randrange(p) for _ in range(3000)]
    assert run(shim, fid, 0, a, b) == [(x + y) % p for x, y in zip(a, b)]
    assert run(shim, fid, 1, a, b) == [(x - y) % p for x, y in zip(a, b)]
    want = [x * y * rinv % p for x, y in zip(a, b)]
    assert run(shim, fid, 2, a, b) == want      # IMAD.WIDE formulation
    assert run(shim, fid, 3, a, b) == want      # 32-bit lo/hi formulation
    assert run(shim, fid, 4, a, b) == [x * R.R256 % p for x in a]
    assert run(shim, fid, 5, a, b) == [x * rinv % p for x in a]
    assert run(shim, fid, 7, a, b) == want      # fixed-multiplicand product (challenge table)
    n = len(a)
    assert run(shim, fid, 8, a, b) == [sum(a[(i + k) % n] * b[(i + k) % n] for k in range(4)) * rinv % p for i in range(n)]
    assert run(shim, fid, 9, a[:200], b[:200]) == [5000 * x * y * rinv % p for x, y in zip(a[:200], b[:200])]
    assert run(shim, fid, 10, a, b) == [(x - y) % p for x, y in zip(a, b)]      # unreduced difference
    assert run(shim, fid, 11, a, b) == [(2 * y - x) ** 2 * rinv % p for x, y in zip(a, b)]  # unreduced line value at t = 2
    # fold(a, b, r=to_mont(b)) in Montgomery arithmetic == a + b*(b-a) on raw residues
    assert run(shim, fid, 6, a, b) == [(x + y * (y - x)) % p for x, y in zip(a, b)]
