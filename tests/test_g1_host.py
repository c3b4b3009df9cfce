"""g1.cuh compiled for the HOST (carry flag emulated): the 12-limb Montgomery multiplier of BLS12-381 Fq and the G1
group law used by the KZG kernels, against Python integers / the affine oracle (oracle/kzg_ref.py)."""
import ctypes as C
import os
import random
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from oracle import kzg_ref as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "zk-research-implementations_b200", "csrc")
Q = K.Q
RM = 1 << 384


@pytest.fixture(scope="module")
def shim():
    d = tempfile.mkdtemp(prefix="zkb_host_g1_")
    so = os.path.join(d, "host_g1.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", CSRC,
                           os.path.join(ROOT, "tests", "host_g1_shim.cpp"), "-o", so])
    lib = C.CDLL(so)
    for f in (lib.host_fq_op, lib.host_g1_op):
        f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    return lib


def to_arr(vals, words):
    out = np.zeros((len(vals), words), dtype=np.uint32)
    for i, v in enumerate(vals):
        for k in range(words):
            out[i, k] = (v >> (32 * k)) & 0xFFFFFFFF
    return out


def from_arr(a):
    return [sum(int(a[i, k]) << (32 * k) for k in range(a.shape[1])) for i in range(a.shape[0])]


def fq(lib, op, a, b):
    A, B = to_arr(a, 12), to_arr(b, 12)
    out = np.zeros_like(A)
    lib.host_fq_op(op, A.ctypes.data, B.ctypes.data, out.ctypes.data, len(a))
    return from_arr(out)


def pts_arr(pts):
    return to_arr([0 if p is None else p[0] | (p[1] << 384) for p in pts], 24)


def g1(lib, op, P, Qs):
    A, B = pts_arr(P), pts_arr(Qs) if not isinstance(Qs, np.ndarray) else Qs
    out = np.zeros_like(A)
    lib.host_g1_op(op, A.ctypes.data, B.ctypes.data, out.ctypes.data, len(P))
    res = []
    for v in from_arr(out):
        x, y = v & ((1 << 384) - 1), v >> 384
        res.append(None if x == 0 and y == 0 else (x, y))
    return res


def test_header_constants():
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_fq_constants.py"), "--check"])


def test_fq_limb_arithmetic(shim):
    rinv = pow(RM, -1, Q)
    rng = random.Random(1)
    edge = [0, 1, 2, Q - 1, Q - 2, RM % Q, Q >> 1, (1 << 32) - 1, 1 << 32, (1 << 380) - 1, Q - (1 << 32), (1 << 352) + 12345]
    a = [x for x in edge for _ in edge] + [rng.randrange(Q) for _ in range(2000)]
    b = [y for _ in edge for y in edge] + [rng.randrange(Q) for _ in range(2000)]
    assert fq(shim, 0, a, b) == [(x + y) % Q for x, y in zip(a, b)]
    assert fq(shim, 1, a, b) == [(x - y) % Q for x, y in zip(a, b)]
    assert fq(shim, 2, a, b) == [x * y * rinv % Q for x, y in zip(a, b)]
    assert fq(shim, 3, a, b) == [x * RM % Q for x in a]
    assert fq(shim, 4, a, b) == [x * rinv % Q for x in a]
    nz = [x for x in a[:60] if x]
    assert fq(shim, 5, nz, nz) == [pow(x * rinv % Q, -1, Q) * RM % Q for x in nz]  # Montgomery in, Montgomery out


def test_g1_group_law(shim):
    rng = random.Random(2)
    ks = [1, 2, 3, 5, K.R - 1, K.R - 2] + [rng.randrange(K.R) for _ in range(10)]
    P = [K.g1_mul(K.G1, k) for k in ks]
    Qp = [K.g1_mul(K.G1, rng.randrange(K.R)) for _ in ks]
    assert g1(shim, 4, P[:1], P[:1]) == [K.G1]
    assert g1(shim, 0, P, Qp) == [K.g1_add(p, q) for p, q in zip(P, Qp)]
    assert g1(shim, 1, P, Qp) == [K.g1_add(p, q) for p, q in zip(P, Qp)]
    assert g1(shim, 2, P, P) == [K.g1_add(p, p) for p in P]
    # special cases: P + P (doubling through add), P + (-P), infinity on either side
    neg = [K.g1_neg(p) for p in P]
    inf = [None] * len(P)
    for op in (0, 1):
        assert g1(shim, op, P, P) == [K.g1_add(p, p) for p in P]
        assert g1(shim, op, P, neg) == inf
        assert g1(shim, op, P, inf) == P
        assert g1(shim, op, inf, P) == P
    assert g1(shim, 2, inf, inf) == inf
    # small scalar multiples and a chain through Jacobian intermediates
    small = [0, 1, 2, 3, 0xFFFF, 0x10000, 0xFFFFFFFF, 12345]
    Pb = [P[i % len(P)] for i in range(len(small))]
    karr = np.zeros((len(small), 24), dtype=np.uint32)
    karr[:, 0] = small
    assert g1(shim, 3, Pb, karr) == [K.g1_mul(p, k) for p, k in zip(Pb, small)]
    assert g1(shim, 5, P, Qp) == [K.g1_add(K.g1_mul(p, (1 << 32) + 2), K.g1_mul(q, 2)) for p, q in zip(P, Qp)]
