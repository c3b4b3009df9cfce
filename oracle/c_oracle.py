"""ctypes loader for oracle/libzkoracle.so (the C restatement, zk_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of zk_oracle.c.  Field elements are
numpy uint64 arrays of shape (..., 4): little-endian limbs of the CANONICAL
integer.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzkoracle.so")
_lib = None

u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "zk_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libzkoracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.zko_transcript_new.restype = C.c_void_p
        _lib.zko_transcript_new.argtypes = [C.c_int]
        _lib.zko_transcript_free.argtypes = [C.c_void_p]
        _lib.zko_transcript_append.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        _lib.zko_transcript_challenge.argtypes = [C.c_void_p, u64p]
        _lib.zko_keccak256.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        _lib.zko_to_mont.argtypes = [C.c_int, u64p, u64p, C.c_size_t]
        _lib.zko_from_mont.argtypes = [C.c_int, u64p, u64p, C.c_size_t]
        _lib.zko_synth_table.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64,
                                         C.c_uint64, u64p]
        _lib.zko_mle_partial_evaluate.argtypes = [C.c_int, u64p, C.c_uint32, C.c_uint32, u64p, u64p]
        _lib.zko_mle_evaluate.argtypes = [C.c_int, u64p, C.c_uint32, u64p, u64p]
        _lib.zko_uni_interpolate.argtypes = [C.c_int, u64p, u64p, C.c_int, u64p]
        _lib.zko_sumcheck_prove.argtypes = [C.c_int, u64p, C.c_uint32, C.c_int, u64p, u64p, u64p]
        _lib.zko_sumcheck_verify.argtypes = [C.c_int, u64p, C.c_uint32, C.c_int, C.c_int, u64p, u64p, C.c_uint32]
        _lib.zko_gkr_sumcheck_prove.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(u64p), C.c_uint32,
                                                u64p, i32p, u64p, u64p, u64p]
        _lib.zko_gkr_sumcheck_verify.argtypes = [C.c_void_p, C.c_int, C.c_int, u64p, i32p, u64p, u64p, u64p]
        _lib.zko_circuit_evaluate.argtypes = [C.c_int, C.c_int, u32p, u8p, u64p, C.c_size_t, u64p]
        _lib.zko_gkr_prove.argtypes = [C.c_int, C.c_int, u32p, u8p, u64p, C.c_size_t, u64p, u64p, i32p, u64p, u64p,
                                       u64p]
        _lib.zko_gkr_prove_wired.argtypes = [C.c_int, C.c_int, u32p, C.c_size_t, u8p, u32p, u32p, u64p, u64p, u64p, i32p, u64p,
                                             u64p, u64p]
        _lib.zko_vec_op.argtypes = [C.c_int, C.c_int, u64p, u64p, u64p, C.c_size_t]
        _lib.zko_set_threads.argtypes = [C.c_int]
        _lib.zko_max_threads.restype = C.c_int
    return _lib


# ---------------------------------------------------------------- helpers
def ints_to_arr(vals: Sequence[int]) -> np.ndarray:
    out = np.empty((len(vals), 4), dtype=np.uint64)
    m = (1 << 64) - 1
    for i, v in enumerate(vals):
        v = int(v)
        out[i, 0] = v & m
        out[i, 1] = (v >> 64) & m
        out[i, 2] = (v >> 128) & m
        out[i, 3] = (v >> 192) & m
    return out


def arr_to_ints(a: np.ndarray) -> List[int]:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    return [int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192) for r in a]


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def set_threads(n: int) -> None:
    lib().zko_set_threads(int(n))


def max_threads() -> int:
    return int(lib().zko_max_threads())


def keccak256(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    lib().zko_keccak256(data, len(data), out)
    return out.raw


def to_mont(field: int, a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().zko_to_mont(field, _p(a), _p(out), a.size // 4)
    return out


def from_mont(field: int, a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().zko_from_mont(field, _p(a), _p(out), a.size // 4)
    return out


def synth_table(field: int, seed: int, table: int, n_vars: int, first: int = 0, stride: int = 1,
                count: int | None = None) -> np.ndarray:
    if count is None:
        count = 1 << n_vars
    out = np.empty((count, 4), dtype=np.uint64)
    lib().zko_synth_table(field, seed, table, n_vars, first, stride, count, _p(out))
    return out


def mle_partial_evaluate(field: int, table: np.ndarray, bit: int, r: int) -> np.ndarray:
    table = np.ascontiguousarray(table, dtype=np.uint64)
    n = table.shape[0]
    nv = n.bit_length() - 1
    out = np.empty((n // 2, 4), dtype=np.uint64)
    rc = lib().zko_mle_partial_evaluate(field, _p(table), nv, bit, _p(ints_to_arr([r])), _p(out))
    assert rc == 0
    return out


def mle_evaluate(field: int, table: np.ndarray, rs: Sequence[int]) -> int:
    table = np.ascontiguousarray(table, dtype=np.uint64)
    nv = table.shape[0].bit_length() - 1
    assert len(rs) == nv
    out = np.zeros((1, 4), dtype=np.uint64)
    ra = ints_to_arr(rs) if nv else np.zeros((1, 4), dtype=np.uint64)
    lib().zko_mle_evaluate(field, _p(table), nv, _p(ra), _p(out))
    return arr_to_ints(out)[0]


def uni_interpolate(field: int, xs: Sequence[int], ys: Sequence[int]) -> List[int]:
    out = np.zeros((8, 4), dtype=np.uint64)
    ln = lib().zko_uni_interpolate(field, _p(ints_to_arr(xs)), _p(ints_to_arr(ys)), len(xs), _p(out))
    return arr_to_ints(out[:ln])


def sumcheck_prove(field: int, table: np.ndarray, absorb_table: bool = True):
    """-> (claimed_sum, msgs[n][2], challenges[n]) as Python ints."""
    table = np.ascontiguousarray(table, dtype=np.uint64)
    nv = table.shape[0].bit_length() - 1
    claimed = np.zeros((1, 4), dtype=np.uint64)
    msgs = np.zeros((max(nv, 1), 2, 4), dtype=np.uint64)
    chals = np.zeros((max(nv, 1), 4), dtype=np.uint64)
    lib().zko_sumcheck_prove(field, _p(table), nv, int(absorb_table), _p(claimed), _p(msgs), _p(chals))
    m = arr_to_ints(msgs[:nv])
    return arr_to_ints(claimed)[0], [m[2 * i: 2 * i + 2] for i in range(nv)], arr_to_ints(chals[:nv])


def sumcheck_verify(field: int, table: np.ndarray, claimed: int, msgs: Sequence[Sequence[int]],
                    absorb_table: bool = True, redundant_fold: bool = False) -> bool:
    table = np.ascontiguousarray(table, dtype=np.uint64)
    nv = table.shape[0].bit_length() - 1
    flat = ints_to_arr([x for m in msgs for x in m]) if msgs else np.zeros((1, 4), dtype=np.uint64)
    return bool(lib().zko_sumcheck_verify(field, _p(table), nv, int(absorb_table), int(redundant_fold),
                                          _p(ints_to_arr([claimed])), _p(flat), len(msgs)))


class Transcript:
    def __init__(self, field: int):
        self.field = field
        self.h = lib().zko_transcript_new(field)

    def append(self, data: bytes) -> None:
        lib().zko_transcript_append(self.h, data, len(data))

    def challenge(self) -> int:
        out = np.zeros((1, 4), dtype=np.uint64)
        lib().zko_transcript_challenge(self.h, _p(out))
        return arr_to_ints(out)[0]

    def __del__(self):
        try:
            lib().zko_transcript_free(self.h)
        except Exception:
            pass


def gkr_sumcheck_prove(tr: Transcript, mode: int, P: int, d: int, tables: Sequence[np.ndarray]):
    """tables: P*d canonical tables, product-major.  -> dict(coeffs (trimmed lists),
    challenges, evals[n][d+1], final_vals[P*d])."""
    tabs = [np.ascontiguousarray(t, dtype=np.uint64) for t in tables]
    assert len(tabs) == P * d
    nv = tabs[0].shape[0].bit_length() - 1
    ptrs = (u64p * len(tabs))(*[_p(t) for t in tabs])
    coeffs = np.zeros((max(nv, 1), d + 1, 4), dtype=np.uint64)
    lens = np.zeros(max(nv, 1), dtype=np.int32)
    chals = np.zeros((max(nv, 1), 4), dtype=np.uint64)
    evals = np.zeros((max(nv, 1), d + 1, 4), dtype=np.uint64)
    fin = np.zeros((P * d, 4), dtype=np.uint64)
    rc = lib().zko_gkr_sumcheck_prove(tr.h, mode, P, d, ptrs, nv, _p(coeffs), lens.ctypes.data_as(i32p), _p(chals),
                                      _p(evals), _p(fin))
    if rc != 0:
        raise ValueError(f"zko_gkr_sumcheck_prove rc={rc}")
    cl = [arr_to_ints(coeffs[k, : lens[k]]) for k in range(nv)]
    ev = [arr_to_ints(evals[k]) for k in range(nv)]
    return dict(coeffs=cl, challenges=arr_to_ints(chals[:nv]), evals=ev, final_vals=arr_to_ints(fin))


def gkr_sumcheck_verify(tr: Transcript, coeffs: Sequence[Sequence[int]], claimed: int):
    n = len(coeffs)
    slots = max([len(c) for c in coeffs] + [1])
    ca = np.zeros((max(n, 1), slots, 4), dtype=np.uint64)
    lens = np.zeros(max(n, 1), dtype=np.int32)
    for k, c in enumerate(coeffs):
        lens[k] = len(c)
        if c:
            ca[k, : len(c)] = ints_to_arr(c)
    fin = np.zeros((1, 4), dtype=np.uint64)
    chals = np.zeros((max(n, 1), 4), dtype=np.uint64)
    ok = lib().zko_gkr_sumcheck_verify(tr.h, n, slots, _p(ca), lens.ctypes.data_as(i32p), _p(ints_to_arr([claimed])),
                                       _p(fin), _p(chals))
    return bool(ok), arr_to_ints(fin)[0], arr_to_ints(chals[:n] if ok else chals[:1])


def circuit_evaluate(field: int, gates: Sequence[int], ops: np.ndarray, inputs: np.ndarray) -> List[np.ndarray]:
    g = np.asarray(gates, dtype=np.uint32)
    ops = np.ascontiguousarray(ops, dtype=np.uint8)
    inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
    out = np.zeros((int(g.sum()), 4), dtype=np.uint64)
    lib().zko_circuit_evaluate(field, len(g), g.ctypes.data_as(u32p), ops.ctypes.data_as(u8p), _p(inputs),
                               inputs.shape[0], _p(out))
    res, off = [], 0
    for k in g:
        res.append(out[off: off + int(k)])
        off += int(k)
    return res


def gkr_prove(field: int, gates: Sequence[int], ops: np.ndarray, inputs: np.ndarray):
    """gates: per-layer gate counts, input side first.  -> dict as GkrProof of pyref."""
    g = np.asarray(gates, dtype=np.uint32)
    ops = np.ascontiguousarray(ops, dtype=np.uint8)
    inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
    L = len(g)
    rounds_per_layer = [2 * max(1, int(2 * k).bit_length() - 1) for k in g[::-1]]
    total = sum(rounds_per_layer)
    w0 = np.zeros((2, 4), dtype=np.uint64)
    coeffs = np.zeros((total, 3, 4), dtype=np.uint64)
    lens = np.zeros(total, dtype=np.int32)
    chals = np.zeros((total, 4), dtype=np.uint64)
    claimed = np.zeros((max(L - 1, 1), 2, 4), dtype=np.uint64)
    fin = np.zeros((2, 4), dtype=np.uint64)
    rc = lib().zko_gkr_prove(field, L, g.ctypes.data_as(u32p), ops.ctypes.data_as(u8p), _p(inputs), inputs.shape[0],
                             _p(w0), _p(coeffs), lens.ctypes.data_as(i32p), _p(chals), _p(claimed), _p(fin))
    if rc < 0:
        raise ValueError(f"zko_gkr_prove rc={rc}")
    assert rc == total, (rc, total)
    polys, ch, off = [], [], 0
    for nr in rounds_per_layer:
        polys.append([arr_to_ints(coeffs[off + k, : lens[off + k]]) for k in range(nr)])
        ch.append(arr_to_ints(chals[off: off + nr]))
        off += nr
    ce = arr_to_ints(claimed[: L - 1]) if L > 1 else []
    return dict(output_poly=arr_to_ints(w0), proof_polynomials=polys,
                claimed_evaluations=[(ce[2 * i], ce[2 * i + 1]) for i in range(L - 1)],
                final_openings=tuple(arr_to_ints(fin)), challenges=ch)


def gkr_prove_wired(field: int, n_inputs: int, layers, inputs: np.ndarray, want_challenges: bool = True):
    """General wiring (extension): layers = [(ops, in1, in2), ...] input side first.  -> dict as gkr_prove."""
    g = np.array([len(l[0]) for l in layers], dtype=np.uint32)
    ops = np.ascontiguousarray(np.concatenate([np.asarray(l[0], dtype=np.uint8) for l in layers]))
    in1 = np.ascontiguousarray(np.concatenate([np.asarray(l[1], dtype=np.uint32) for l in layers]))
    in2 = np.ascontiguousarray(np.concatenate([np.asarray(l[2], dtype=np.uint32) for l in layers]))
    inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
    L = len(g)
    widths = [int(n_inputs)] + [int(x) for x in g[:-1]]
    rounds_per_layer = [2 * (w.bit_length() - 1) for w in widths[::-1]]
    total = sum(rounds_per_layer)
    n0 = max(int(g[-1]), 2)
    w0 = np.zeros((n0, 4), dtype=np.uint64)
    coeffs = np.zeros((total, 3, 4), dtype=np.uint64)
    lens = np.zeros(total, dtype=np.int32)
    chals = np.zeros((total, 4), dtype=np.uint64)
    claimed = np.zeros((max(L - 1, 1), 2, 4), dtype=np.uint64)
    fin = np.zeros((2, 4), dtype=np.uint64)
    rc = lib().zko_gkr_prove_wired(field, L, g.ctypes.data_as(u32p), int(n_inputs), ops.ctypes.data_as(u8p), in1.ctypes.data_as(u32p),
                                   in2.ctypes.data_as(u32p), _p(inputs), _p(w0), _p(coeffs), lens.ctypes.data_as(i32p), _p(chals),
                                   _p(claimed), _p(fin))
    if rc < 0:
        raise ValueError(f"zko_gkr_prove_wired rc={rc}")
    assert rc == total, (rc, total)
    polys, ch, off = [], [], 0
    for nr in rounds_per_layer:
        polys.append([arr_to_ints(coeffs[off + k, : lens[off + k]]) for k in range(nr)])
        if want_challenges:
            ch.append(arr_to_ints(chals[off: off + nr]))
        off += nr
    ce = arr_to_ints(claimed[: L - 1]) if L > 1 else []
    return dict(output_poly=arr_to_ints(w0) if n0 <= (1 << 16) else w0, proof_polynomials=polys,
                claimed_evaluations=[(ce[2 * i], ce[2 * i + 1]) for i in range(L - 1)],
                final_openings=tuple(arr_to_ints(fin)), challenges=ch)


def vec_op(field: int, op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty_like(a)
    lib().zko_vec_op(field, op, _p(a), _p(b), _p(out), a.size // 4)
    return out
