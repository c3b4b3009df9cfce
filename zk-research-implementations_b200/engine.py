"""ctypes binding of libzkb200.so (include/zkb200.h) and the device context.

This is the only place Python touches the C ABI.  There is NO fallback: if the
shared library is missing, or no CUDA device can be opened, the call raises.
Field elements are Python ints (canonical) at this level and numpy uint64
arrays of shape (n, 4) -- little-endian limbs -- on the way to the library,
which speaks ark-ff Montgomery residues (zkb200.h "Conventions").
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzkb200.so")

BN254_FR, BN254_FQ, BLS12_381_FR = 0, 1, 2
MODE_COMPAT, MODE_FULL = 0, 1
OP_ADD, OP_MUL, OP_SUB = 0, 1, 2

MODULI = {
    BN254_FR: 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
    BN254_FQ: 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
    BLS12_381_FR: 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
}

# zkb_status -> the reference's panic strings
_PANICS = {
    -2: "Invalid evaluations",
    -3: "Invalid number of values",
    -4: "all evaluations must have same length",
    -5: "all product polys must have same degree",
}

u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


class ZkbError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"zkb200 status {status}: {msg}")
        self.status = status


_lib = None


def lib() -> C.CDLL:
    """Load libzkb200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C zk-research-implementations_b200/csrc); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64, sz = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_size_t
    sig = {
        "zkb_strerror": (C.c_char_p, [i32]),
        "zkb_version": (C.c_char_p, []),
        "zkb_ctx_create": (i32, [i32, i32, i32, C.POINTER(vp)]),
        "zkb_ctx_destroy": (i32, [vp]),
        "zkb_ctx_last_error": (C.c_char_p, [vp]),
        "zkb_ctx_stream": (vp, [vp]),
        "zkb_ctx_launch_count": (u64, [vp]),
        "zkb_ctx_sync": (i32, [vp]),
        "zkb_ctx_profile": (i32, [vp, i32]),
        "zkb_ctx_profile_read": (i32, [vp, i32, C.POINTER(u64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "zkb_kernel_name": (C.c_char_p, [i32]),
        "zkb_sumpoly_reset": (i32, [vp, u64]),
        "zkb_comm_unique_id": (i32, [u8p]),
        "zkb_ctx_comm_init": (i32, [vp, i32, i32, u8p]),
        "zkb_ctx_set_gather_threshold": (i32, [vp, u32]),
        "zkb_ctx_set_tail_threshold": (i32, [vp, u32]),
        "zkb_ctx_set_small_threshold": (i32, [vp, u32]),
        "zkb_ctx_set_device_transcript": (i32, [vp, i32]),
        "zkb_ctx_device_transcript_stats": (i32, [vp, u64p, u64p]),
        "zkb_mle_upload": (i32, [vp, vp, u64, u64p]),
        "zkb_mle_upload_shard": (i32, [vp, vp, u64, u64p]),
        "zkb_mle_generate": (i32, [vp, u64, u64, u32, u64p]),
        "zkb_mle_download": (i32, [vp, u64, vp]),
        "zkb_mle_download_canonical": (i32, [vp, u64, vp]),
        "zkb_mle_clone": (i32, [vp, u64, u64p]),
        "zkb_mle_free": (i32, [vp, u64]),
        "zkb_mle_num_vars": (i32, [vp, u64, u32p]),
        "zkb_mle_partial_evaluate": (i32, [vp, u64, u32, u64p, u64p]),
        "zkb_mle_multi_partial_evaluate": (i32, [vp, u64, u64p, u32, u64p]),
        "zkb_mle_evaluate": (i32, [vp, u64, u64p, u32, u64p]),
        "zkb_mle_sum_halves": (i32, [vp, u64, u64p]),
        "zkb_mle_scale": (i32, [vp, u64, u64p, u64p]),
        "zkb_mle_binary": (i32, [vp, u64, u64, i32, u64p]),
        "zkb_mle_tensor": (i32, [vp, u64, u64, i32, u64p]),
        "zkb_sumpoly_create": (i32, [vp, u64p, u32, u32, u64p]),
        "zkb_sumpoly_free": (i32, [vp, u64]),
        "zkb_sumpoly_evaluate": (i32, [vp, u64, u64p, u32, u64p]),
        "zkb_sc_round_evals": (i32, [vp, u64, u64p]),
        "zkb_sc_bind_and_next": (i32, [vp, u64, u64p, u64p]),
        "zkb_sc_final_values": (i32, [vp, u64, u64p]),
        "zkb_transcript_new": (i32, [i32, C.POINTER(vp)]),
        "zkb_transcript_free": (i32, [vp]),
        "zkb_transcript_append": (i32, [vp, C.c_char_p, sz]),
        "zkb_transcript_append_elements": (i32, [vp, u64p, sz]),
        "zkb_transcript_challenge": (i32, [vp, u64p]),
        "zkb_keccak256": (i32, [C.c_char_p, sz, C.c_char_p]),
        "zkb_fe_to_mont": (i32, [i32, u64p, u64p, sz]),
        "zkb_fe_from_mont": (i32, [i32, u64p, u64p, sz]),
        "zkb_fe_reduce_wide": (i32, [i32, u64p, u64p, sz]),
        "zkb_uni_interpolate": (i32, [i32, u64p, u64p, u32, u64p, u32p]),
        "zkb_uni_evaluate": (i32, [i32, u64p, u32, u64p, u64p]),
        "zkb_sumcheck_prove": (i32, [vp, u64, u32, u64p, u64p, u64p]),
        "zkb_sumcheck_verify": (i32, [vp, u64, u32, u64p, u64p, u32, i32p]),
        "zkb_gkr_sumcheck_prove": (i32, [vp, vp, u64p, u64, u64p, i32p, u64p, u64p]),
        "zkb_gkr_sumcheck_verify": (i32, [vp, u32, u32, u64p, i32p, u64p, i32p, u64p, u64p]),
        "zkb_circuit_create": (i32, [vp, u32, u32p, u8p, u64p]),
        "zkb_circuit_free": (i32, [vp, u64]),
        "zkb_circuit_evaluate": (i32, [vp, u64, vp, u64, vp]),
        "zkb_layer_add_mul_i": (i32, [vp, u8p, u32, i32, u64p]),
        "zkb_gkr_prove": (i32, [vp, u64, vp, u64, u64p, u64p, i32p, u64p, u64p, u64p, u32p]),
        "zkb_gkr_verify": (i32, [vp, u64, vp, u64, u64p, u64p, i32p, u64p, u64p, i32p]),
        "zkb_gkr_total_rounds": (u32, [u32, u32p]),
        "zkb_circuit_create_wired": (i32, [vp, u32, u32p, u64, u8p, u32p, u32p, u64p]),
        "zkb_circuit_total_rounds": (i32, [vp, u64, u32p]),
        "zkb_gkr_prove_wired": (i32, [vp, u64, vp, u64, u64p, u64, u64p, i32p, u64p, u64p, u64p, u32p]),
        "zkb_gkr_verify_wired": (i32, [vp, u64, vp, u64, u64p, u64, u64p, i32p, u64p, u64p, i32p]),
        "zkb_proof_encode": (i32, [i32, i32, u32, u32, u64p, i32p, u64p, vp, sz, C.POINTER(sz)]),
        "zkb_proof_decode": (i32, [vp, sz, i32p, i32p, u32p, u32, u64p, i32p, u64p]),
        "zkb_kzg_setup": (i32, [vp, u32, u64p, u64p]),
        "zkb_kzg_free": (i32, [vp, u64]),
        "zkb_kzg_basis": (i32, [vp, u64, u32, u64, u64, vp]),
        "zkb_kzg_commit": (i32, [vp, u64, u64, vp]),
        "zkb_kzg_open": (i32, [vp, u64, u64, u64p, u32, u64p]),
        "zkb_kzg_get_proof": (i32, [vp, u64, u64, u64p, u64p, u32, vp]),
        "zkb_ctx_tensor_cores": (i32, [vp, i32p, i32p]),
        "zkb_ctx_small_cluster_max": (i32, [vp, i32p]),
        "zkb_tc_fold_matrices": (i32, [i32, u64p, vp]),
        "zkb_fft_evaluate": (i32, [vp, u64p, u64, u64p]),
        "zkb_fft_interpolate": (i32, [vp, u64p, u64, u64p]),
        "zkb_mle_ntt": (i32, [vp, u64, i32, u64p]),
        "zkb_merkle_build": (i32, [vp, u64p, u64, u32, u64p]),
        "zkb_merkle_free": (i32, [vp, u64]),
        "zkb_merkle_root": (i32, [vp, u64, u64p]),
        "zkb_merkle_nodes": (i32, [vp, u64, u32, u64, u64, u64p]),
        "zkb_merkle_update_leaf": (i32, [vp, u64, u64, u64p, i32]),
        "zkb_merkle_create_proof": (i32, [vp, u64, u64p, u64, u64p, vp]),
        "zkb_merkle_verify": (i32, [vp, u64, u64p, u64p, vp, u32, i32p]),
        "zkb_bench_modmul": (i32, [vp, i32, u32, C.POINTER(C.c_double)]),
        "zkb_bench_imad": (i32, [vp, i32, u32, C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = the library does not match include/zkb200.h
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTS = None  # filled by tests from include/zkb200.h


def tree_src_hash() -> str:
    """Hash of the library sources as they are in the tree (csrc/src_hash.py; the Makefile bakes the same
    hash into the library at build time)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_zkb_src_hash", os.path.join(_HERE, "csrc", "src_hash.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.src_hash()


def build_info() -> dict:
    """zkb_version() of the loaded library against the tree: `stale` means the .so was built from other sources."""
    v = lib().zkb_version().decode()
    lib_hash = v.rsplit("src:", 1)[1] if "src:" in v else "unknown"
    tree = tree_src_hash()
    return {"version": v, "lib_src_hash": lib_hash, "tree_src_hash": tree, "stale": lib_hash != tree}


# ------------------------------------------------------------- limb helpers
def ints_to_limbs(vals: Iterable[int]) -> np.ndarray:
    vals = [int(v) for v in vals]
    out = np.empty((len(vals), 4), dtype=np.uint64)
    m = (1 << 64) - 1
    for i, v in enumerate(vals):
        out[i, 0] = v & m
        out[i, 1] = (v >> 64) & m
        out[i, 2] = (v >> 128) & m
        out[i, 3] = (v >> 192) & m
    return out


def limbs_to_ints(a: np.ndarray) -> List[int]:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    return [int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192) for r in a]


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def to_mont(field: int, canonical: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(canonical, dtype=np.uint64)
    out = np.empty_like(a)
    _ck(None, lib().zkb_fe_to_mont(field, _p(a), _p(out), a.size // 4))
    return out


def from_mont(field: int, mont: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(mont, dtype=np.uint64)
    out = np.empty_like(a)
    _ck(None, lib().zkb_fe_from_mont(field, _p(a), _p(out), a.size // 4))
    return out


def _ck(ctx: Optional["Context"], status: int) -> None:
    if status == 0:
        return
    if status in _PANICS:  # the reference panics with these strings
        raise ValueError(_PANICS[status])
    msg = lib().zkb_strerror(status).decode()
    if ctx is not None and ctx._h:
        detail = lib().zkb_ctx_last_error(ctx._h).decode()
        if detail:
            msg = f"{msg} ({detail})"
    raise ZkbError(status, msg)


class Context:
    """One CUDA device + stream + scratch (zkb_ctx).  Not thread-safe: one per host thread."""

    def __init__(self, field: int = BN254_FR, device: int = 0, mode: int = MODE_COMPAT):
        self.field, self.device, self.mode = field, device, mode
        self.p = MODULI[field]
        self._h = C.c_void_p()
        self.rank, self.world = 0, 1
        _ck(None, lib().zkb_ctx_create(field, device, mode, C.byref(self._h)))

    # -- lifetime
    def close(self) -> None:
        if self._h:
            lib().zkb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def tensor_cores(self):
        """(enabled, persistent): see zkb_ctx_tensor_cores."""
        a, b = C.c_int32(), C.c_int32()
        _ck(self, lib().zkb_ctx_tensor_cores(self._h, C.byref(a), C.byref(b)))
        return bool(a.value), bool(b.value)

    def small_cluster_max(self) -> int:
        """CTAs of the largest cluster the on-chip round kernel uses: see zkb_ctx_small_cluster_max."""
        a = C.c_int32()
        _ck(self, lib().zkb_ctx_small_cluster_max(self._h, C.byref(a)))
        return a.value

    def sync(self) -> None:
        _ck(self, lib().zkb_ctx_sync(self._h))

    @property
    def stream(self) -> int:
        return int(lib().zkb_ctx_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(lib().zkb_ctx_launch_count(self._h))

    def profile(self, enable: bool) -> None:
        _ck(self, lib().zkb_ctx_profile(self._h, 1 if enable else 0))

    def profile_read(self) -> dict:
        """{kernel name: (launches, total ms, total algorithmic bytes)} since profile(True)."""
        out = {}
        for k in range(11):
            n, ms, by = C.c_uint64(), C.c_double(), C.c_double()
            _ck(self, lib().zkb_ctx_profile_read(self._h, k, C.byref(n), C.byref(ms), C.byref(by)))
            if n.value:
                out[lib().zkb_kernel_name(k).decode()] = (int(n.value), float(ms.value), float(by.value))
        return out

    def comm_init(self, rank: int, world: int, unique_id: bytes) -> None:
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _ck(self, lib().zkb_ctx_comm_init(self._h, rank, world, buf))
        self.rank, self.world = rank, world

    def set_tail_threshold(self, log2_entries: int) -> None:
        _ck(self, lib().zkb_ctx_set_tail_threshold(self._h, log2_entries))

    def set_small_threshold(self, smem_bytes: int) -> None:
        _ck(self, lib().zkb_ctx_set_small_threshold(self._h, smem_bytes))

    def set_device_transcript(self, enable: bool) -> None:
        _ck(self, lib().zkb_ctx_set_device_transcript(self._h, 1 if enable else 0))

    def device_transcript_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        _ck(self, lib().zkb_ctx_device_transcript_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_gather_threshold(self, log2_local: int) -> None:
        _ck(self, lib().zkb_ctx_set_gather_threshold(self._h, log2_local))

    # -- element conversion at the boundary
    def mont(self, vals: Sequence[int]) -> np.ndarray:
        if len(vals) == 0:
            return np.zeros((0, 4), dtype=np.uint64)
        return to_mont(self.field, ints_to_limbs([int(v) % self.p for v in vals]))

    def unmont(self, arr: np.ndarray) -> List[int]:
        if arr.size == 0:
            return []
        return limbs_to_ints(from_mont(self.field, arr))


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _ck(None, lib().zkb_comm_unique_id(buf))
    return bytes(buf)


def keccak256(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    _ck(None, lib().zkb_keccak256(data, len(data), out))
    return out.raw
