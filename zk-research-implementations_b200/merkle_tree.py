"""Device-backed mirror of `merkle_tree/src/merkle_tree.rs` (MerkleTree :24-214): a Keccak-256 Merkle tree over field
elements, leaves and levels resident on the device (one hash per thread)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from .engine import Context, ZkbError, _ck, _p, lib

LEFT, RIGHT = 0, 1  # LeafSide (:6-10)


@dataclass
class MerkleProof:  # :18-22
    data: int
    proof: List[Tuple[int, int]]  # (data_hash, data_side)


class MerkleTree:
    def __init__(self, ctx: Context, depth: int, inputs: Sequence[int] = ()):  # new (:31-50) / new_with_inputs (:52-84)
        self.ctx, self.depth = ctx, depth
        if len(inputs) > (1 << depth):
            raise ValueError("Too many inputs for tree depth")  # :54-56
        h = C.c_uint64()
        arr = ctx.mont(list(inputs)) if len(inputs) else np.zeros((1, 4), dtype=np.uint64)
        _ck(ctx, lib().zkb_merkle_build(ctx.handle, _p(arr), len(inputs), depth, C.byref(h)))
        self._h = h.value

    def free(self) -> None:
        if self._h:
            lib().zkb_merkle_free(self.ctx.handle, self._h)
            self._h = 0

    def nodes(self, level: int, first: int = 0, count: int = None) -> List[int]:
        count = (1 << (self.depth - level)) - first if count is None else count
        out = np.zeros((max(count, 1), 4), dtype=np.uint64)
        _ck(self.ctx, lib().zkb_merkle_nodes(self.ctx.handle, self._h, level, first, count, _p(out)))
        return self.ctx.unmont(out[:count])

    @property
    def leaves(self) -> List[int]:
        return self.nodes(0)

    def get_root_hash(self) -> int:  # :134-136
        out = np.zeros((1, 4), dtype=np.uint64)
        _ck(self.ctx, lib().zkb_merkle_root(self.ctx.handle, self._h, _p(out)))
        return self.ctx.unmont(out)[0]

    def update_leaf(self, leaf_id: int, data: int, is_hash: bool) -> None:  # :86-107
        if leaf_id >= (1 << self.depth):
            raise ValueError("Invalid leaf ID")
        _ck(self.ctx, lib().zkb_merkle_update_leaf(self.ctx.handle, self._h, leaf_id, _p(self.ctx.mont([data])), 1 if is_hash else 0))

    def create_proof(self, data_to_prove: int, leaf_id: int) -> MerkleProof:  # :138-183
        if leaf_id >= (1 << self.depth):
            raise ValueError("Invalid leaf ID")
        sib = np.zeros((self.depth, 4), dtype=np.uint64)
        sides = np.zeros(self.depth, dtype=np.uint8)
        try:
            _ck(self.ctx, lib().zkb_merkle_create_proof(self.ctx.handle, self._h, _p(self.ctx.mont([data_to_prove])), leaf_id, _p(sib), sides.ctypes.data))
        except ZkbError as e:
            if "Data does not match the leaf hash" in str(e):
                raise ValueError("Data does not match the leaf hash") from None
            raise
        return MerkleProof(data_to_prove, list(zip(self.ctx.unmont(sib), [int(s) for s in sides])))

    def verify(self, proof: MerkleProof) -> bool:  # :185-199
        n = len(proof.proof)
        sib = self.ctx.mont([h for h, _ in proof.proof]) if n else np.zeros((1, 4), dtype=np.uint64)
        sides = np.array([s for _, s in proof.proof] or [0], dtype=np.uint8)
        ok = C.c_int32()
        _ck(self.ctx, lib().zkb_merkle_verify(self.ctx.handle, self._h, _p(self.ctx.mont([proof.data])), _p(sib), sides.ctypes.data, n, C.byref(ok)))
        return bool(ok.value)
