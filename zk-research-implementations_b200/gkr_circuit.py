"""Device-backed mirror of `gkr::gkr_circuit` (gkr_circuit.rs:4-143)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from . import engine as E
from .engine import Context, _ck, lib
from .multilinear_polynomial import MultilinearPoly, Operation


@dataclass
class Gate:  # :4-23
    l_input: int
    r_input: int
    output: int
    op: Operation


class Layer:  # :25-104
    def __init__(self, gates: List[Gate], ctx: Context = None):
        self.gates = gates
        self.ctx = ctx

    def get_layer_poly(self) -> List[int]:  # :35-37
        return [g.output for g in self.gates]

    def get_add_mul_i(self, op: Operation) -> MultilinearPoly:  # :39-52 (dense; small layers only)
        ops = np.array([int(g.op) for g in self.gates], dtype=np.uint8)
        h = C.c_uint64()
        _ck(self.ctx, lib().zkb_layer_add_mul_i(self.ctx.handle, ops.ctypes.data_as(C.POINTER(C.c_uint8)), len(ops), int(op), C.byref(h)))
        return MultilinearPoly(self.ctx, _handle=h.value)


class Circuit:
    """Circuit::new(structure) (:113-125): `structure` lists the layers input side first."""

    def __init__(self, ctx: Context, structure: Sequence[Sequence[Operation]]):
        self.ctx = ctx
        self.layers = [Layer([Gate(0, 0, 0, Operation(op)) for op in ops], ctx) for ops in structure]
        self.gates = np.array([len(ops) for ops in structure], dtype=np.uint32)
        self.ops = np.array([int(op) for ops in structure for op in ops], dtype=np.uint8)
        h = C.c_uint64()
        _ck(ctx, lib().zkb_circuit_create(ctx.handle, len(self.gates), self.gates.ctypes.data_as(C.POINTER(C.c_uint32)),
                                          self.ops.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(h)))
        self._h = h.value

    @property
    def handle(self) -> int:
        return self._h

    def evaluate(self, inputs: Sequence[int]) -> List[List[int]]:  # :127-143
        ctx = self.ctx
        arr = ctx.mont(inputs)
        out = np.zeros((int(self.gates.sum()), 4), dtype=np.uint64)
        _ck(ctx, lib().zkb_circuit_evaluate(ctx.handle, self._h, arr.ctypes.data, len(inputs), out.ctypes.data))
        vals = ctx.unmont(out)
        res, off = [], 0
        cur = [int(x) % ctx.p for x in inputs]
        for layer, g in zip(self.layers, self.gates):
            lay = vals[off: off + int(g)]
            for i, gate in enumerate(layer.gates):
                gate.l_input, gate.r_input, gate.output = cur[2 * i], cur[2 * i + 1], lay[i]
            res.append(lay)
            cur = lay
            off += int(g)
        return res

    def free(self) -> None:
        if self._h:
            lib().zkb_circuit_free(self.ctx.handle, self._h)
            self._h = 0


class WiredCircuit:
    """EXTENSION beyond the reference (zkb200.h "general wiring"): gate g of layer l reads wires in1[g], in2[g] of the
    layer below (the inputs for l = 0); layers are listed input side first, every width a power of two, and the
    output layer may be wide.  `layers` = [(ops, in1, in2), ...]."""

    def __init__(self, ctx: Context, n_inputs: int, layers):
        self.ctx, self.n_inputs = ctx, int(n_inputs)
        self.gates = np.array([len(l[0]) for l in layers], dtype=np.uint32)
        self.ops = np.ascontiguousarray(np.concatenate([np.asarray(l[0], dtype=np.uint8) for l in layers]))
        self.in1 = np.ascontiguousarray(np.concatenate([np.asarray(l[1], dtype=np.uint32) for l in layers]))
        self.in2 = np.ascontiguousarray(np.concatenate([np.asarray(l[2], dtype=np.uint32) for l in layers]))
        h = C.c_uint64()
        u32p = C.POINTER(C.c_uint32)
        _ck(ctx, lib().zkb_circuit_create_wired(ctx.handle, len(self.gates), self.gates.ctypes.data_as(u32p), self.n_inputs,
                                                self.ops.ctypes.data_as(C.POINTER(C.c_uint8)), self.in1.ctypes.data_as(u32p),
                                                self.in2.ctypes.data_as(u32p), C.byref(h)))
        self._h = h.value
        nr = C.c_uint32()
        _ck(ctx, lib().zkb_circuit_total_rounds(ctx.handle, self._h, C.byref(nr)))
        self.total_rounds = nr.value
        widths = [self.n_inputs] + [int(g) for g in self.gates[:-1]]
        self.rounds_per_layer = [2 * (w.bit_length() - 1) for w in widths[::-1]]  # output side first
        self.n_w0 = max(int(self.gates[-1]), 2)

    @classmethod
    def binary_tree(cls, ctx: Context, structure: Sequence[Sequence[Operation]]) -> "WiredCircuit":
        """The reference's fixed wiring (gkr_circuit.rs:76-78) expressed in the general form."""
        return cls(ctx, 2 * len(structure[0]),
                   [([int(o) for o in ops], [2 * g for g in range(len(ops))], [2 * g + 1 for g in range(len(ops))]) for ops in structure])

    @property
    def handle(self) -> int:
        return self._h

    def evaluate(self, inputs: Sequence[int]) -> List[List[int]]:
        ctx = self.ctx
        arr = ctx.mont(inputs)
        out = np.zeros((int(self.gates.sum()), 4), dtype=np.uint64)
        _ck(ctx, lib().zkb_circuit_evaluate(ctx.handle, self._h, arr.ctypes.data, len(inputs), out.ctypes.data))
        vals = ctx.unmont(out)
        res, off = [], 0
        for g in self.gates:
            res.append(vals[off: off + int(g)])
            off += int(g)
        return res

    def free(self) -> None:
        if self._h:
            lib().zkb_circuit_free(self.ctx.handle, self._h)
            self._h = 0
