#!/usr/bin/env python
"""Where the GKR prove time goes: input upload + circuit evaluation vs the layer sumchecks, pageable vs pinned input."""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
z = importlib.import_module("zk-research-implementations_b200")
from oracle import c_oracle as O
log_in = int(sys.argv[1]) if len(sys.argv) > 1 else 21
ctx = z.Context(0, 0, 0)
rng = np.random.default_rng(7)
structure = [[z.Operation(int(b)) for b in rng.integers(0, 2, size=1 << (log_in - 1 - l))] for l in range(log_in)]
circ = z.gkr_circuit.Circuit(ctx, structure)
inputs = z.engine.to_mont(0, O.synth_table(0, 5, 0, log_in))
pinned = torch.from_numpy(inputs.view(np.int64)).pin_memory()
L = z.engine.lib()
def t(f, reps=5):
    f(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    ctx.sync()
    return (time.perf_counter() - t0) * 1e3 / reps
ev_pageable = t(lambda: L.zkb_circuit_evaluate(ctx.handle, circ.handle, inputs.ctypes.data, inputs.shape[0], None))
ev_pinned = t(lambda: L.zkb_circuit_evaluate(ctx.handle, circ.handle, C.c_void_p(pinned.data_ptr()), inputs.shape[0], None))
p1 = z.gkr_protocol.RawGkrProver(circ, inputs)
pr_pageable = t(p1.prove)
class P2(z.gkr_protocol.RawGkrProver):
    pass
p2 = z.gkr_protocol.RawGkrProver(circ, inputs)
p2.inputs = pinned.numpy().view(np.uint64)
pr_pinned = t(p2.prove)
ctx.profile(True); p2.prove(); prof = ctx.profile_read(); ctx.profile(False)
print({"evaluate_ms_pageable": ev_pageable, "evaluate_ms_pinned": ev_pinned, "prove_ms_pageable": pr_pageable, "prove_ms_pinned": pr_pinned,
       "rounds": p1.total, "launches_per_prove": sum(v[0] for v in prof.values()), "profile": prof})
