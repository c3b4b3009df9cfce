#!/usr/bin/env python
"""BASELINE configs[4]: MultilinearPoly partial_evaluate / evaluate sweep over 2^16 .. 2^30 entries, against the HBM
roofline (algorithmic bytes: fold 48 N, evaluate 32 N) and the multiplier roofline (fold: 80 wide
MACs per output entry; evaluate: 80 per input entry).

One GPU:  python tools/mle_sweep.py [lo] [hi]
N GPUs :  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/mle_sweep.py [lo] [hi]
          (lo/hi are GLOBAL variable counts; rank j holds the entries i with i mod N == j, so the fold of variable 0 and
          the first n - log2 N variables of evaluate are local; evaluate ends with one all-gather of N field elements)
Kernel times are CUDA-event times summed per call, maximum over ranks; api_ms is the host wall clock of the C-ABI call.
`sweep()` is also what bench.py's `mle_sweep` leg runs (with the clock sampler of that leg around it); every size is
checked against the oracle-independent identity evaluate(r) == evaluate(partial_evaluate(0, r_0), r_1..) and, up to
2^20 entries, bit for bit against the CPU oracle."""
import ctypes as C
import importlib
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def sweep(z, ctx, sizes, hbm_gbs, dist=None, oracle=None, emit=None):
    """Run the sweep on an existing context (sharded if ctx has a communicator).  Returns one dict per size."""
    import torch

    world = ctx.world
    log2w = world.bit_length() - 1
    L, p = z.engine.lib(), z.engine.MODULI[ctx.field]

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rng = random.Random(3)
    rows = []
    for n in sizes:
        if n - log2w < 10:
            continue
        m = z.MultilinearPoly.generate(ctx, 11, 0, n)  # this rank's shard of the global 2^n table
        rs = [rng.randrange(p) for _ in range(n)]
        arr = ctx.mont(rs)
        out = (C.c_uint64 * 4)()
        h = C.c_uint64()
        reps = 5 if n - log2w <= 26 else 2

        def fold(keep=False):
            z.engine._ck(ctx, L.zkb_mle_partial_evaluate(ctx.handle, m.handle, 0, z.engine._p(arr[:1].copy()), C.byref(h)))
            if not keep:
                L.zkb_mle_free(ctx.handle, h.value)

        def evaluate():
            z.engine._ck(ctx, L.zkb_mle_evaluate(ctx.handle, m.handle, z.engine._p(arr), n, out))

        res = {"n_vars": n, "entries": 1 << n, "n_gpus": world}
        for name, fn, alg in (("partial_evaluate", fold, 48.0 * (1 << n)), ("evaluate", evaluate, 32.0 * (1 << n))):
            fn()
            ctx.sync()
            if dist is not None:
                dist.barrier()
            ctx.profile(True)
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            ctx.sync()
            wall = max_over_ranks((time.perf_counter() - t0) / reps)
            prof = ctx.profile_read()
            ctx.profile(False)
            kms = max_over_ranks(sum(v[1] for v in prof.values()) / reps)
            res[name] = {"kernel_ms": round(kms, 4), "api_ms": round(wall * 1e3, 4), "GBps_kernel": round(alg / (kms * 1e-3) / 1e9, 1),
                         "frac_hbm": round(alg / (kms * 1e-3) / 1e9 / (hbm_gbs * world), 3), "launches": sum(v[0] for v in prof.values()) // reps}
        # parity at this size: the one-pass evaluate equals evaluate-after-one-fold (different kernels), and the oracle where it is fast
        evaluate()
        v_full = [int(out[i]) for i in range(4)]
        fold(keep=True)
        out2 = (C.c_uint64 * 4)()
        z.engine._ck(ctx, L.zkb_mle_evaluate(ctx.handle, h.value, z.engine._p(arr[1:].copy()), n - 1, out2))
        L.zkb_mle_free(ctx.handle, h.value)
        res["parity"] = {"evaluate_equals_fold_then_evaluate": v_full == [int(out2[i]) for i in range(4)]}
        if oracle is not None and n <= 20:
            want = oracle.mle_evaluate(ctx.field, oracle.synth_table(ctx.field, 11, 0, n), rs)
            import numpy as np

            got = ctx.unmont(np.array([v_full], dtype=np.uint64))[0]
            res["parity"]["evaluate_equals_oracle"] = got == want
        rows.append(res)
        if emit:
            emit(res)
        m.free()
    return rows


def main():
    z = importlib.import_module("zk-research-implementations_b200")
    lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 30
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    import torch

    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = z.Context(0, local, 0)
    if world > 1:
        box = [z.engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(rank, world, box[0])
    from oracle import c_oracle as O

    sweep(z, ctx, range(lo, hi + 1, 2), peaks, dist, O, emit=(lambda r: print(json.dumps(r), flush=True)) if rank == 0 else None)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
