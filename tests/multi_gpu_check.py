#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on a box with N GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Every rank holds its low-bit shard of the global tables (rank j owns entries i with i mod N == j); the
sharded proofs (plain and composed, several shapes, several gather thresholds) must be bit-identical to the
single-GPU proof of the same global table and to the CPU oracle where it is small enough."""
import importlib
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

z = importlib.import_module("zk-research-implementations_b200")
from oracle import c_oracle as O


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fid, p = z.BN254_FR, z.engine.MODULI[z.BN254_FR]
    box = [z.engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx = z.Context(fid, local, z.MODE_FULL)
    ctx.comm_init(rank, world, box[0])
    solo = z.Context(fid, local, z.MODE_FULL)  # no communicator: the single-GPU engine on the whole table
    S, T = z.sum_check_protocol, z.fiat_shamir.Transcript
    ok = True
    for n, P, D, thr in ((6, 1, 2, 1), (10, 2, 3, 3), (14, 1, 2, 12), (16, 2, 2, 8), (18, 1, 3, 12), (15, 1, 2, 0), (17, 2, 3, 0)):  # 0 = automatic
        ctx.set_gather_threshold(thr)
        full = [O.synth_table(fid, 77 + n, t, n) for t in range(P * D)]
        mont = [z.engine.to_mont(fid, f) for f in full]
        # sharded: upload_shard picks this rank's entries out of the full host table
        sh = [z.MultilinearPoly.from_montgomery(ctx, m, shard=True) for m in mont]
        sp = z.SumPoly(ctx, [z.ProductPoly.from_polys(ctx, sh[q * D:(q + 1) * D]) for q in range(P)])
        pr = S.gkr_prove(0, sp, T(fid))
        # single GPU
        so = [z.MultilinearPoly.from_montgomery(solo, m) for m in mont]
        sp1 = z.SumPoly(solo, [z.ProductPoly.from_polys(solo, so[q * D:(q + 1) * D]) for q in range(P)])
        pr1 = S.gkr_prove(0, sp1, T(fid))
        same = ([q.coefficients for q in pr.proof_polynomials] == [q.coefficients for q in pr1.proof_polynomials]
                and pr.random_challenges == pr1.random_challenges and pr.final_values == pr1.final_values)
        ref = O.gkr_sumcheck_prove(O.Transcript(fid), 1, P, D, full)
        same_o = [q.coefficients for q in pr.proof_polynomials] == ref["coeffs"] and pr.final_values == ref["final_vals"]
        # device-generated shards equal uploaded shards
        g = z.MultilinearPoly.generate(ctx, 77 + n, 0, n)
        same_g = np.array_equal(g.montgomery(), sh[0].montgomery())
        # sharded evaluate and plain sumcheck
        rng = random.Random(n)
        rs = [rng.randrange(p) for _ in range(n)]
        ev = sh[0].evaluate(rs) == so[0].evaluate(rs) == O.mle_evaluate(fid, full[0], rs)
        pl, pl1 = S.prove(sh[0], absorb_table=False), S.prove(so[0], absorb_table=False)
        same_p = (pl.claimed_sum, pl.proof_polynomials) == (pl1.claimed_sum, pl1.proof_polynomials)
        # repeat the sharded proof: the shared-memory exchange and the mailbox must give the same bytes every time
        rep_ok = True
        rawp = S.RawGkrProver(sp)
        rawp.prove(T(fid))
        want = rawp.coeffs.copy()
        for _ in range(int(os.environ.get("ZKB_REPEAT", "20"))):
            rawp.coeffs[:] = 0
            rawp.prove(T(fid))
            rep_ok &= bool(np.array_equal(rawp.coeffs, want))
        same &= rep_ok
        line = f"rank {rank}: n={n} P={P} D={D} thr={thr}: sharded==single {same}, ==oracle {same_o}, generate {same_g}, evaluate {ev}, plain {same_p}"
        print(line, flush=True)
        ok &= same and same_o and same_g and ev and same_p
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_PARITY", "OK" if int(t.item()) else "FAILED", flush=True)
    dist.destroy_process_group()
    return 0 if int(t.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
