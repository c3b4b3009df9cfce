"""The N > 1 path on CPU: two gloo processes play two GPUs.  Each holds its LOW-BIT shard of the tables
(rank j owns entries i with i mod 2 == j), computes its partial round sums (here with the CPU
oracle standing in for the kernels), and the ranks combine them exactly as the engine does: an integer
all-reduce over zero-extended 32-bit limbs followed by libzkb200's host-side carry/reduce
(zkb_fe_reduce_wide), then every rank runs the same transcript.  The sharded run must reproduce the
single-process proof bit for bit, and the shard layout must keep both halves of every bound variable local."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, P, D, ret):
    sys.path.insert(0, ROOT)
    import importlib

    from oracle import c_oracle as O
    from oracle import pyref as R

    z = importlib.import_module("zk-research-implementations_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fid, p = 0, R.BN254_FR
    E = z.engine
    full = [O.arr_to_ints(O.synth_table(fid, 321, t, n)) for t in range(P * D)]
    local = [t[rank::world] for t in full]  # low-bit shard: local index = global >> log2(world)
    # the single-process proof every rank must reproduce
    sp_ref = R.SumPoly([R.ProductPoly(full[q * D:(q + 1) * D], p) for q in range(P)])
    ref = R.gkr_prove(0, sp_ref, R.Transcript(p), "full")
    tr = z.fiat_shamir.Transcript(fid)
    U = z.univariate_polynomial.UnivariatePoly
    coeffs, chals = [], []
    cur = local
    log2w = world.bit_length() - 1
    for rnd in range(n):
        if len(cur[0]) == 1:  # the engine gathers before this point (C2); emulate the gather
            gathered = []
            for t in cur:
                buf = [None] * world
                dist.all_gather_object(buf, t)
                gathered.append([buf[r][i] for i in range(len(t)) for r in range(world)])
            cur = gathered
            world_now = 1
        else:
            world_now = world if len(cur[0]) * world == (1 << (n - rnd)) else 1
        tabs3 = [cur[q * D:(q + 1) * D] for q in range(P)]
        part = R.round_evals_full(tabs3, p)  # this rank's partial s(0..d)
        if world_now > 1:
            wide = np.zeros((len(part), 8), dtype=np.int64)
            for i, v in enumerate(part):
                for k in range(8):
                    wide[i, k] = (v >> (32 * k)) & 0xFFFFFFFF
            t = torch.from_numpy(wide)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)  # C1: exact integer sum of the ranks' limbs
            out = np.zeros((len(part), 4), dtype=np.uint64)
            w64 = np.ascontiguousarray(t.numpy().astype(np.uint64))
            assert E.lib().zkb_fe_reduce_wide(fid, w64.ctypes.data_as(E.u64p), out.ctypes.data_as(E.u64p), len(part)) == 0
            evals = E.limbs_to_ints(out)
        else:
            evals = part
        poly = U.interpolate(list(enumerate(evals)), fid)
        coeffs.append(poly.coefficients)
        tr.append(z.fiat_shamir.fq_vec_to_bytes(poly.coefficients))
        r = tr.get_random_challenge()
        chals.append(r)
        # fold variable 0 = the most significant LOCAL bit: both halves are local on every rank
        h = len(cur[0]) // 2
        cur = [[(t[i] + r * (t[i + h] - t[i])) % p for i in range(h)] for t in cur]
    ok = coeffs == ref.proof_polynomials and chals == ref.random_challenges
    ret[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("n,P,D", [(6, 1, 2), (5, 2, 3)])
def test_two_rank_sharded_sumcheck_matches_single_process(n, P, D):
    world = 2
    port = 29600 + (os.getpid() % 200) + n
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, P, D, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]
