#!/usr/bin/env python
"""bench.py -- the benchmark of the sumcheck / GKR prover path (contract: see the task statement / DESIGN.md section 10).

Metric (BASELINE.json): sumcheck prover Fr evals/s at 2^k variables.
Headline workload at N=1 = BASELINE configs[1]: composed sumcheck over ONE ProductPoly of 2 multilinear polynomials at
24 variables over BN254 Fr, `full` mode.  At N GPUs the tables are sharded on low index bits
with 2^24 entries per factor per GPU (weak scaling): n = 24 + log2(N) variables.

A "step" is one complete proof (all n rounds: kernels + host Keccak transcript + interpolation).
  value      2^n hypercube points / step time, tables resident in HBM (CUDA events on the engine's stream, max over ranks)
  e2e        the same proof through the C ABI from pinned HOST buffers of ark-ff Montgomery limbs (H2D upload, layout
             transpose, proof, D2H of the round messages inside the timed region)
  roofline   the dominant kernel of the timed region, per launch, against MEASURED_PEAKS.json (HBM) and against the
             multiplier rate measured in this run (zkb_bench_imad)
  parity_checked   the proof bytes of the timed workload compared with the CPU oracle's proof of the same tables
                   (N=1) / with a single-GPU proof of the same global tables (N>1) BEFORE the line is printed
  cpu_baseline     oracle/zk_oracle.c (a C restatement of the reference's loops and schedule; the Rust reference cannot be
                   built here) at the SAME n = 24: single thread (the reference is single-threaded) and all host cores
Further legs in the same JSON line (each with its own clocks record and parity statement):
  target       north-star target shape: ONE product of 3 MLEs at 2^28 variables on one B200
  plain_n20    BASELINE configs[0]: sum_check::prove + verify at n = 20, core vs whole-table Keccak, CPU port beside it
  gkr, gkr_uniform   BASELINE configs[2] (reference-legal tree / as written with general wiring)
  kzg          SURVEY 8f-3: the input-layer commitment (multilinear KZG over BLS12-381 G1) at 2^20 inputs
  ntt, merkle  SURVEY 8f-4: fft/src/fft.rs at 2^24 coefficients, merkle_tree/src/merkle_tree.rs at depth 20
  config3      BASELINE configs[3] (N >= 2): SumPoly of 2 products x 3 factors at 28 variables sharded over the GPUs
  mle_sweep    BASELINE configs[4]: partial_evaluate / evaluate from 2^16 to 2^30 entries
`--impl reference` times the CPU port alone (rank 0 only) on the same workload description.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "zk-research-implementations_b200"
SEED = 0xB2000002
N_VARS_PER_GPU = 24
METRIC = "sumcheck prover Fr evals/sec at 2^k vars"
UNIT = "Fr evals/s"
DTYPE = "u32x8 Montgomery (BN254 Fr)"

# Wide (32x32->64) multiply-accumulates of the round kernels per unit of work:
#   fold with a fixed multiplicand 80, lazy 512-bit product 64, full Montgomery product 136.
#   A fold on the tensor cores (tcfold.cuh: u8 x u8 -> s32 tcgen05.mma, throughput-bound rounds of products of >= 2
#   factors and <= 4 points) leaves the CUDA cores ONE Montgomery row = 8 wide multiplies.
MACS_FOLD, MACS_FOLD_TC, MACS_LAZY, MACS_FULL = 80, 8, 64, 136


def tc_folds(D: int) -> bool:
    return 2 <= D <= 3 and os.environ.get("ZKB200_NO_TC") is None


def macs_per_quad(P: int, D: int) -> int:
    """k_sc_fold_eval / k_sc_tail, large rounds: per quad and product, 2*D folds and D sums (s(1) comes from the claim),
    each sum one product of D factors = (D-2) full products + one lazy product.  With the tensor-core path a fold costs one
    Montgomery row, and for >= 3 factors the last (lazy) product of every sum is a Gram-matrix block on the tensor cores."""
    if tc_folds(D):
        if D == 3:  # Gram accumulation: one raw (64-multiply) and two Montgomery partial products per quad, no lazy product
            return P * (2 * D * MACS_FOLD_TC + MACS_LAZY + 2 * MACS_FULL)
        return P * (2 * D * MACS_FOLD_TC + D * ((D - 2) * MACS_FULL + MACS_LAZY))
    return P * (2 * D * MACS_FOLD + D * ((D - 2) * MACS_FULL + MACS_LAZY))


def macs_per_pair_round0(P: int, D: int) -> int:
    """k_sc_eval: D+1 points of a product of D factors per pair position.  For D = 3 the product of the first two
    factors is a quadratic in t, interpolated from three full products (kernels.cuh RoundAcc::add_product).  Tensor-core
    path: D = 2 is a pure Gram matrix (no CUDA-core multiply), D = 3 keeps four raw 64-multiply products."""
    if D == 1:
        return 0
    if tc_folds(D):
        return 0 if D == 2 else P * 4 * MACS_LAZY  # D = 3: four raw products of the first two factors
    if D == 3:
        return P * (3 * MACS_FULL + 4 * MACS_LAZY)
    return P * (D + 1) * ((D - 2) * MACS_FULL + MACS_LAZY)


def workload_config(world: int, n_vars_per_gpu: int) -> dict:
    """The `config` object, identical for both arms."""
    log2w = world.bit_length() - 1
    return {"workload": "configs[1]: composed sumcheck, ProductPoly of 2 MLEs, 24 variables per GPU, BN254 Fr, full mode",
            "n_vars": n_vars_per_gpu + log2w, "n_vars_per_gpu": n_vars_per_gpu, "products": 1, "factors": 2,
            "l2": "inputs (1 GiB per GPU) larger than the 126 MB L2; no flush",
            "parallelism": f"low-bit table sharding x{world}: round sums combined per round, one all-gather" if world > 1 else "single GPU"}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def host_cores() -> int:
    """Host threads this process may use.  NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the bench runs (B200_PROFILING.md); one process for the whole run,
    each leg reads the samples of its own time window."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout_s: float) -> None:
        """nvidia-smi's start-up (NVML initialisation touches every GPU of the box) must not fall into a timed region."""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.01)

    def window(self, t0: float, t1: float) -> dict:
        """Summary of the samples taken in [t0, t1] (perf_counter); falls back to the nearest sample for short legs."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        while time.perf_counter() - t_end < 0.12 and not any(t >= t1 for t, _ in self.rows):
            time.sleep(0.01)  # let the sample that covers the end of the window arrive
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.06]
        if not rows and self.rows:
            rows = [min(self.rows, key=lambda tr: abs(tr[0] - t1))[1]]
        sm, mx, pw, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}

    def stop(self) -> None:
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()


# --------------------------------------------------------------------------------------------- CPU port (oracle)
def cpu_port_proof(n_vars: int, threads: int, reps: int = 1, P: int = 1, D: int = 2, seed: int = SEED):
    """The oracle's composed-sumcheck prover (the reference's schedule) on the synthetic tables of the bench.
    Returns (evals/s of the best run, seconds, the proof dict of the last run)."""
    from oracle import c_oracle as O

    O.build()
    O.set_threads(threads)
    tabs = [O.synth_table(0, seed, t, n_vars) for t in range(P * D)]
    best, ref = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        ref = O.gkr_sumcheck_prove(O.Transcript(0), 1, P, D, tabs)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return (1 << n_vars) / best, best, ref


def proof_equals_oracle(pr, ref) -> bool:
    return ([q.coefficients for q in pr.proof_polynomials] == ref["coeffs"] and pr.random_challenges == ref["challenges"]
            and pr.final_values == ref["final_vals"])


def proof_bytes(n_vars: int, D: int, T: int) -> int:
    return 32 * (n_vars * (D + 1) + n_vars + T)  # coefficients + challenges + bound values


def run_reference(args):
    """The reference arm: the CPU port alone, rank 0 only, all host cores, always the full n = 24 workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import c_oracle as O

    O.build()
    cores = host_cores()
    n_s = args.n_vars
    O.set_threads(cores)
    tabs = [O.synth_table(0, SEED, t, n_s) for t in range(2)]
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.gkr_sumcheck_prove(O.Transcript(0), 1, 1, 2, tabs)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    tot = sum(times)
    value = (1 << n_s) * len(times) / tot
    v1, dt1, _ = cpu_port_proof(n_s, 1, 1)
    sample = (f"composed sumcheck (1 product x 2 factors, BN254 Fr) at n={n_s} variables per step (one GPU's share of the workload; per-entry "
              f"cost is size-independent), C restatement of the reference's loops and schedule, OpenMP over {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args.gpus, args.n_vars),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "single_thread": {"value": v1, "unit": UNIT, "cores": 1, "seconds": dt1,
                                           "note": "the reference itself is single-threaded (SURVEY F1): this is the like-for-like figure"},
                         "all_cores": {"value": value, "unit": UNIT, "cores": cores}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------- GPU legs
class Bench:
    def __init__(self, args):
        import numpy as np
        import torch

        self.np, self.torch, self.args = np, torch, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            # One process per GPU, each with a host thread that spins on its GPU's mailbox once per round: give every rank
            # its own slice of the host cores so that the ranks (and their helper threads) do not migrate onto each other.
            try:
                cpus = sorted(os.sched_getaffinity(0))
                per = len(cpus) // self.world
                if per >= 1:
                    os.sched_setaffinity(0, set(cpus[self.local * per:(self.local + 1) * per]))
            except (AttributeError, OSError):
                pass
            import torch.distributed as dist

            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.ensure_library()
        self.z = importlib.import_module(PKG)
        self.log2w = self.world.bit_length() - 1
        assert self.world == 1 << self.log2w, "the number of GPUs must be a power of two"
        self.peaks, self.peak_kind = measured_peaks()
        self.sampler = ClockSampler(self.local) if self.rank == 0 else None
        if self.sampler:
            self.sampler.wait_first(3.0)

    def ensure_library(self):
        """The library is a build artefact: build it on a fresh checkout, rebuild it if it is stale against the tree."""
        if self.local == 0:
            so = os.path.join(ROOT, PKG, "libzkb200.so")
            stale = not os.path.exists(so)
            if not stale:  # built from other sources than the tree's?  (the Makefile bakes csrc/src_hash.py's hash into the library)
                import importlib.util

                spec = importlib.util.spec_from_file_location("_zkb_src_hash", os.path.join(ROOT, PKG, "csrc", "src_hash.py"))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                with open(so, "rb") as f:
                    stale = ("src:" + mod.src_hash()).encode() not in f.read()
            if stale:
                import __graft_entry__

                __graft_entry__.build()
        if self.dist is not None:
            self.dist.barrier()

    # -- helpers
    def new_ctx(self, mode=None, comm=True):
        z = self.z
        ctx = z.Context(z.BN254_FR, self.local, z.MODE_FULL if mode is None else mode)
        if self.world > 1 and comm:
            box = [z.engine.comm_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(box, src=0)
            ctx.comm_init(self.rank, self.world, box[0])
        if self.args.tail_log2 is not None:
            ctx.set_tail_threshold(self.args.tail_log2)
        if self.args.small_bytes is not None:
            ctx.set_small_threshold(self.args.small_bytes)
        return ctx

    def barrier(self, ctx):
        ctx.sync()
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()

    def reduce(self, x: float, op: str) -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return float(t.item())

    class _HostPriority:
        """The persistent kernels wait for this thread once per round (it runs the Keccak transcript): a proof of 1 ms has 24
        such hand-offs, and one pre-emption of the thread stalls the GPU for a scheduler quantum.  For the timed region the
        thread asks for SCHED_FIFO (the box runs the bench as root); where that is refused nothing changes.  Recorded in
        the JSON line as `host_sched`."""
        policy = "default"

        def __enter__(self):
            try:
                if len(os.sched_getaffinity(0)) < 2:  # the other threads of this rank (NCCL proxy, sampler) need a core too
                    raise OSError("single core")
                self.old = (os.sched_getscheduler(0), os.sched_getparam(0))
                os.sched_setscheduler(0, os.SCHED_FIFO, os.sched_param(10))
                Bench._HostPriority.policy = "SCHED_FIFO during the timed proofs"
            except Exception:
                self.old = None
            return self

        def __exit__(self, *exc):
            if self.old is not None:
                try:
                    os.sched_setscheduler(0, self.old[0], self.old[1])
                except Exception:
                    pass
            return False

    def timed_proofs(self, ctx, raw, warmup: int, steps: int):
        """`steps` proofs bracketed by barrier + synchronize, CUDA events on the engine's stream, max over ranks.
        Returns (ms per step, kernel launches in the timed region, per-kernel profile, clocks of the window)."""
        torch, z = self.torch, self.z
        T = z.fiat_shamir.Transcript
        stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", self.local))
        self.barrier(ctx)
        for _ in range(warmup):
            raw.prove(T(z.BN254_FR))
        self.barrier(ctx)
        ctx.profile(True)
        l0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with Bench._HostPriority():
            t0 = time.perf_counter()
            e0.record(stream)
            for _ in range(steps):
                raw.prove(T(z.BN254_FR))
            e1.record(stream)
            self.barrier(ctx)
            t1 = time.perf_counter()
        ms = self.reduce(e0.elapsed_time(e1), "max") / steps
        launches = ctx.launch_count - l0
        prof = ctx.profile_read()
        ctx.profile(False)
        clocks = self.sampler.window(t0, t1) if self.sampler else None
        return ms, launches, prof, clocks

    def imad_rate(self, ctx) -> float:
        """IMAD.WIDE.U32.X (the multiplier's carry-chained wide multiply-accumulate) per second, measured now."""
        import ctypes as C

        v = C.c_double()
        self.z.engine._ck(ctx, self.z.engine.lib().zkb_bench_imad(ctx.handle, 3, 4096, C.byref(v)))
        return v.value

    def properties_hold(self, ctx, tabs, P, D, pr) -> dict:
        """Size-independent parity properties of a composed proof (used where the oracle cannot replay the size):
        the host verifier accepts the transcript, the final claim equals sum of products of the bound values, and every
        bound value equals the MLE of its table at the challenge point (independent k_multifold kernel)."""
        z = self.z
        p = z.engine.MODULI[z.BN254_FR]
        S = z.sum_check_protocol
        polys = pr.proof_polynomials
        claim = (polys[0].evaluate(0) + polys[0].evaluate(1)) % p
        v = S.gkr_verify(polys, claim, z.fiat_shamir.Transcript(z.BN254_FR))
        fin = pr.final_values
        tot = 0
        for q in range(P):
            m = 1
            for f in range(D):
                m = m * fin[q * D + f] % p
            tot = (tot + m) % p
        evals_ok = all(tabs[t].evaluate(pr.random_challenges) == fin[t] for t in range(P * D))
        return {"verifier_accepts": bool(v.verified), "challenges_replayed": v.random_challenges == pr.random_challenges,
                "final_claim_is_product_of_bound_values": v.final_claimed_sum == tot, "bound_values_equal_mle_evaluations": bool(evals_ok)}

    # -- legs
    def headline(self):
        z, np, torch, args = self.z, self.np, self.torch, self.args
        S, T = z.sum_check_protocol, z.fiat_shamir.Transcript
        warmup = args.warmup if args.quick else max(args.warmup, 3)
        n = args.n_vars + self.log2w
        P_, D_ = args.products, args.factors
        ctx = self.ctx = self.new_ctx()
        tabs = [z.MultilinearPoly.generate(ctx, SEED, t, n) for t in range(P_ * D_)]
        sp = z.SumPoly(ctx, [z.ProductPoly.from_polys(ctx, tabs[q * D_:(q + 1) * D_]) for q in range(P_)])
        raw = S.RawGkrProver(sp)  # the bare C-ABI call; outputs are Montgomery limbs as a Rust caller receives them
        ms_step, launches, prof, clocks = self.timed_proofs(ctx, raw, warmup, args.steps)
        value = (1 << n) / (ms_step * 1e-3)
        if args.quick:
            if self.rank == 0:
                print(json.dumps({"quick": True, "ms_per_step": ms_step, "value": value, "gpu_launches": launches, "clocks": clocks,
                                  "profile": {k: v for k, v in prof.items()},
                                  "gkr": gkr_leg(z, args) if self.world == 1 and args.gkr_log_inputs > 0 else None}))
            return None
        proof = raw.proof()

        # ------------------------------------------------------------ parity of the timed workload, before anything is printed
        parity = {"n_vars": n, "bytes": proof_bytes(n, D_, P_ * D_)}
        if self.world > 1:
            # the same GLOBAL tables on ONE GPU (rank 0, a context without communicator); every rank compares its
            # sharded proof with it
            box = [None]
            if self.rank == 0:
                solo = self.new_ctx(comm=False)
                stabs = [z.MultilinearPoly.generate(solo, SEED, t, n) for t in range(P_ * D_)]
                ssp = z.SumPoly(solo, [z.ProductPoly.from_polys(solo, stabs[q * D_:(q + 1) * D_]) for q in range(P_)])
                sraw = S.RawGkrProver(ssp)
                sraw.prove(T(z.BN254_FR))
                box[0] = (sraw.coeffs.tobytes(), sraw.lens.tobytes(), sraw.chals.tobytes(), sraw.fin.tobytes())
                ssp.free()
                for t in stabs:
                    t.free()
                solo.close()
            self.dist.broadcast_object_list(box, src=0)
            mine = (raw.coeffs.tobytes(), raw.lens.tobytes(), raw.chals.tobytes(), raw.fin.tobytes())
            ok = self.reduce(1.0 if mine == box[0] else 0.0, "min") == 1.0
            assert ok, "sharded proof differs from the single-GPU proof of the same global tables"
            parity.update({"against": "single-GPU proof of the same global tables (rank 0, no communicator)", "ranks_equal": True})

        # ------------------------------------------------------------ end to end from host buffers (`e2e`)
        n_local = 1 << args.n_vars
        host = []
        for t in tabs[:2]:
            h = torch.empty((n_local, 4), dtype=torch.int64).pin_memory()
            h.numpy().view(np.uint64)[:] = t.montgomery()  # this rank's shard as a Rust Vec<F> would hold it
            host.append(h)

        def e2e_step():
            ta = z.MultilinearPoly.from_host_pointer(ctx, host[0].data_ptr(), n_local)
            tb = z.MultilinearPoly.from_host_pointer(ctx, host[1].data_ptr(), n_local)
            s2 = z.SumPoly(ctx, [z.ProductPoly.from_polys(ctx, [ta, tb])])
            r2 = S.RawGkrProver(s2)
            r2.prove(T(z.BN254_FR))
            s2.free()
            ta.free()
            tb.free()
            return r2

        e2e_steps = min(args.steps, 5)
        pr2 = e2e_step()
        assert np.array_equal(pr2.coeffs, raw.coeffs) and np.array_equal(pr2.fin, raw.fin), "host-buffer path and resident path disagree"
        stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", self.local))
        self.barrier(ctx)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(e2e_steps):
            e2e_step()
        f1.record(stream)
        self.barrier(ctx)
        e2e_ms = self.reduce(f0.elapsed_time(f1), "max") / e2e_steps
        h2d = self.reduce(2.0 * n_local * 32, "sum")
        d2h = float(n * 3 * 32 + 2 * 32)  # per round 3 evaluations, plus the two bound values (every rank reads the same)
        del host

        if self.rank != 0:
            return None
        peaks = self.peaks
        imad = self.imad_rate(ctx)
        # dominant kernel of the timed region = the one with the largest summed CUDA-event duration
        round_kernels = {k: v for k, v in prof.items() if k.startswith("k_sc_")}
        dom = max(round_kernels, key=lambda k: round_kernels[k][1])
        dl, dms, dby = round_kernels[dom]
        achieved = dby / (dms * 1e-3) / 1e9 if dms > 0 else 0.0
        # The largest single round, timed alone through the step API (one k_sc_fold_eval launch: the same
        # round_pass code the persistent kernel runs, without the host mailbox waits inside the launch).
        L_ = z.engine.lib()
        big = {}
        if self.world == 1:
            sph = sp.handle()
            ev_buf = np.zeros((4, 4), dtype=np.uint64)
            r_buf = ctx.mont([0x1234567890ABCDEF1234567890ABCDEF])
            tot_ms = tot_by = 0.0
            for _ in range(5):
                L_.zkb_sumpoly_reset(ctx.handle, sph)
                L_.zkb_sc_round_evals(ctx.handle, sph, z.engine._p(ev_buf))
                ctx.profile(True)
                L_.zkb_sc_bind_and_next(ctx.handle, sph, z.engine._p(r_buf), z.engine._p(ev_buf))
                pr_ = ctx.profile_read()
                ctx.profile(False)
                tot_ms += pr_["k_sc_fold_eval"][1]
                tot_by += pr_["k_sc_fold_eval"][2]
            L_.zkb_sumpoly_reset(ctx.handle, sph)
            quads = (1 << args.n_vars) / 4
            mac_floor_us = quads * macs_per_quad(P_, D_) / imad * 1e6
            big = {"kernel": "k_sc_fold_eval (round 1: 2^%d -> 2^%d entries per table)" % (args.n_vars, args.n_vars - 1),
                   "us": 1e3 * tot_ms / 5, "alg_bytes": tot_by / 5, "achieved": tot_by / (tot_ms * 1e-3) / 1e9,
                   "frac": tot_by / (tot_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                   "multiplier_floor_us": mac_floor_us, "frac_of_multiplier_floor": mac_floor_us / (1e3 * tot_ms / 5)}
        traffic, traffic_note = None, "no ncu --set full capture of this build is committed"
        tp = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                traffic_note = ("NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of %s from the committed capture %s (algorithmic "
                                "bytes of that launch: %d)" % (tj["kernel"], tj["source"], tj["alg_bytes"]))
            except Exception:
                pass
        roof = {
            "bound": "hbm", "kernel": dom + f"<BN254Fr,PROD,D={D_},NPTS={D_ + 1}>", "achieved": achieved, "peak": peaks["hbm_gbs"],
            "peak_source": f"MEASURED_PEAKS.json ({self.peak_kind})", "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "traffic": traffic, "traffic_note": traffic_note,
            "launches": dl, "kernel_ms_per_step": dms / args.steps, "alg_bytes_per_step": dby / args.steps,
            "share_of_step": (dms / args.steps) / ms_step,
            "note": "the persistent kernel's duration includes its per-round waits for the host transcript (mailbox); "
                    "largest_round isolates one round of the same code",
            "largest_round": big,
            "fold_engine": ("tcgen05.mma kind::i8 (u8 x u8 -> s32 in TMEM, operands by TMA bulk copy; csrc/tcfold.cuh): bit-identical to the "
                            "CUDA-core fold" if tc_folds(D_) else "CUDA cores (fixed-multiplicand Montgomery product)"),
            "tensor_cores": dict(zip(("enabled", "persistent_kernel"), ctx.tensor_cores())),
            "imad": {"wide_macs_per_quad": macs_per_quad(P_, D_), "bytes_per_quad": 96 * P_ * D_ * 2,
                     "measured_imad_wide_x_per_s": imad, "measured": "zkb_bench_imad in this run"},
            "kernels": {k: {"launches": v[0], "ms_per_step": v[1] / args.steps, "GBps": v[2] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0}
                        for k, v in prof.items()},
        }
        cfg = workload_config(self.world, args.n_vars)
        cfg.update({"products": P_, "factors": D_})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": self.world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic", "config": cfg,
            "e2e": {"value": (1 << n) / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "steps": e2e_steps},
            "gpu_launches": launches, "roofline": roof, "clocks": clocks, "build": z.engine.build_info(), "host_sched": Bench._HostPriority.policy,
            "table_entries_per_s": P_ * D_ * value,
        }
        # CPU port at the SAME size: its proof is the parity check of the timed workload (N = 1)
        if self.world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            v_all, dt_all, ref = cpu_port_proof(args.n_vars, cores, 2, P_, D_)
            assert proof_equals_oracle(proof, ref), "GPU proof differs from the CPU oracle's proof of the same tables"
            parity.update({"against": "oracle/zk_oracle.c proof of the same synthetic tables (coefficients, challenges, bound values)", "equal": True})
            v_1, dt_1, ref1 = cpu_port_proof(args.n_vars, 1, 1, P_, D_)
            assert ref1["coeffs"] == ref["coeffs"]
            line["cpu_baseline"] = {
                "value": v_all, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"the same composed sumcheck at n={args.n_vars} variables (the full workload; best of 2, {dt_all:.2f} s each), oracle/zk_oracle.c "
                          f"restating the reference's loops and schedule, OpenMP over {cores} threads",
                "single_thread": {"value": v_1, "unit": UNIT, "cores": 1, "seconds": dt_1,
                                  "note": "the reference itself is single-threaded (SURVEY F1): this is the like-for-like figure"},
                "all_cores": {"value": v_all, "unit": UNIT, "cores": cores, "seconds": dt_all}}
        line["parity_checked"] = parity
        sp.free()
        for t in tabs:
            t.free()
        return line

    def target_leg(self):
        """North-star target: ONE product of 3 MLEs at 2^28 variables on one B200, tables resident."""
        z, args = self.z, self.args
        S, T = z.sum_check_protocol, z.fiat_shamir.Transcript
        n, P_, D_ = args.target_n_vars, 1, 3
        ctx = self.ctx
        tabs = [z.MultilinearPoly.generate(ctx, SEED + 3, t, n) for t in range(D_)]
        sp = z.SumPoly(ctx, [z.ProductPoly.from_polys(ctx, tabs)])
        raw = S.RawGkrProver(sp)
        steps = max(3, min(args.steps, 5))
        ms, launches, prof, clocks = self.timed_proofs(ctx, raw, 2, steps)
        pr = raw.proof()
        props = self.properties_hold(ctx, tabs, P_, D_, pr)
        assert all(props.values()), f"target leg: parity properties failed: {props}"
        sp.free()
        for t in tabs:
            t.free()
        # the same shape at a size the oracle replays in a fraction of a second: bit-exact comparison
        small = {}
        if not args.no_cpu_baseline:
            ns = 20
            stabs = [z.MultilinearPoly.generate(ctx, SEED + 3, t, ns) for t in range(D_)]
            ssp = z.SumPoly(ctx, [z.ProductPoly.from_polys(ctx, stabs)])
            spr = S.gkr_prove(0, ssp, T(z.BN254_FR))
            _, _, ref = cpu_port_proof(ns, host_cores(), 1, P_, D_, SEED + 3)
            assert proof_equals_oracle(spr, ref), "target shape at n=20: GPU proof differs from the oracle"
            small = {"n_vars": ns, "equal_to_oracle": True, "bytes": proof_bytes(ns, D_, D_)}
            ssp.free()
            for t in stabs:
                t.free()
        imad = self.imad_rate(ctx)
        N = float(1 << n)
        alg_bytes = 128.0 * D_ * N  # composed sumcheck prove = 128*T*N
        macs = (N / 2) * macs_per_pair_round0(P_, D_) + (N / 2) * macs_per_quad(P_, D_)  # round 0 + sum over rounds of N_k/4 quads
        hbm_floor = alg_bytes / (self.peaks["hbm_gbs"] * 1e9) * 1e3
        mac_floor = macs / imad * 1e3
        kern = {k: {"launches": v[0] // steps, "ms_per_proof": v[1] / steps, "GBps": v[2] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0} for k, v in prof.items()}
        return {"workload": f"north-star target: composed sumcheck, ONE ProductPoly of 3 MLEs, {n} variables, BN254 Fr, full mode, one B200, tables resident "
                            f"({D_} x {32 * (1 << n) >> 30} GiB + {D_} x {16 * (1 << n) >> 30} GiB work)",
                "ms_per_proof": ms, "value": N / (ms * 1e-3), "unit": UNIT, "steps": steps, "warmup": 2, "gpu_launches": launches,
                "roofline": {"alg_bytes": alg_bytes, "achieved_GBps": alg_bytes / (ms * 1e-3) / 1e9, "hbm_peak_GBps": self.peaks["hbm_gbs"],
                             "frac_hbm": hbm_floor / ms, "hbm_floor_ms": hbm_floor,
                             "wide_macs": macs, "measured_imad_wide_x_per_s": imad, "multiplier_floor_ms": mac_floor,
                             "frac_multiplier": mac_floor / ms, "bound": "multiplier" if mac_floor > hbm_floor else "hbm",
                             "frac": max(mac_floor, hbm_floor) / ms,
                             "note": "frac = slower of the two floors / measured time (north_star: judged against the slower of HBM bandwidth and "
                                     "32-bit integer multiply throughput); both floors are reproducible from the fields of this object"},
                "kernels": kern, "clocks": clocks,
                "parity": {"properties_at_full_size": props, "bit_exact_vs_oracle_same_shape": small}}

    def plain_leg(self):
        """BASELINE configs[0]: sum_check::prove + verify on a random 20-variable MLE over BN254 Fr (the reference
        hashes the whole 32 MiB table into the transcript; `core` is the same call with a seeded transcript)."""
        z, np, args = self.z, self.np, self.args
        from oracle import c_oracle as O

        S = z.sum_check_protocol
        n = 20
        ctx = self.ctx
        m = z.MultilinearPoly.generate(ctx, SEED + 4, 0, n)

        def t_ms(fn, reps):
            fn()
            ctx.sync()
            t0 = time.perf_counter()
            for _ in range(reps):
                r = fn()
            ctx.sync()
            return (time.perf_counter() - t0) * 1e3 / reps, r

        ta = time.perf_counter()
        prove_ms, pr = t_ms(lambda: S.prove(m, absorb_table=True), 3)
        verify_ms, ok = t_ms(lambda: S.verify(m, pr, absorb_table=True), 3)
        core_ms, prc = t_ms(lambda: S.prove(m, absorb_table=False), 10)
        vcore_ms, okc = t_ms(lambda: S.verify(m, prc, absorb_table=False), 10)
        tb = time.perf_counter()
        keccak_ms, _ = t_ms(lambda: z.engine.keccak256(b"\0" * (32 << n)), 2)
        clocks = self.sampler.window(ta, tb) if self.sampler else None
        out = {"workload": "configs[0]: sum_check::prove + verify, one random 20-variable MLE, BN254 Fr (host wall clock around the C-ABI calls, table resident)",
               "prove_ms": prove_ms, "verify_ms": verify_ms, "verify_accepts": bool(ok),
               "core": {"prove_ms": core_ms, "verify_ms": vcore_ms, "verify_accepts": bool(okc), "evals_per_s": (1 << n) / (core_ms * 1e-3),
                        "note": "seeded transcript: the 32 MiB of table bytes are not absorbed (SURVEY F9)"},
               "host_keccak_of_table_ms": keccak_ms, "clocks": clocks}
        if not args.no_cpu_baseline:
            tab = O.synth_table(0, SEED + 4, 0, n)
            cpu = {}
            for label, thr in (("single_thread", 1), ("all_cores", host_cores())):
                O.set_threads(thr)
                t0 = time.perf_counter()
                claimed, msgs, chals = O.sumcheck_prove(0, tab)
                t1 = time.perf_counter()
                okv = O.sumcheck_verify(0, tab, claimed, msgs)
                t2 = time.perf_counter()
                cpu[label] = {"prove_ms": (t1 - t0) * 1e3, "verify_ms": (t2 - t1) * 1e3, "cores": thr, "verify_accepts": bool(okv)}
            assert (pr.claimed_sum, pr.proof_polynomials) == (claimed, msgs), "plain sumcheck at n=20: GPU proof differs from the oracle"
            cpu["kind"] = "port"
            out["cpu_baseline"] = cpu
            out["parity_checked"] = {"n_vars": n, "equal_to_oracle": True, "bytes": 32 * (1 + 2 * n)}
        m.free()
        return out

    def config3_leg(self):
        """BASELINE configs[3]: SumPoly of 2 ProductPolys x 3 factors at 28 variables, sharded over the N GPUs."""
        z, args = self.z, self.args
        S, T = z.sum_check_protocol, z.fiat_shamir.Transcript
        n, P_, D_ = args.config3_n_vars, 2, 3
        ctx = self.ctx
        tabs = [z.MultilinearPoly.generate(ctx, SEED + 5, t, n) for t in range(P_ * D_)]
        sp = z.SumPoly(ctx, [z.ProductPoly.from_polys(ctx, tabs[q * D_:(q + 1) * D_]) for q in range(P_)])
        raw = S.RawGkrProver(sp)
        steps = max(3, min(args.steps, 5))
        ms, launches, prof, clocks = self.timed_proofs(ctx, raw, 2, steps)
        pr = raw.proof()
        props = self.properties_hold(ctx, tabs, P_, D_, pr)
        ok = self.reduce(1.0 if all(props.values()) else 0.0, "min") == 1.0
        assert ok, f"config3: parity properties failed: {props}"
        sp.free()
        for t in tabs:
            t.free()
        # bit-exact: the same shape at n = 20 sharded over the same ranks vs the oracle
        small = {}
        if not args.no_cpu_baseline:
            ns = 20
            stabs = [z.MultilinearPoly.generate(ctx, SEED + 5, t, ns) for t in range(P_ * D_)]
            ssp = z.SumPoly(ctx, [z.ProductPoly.from_polys(ctx, stabs[q * D_:(q + 1) * D_]) for q in range(P_)])
            spr = S.gkr_prove(0, ssp, T(z.BN254_FR))
            eq = 1.0
            if self.rank == 0:
                _, _, ref = cpu_port_proof(ns, host_cores(), 1, P_, D_, SEED + 5)
                eq = 1.0 if proof_equals_oracle(spr, ref) else 0.0
            assert self.reduce(eq, "min") == 1.0, "config3 shape at n=20: sharded GPU proof differs from the oracle"
            small = {"n_vars": ns, "equal_to_oracle": True, "bytes": proof_bytes(ns, D_, P_ * D_)}
            ssp.free()
            for t in stabs:
                t.free()
        if self.rank != 0:
            return None
        imad = self.imad_rate(ctx)
        N = float(1 << n)
        alg_bytes = 128.0 * P_ * D_ * N
        macs = (N / 2) * macs_per_pair_round0(P_, D_) + (N / 2) * macs_per_quad(P_, D_)
        hbm_floor = alg_bytes / (self.peaks["hbm_gbs"] * 1e9 * self.world) * 1e3
        mac_floor = macs / (imad * self.world) * 1e3
        return {"workload": f"configs[3]: SumPoly of 2 ProductPolys x 3 MLEs, {n} variables, BN254 Fr, full mode, low-bit sharding over {self.world} GPUs",
                "n_gpus": self.world, "ms_per_proof": ms, "value": N / (ms * 1e-3), "unit": UNIT, "steps": steps, "warmup": 2, "gpu_launches": launches,
                "roofline": {"alg_bytes": alg_bytes, "achieved_GBps": alg_bytes / (ms * 1e-3) / 1e9, "hbm_peak_GBps": self.peaks["hbm_gbs"] * self.world,
                             "hbm_floor_ms": hbm_floor, "wide_macs": macs, "measured_imad_wide_x_per_s_per_gpu": imad, "multiplier_floor_ms": mac_floor,
                             "bound": "multiplier" if mac_floor > hbm_floor else "hbm", "frac": max(mac_floor, hbm_floor) / ms},
                "kernels": {k: {"launches": v[0] // steps, "ms_per_proof": v[1] / steps} for k, v in prof.items()}, "clocks": clocks,
                "parity": {"properties_at_full_size": props, "bit_exact_vs_oracle_same_shape_sharded": small}}

    def mle_sweep_leg(self):
        """BASELINE configs[4]: partial_evaluate / evaluate sweep (tools/mle_sweep.py), with the clocks of its window."""
        from tools import mle_sweep
        from oracle import c_oracle as O

        hi = self.args.mle_sweep_hi
        sizes = [s for s in (16, 20, 24, 28, 30, 32) if s <= hi + self.log2w and s - self.log2w <= 30]
        t0 = time.perf_counter()
        rows = mle_sweep.sweep(self.z, self.ctx, sizes, self.peaks["hbm_gbs"], self.dist, None if self.args.no_cpu_baseline else O)
        t1 = time.perf_counter()
        if self.rank != 0:
            return None
        for r in rows:
            assert all(r["parity"].values()), f"mle sweep parity failed at n={r['n_vars']}: {r['parity']}"
        return {"workload": "configs[4]: MultilinearPoly::partial_evaluate(0, r) and ::evaluate(r) on a random table, BN254 Fr; algorithmic bytes 48 N / 32 N; "
                            "frac_hbm against N_gpus x MEASURED_PEAKS hbm_gbs; evaluate is multiplier-bound (7 fixed folds per 8 entries, DESIGN.md section 6)",
                "rows": rows, "clocks": self.sampler.window(t0, t1) if self.sampler else None}


def gkr_leg(z, args):
    """BASELINE configs[2] in its reference-legal form (SURVEY F8 i): binary-tree add/mul circuit, 2^21 inputs,
    21 layers (widest 2^20 gates), BN254 Fr.  Times zkb_gkr_prove (circuit evaluation + all layer sumchecks,
    inputs uploaded from host memory every call) and checks the proof with zkb_gkr_verify."""
    import numpy as np
    import torch
    from oracle import c_oracle as O

    log_in = args.gkr_log_inputs
    L = log_in
    rng = np.random.default_rng(7)
    ctx = z.Context(z.BN254_FR, 0, z.MODE_COMPAT)
    structure = [[z.Operation(int(b)) for b in rng.integers(0, 2, size=1 << (L - 1 - l))] for l in range(L)]
    circ = z.gkr_circuit.Circuit(ctx, structure)
    inputs = z.engine.to_mont(z.BN254_FR, O.synth_table(0, SEED + 1, 0, log_in))
    reps = max(3, min(args.steps, 10))

    def run(buf):
        prover = z.gkr_protocol.RawGkrProver(circ, buf)
        for _ in range(2):
            prover.prove()
        with Bench._HostPriority():
            t0 = time.perf_counter()
            for _ in range(reps):
                prover.prove()
            ms_ = (time.perf_counter() - t0) * 1e3 / reps
        return prover, ms_

    _, ms_pageable = run(inputs)
    pinned = torch.from_numpy(inputs.view(np.int64)).pin_memory()  # keep the tensor alive: the prover reads its memory
    prover, ms = run(pinned.numpy().view(np.uint64))
    ctx.profile(True)
    prover.prove()
    prof = ctx.profile_read()
    ctx.profile(False)
    ok = prover.verify()
    ctx.close()
    cpu = None
    if not args.no_cpu_baseline:
        # CPU figure beside it: the oracle's O(G) two-phase prover on the SAME circuit and inputs.  The reference's own
        # dense construction (2^(3g+2)-entry add_i/mul_i tables, gkr_circuit.rs:39-65) cannot run at this size at all.
        O.build()
        cores = host_cores()
        O.set_threads(cores)
        flat = np.concatenate([np.array([int(o) for o in layer], dtype=np.uint8) for layer in structure])
        t0 = time.perf_counter()
        ref = O.gkr_prove(0, [len(layer) for layer in structure], flat, O.synth_table(0, SEED + 1, 0, log_in))
        cpu_ms = (time.perf_counter() - t0) * 1e3
        same = list(ref["final_openings"]) == O.arr_to_ints(z.engine.from_mont(z.BN254_FR, prover.fin))
        assert same, "GKR tree leg: the device's final openings differ from the oracle's"
        cpu = {"prove_ms": cpu_ms, "cores": cores, "kind": "port",
               "note": "oracle/zk_oracle.c two-phase (sparse) GKR prover, one run, OpenMP over all host threads (the reference itself "
                       "is single-threaded, and its dense construction is infeasible at this size)",
               "same_final_openings_as_device": same}
    rounds = int(prover.total)
    kernel_ms = sum(v[1] for v in prof.values())
    return {"prove_ms": ms, "prove_ms_pageable_input": ms_pageable, "verify_accepts": ok, "cpu_baseline": cpu,
            "kernel_ms": {k: round(v[1], 4) for k, v in prof.items()}, "launches": sum(v[0] for v in prof.values()), "rounds": rounds, "layers": L, "inputs": 1 << log_in,
            "latency": {"us_per_round": 1e3 * ms / rounds, "kernel_ms_total": kernel_ms, "kernel_share": kernel_ms / ms,
                        "note": "latency-bound: every round is a host transcript step (Keccak) between two dependent device passes; "
                                "the figure to compare is us_per_round against the ~6 us of two PCIe crossings + host Keccak (DESIGN.md section 7)"},
            "workload": "configs[2] (reference-legal form): binary-tree circuit, 2^%d inputs, %d layers, widest layer 2^%d gates, BN254 Fr; "
                        "KZG input commitment excluded (SURVEY F11); host wall clock around zkb_gkr_prove incl. the upload of the inputs "
                        "from pinned host memory (prove_ms) or pageable memory (prove_ms_pageable_input)" % (log_in, L, log_in - 1)}


def gkr_uniform_leg(z, args):
    """BASELINE configs[2] AS WRITTEN (SURVEY 8d C3 ii): 16 layers of 2^20 add/mul gates each with random wiring over
    2^20 inputs -- not expressible by the reference's (2i, 2i+1) wiring, so this runs the general-wiring extension
    (zkb_gkr_prove_wired; same protocol, log2(outputs) challenges for the output MLE).  The 2^20-element output layer is
    part of the proof: its download and its Keccak absorption on the host are inside prove_ms."""
    import numpy as np
    import torch
    from oracle import c_oracle as O

    lg, L = args.gkr_uniform_log_gates, args.gkr_uniform_layers
    G = 1 << lg
    rng = np.random.default_rng(11)
    ctx = z.Context(z.BN254_FR, 0, z.MODE_COMPAT)
    spec = [(rng.integers(0, 2, size=G, dtype=np.uint8), rng.integers(0, G, size=G, dtype=np.uint32),
             rng.integers(0, G, size=G, dtype=np.uint32)) for _ in range(L)]
    circ = z.gkr_circuit.WiredCircuit(ctx, G, spec)
    inputs = z.engine.to_mont(z.BN254_FR, O.synth_table(0, SEED + 2, 0, lg))
    pinned = torch.from_numpy(inputs.view(np.int64)).pin_memory()
    prover = z.gkr_protocol.RawWiredGkrProver(circ, pinned.numpy().view(np.uint64))
    reps = max(2, min(args.steps, 5))
    prover.prove()
    t0 = time.perf_counter()
    for _ in range(reps):
        prover.prove()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    ctx.profile(True)
    prover.prove()
    prof = ctx.profile_read()
    ctx.profile(False)
    t0 = time.perf_counter()
    ok = prover.verify()
    verify_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    z.engine.keccak256(b"\0" * (32 * G))
    keccak_ms = (time.perf_counter() - t0) * 1e3
    circ.free()
    ctx.close()
    cpu = None
    if not args.no_cpu_baseline:
        # bounded CPU sample: the same first two layers (identical per-layer cost: uniform widths), scaled to L layers
        O.build()
        O.set_threads(host_cores())
        ls = min(2, L)
        t0 = time.perf_counter()
        O.gkr_prove_wired(0, G, spec[:ls], O.synth_table(0, SEED + 2, 0, lg), want_challenges=False)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        cpu = {"prove_ms_sample": cpu_ms, "sample": "%d of the %d layers (uniform layers: per-layer cost is constant)" % (ls, L),
               "prove_ms_scaled": cpu_ms * L / ls, "cores": host_cores(), "kind": "port",
               "note": "oracle/zk_oracle.c two-phase general-wiring prover; sumcheck rounds and folds OpenMP over all host threads, "
                       "the per-gate table accumulation serial"}
    return {"prove_ms": ms, "prove_ms_excl_output_absorb": ms - keccak_ms, "verify_ms": verify_ms, "verify_accepts": ok,
            "host_keccak_of_output_layer_ms": keccak_ms, "cpu_baseline": cpu,
            "kernel_ms": {k: round(v[1], 4) for k, v in prof.items()}, "launches": sum(v[0] for v in prof.values()),
            "rounds": int(prover.total), "layers": L, "gates_per_layer": G,
            "latency": {"us_per_round_excl_output_absorb": 1e3 * (ms - keccak_ms) / int(prover.total), "kernel_ms_total": sum(v[1] for v in prof.values())},
            "workload": "configs[2] as written: %d layers x 2^%d gates, random wiring, 2^%d inputs, BN254 Fr, general-wiring "
                        "extension (the reference's fixed wiring cannot express uniform layers); pinned input; KZG excluded" % (L, lg, lg)}


def kzg_leg(z, args):
    """the input-layer commitment of gkr_protocol::prove (gkr_protocol.rs:92-118) at the configs[2] input size:
    multilinear KZG over BLS12-381 G1 (pcs/src/kzg_pcs/kzg.rs:17-95), taus as an input.  Host wall clock around the C-ABI
    calls, table resident; parity is the test suite's job (tests/test_gpu_kzg.py against oracle/kzg_ref.py, which is pinned to
    the reference's known answers)."""
    import random

    n = args.kzg_log_inputs
    fid = 2
    p = z.engine.MODULI[fid]
    rng = random.Random(11)
    out = {"workload": f"multilinear KZG of a random {n}-variable MLE over BLS12-381 (G1 side): setup from the taus, commit, open, get_proof "
                       "(one quotient commitment per variable)", "n_vars": n}
    with z.Context(fid, 0, 1) as ctx:
        m = z.MultilinearPoly.generate(ctx, SEED + 9, 0, n)
        taus = [rng.randrange(p) for _ in range(n)]
        t0 = time.perf_counter()
        k = z.kzg.KZG(m, taus)
        t1 = time.perf_counter()
        c1 = k.commit(m)  # first call builds the fixed-base tables
        t2 = time.perf_counter()
        c2 = k.commit(m)
        t3 = time.perf_counter()
        r = [rng.randrange(p) for _ in range(n)]
        v = k.open(r, m)
        t4b = time.perf_counter()
        k.get_proof(v, r, m)  # (first call: side streams, lazily loaded kernels)
        t4 = time.perf_counter()
        pr = k.get_proof(v, r, m)
        t5 = time.perf_counter()
        out.update({"setup_ms": 1e3 * (t1 - t0), "commit_first_ms": 1e3 * (t2 - t1), "commit_ms": 1e3 * (t3 - t2), "open_ms": 1e3 * (t4b - t3),
                    "get_proof_ms": 1e3 * (t5 - t4), "msm_points_per_s": (1 << n) / (t3 - t2), "commit_deterministic": c1 == c2,
                    "quotients": len(pr)})
        k.free()
        m.free()
    return out


def fft_merkle_leg(z, args):
    """the NTT of fft/src/fft.rs and the Keccak Merkle tree of merkle_tree/src/merkle_tree.rs on device-resident data
    (host wall clock around the C-ABI calls after a warm-up call; parity against the oracle is tests/test_gpu_fft_merkle.py, here
    the round trip interpolate(evaluate(x)) == x at the timed size).  No CPU figure: the oracle restatement is pure Python."""
    out = {}
    with z.Context(0, 0, 1) as ctx:
        n = args.ntt_log_n
        t = z.MultilinearPoly.generate(ctx, SEED + 12, 0, n)
        z.fft.ntt(t).free()
        ctx.sync()
        t0 = time.perf_counter()
        ev = z.fft.ntt(t)
        ctx.sync()
        t1 = time.perf_counter()
        back = z.fft.ntt(ev, inverse=True)
        ctx.sync()
        t2 = time.perf_counter()
        ok = (back - t).sum_halves() == [0, 0]
        passes = 1 + max(0, -(-(n - 9) // 6))
        out["ntt"] = {"workload": f"fft_evaluate / fft_interpolate of 2^{n} BN254 Fr coefficients, device-resident", "log_n": n,
                      "evaluate_ms": 1e3 * (t1 - t0), "interpolate_ms": 1e3 * (t2 - t1), "passes": passes,
                      "GBps_streamed": passes * 64.0 * (1 << n) / (t1 - t0) / 1e9, "butterflies_per_s": (n << (n - 1)) / (t1 - t0),
                      "roundtrip_is_identity": ok}
        for m in (t, ev, back):
            m.free()
        d = args.merkle_depth
        import numpy as np
        rng = np.random.default_rng(3)
        raw = rng.integers(0, 1 << 62, size=(1 << d, 4), dtype=np.uint64)  # any 248-bit values < p as Montgomery limbs
        raw[:, 3] &= (1 << 56) - 1
        L = z.engine.lib()
        h = z.engine.C.c_uint64()

        def build():
            z.engine._ck(ctx, L.zkb_merkle_build(ctx.handle, z.engine._p(raw), 1 << d, d, z.engine.C.byref(h)))

        build()
        L.zkb_merkle_free(ctx.handle, h.value)
        t0 = time.perf_counter()
        build()
        t1 = time.perf_counter()
        root = np.zeros((1, 4), dtype=np.uint64)
        z.engine._ck(ctx, L.zkb_merkle_root(ctx.handle, h.value, z.engine._p(root)))
        L.zkb_merkle_free(ctx.handle, h.value)
        out["merkle"] = {"workload": f"MerkleTree::new_with_inputs, depth {d}, 2^{d} inputs from host memory (32 MiB H2D included)", "depth": d,
                         "build_ms": 1e3 * (t1 - t0), "hashes": (2 << d) - 1, "hashes_per_s": ((2 << d) - 1) / (t1 - t0)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-vars", type=int, default=N_VARS_PER_GPU, help="variables per GPU shard")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip every CPU-oracle leg (and with it the oracle parity checks)")
    ap.add_argument("--gkr-log-inputs", type=int, default=21, help="0 = skip the GKR tree leg")
    ap.add_argument("--gkr-uniform-log-gates", type=int, default=20, help="0 = skip the uniform-layer GKR leg")
    ap.add_argument("--gkr-uniform-layers", type=int, default=16)
    ap.add_argument("--ntt-log-n", type=int, default=24, help="NTT / Merkle leg (SURVEY 8f-4); 0 = skip")
    ap.add_argument("--merkle-depth", type=int, default=20)
    ap.add_argument("--kzg-log-inputs", type=int, default=20, help="input-layer commitment leg (multilinear KZG, BLS12-381); 0 = skip")
    ap.add_argument("--target-n-vars", type=int, default=28, help="north-star target leg (1 x 3 factors); 0 = skip")
    ap.add_argument("--config3-n-vars", type=int, default=28, help="configs[3] leg at N >= 2 (2 x 3 factors, global variables); 0 = skip")
    ap.add_argument("--mle-sweep-hi", type=int, default=30, help="largest per-GPU log2 size of the configs[4] sweep; 0 = skip")
    ap.add_argument("--no-plain", action="store_true", help="skip the configs[0] leg")
    ap.add_argument("--products", type=int, default=1, help="ProductPolys in the SumPoly (default: BASELINE configs[1])")
    ap.add_argument("--factors", type=int, default=2, help="factors per ProductPoly")
    ap.add_argument("--tail-log2", type=int, default=None, help="persistent-kernel threshold (0 = one launch per round)")
    ap.add_argument("--small-bytes", type=int, default=None, help="shared-memory budget of the small-table kernel (0 = off)")
    ap.add_argument("--quick", action="store_true", help="profiling run: resident leg only, warm-up as given (numbers are not bench values)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    b = Bench(args)
    line = b.headline()
    if not args.quick:
        extra = {}
        if b.world == 1:
            if args.target_n_vars > 0:
                extra["target"] = b.target_leg()
            if not args.no_plain:
                extra["plain_n20"] = b.plain_leg()
        elif args.config3_n_vars > 0:
            extra["config3"] = b.config3_leg()
        if args.mle_sweep_hi > 0:
            extra["mle_sweep"] = b.mle_sweep_leg()
        if b.rank == 0:
            line.update({k: v for k, v in extra.items() if v is not None})
            if b.world == 1:
                if args.gkr_log_inputs > 0:
                    line["gkr"] = gkr_leg(b.z, args)
                if args.gkr_uniform_log_gates > 0:
                    line["gkr_uniform"] = gkr_uniform_leg(b.z, args)
                if args.kzg_log_inputs > 0:
                    line["kzg"] = kzg_leg(b.z, args)
                if args.ntt_log_n > 0:
                    line.update(fft_merkle_leg(b.z, args))
            print(json.dumps(line))
    if b.sampler:
        b.sampler.stop()
    if b.dist is not None:
        b.dist.barrier()
        b.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
