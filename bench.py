#!/usr/bin/env python
"""bench.py -- propagator slices/s on the BASELINE.json headline workload.

    python bench.py --gpus N --steps K --warmup W            # this engine (default N=1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)
    torchrun ... bench.py --gpus N ...                        # N > 1: one rank per GPU over NCCL

Workload (BASELINE.json `metric`): two 3-level transmons, d = 9, K = 2
control lines, N = 1000 PWC slices, B = 4096 control-signal sets PER GPU (the batch axis shards
across GPUs with no data-path collective; one all-gather of the final unitaries per step is
included for N > 1, as BASELINE.json's north_star specifies).  A "step" is one batched
`compute_propagators`-equivalent pass: signals[B,K,N] -> U[B,9,9].

One JSON line on stdout (rank 0).  `value` = whole-job slices/s with inputs resident in HBM;
`e2e` = the same through the public API with HOST buffers (H2D of the signals and D2H of the
unitaries inside the timed region); `roofline` = the fused kernel against the fp64 DFMA peak
measured in this process (MEASURED_PEAKS.json carries no fp64 figure) and against measured
HBM; `cpu_baseline` = the oracle port (the reference restated op-for-op in numpy; TensorFlow
is not installable here) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D, K, N, B_PER_GPU, DT = 9, 2, 1000, 4096, 1e-11
N_ROTATE = 3  # rotating input buffers: 3 x 65.5 MB > 126 MB L2
WORKLOAD = "two-transmon d=9, K=2 controls, N=1000 PWC slices, B=4096 signal sets per GPU (cfg2 shape at the metric's 4k batch)"


# ------------------------------------------------------------------------------------------------
# CPU side (oracle port of the reference) -- the ONLY part of this file that touches oracle/
# ------------------------------------------------------------------------------------------------

def _cpu_worker(args):
    seed, nb = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import numpy as np  # noqa: F401
    from oracle import c3_oracle as orc
    from c3_b200 import synth
    m = synth.two_transmon()
    sig = synth.controls_fast(m, nb, N, DT, seed=seed)
    t0 = time.perf_counter()
    U = orc.propagate_batch(m.h0, m.hks, sig, DT)
    return time.perf_counter() - t0, float(abs(U).sum())


class CpuPool:
    """One worker process per host core, kept alive across steps (spawned, so it is safe to use
    before or after CUDA initialisation)."""

    def __init__(self, n_proc: int):
        import multiprocessing as mp
        self.n = n_proc
        self.pool = mp.get_context("spawn").Pool(n_proc)
        self.pool.map(_cpu_worker, [(10_000 + i, 1) for i in range(n_proc)])  # import warm-up

    def run(self, per_proc: int, seed: int = 0):
        """B = n_proc * per_proc signal sets through the oracle (reference restatement: per
        signal set, batched expm over the N slices then the pairwise product tree, exactly the
        reference's serial loop).  Returns (slices_per_s, wall_s)."""
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, [(seed + i, per_proc) for i in range(self.n)], chunksize=1)
        wall = time.perf_counter() - t0
        return self.n * per_proc * N / wall, wall

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_proc = 6
    vals = []
    pool = CpuPool(cores)
    for _ in range(args.warmup):
        pool.run(1)
    t_all = 0.0
    for s in range(args.steps):
        v, wall = pool.run(per_proc, seed=100 * s)
        vals.append(v)
        t_all += wall
    pool.close()
    value = args.steps * cores * per_proc * N / t_all if t_all > 0 else 0.0
    sample = f"{cores * per_proc} signal sets x {N} slices per step ({cores} processes x {per_proc})"
    line = {
        "impl": "reference", "metric": "propagator slices/sec", "value": value, "unit": "slices/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_all / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": WORKLOAD, "d": D, "K": K, "N": N, "B_per_gpu": B_PER_GPU, "global_batch": B_PER_GPU * max(args.gpus, 1),
                   "dt": DT, "parallelism": f"{cores} host processes", "l2": "n/a (host arm)"},
        "sampled": f"each step times {cores * per_proc} of the {B_PER_GPU} signal sets of the workload (rate-normalised)",
        "cpu_baseline": {"value": value, "unit": "slices/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_note": "TensorFlow (the reference's arithmetic backend) is not installable in this image; "
                          "this is oracle/c3_oracle.py, the numpy restatement pinned to the reference's golden vectors",
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": max(power) if power else None}
        return out



# ------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations and the strong-scaling point (measured AFTER the headline region)
# ------------------------------------------------------------------------------------------------

def _timed(fn, reps, torch, dist, world):
    """Device time of `reps` calls of fn() in ms per call, max over ranks (barrier + synchronize on both sides)."""
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], device=torch.device("cuda", torch.cuda.current_device()), dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms, out


def run_configs(engine, synth, dev, rank, world, dfma_peak, all_gather, check):
    """configs[cfgN] = throughput, roofline fraction and sampled parity of every BASELINE.json configuration at its stated
    size; `strong` = the headline batch (4096 signal sets in TOTAL) split over the ranks.  Multi-GPU configs (cfg3, cfg4,
    cfg5) shard their batch over the ranks with one all-gather of the result, as BASELINE.json words them; cfg1 / cfg2 are
    single-GPU cases and run on every rank's own GPU (rank 0's number is reported)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import scipy.linalg
    from c3_b200 import flops
    from c3_b200.distributed import shard_bounds

    def expm_each(a):
        return np.stack([scipy.linalg.expm(x) for x in a])

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    def gather(U, total):
        return all_gather(U, total) if world > 1 else U

    out = {}
    orc = None
    if check:
        from oracle import c3_oracle as orc            # checker only: sampled rows, never inside a timed region

    # ---- cfg1: single-qubit 3-level, batch 1: latency of one call (prepared model, captured graph) -----------------
    m1 = synth.one_qubit()
    pm1 = engine.prepare_model(m1.h0, m1.hks, DT)
    c1 = {}
    for n1 in (50, 800):
        sig = torch.as_tensor(synth.controls(m1, 1, n1)).to(dev)
        gp = engine.GraphedPwc(pm1, 1, n1)
        ms_graph, U = _timed(lambda: gp.run(), 200, torch, dist, 1)
        ms_call, _ = _timed(lambda: engine.pwc_prepared(pm1, sig), 200, torch, dist, 1)
        ms_cold, _ = _timed(lambda: engine.pwc_closed(m1.h0, m1.hks, sig, DT), 50, torch, dist, 1)
        gp.run(sig)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            gp.run()
        torch.cuda.synchronize()
        host_us = (time.perf_counter() - t0) / 200 * 1e6
        c1[f"N{n1}"] = {"us_per_call_graph_replay": ms_graph * 1e3, "us_per_call_graph_replay_host_clock": host_us,
                        "us_per_call_prepared": ms_call * 1e3, "us_per_call_unprepared": ms_cold * 1e3,
                        "slices_per_s": n1 / (ms_graph * 1e-3)}
        if check:
            c1[f"N{n1}"]["parity_rel_fro_max"] = rel(gp.run(sig).cpu().numpy(), orc.propagate_batch(m1.h0, m1.hks, sig.cpu().numpy(), DT))
    out["cfg1"] = {"workload": "single-qubit 3-level rx90p, batch = 1, N = 50 (BASELINE) and N = 800 (what test/one_qubit.hjson yields)",
                   "launches_per_call": {"unprepared": 5, "prepared": 2, "graph": 1}, **c1}

    # ---- cfg2: two-qubit d = 9, 1000 slices, batch 256 on one GPU -----------------------------------------------------
    m2 = synth.two_transmon()
    sig2_h = synth.controls_fast(m2, 256, N, DT, seed=99)
    sig2 = torch.as_tensor(sig2_h).to(dev)
    pm2 = engine.prepare_model(m2.h0, m2.hks, DT)
    ms, U = _timed(lambda: engine.pwc_prepared(pm2, sig2), 20, torch, dist, 1)
    f2 = flops.flops_per_slice_closed(m2.h0, m2.hks, sig2_h[:4], DT)
    out["cfg2"] = {"workload": "two-qubit d = 9, N = 1000, batch = 256, one GPU", "ms": ms, "slices_per_s": 256 * N / (ms * 1e-3),
                   "roofline": {"bound": "fp64", "achieved": 256 * N * f2 / (ms * 1e-3) / 1e12, "peak": dfma_peak, "unit": "TFLOP/s",
                                "frac": 256 * N * f2 / (ms * 1e-3) / 1e12 / dfma_peak, "flops_per_slice": f2}}
    if check:
        rows = [0, 100, 255]
        want = orc.propagate_batch(m2.h0, m2.hks, sig2_h[rows], DT)
        out["cfg2"]["parity_rel_fro_max"] = max(rel(U[r].cpu().numpy(), want[i]) for i, r in enumerate(rows))

    # ---- cfg3: Lindblad D = 81, 1000 slices, batch 1024 (sharded over the ranks) ------------------------------------
    B3 = 1024
    lo, hi = shard_bounds(B3, world, rank)
    sig3_h = synth.controls_fast(m2, B3, N, DT, seed=7)
    sig3 = torch.as_tensor(sig3_h[lo:hi]).to(dev)
    pm3 = engine.prepare_model(m2.h0, m2.hks, DT, col_ops=m2.col_ops, lindblad=True)
    ms, U3 = _timed(lambda: gather(engine.pwc_prepared(pm3, sig3), B3), 2, torch, dist, world)
    H = m2.h0[None] + np.einsum("kn,kij->nij", sig3_h[0, :, ::50], m2.hks)
    n1 = 2.0 * np.abs(H * DT).sum(axis=-2).max(axis=-1).max()            # superoperator norm ~ 2 x closed
    f3 = flops.flops_lindblad(9, *flops.higham_order(float(n1)))
    out["cfg3"] = {"workload": "two-qubit Lindblad, D = d^2 = 81, N = 1000, batch = 1024" + (f", sharded x{world}" if world > 1 else ""),
                   "ms": ms, "slices_per_s": B3 * N / (ms * 1e-3),
                   "roofline": {"bound": "fp64 (DMMA)", "achieved": (hi - lo) * N * f3 / (ms * 1e-3) / 1e12, "peak": dfma_peak,
                                "unit": "TFLOP/s per GPU", "frac": (hi - lo) * N * f3 / (ms * 1e-3) / 1e12 / dfma_peak, "flops_per_slice": f3}}
    if check:
        rows = [0, B3 - 1]
        want = orc.propagate_batch(m2.h0, m2.hks, sig3_h[rows], DT, col_ops=m2.col_ops, lindbladian=True, expm=expm_each)
        out["cfg3"]["parity_rel_fro_max"] = max(rel(U3[r].cpu().numpy(), want[i]) for i, r in enumerate(rows))
        out["cfg3"]["parity_rows"] = len(rows)
    del U3, sig3

    # ---- cfg4: ORBIT, 4096 random Clifford sequences x 20 Cliffords, d = 9 (sharded over the ranks) ---------------------
    S4 = 4096
    gates = engine.pwc_prepared(pm2, torch.as_tensor(synth.controls(m2, 5, 70)).to(dev))
    idx, lens = synth.rb_sequences(S4, 20, 5, seed=0)
    lo, hi = shard_bounds(S4, world, rank)
    idx_d, lens_d = torch.as_tensor(idx[lo:hi]).to(dev), torch.as_tensor(lens[lo:hi]).to(dev)
    ms, U4 = _timed(lambda: gather(engine.seq_product(gates, idx_d, lens_d), S4), 50, torch, dist, world)
    out["cfg4"] = {"workload": "ORBIT: 4096 sequences x 20 Cliffords (mean 45 native gates), d = 9" + (f", sharded x{world}" if world > 1 else ""),
                   "ms": ms, "sequences_per_s": S4 / (ms * 1e-3), "gate_products_per_s": float(lens.sum()) / (ms * 1e-3),
                   "roofline": {"bound": "latency (1 Gflop in total)", "achieved": float(lens.sum()) * 8 * 729 / (ms * 1e-3) / 1e12,
                                "peak": dfma_peak, "unit": "TFLOP/s", "frac": float(lens.sum()) * 8 * 729 / (ms * 1e-3) / 1e12 / dfma_peak}}
    if check:
        gh = gates.cpu().numpy()
        names = [f"g{i}" for i in range(5)]
        rows = [0, 1, S4 // 2, S4 - 1]
        want = orc.evaluate_sequences({n: gh[i] for i, n in enumerate(names)}, [[names[j] for j in idx[r, :lens[r]]] for r in rows])
        out["cfg4"]["parity_rel_fro_max"] = max(rel(U4[r].cpu().numpy(), want[i]) for i, r in enumerate(rows))

    # ---- cfg5: tunable coupler d = 27, K = 3, 2000 slices, 8192 samples (sharded over the ranks) ------------------------
    m5 = synth.tunable_coupler()
    B5, N5 = 8192, 2000
    lo, hi = shard_bounds(B5, world, rank)
    sig5_h = synth.controls_fast(m5, hi - lo, N5, DT, seed=5, b_offset=rank * 1000003)
    sig5 = torch.as_tensor(sig5_h).to(dev)
    pm5 = engine.prepare_model(m5.h0, m5.hks, DT)
    ms, U5 = _timed(lambda: gather(engine.pwc_prepared(pm5, sig5), B5), 2, torch, dist, world)
    f5 = flops.flops_per_slice_closed(m5.h0, m5.hks, sig5_h[:2, :, ::20], DT)
    out["cfg5"] = {"workload": "tunable coupler d = 27, K = 3, N = 2000, 8192 parameter samples" + (f", sharded x{world}" if world > 1 else ""),
                   "ms": ms, "slices_per_s": B5 * N5 / (ms * 1e-3),
                   "roofline": {"bound": "fp64 (DMMA)", "achieved": (hi - lo) * N5 * f5 / (ms * 1e-3) / 1e12, "peak": dfma_peak,
                                "unit": "TFLOP/s per GPU", "frac": (hi - lo) * N5 * f5 / (ms * 1e-3) / 1e12 / dfma_peak, "flops_per_slice": f5}}
    if check:
        rows = [0, hi - lo - 1]                        # rank 0's shard sits at the head of the gathered batch
        want = orc.propagate_batch(m5.h0, m5.hks, sig5_h[rows], DT, expm=expm_each)
        out["cfg5"]["parity_rel_fro_max"] = max(rel(U5[r].cpu().numpy(), want[i]) for i, r in enumerate(rows))
    del U5, sig5

    # ---- f-1: forward + gradient w.r.t. the control fields against the forward pass alone (rank 0's GPU, reduced batches) ----
    if rank == 0:
        gr = {}
        rng = np.random.default_rng(0)
        for name, mm, Bg, Ng, lind in (("d9", m2, 1024, N, False), ("d27", m5, 296, 200, False), ("D81_lindblad", m2, 148, 40, True)):
            sg = torch.as_tensor(synth.controls_fast(mm, Bg, Ng, DT, seed=3)).to(dev)
            Dg = mm.d * mm.d if lind else mm.d
            Ub = torch.as_tensor(rng.normal(size=(Bg, Dg, Dg)) + 1j * rng.normal(size=(Bg, Dg, Dg))).to(dev)
            if lind:
                fwd = lambda: engine.pwc_lindblad(mm.h0, mm.hks, mm.col_ops, sg, DT)
                bwd = lambda: engine.pwc_lindblad_grad(mm.h0, mm.hks, mm.col_ops, sg, DT, Ub, max_workspace_bytes=24 << 30)
            else:
                fwd = lambda: engine.pwc_closed(mm.h0, mm.hks, sg, DT)
                bwd = lambda: engine.pwc_closed_grad(mm.h0, mm.hks, sg, DT, Ub, max_workspace_bytes=24 << 30)
            ms_f, _ = _timed(fwd, 2, torch, dist, 1)
            ms_g, _ = _timed(bwd, 2, torch, dist, 1)
            gr[name] = {"B": Bg, "N": Ng, "forward_ms": ms_f, "forward_plus_gradient_ms": ms_g, "ratio": ms_g / ms_f,
                        "slices_per_s_with_gradient": Bg * Ng / (ms_g * 1e-3)}
            if not lind:        # the autograd node's pair: a forward pass that keeps its chunk products + the backward pass from them
                def step():
                    U, saved = engine.pwc_closed_saving(mm.h0, mm.hks, sg, DT)
                    return engine.pwc_closed_grad_saved(sg, Ub, saved)
                ms_s, _ = _timed(step, 2, torch, dist, 1)
                gr[name]["forward_saving_plus_backward_ms"] = ms_s
                gr[name]["ratio_saving"] = ms_s / ms_f
        out["gradient"] = {"what": "U and dL/d signals[B,K,N] for a given cotangent of U.  forward_plus_gradient: one call "
                                   "(c3b_pwc_closed_grad / c3b_pwc_lindblad_grad); forward_saving_plus_backward: the two calls of an autograd "
                                   "node (c3b_pwc_closed_fwd_saved + c3b_pwc_closed_bwd_saved), i.e. a whole optimiser step.  Closed d = 9 "
                                   "and d = 27: fused unitary-recurrence kernels without stored propagators; Lindblad D = 81: stored "
                                   "propagators, sweeps and Frechet derivative on the DMMA product", **gr}
        engine.release_workspaces()

    # ---- strong scaling: the headline batch of 4096 signal sets in TOTAL, split over the ranks --------------------------
    lo, hi = shard_bounds(B_PER_GPU, world, rank)
    sig_s = torch.as_tensor(synth.controls_fast(m2, B_PER_GPU, N, DT, seed=4242)[lo:hi]).to(dev)
    ms, _ = _timed(lambda: gather(engine.pwc_prepared(pm2, sig_s), B_PER_GPU), 10, torch, dist, world)
    strong = {"B_total": B_PER_GPU, "B_per_gpu": hi - lo, "n_gpus": world, "ms_per_step": ms, "slices_per_s": B_PER_GPU * N / (ms * 1e-3),
              "note": "same d = 9, N = 1000 workload with the batch FIXED at 4096 in total.  Limiter when the efficiency drops: the per-GPU "
                      "batch (512 rows at 8 GPUs) gives the persistent kernel 12 waves of warp units instead of 100, so its tail, the "
                      "segment-fold launch, the all-gather and ~10 us of launch latency weigh on a 1.1 ms kernel (a B = 512 launch alone "
                      "runs about 10 % below the B = 4096 rate)"}
    return out, strong


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

def run_engine(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and rank == 0 and world > 1:
        print(f"[bench] WORLD_SIZE={world} != --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
    n_gpus = world if world > 1 else 1

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised in this process
    cpu_baseline = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        per_proc = 40
        pool = CpuPool(cores)
        v, wall = pool.run(per_proc)
        pool.close()
        cpu_baseline = {"value": v, "unit": "slices/s", "cores": cores, "kind": "port",
                        "sample": f"{cores * per_proc} signal sets x {N} slices ({cores} processes x {per_proc}), "
                                  f"{wall:.1f} s wall; oracle/c3_oracle.py (TensorFlow not installable)"}

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from c3_b200 import engine, synth, propagation
    from c3_b200.distributed import all_gather_unitaries

    model = synth.two_transmon()
    B = B_PER_GPU
    h0 = torch.as_tensor(model.h0, device=dev)
    hks = torch.as_tensor(model.hks, device=dev)
    host_sig = [torch.as_tensor(synth.controls_fast(model, B, N, DT, seed=1234 + 17 * i, b_offset=rank * 1000003))
                .pin_memory() for i in range(N_ROTATE)]
    dev_sig = [h.to(dev) for h in host_sig]
    U_host = torch.empty((B, D, D), dtype=torch.complex128).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()

    def step(i):
        U = engine.pwc_closed(h0, hks, dev_sig[i % N_ROTATE], DT)
        if world > 1:
            return all_gather_unitaries(U)
        return U

    # ---- parity in the same run: sampled rows against the oracle ---------------------------
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import c3_oracle as orc
        rows = [0, 1, B // 2, B - 1]
        Ug = engine.pwc_closed(h0, hks, dev_sig[0], DT)[rows].cpu().numpy()
        want = orc.propagate_batch(model.h0, model.hks, host_sig[0][rows].numpy(), DT)
        parity = max(float(np.linalg.norm(Ug[i] - want[i]) / np.linalg.norm(want[i])) for i in range(len(rows)))

    # ---- fp64 peak (roofline denominator), measured live -----------------------------------
    dfma_peak = engine.measure_fp64_peak("dfma", 0.5, device=local_rank)
    dmma_peak = engine.measure_fp64_peak("dmma", 0.3, device=local_rank)

    for i in range(max(args.warmup, 3)):  # never fewer than 3 warm-up steps
        step(i)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    launches0 = engine.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = engine.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)

    # ---- dominant kernel alone (events inside the library, on the launching stream) ----------
    engine.set_tuning("profile", 1)
    kms = []
    for i in range(min(args.steps, 10)):
        engine.pwc_closed(h0, hks, dev_sig[i % N_ROTATE], DT)
        kms.append(engine.last_kernel_ms())
    engine.set_tuning("profile", 0)
    kernel_ms = statistics.mean(kms)

    # ---- end to end through the public API with host buffers ---------------------------------
    def e2e_step(i):
        U = propagation.pwc_batch(h0, hks, host_sig[i % N_ROTATE], DT)   # H2D inside
        if world > 1:
            U = all_gather_unitaries(U)[rank * B:(rank + 1) * B]
        U_host.copy_(U, non_blocking=True)                                # D2H inside
        torch.cuda.synchronize()
        return U_host

    for i in range(2):
        e2e_step(i)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    torch.cuda.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    engine.check_gated_launches()          # a gated launch that gave up on a row would have left NaNs: fail loudly

    # ---- end to end from pulse PARAMETERS: the control fields are generated on the device (f-2) --------
    # same gate length / slice count, one DRAG Gaussian per line, per-sample amplitude, angle and detuning
    rng = np.random.default_rng(4321 + rank)
    T = N * DT
    envp = np.zeros((B, K, 1, 9))
    envp[..., 0, 0] = rng.uniform(0.2, 0.5, (B, K))              # amp
    envp[..., 0, 1] = T                                          # t_final
    envp[..., 0, 2] = T / 4                                      # sigma
    envp[..., 0, 3] = rng.uniform(0, 2 * np.pi, (B, K))          # xy_angle
    envp[..., 0, 4] = -2 * np.pi * 53e6                          # freq_offset
    envp[..., 0, 5] = -1.0                                       # delta
    envp[..., 0, 8] = 1.0
    env_host = torch.as_tensor(envp).pin_memory()
    lo_host = torch.as_tensor(np.broadcast_to(2 * np.pi * np.array([5.05e9, 5.65e9]), (B, K)).copy()).pin_memory()
    shape_t = torch.full((K, 1), 2, dtype=torch.int32, device=dev)         # gaussian_nonorm
    flags_t = torch.ones((K, 1), dtype=torch.int32, device=dev)            # DRAG quadrature
    chain_t = torch.as_tensor(np.tile([1.0 / DT, 2e9, 0.3e-9, 1, 0, 1e9, 0, 1, 0, 0, np.nan], (K, 1)), device=dev)
    sig_buf = torch.empty((B, K, N), dtype=torch.float64, device=dev)

    def params_step(i):
        sig = engine.generate_signals(env_host, shape_t, flags_t, lo_host, chain_t, 0.0, T, out=sig_buf)   # H2D of the parameters inside
        U = engine.pwc_closed(h0, hks, sig, DT)
        if world > 1:
            U = all_gather_unitaries(U)[rank * B:(rank + 1) * B]
        U_host.copy_(U, non_blocking=True)
        torch.cuda.synchronize()
        return U_host

    for i in range(2):
        params_step(i)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        params_step(i)
    torch.cuda.synchronize()
    barrier()
    par_s = time.perf_counter() - t0
    clocks = sampler.stop()

    configs, strong = None, None
    if not args.no_configs:
        configs, strong = run_configs(engine, synth, dev, rank, world, dfma_peak, all_gather_unitaries,
                                      check=(rank == 0 and not args.no_parity))

    if world > 1:
        t = torch.tensor([ms_total, e2e_s * 1e3, kernel_ms, par_s * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, kernel_ms, par_ms = [float(x) for x in t.tolist()]
        e2e_s, par_s = e2e_ms * 1e-3, par_ms * 1e-3

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = n_gpus * B * N / (ms_per_step * 1e-3)
        e2e_value = n_gpus * B * N * args.steps / e2e_s
        # algorithmic flops per slice: Higham's minimal (m, s) for the slice norms of this workload
        from c3_b200.flops import flops_per_slice_closed
        fl = flops_per_slice_closed(model.h0, model.hks, host_sig[0][:8].numpy(), DT)
        achieved_tf = B * N * fl / (kernel_ms * 1e-3) / 1e12
        alg_bytes = B * N * (8 * K) + B * D * D * 16
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic, traffic_src = None, None
        try:   # DRAM bytes per launch of the fused kernel from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json" if os.path.exists(os.path.join(ROOT, "profiles", "r02_traffic.json")) else "r01_traffic_s3.json")))
            traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
            traffic_src = tj["source"]
        except Exception:
            pass
        line = {
            "metric": "propagator slices/sec", "value": value, "unit": "slices/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64 FMA)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "d": D, "K": K, "N": N, "B_per_gpu": B, "global_batch": B * n_gpus,
                       "dt": DT, "parallelism": f"batch-sharded x{n_gpus}" + (", one all-gather of U per step" if n_gpus > 1 else ""),
                       "l2": f"{N_ROTATE} rotating input buffers ({N_ROTATE * B * K * N * 8 / 1e6:.0f} MB > 126 MB L2)"},
            "kernel": "pwc_blk9_taylor_kernel<4,2,true,0> (fused assemble + degree-15+ Taylor expm in 4 products + ordered product; 3x3 lane blocks, own-block operands from registers, element-major conflict-free shared layout)",
            "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": B * K * N * 8,
                    "d2h_bytes_per_step": B * D * D * 16, "api": "c3_b200.propagation.pwc_batch(host signals) -> U.cpu()"},
            "e2e_from_params": {"value": n_gpus * B * N * args.steps / par_s, "unit": "slices/s",
                                "h2d_bytes_per_step": int(env_host.numel() * 8 + lo_host.numel() * 8),
                                "d2h_bytes_per_step": B * D * D * 16,
                                "api": "engine.generate_signals(host pulse parameters) -> engine.pwc_closed -> U.cpu(): "
                                       "control fields generated on the device (SURVEY 8f row f-2)"},
            "gpu_launches": launches,
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": dfma_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / dfma_peak, "traffic": traffic, "traffic_unit": "bytes per launch",
                         "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "flops_per_slice": fl, "kernel_ms": kernel_ms,
                         "peak_source": "DFMA micro-benchmark in this process (c3b_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no fp64 entry",
                         "peak_detail": {"dfma_tflops": dfma_peak, "dmma_m8n8k4_tflops": dmma_peak,
                                         "theory_tflops": 148 * 64 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12,
                                         "theory": "148 SMs x 64 DFMA/clk x 2 flop x max SM clock"},
                         "hbm": {"achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
            "configs": configs,
            "strong": strong,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "parity_rel_fro_max": parity,
            "wall_s_timed_region": t_wall,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """Keep fd 1 for the ONE JSON line: everything else that writes to stdout from native code
    (e.g. NCCL's version banner) is sent to stderr."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline only: skip the other BASELINE configs and the strong-scaling point")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
