#!/usr/bin/env python
"""bench.py -- fp64 Boltzmann shooting solves/sec (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic problems.  The default workload is
BASELINE config C2, the 10^5-problem van Genuchten (alpha, n, Ks) sweep for horizontal imbibition
(b = 0.7 - 1e-7, i = 0.025, itol = 1e-3), per GPU: with N GPUs every rank solves its own 10^5-problem
sweep (weak scaling, no data-path collective) and the 64-byte result records are all-gathered over NCCL
inside the timed region.  The kernel is the persistent shot-queue kernel of frontx_b200/csrc/fx_solve.cuh
(one launch per step does the whole sweep).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --config C3|C4|C5 ...    # the other BASELINE configs as the timed workload
    python bench.py --impl reference ...     # the CPU arm: the oracle on all host cores

Beside the metric the C2 line carries, measured after the timed region:
  c3_strong   BASELINE config C3 AS CONFIGURED: 10^7 mixed Brooks-Corey / LETd problems STRONG-scaled over the
              N GPUs through frontx_b200._dist.solve_shards_distributed (every rank builds only its own
              shard, block-seeded; one all-gather + one indexed copy), device-timed and end to end;
  latency     config C1 (one power-law solve with dense output) and a 64-candidate `fit` generation;
  inverse_c5  the inverse scan on config C5 against the HBM copy bandwidth;
  cpu_baseline the CPU restatement on the host cores (-O3 -march=native, built on this box).

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, all host threads), NOT
the reference's JAX executable: jax, diffrax and optimistix are not installable in this image (no wheels,
no network).
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

B_BOUNDARY = 0.7 - 1e-7
I_INITIAL = 0.025
ITOL = 1e-3
BROOKS_COREY, VAN_GENUCHTEN, LETD = 2, 3, 5
C3_BLOCK = 4096
OB, HEIGHT = 1e-6, 1e-3  # C4: cylindrical, angle 2 pi

# algorithmic (weighted) flops, SURVEY.md section 8d: add/mul/cmp = 1, FMA = 2, div/sqrt = 4,
# exp/log = 8, pow = 17.  Per right-hand side: D and D' (F_D1) + RHS and chord update/norm (44); per
# Jacobian: D'' (F_D2) + assembly and 2x2 inverse (32); per step attempt: stage sums, error norm, controller (40).
F_D1 = {0: 0, 1: 25, 2: 25, 3: 82, 4: 140, 5: 58, 6: 25}
F_D2 = {0: 6, 1: 6, 2: 6, 3: 40, 4: 60, 5: 30, 6: 6}
F_ATTEMPT = 40
# algorithmic bytes per problem: model id + 12 parameters + b + i in, summary + counters out
BYTES_IN, BYTES_OUT = 4 + 96 + 8 + 8, 64 + 32

METRIC = "fp64 Boltzmann shooting solves/sec"
WORKLOADS = {
    "C2": "C2: 10^5 van Genuchten (alpha, n, Ks) sweep, horizontal imbibition, b=0.7-1e-7, i=0.025, itol=1e-3",
    "C3": "C3: mixed Brooks-Corey/LETd batch, horizontal imbibition, b=0.7-1e-7, i=0.025, itol=1e-3",
    "C4": "C4: van Genuchten (paper set), cylindrical flow-rate boundary, ob=1e-6, Qb ~ logU[1e-13, 5e-10]",
    "C5": "C5: profile -> D(theta) inverse scan, profiles of 10^6 points, theta = exp(-a o), D at the knots",
}


def c2_workload(n: int, seed: int):
    """van Genuchten sweep of SURVEY.md section 8d (C2): n ~ U[1.5, 9], alpha ~ logU[0.5, 5],
    Ks ~ logU[1e-7, 1e-4], l = 0.5, theta_range = (0.005, 0.7)."""
    from frontx_b200 import models as M

    rng = np.random.default_rng(seed)
    nn = rng.uniform(1.5, 9.0, n)
    alpha = np.exp(rng.uniform(np.log(0.5), np.log(5.0), n))
    Ks = np.exp(rng.uniform(np.log(1e-7), np.log(1e-4), n))
    return M.VanGenuchten.batch(n=nn, alpha=alpha, Ks=Ks, l=0.5, theta_range=(0.005, 0.7))


def c3_rows(idx: np.ndarray):
    """Rows ``idx`` (global problem indices, ascending) of config C3: problem k is
    Brooks-Corey when k is even, LETd when odd; BC n ~ U[0.2, 0.6], l ~ U[1, 5], Ks ~ logU[1e-6, 1e-5],
    theta_range (2.378e-5, 0.7); LETd L ~ U[0.004, 1.5], E ~ logU[1e3, 2e4], T ~ U[1.0, 1.6],
    Dwt ~ logU[1e-4, 2e-3], theta_range (0.0198, 0.7).  Every block of 4096 problems has its own generator
    (seed (1, block)), so any rank can build any shard without the rest of the batch."""
    from frontx_b200 import ModelBatch

    idx = np.asarray(idx, dtype=np.int64)
    ids = np.empty(idx.size, dtype=np.int32)
    par = np.zeros((idx.size, 12))
    blocks = idx // C3_BLOCK
    starts = np.flatnonzero(np.r_[True, blocks[1:] != blocks[:-1]])  # idx ascending: one contiguous run per block
    for a, z in zip(starts, np.r_[starts[1:], idx.size]):
        blk = int(blocks[a])
        u = np.random.default_rng([1, blk]).random((C3_BLOCK, 4))[idx[a:z] - blk * C3_BLOCK]
        even = idx[a:z] % 2 == 0
        odd = ~even
        p = par[a:z]
        # Brooks-Corey {n, l, Ks, alpha, theta_r, theta_s}
        p[even, 0] = 0.2 + 0.4 * u[even, 0]
        p[even, 1] = 1.0 + 4.0 * u[even, 1]
        p[even, 2] = np.exp(np.log(1e-6) + np.log(10.0) * u[even, 2])
        p[even, 3], p[even, 4], p[even, 5] = 1.0, 2.378e-5, 0.7
        # LETd {L, E, T, Dwt, theta_r, theta_s}
        p[odd, 0] = 0.004 + 1.496 * u[odd, 0]
        p[odd, 1] = np.exp(np.log(1e3) + np.log(20.0) * u[odd, 1])
        p[odd, 2] = 1.0 + 0.6 * u[odd, 2]
        p[odd, 3] = np.exp(np.log(1e-4) + np.log(20.0) * u[odd, 3])
        p[odd, 4], p[odd, 5] = 0.0198, 0.7
        ids[a:z] = np.where(even, BROOKS_COREY, LETD)
    return ModelBatch(ids, par)


def c4_workload(n: int, seed: int = 2):
    """C4: the paper's van Genuchten set (tests/test_solve.py:73-78 of the reference), cylindrical, ob = 1e-6,
    angle 2 pi, height 1e-3, Qb ~ logU[1e-13, 5e-10] (the range a boundary value below saturation can carry)."""
    from frontx_b200 import ModelBatch
    from frontx_b200 import models as M

    vg = M.VanGenuchten(n=8.093, l=2.344, Ks=2.079e-6, theta_range=(0.004943, 0.7))
    rng = np.random.default_rng(seed)
    Qb = np.exp(rng.uniform(np.log(1e-13), np.log(5e-10), n))
    return ModelBatch(np.full(n, vg.model_id, dtype=np.int32), np.tile(vg.params(), (n, 1))), Qb


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []
        self.thread = None

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            assert self.proc and self.proc.stdout
            for line in self.proc.stdout:
                self.lines.append(line.strip())

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


NCU_RAW = ROOT / "profiles" / "r2_vg_kernel_ncu_raw_n100k.txt"


def ncu_metrics(n: int):
    """From the committed `ncu --set full` capture of the dominant kernel (profiles/, taken at 10^5 problems per
    launch; None for other sizes): DRAM bytes per launch, FP64-pipe and issue utilisation."""
    if n != 100_000 or not NCU_RAW.exists():
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    want = {"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct"}
    total, util = 0.0, {}
    for line in NCU_RAW.read_text().splitlines():
        f = line.split()
        if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[1] in unit:
            total += float(f[2]) * unit[f[1]]
        if len(f) == 3 and f[0] in want:
            util[want[f[0]]] = float(f[2])
    return total or None, util or None


def hbm_peak():
    peaks = ROOT / "MEASURED_PEAKS.json"
    if peaks.exists():
        try:
            return float(json.loads(peaks.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:  # noqa: BLE001
            pass
    return 6527.5, "fallback 6527.5 GB/s (copy bandwidth measured on this pool earlier)"


def flops_of(ids: np.ndarray, cnt: np.ndarray) -> float:
    """F = n_rhs (F_D1 + 44) + n_jac (F_D2 + 32) + n_attempts 40, from the kernel's own counters."""
    f1 = np.vectorize(F_D1.get)(ids).astype(np.float64) + 44.0
    f2 = np.vectorize(F_D2.get)(ids).astype(np.float64) + 32.0
    ok = cnt[:, 0] >= 0
    return float((cnt[ok, 5] * f1[ok] + cnt[ok, 6] * f2[ok] + cnt[ok, 3] * F_ATTEMPT).sum())


def c5_inputs(dev, batch: int, npts: int):
    import torch

    gen = torch.Generator(device=dev).manual_seed(3)
    a = 0.5 + 1.5 * torch.rand((batch, 1), dtype=torch.float64, device=dev, generator=gen)
    o = (torch.linspace(0, 20, npts, dtype=torch.float64, device=dev).unsqueeze(0) / a).contiguous()
    return a, o, torch.exp(-a * o)


def inverse_c5_probe(dev, stream, batch: int = 1000, npts: int = 1_000_000, reps: int = 3):
    """BASELINE config C5 (10^3 profiles x 10^6 points, theta = exp(-a o)): the profile -> D inverse scan,
    inputs resident, CUDA events, against the HBM copy bandwidth (24 algorithmic bytes per point)."""
    import torch

    from frontx_b200 import _lib

    peak, src = hbm_peak()
    L = _lib.lib()
    a, o, theta = c5_inputs(dev, batch, npts)
    D = torch.empty_like(o)
    sorp = torch.empty((batch, 4), dtype=torch.float64, device=dev)
    status = torch.empty((batch,), dtype=torch.int32, device=dev)
    best = {False: float("inf"), True: float("inf")}
    for with_sorp in (False, True):
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            _lib.check(L.fx_inverse_batch(batch, npts, o.data_ptr(), theta.data_ptr(), D.data_ptr(),
                                          sorp.data_ptr() if with_sorp else None, status.data_ptr(), None, None,
                                          stream.cuda_stream))
            e1.record(stream)
            torch.cuda.synchronize()
            best[with_sorp] = min(best[with_sorp], e0.elapsed_time(e1))
    want = (1 - torch.log(theta[:, npts // 100: npts // 2])) / (2 * a * a)
    err = float(((D[:, npts // 100: npts // 2] - want).abs() / want).max())
    gbs = batch * npts * 24 / best[False] / 1e6
    out = {"workload": f"C5: {batch} profiles x {npts} points, D at the knots", "ms": best[False], "bound": "hbm",
           "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": src,
           "with_sorptivity_ms": best[True], "with_sorptivity_frac": batch * npts * 24 / best[True] / 1e6 / peak,
           "status_ok": int((status == 0).sum().item()), "max_rel_dev_from_analytic_D": err}
    del o, theta, D, want
    torch.cuda.empty_cache()
    return out


def latency_probe(dev, stream):
    """Config C1 (one frontx.solve: D = theta^4, b = 1, i = 0.1, dense output) and a 64-candidate generation of
    `fit` with dense output, device-timed through fx_solve_batch (median of 10)."""
    import torch

    from frontx_b200 import _lib
    from frontx_b200 import models as M

    L = _lib.lib()

    def timed(models, b, i, k_max):
        n = len(models)
        ids, par = models.to_device(dev)
        bt = torch.full((n,), b, dtype=torch.float64, device=dev)
        it = torch.full((n,), i, dtype=torch.float64, device=dev)
        summary = torch.empty((n, _lib.NSUM), dtype=torch.float64, device=dev)
        counters = torch.zeros((n, _lib.NCNT), dtype=torch.int32, device=dev)
        knots = torch.zeros((n, k_max, 4), dtype=torch.float64, device=dev)
        opt = _lib.default_options(itol=ITOL, model_mask=models.mask())
        ms = []
        for _ in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            _lib.check(L.fx_solve_batch(n, ids.data_ptr(), par.data_ptr(), bt.data_ptr(), it.data_ptr(), opt,
                                        summary.data_ptr(), counters.data_ptr(), k_max, knots.data_ptr(),
                                        stream.cuda_stream))
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms[2:])), float(counters[:, 1].double().mean().item())

    c1, shots1 = timed(M.ModelBatch.from_models([M.PowerLaw(k=4.0)]), 1.0, 0.1, 512)
    gen, shots64 = timed(c2_workload(64, seed=7), B_BOUNDARY, I_INITIAL, 512)
    return {"c1_single_solve_ms": c1, "c1_shots": shots1, "fit_generation_64_ms": gen, "fit_generation_shots": shots64,
            "note": "device time of one fx_solve_batch call with dense output (k_max=512), median of 10; round 1 "
                    "(no speculation with dense output, all seven model grids launched): 10.4 ms and 34.8 ms"}


def cpu_arm(models, n_sample: int, threads: int, twin: bool = False):
    """Time the oracle (CPU restatement of the reference algorithm) on the first n_sample problems."""
    from oracle import fxo

    fn = fxo.twin_solve_batch if twin else fxo.solve_batch
    t0 = time.perf_counter()
    out = fn(models.model_id[:n_sample], models.params[:n_sample], B_BOUNDARY, I_INITIAL, itol=ITOL, nthreads=threads)
    dt = time.perf_counter() - t0
    return n_sample / dt, dt, out


CPU_LABEL = ("CPU restatement of the reference algorithm (oracle/fx_oracle.c: libm pow, Taylor jets, LU, plain nested "
             "loops), not the reference's JAX executable (jax/diffrax not installable here)")


def run_reference(args) -> None:
    """--impl reference: the CPU arm, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import fxo

    fxo.build()
    flags = fxo.use_native()
    threads = fxo.num_threads()
    models = c2_workload(args.problems, seed=0)
    # bounded sample per step: ~4 s of CPU work, sized from a short probe
    probe = min(64 * threads, len(models))
    rate, _, _ = cpu_arm(models, probe, threads)
    sample = int(min(len(models), max(probe, rate * 4.0)))
    for _ in range(args.warmup):
        cpu_arm(models, min(sample, probe), threads)
    t_total = 0.0
    for _ in range(args.steps):
        _, dt, _ = cpu_arm(models, sample, threads)
        t_total += dt
    value = sample * args.steps / t_total
    twin_rate, _, _ = cpu_arm(models, min(len(models), 4 * sample), threads, twin=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS["C2"], "problems_per_gpu": args.problems, "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": "port",
                         "per_core": value / threads, "build": flags,
                         "sample": f"first {sample} problems of the C2 sweep per step, {threads} threads; " + CPU_LABEL,
                         "twin": {"value": twin_rate, "per_core": twin_rate / threads,
                                  "what": "the GPU kernel's own division-free arithmetic headers built for the host "
                                          "(oracle/fx_twin.cpp), same problems: the faster CPU formulation"}},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def c3_strong_probe(rank: int, world: int, dev, n_total: int, reps: int = 2):
    """BASELINE config C3 as configured: n_total mixed Brooks-Corey / LETd problems sharded over the ranks through
    frontx_b200._dist.solve_shards_distributed.  Device-timed (max over ranks): solve + all-gather of summaries and
    counters + the indexed copy into the caller's order; end to end adds the H2D copy of every rank's shard from
    pinned host memory and the D2H read of the global records on rank 0."""
    import torch
    import torch.distributed as dist

    from frontx_b200 import ModelBatch, _dist, _lib

    mine = _dist.shard_indices(n_total, rank, world, C3_BLOCK)
    t0 = time.perf_counter()
    shard = c3_rows(mine)
    gen_s = time.perf_counter() - t0
    order = torch.as_tensor(_dist.gather_order(n_total, world, C3_BLOCK), device=dev)
    h_ids = torch.from_numpy(shard.model_id).pin_memory()
    h_par = torch.from_numpy(shard.params).pin_memory()
    stream = torch.cuda.current_stream(dev)
    out = {}

    def build_resident(idx):
        return shard, B_BOUNDARY, I_INITIAL  # ids/params cached on the device by ModelBatch.to_device

    def build_from_host(idx):
        mb = ModelBatch(shard.model_id, shard.params)
        mb._dev = (h_ids.to(dev, non_blocking=True), h_par.to(dev, non_blocking=True))
        return mb, B_BOUNDARY, I_INITIAL

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    pinned = None
    if rank == 0:  # where the global records land on the host (allocated outside the timed region, like the inputs)
        pinned = (torch.empty((n_total, _lib.NSUM), dtype=torch.float64).pin_memory(),
                  torch.empty((n_total, _lib.NCNT), dtype=torch.int32).pin_memory())

    def run(builder, pull: bool):
        s, c, _ = _dist.solve_shards_distributed(builder, n_total, itol=ITOL, block=C3_BLOCK, order=order)
        if pull and rank == 0:
            pinned[0].copy_(s, non_blocking=True)
            pinned[1].copy_(c, non_blocking=True)
        return s, c

    run(build_resident, False)  # warm-up
    sync()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        summary, counters = run(build_resident, False)
        e1.record(stream)
        sync()
        ms.append(e0.elapsed_time(e1))
    t = torch.tensor([min(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    cnt = counters.cpu().numpy()
    sync()
    t0 = time.perf_counter()
    run(build_from_host, True)
    torch.cuda.synchronize(dev)
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    out = {"workload": WORKLOADS["C3"], "problems_total": n_total, "n_gpus": world, "scaling": "strong",
           "ms": dev_ms, "value": n_total / dev_ms * 1e3, "unit": "solves/s",
           "e2e": {"value": n_total / e2e_s, "unit": "solves/s",
                   "h2d_bytes_per_rank": int(mine.size) * (4 + 96), "d2h_bytes_rank0": n_total * BYTES_OUT},
           "success_fraction": float((cnt[:, 0] == 0).mean()), "shots_mean": float(cnt[:, 1].mean()),
           "rhs_mean": float(cnt[:, 5].mean()), "rhs_max": int(cnt[:, 5].max()),
           "shard_build_s_per_rank": gen_s,
           "path": "frontx_b200._dist.solve_shards_distributed: block-cyclic shards of 4096 built per rank (generator "
                   "seeded by block), fx_solve_batch per rank, all_gather_into_tensor of summaries and counters, one "
                   "indexed copy into the caller's order (all inside the timed region)"}
    del summary, counters
    torch.cuda.empty_cache()
    return out


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from frontx_b200 import _dist, _lib

    rank, world, local = _dist.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py measures the CUDA path; no CUDA device is visible (no CPU fallback)")
    dev = torch.device("cuda", local)
    L = _lib.lib()
    cfg = args.config
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    opt = _lib.default_options(itol=ITOL)
    unit, metric = "solves/s", METRIC
    q_t = lo_t = hi_t = None
    roof_model = VAN_GENUCHTEN

    if cfg == "C5":
        batch = args.problems if args.problems != 100_000 else 1000
        npts = 1_000_000
        a5, o5, th5 = c5_inputs(dev, batch, npts)
        D5 = torch.empty_like(o5)
        st5 = torch.empty((batch,), dtype=torch.int32, device=dev)
        n, unit, metric = batch, "profiles/s", "profile -> D(theta) inversions/sec (10^6-point profiles)"

        def step() -> None:
            _lib.check(L.fx_inverse_batch(batch, npts, o5.data_ptr(), th5.data_ptr(), D5.data_ptr(), None,
                                          st5.data_ptr(), None, None, stream.cuda_stream))
    else:
        n = args.problems
        if cfg == "C2":
            models = c2_workload(n, seed=rank)
        elif cfg == "C3":
            models = c3_rows(np.arange(n, dtype=np.int64) + rank * ((n + C3_BLOCK - 1) // C3_BLOCK) * C3_BLOCK)
        else:
            models, Qb = c4_workload(n, seed=2 + rank)
            q_t = torch.as_tensor(Qb / (2 * np.pi * OB * HEIGHT), device=dev)
            hi_t = torch.as_tensor(models.saturation(), device=dev)
            opt = _lib.default_options(itol=ITOL, radial_k=1, ob=OB)
        opt.model_mask = models.mask()
        ids, par = models.to_device(dev)
        b_t = torch.full((n,), B_BOUNDARY, dtype=torch.float64, device=dev)
        i_t = torch.full((n,), I_INITIAL, dtype=torch.float64, device=dev)
        lo_t = i_t
        summary = torch.empty((n, _lib.NSUM), dtype=torch.float64, device=dev)
        counters = torch.zeros((n, _lib.NCNT), dtype=torch.int32, device=dev)
        gathered = torch.empty((world * n, _lib.NSUM), dtype=torch.float64, device=dev) if world > 1 else None
        counts = np.bincount(models.model_id, minlength=7)
        roof_model = int(np.argmax(counts))

        def step() -> None:
            if cfg == "C4":
                _lib.check(L.fx_solve_flowrate_batch(n, ids.data_ptr(), par.data_ptr(), q_t.data_ptr(), i_t.data_ptr(),
                                                     lo_t.data_ptr(), hi_t.data_ptr(), opt, summary.data_ptr(),
                                                     counters.data_ptr(), 0, None, stream.cuda_stream))
            else:
                _lib.check(L.fx_solve_batch(n, ids.data_ptr(), par.data_ptr(), b_t.data_ptr(), i_t.data_ptr(), opt,
                                            summary.data_ptr(), counters.data_ptr(), 0, None, stream.cuda_stream))
            if world > 1:
                dist.all_gather_into_tensor(gathered, summary)

    def sync_all() -> None:
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    L.fx_set_kernel_timing(1)
    for _ in range(args.warmup):
        flush.zero_()
        step()
    sync_all()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    sync_all()
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event bracket)
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
        if cfg != "C5":
            ms = ctypes.c_float()
            _lib.check(L.fx_last_solve_kernel_ms(roof_model, ctypes.byref(ms)))
            kernel_ms.append(ms.value)
    sync_all()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- end to end through the host-buffer entry point (pinned host memory) ----
    if cfg in ("C2", "C3"):
        h_ids = torch.from_numpy(models.model_id).pin_memory()
        h_par = torch.from_numpy(models.params).pin_memory()
        h_b = torch.full((n,), B_BOUNDARY, dtype=torch.float64).pin_memory()
        h_i = torch.full((n,), I_INITIAL, dtype=torch.float64).pin_memory()
        h_sum = torch.empty((n, _lib.NSUM), dtype=torch.float64).pin_memory()
        h_cnt = torch.zeros((n, _lib.NCNT), dtype=torch.int32).pin_memory()
        h2d, d2h = n * BYTES_IN, n * BYTES_OUT

        def e2e_step() -> None:
            _lib.check(L.fx_solve_batch_host(n, h_ids.data_ptr(), h_par.data_ptr(), h_b.data_ptr(), h_i.data_ptr(),
                                             opt, h_sum.data_ptr(), h_cnt.data_ptr(), 0, None))
    elif cfg == "C4":
        h_q, h_hi = q_t.cpu().pin_memory(), hi_t.cpu().pin_memory()
        h_ids = torch.from_numpy(models.model_id).pin_memory()
        h_par = torch.from_numpy(models.params).pin_memory()
        h_i = torch.full((n,), I_INITIAL, dtype=torch.float64).pin_memory()
        h_sum = torch.empty((n, _lib.NSUM), dtype=torch.float64).pin_memory()
        h_cnt = torch.zeros((n, _lib.NCNT), dtype=torch.int32).pin_memory()
        h2d, d2h = n * (BYTES_IN + 16), n * BYTES_OUT

        def e2e_step() -> None:  # the public API's path: pinned host arrays in, records back on the host
            d = [x.to(dev, non_blocking=True) for x in (h_ids, h_par, h_q, h_i, h_hi)]
            _lib.check(L.fx_solve_flowrate_batch(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                                                 d[3].data_ptr(), d[4].data_ptr(), opt, summary.data_ptr(),
                                                 counters.data_ptr(), 0, None, stream.cuda_stream))
            h_sum.copy_(summary, non_blocking=True)
            h_cnt.copy_(counters, non_blocking=True)
            torch.cuda.synchronize(dev)
    else:
        h_o, h_th = o5.cpu().pin_memory(), th5.cpu().pin_memory()
        h_D = torch.empty(o5.shape, dtype=torch.float64).pin_memory()
        h2d, d2h = n * npts * 16, n * npts * 8

        def e2e_step() -> None:
            o5.copy_(h_o, non_blocking=True)
            th5.copy_(h_th, non_blocking=True)
            step()
            h_D.copy_(D5, non_blocking=True)
            torch.cuda.synchronize(dev)

    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / float(t.item())

    # ---- roofline of the dominant kernel ----
    config = {"workload": WORKLOADS[cfg], "parallelism": f"problems sharded over {world} GPU(s), no data-path collective; "
              "all-gather of 64-byte summaries" if world > 1 else "1 GPU",
              "l2": "256 MB buffer written between timed iterations (L2 flush)"}
    if cfg == "C5":
        peak, src = hbm_peak()
        gbs = n * npts * 24 * args.steps / (total_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "fx::inverse_scan_kernel<false, true>", "achieved": gbs, "peak": peak,
                    "unit": "GB/s", "frac": gbs / peak, "traffic": None, "peak_source": src,
                    "algorithmic_bytes_per_launch": n * npts * 24}
        config.update({"profiles_per_gpu": n, "points_per_profile": npts, "status_ok": int((st5 == 0).sum().item())})
        ok_frac, shots = None, None
    else:
        cnt = counters.cpu().numpy().astype(np.float64)
        own = models.model_id == roof_model
        flops = flops_of(models.model_id[own], cnt[own])
        k_ms = float(np.mean(kernel_ms))
        peak = ctypes.c_double()
        _lib.check(L.fx_measure_fp64_peak(ctypes.byref(peak), None))
        achieved = flops / (k_ms * 1e-3)
        traffic, util = ncu_metrics(n) if cfg == "C2" else (None, None)
        roofline = {
            "bound": "fp64", "kernel": f"fx::shoot_bisect_kernel<{_lib.MODEL_NAMES[roof_model]}>",
            "achieved": achieved / 1e12, "peak": peak.value / 1e12, "unit": "TFLOP/s",
            "frac": achieved / peak.value, "traffic": traffic, "ncu": util,
            "kernel_ms": k_ms, "algorithmic_flops_per_launch": flops,
            "algorithmic_flops_per_solve": flops / max(int(own.sum()), 1),
            "hbm_gbs": n * (BYTES_IN + BYTES_OUT) / (k_ms * 1e-3) / 1e9,
            "peak_source": "DFMA micro-benchmark run in this process (MEASURED_PEAKS.json has no FP64 entry); "
                           "weights: add/mul=1, FMA=2, div/sqrt=4, exp/log=8",
        }
        ok_frac = float((cnt[:, 0] == 0).mean())
        shots = {"mean": float(cnt[:, 1].mean()), "min": int(cnt[:, 1].min()), "max": int(cnt[:, 1].max())}
        config.update({"problems_per_gpu": n, "success_fraction": ok_frac, "shots": shots})

    # ---- the other configs of the path, for the record (not the metric), on the C2 line only ----
    extra = {}
    if cfg == "C2" and not args.no_extras:
        try:
            extra["c3_strong"] = c3_strong_probe(rank, world, dev, args.c3_problems)
        except Exception as exc:  # noqa: BLE001  the metric line must not depend on it
            extra["c3_strong"] = {"error": str(exc)[:300]}
        if rank == 0 and world == 1:
            for key, fn in (("latency", latency_probe), ("inverse_c5", inverse_c5_probe)):
                try:
                    extra[key] = fn(dev, stream)
                except Exception as exc:  # noqa: BLE001
                    extra[key] = {"error": str(exc)[:300]}

    cpu_baseline = None
    if rank == 0 and world == 1 and cfg == "C2" and not args.no_cpu_baseline:
        from oracle import fxo

        fxo.build()
        flags = fxo.use_native()
        threads = fxo.num_threads()
        probe = min(64 * threads, n)
        rate, _, _ = cpu_arm(models, probe, threads)
        sample = int(min(n, max(probe, rate * 15.0)))  # ~15 s of CPU work
        rate, dt, ref = cpu_arm(models, sample, threads)
        twin_rate, twin_dt, tw = cpu_arm(models, min(n, 4 * sample), threads, twin=True)
        got = summary[:sample, 0].cpu().numpy()
        cpu_baseline = {"value": rate, "unit": "solves/s", "cores": threads, "kind": "port", "per_core": rate / threads,
                        "build": flags,
                        "sample": f"first {sample} problems of the C2 sweep, {threads} threads, {dt:.1f} s; " + CPU_LABEL,
                        "d_dob_agreement_1e-8": float(np.mean(np.abs(ref["summary"][:, 0] / got - 1) < 1e-8)),
                        "twin": {"value": twin_rate, "per_core": twin_rate / threads,
                                 "sample": f"first {min(n, 4 * sample)} problems, {twin_dt:.1f} s",
                                 "what": "the GPU kernel's own division-free arithmetic headers built for the host "
                                         "(oracle/fx_twin.cpp): the faster CPU formulation, bit-equal to the GPU",
                                 "bit_equal_to_gpu": bool(np.array_equal(
                                     tw["summary"], summary[: tw["summary"].shape[0]].cpu().numpy()))}}

    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", choices=list(WORKLOADS), default="C2", help="the timed workload (BASELINE.json configs)")
    ap.add_argument("--problems", type=int, default=100_000, help="problems per GPU per step (C5: profiles)")
    ap.add_argument("--c3-problems", type=int, default=10_000_000, help="total problems of the c3_strong probe")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip c3_strong / latency / inverse_c5")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
