#!/usr/bin/env python
"""bench.py -- fp64 Boltzmann shooting solves/sec (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic problems:
BASELINE config C2, the 10^5-problem van Genuchten (alpha, n, Ks) sweep for
horizontal imbibition (b = 0.7 - 1e-7, i = 0.025, itol = 1e-3), per GPU.  With
N GPUs every rank solves its own 10^5-problem sweep (weak scaling, no data-path
collective) and the 64-byte result records are all-gathered over NCCL inside
the timed region.  The kernel is the persistent shot-queue kernel of
frontx_b200/csrc/fx_solve.cuh (one launch per step does the whole sweep).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...     # the CPU arm: the oracle on all host cores

`--impl reference` times the CPU restatement of the reference algorithm
(oracle/, all host threads), NOT the reference's JAX executable: jax, diffrax
and optimistix are not installable in this image (no wheels, no network).
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
VAN_GENUCHTEN = 3

# algorithmic (weighted) flops, SURVEY.md section 8d: add/mul/cmp = 1, FMA = 2, div/sqrt = 4,
# exp/log = 8, pow = 17.  Per right-hand side: D and D' (van Genuchten 82) + RHS and chord
# update/norm (44); per Jacobian: D'' (+40) + assembly and 2x2 inverse (32); per step attempt:
# stage sums, error norm, controller (40).
F_RHS, F_JAC, F_ATTEMPT = 82 + 44, 40 + 32, 40
# algorithmic bytes per problem: model id + 12 parameters + b + i in, summary + counters out
BYTES_IN, BYTES_OUT = 4 + 96 + 8 + 8, 64 + 32


def c2_workload(n: int, seed: int):
    """van Genuchten sweep of SURVEY.md section 8d (C2): n ~ U[1.5, 9], alpha ~ logU[0.5, 5],
    Ks ~ logU[1e-7, 1e-4], l = 0.5, theta_range = (0.005, 0.7)."""
    from frontx_b200 import models as M

    rng = np.random.default_rng(seed)
    nn = rng.uniform(1.5, 9.0, n)
    alpha = np.exp(rng.uniform(np.log(0.5), np.log(5.0), n))
    Ks = np.exp(rng.uniform(np.log(1e-7), np.log(1e-4), n))
    return M.VanGenuchten.batch(n=nn, alpha=alpha, Ks=Ks, l=0.5, theta_range=(0.005, 0.7))


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


def ncu_dram_bytes_per_launch(n: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed
    `ncu --set full` capture (profiles/, taken at 10^5 problems per launch); None for other sizes."""
    path = ROOT / "profiles" / "r1_vg_kernel_ncu_raw_n100k.txt"
    if n != 100_000 or not path.exists():
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = 0.0
    for line in path.read_text().splitlines():
        f = line.split()
        if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[1] in unit:
            total += float(f[2]) * unit[f[1]]
    return total or None


def ncu_pipe_utilisation(n: int):
    """FP64-pipe and issue utilisation of the dominant kernel from the committed ncu capture (percent of
    cycles; measured under ncu at 10^5 problems per launch, so informative only); None for other sizes."""
    path = ROOT / "profiles" / "r1_vg_kernel_ncu_raw_n100k.txt"
    if n != 100_000 or not path.exists():
        return None
    want = {"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct"}
    out = {}
    for line in path.read_text().splitlines():
        f = line.split()
        if len(f) == 3 and f[0] in want:
            out[want[f[0]]] = float(f[2])
    return out or None


def inverse_c5_probe(dev, stream, batch: int = 1000, npts: int = 1_000_000, reps: int = 3):
    """BASELINE config C5 (10^3 profiles x 10^6 points, theta = exp(-a o)): the profile -> D inverse scan,
    inputs resident, CUDA events, against the HBM copy bandwidth (24 algorithmic bytes per point)."""
    import torch

    from frontx_b200 import _lib

    peaks = Path(__file__).resolve().parent / "MEASURED_PEAKS.json"
    peak, src = 6527.5, "fallback 6527.5 GB/s (copy bandwidth measured on this pool earlier)"
    if peaks.exists():
        try:
            peak, src = float(json.loads(peaks.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    L = _lib.lib()
    gen = torch.Generator(device=dev).manual_seed(3)
    a = 0.5 + 1.5 * torch.rand((batch, 1), dtype=torch.float64, device=dev, generator=gen)
    o = (torch.linspace(0, 20, npts, dtype=torch.float64, device=dev).unsqueeze(0) / a).contiguous()
    theta = torch.exp(-a * o)
    D = torch.empty_like(o)
    status = torch.empty((batch,), dtype=torch.int32, device=dev)
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _lib.check(L.fx_inverse_batch(batch, npts, o.data_ptr(), theta.data_ptr(), D.data_ptr(), None,
                                      status.data_ptr(), None, None, stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    want = (1 - torch.log(theta[:, npts // 100: npts // 2])) / (2 * a * a)
    err = float(((D[:, npts // 100: npts // 2] - want).abs() / want).max())
    gbs = batch * npts * 24 / best / 1e6
    out = {"workload": f"C5: {batch} profiles x {npts} points, D at the knots", "ms": best, "bound": "hbm",
           "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": src,
           "status_ok": int((status == 0).sum().item()), "max_rel_dev_from_analytic_D": err}
    del o, theta, D, want
    torch.cuda.empty_cache()
    return out


def cpu_arm(models, n_sample: int, threads: int):
    """Time the oracle (CPU restatement of the reference algorithm) on the first n_sample problems."""
    from oracle import fxo

    t0 = time.perf_counter()
    out = fxo.solve_batch(models.model_id[:n_sample], models.params[:n_sample], B_BOUNDARY, I_INITIAL,
                          itol=ITOL, nthreads=threads)
    dt = time.perf_counter() - t0
    return n_sample / dt, dt, out


def run_reference(args) -> None:
    """--impl reference: the CPU arm, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import fxo

    fxo.build()
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
    line = {
        "impl": "reference", "metric": "fp64 Boltzmann shooting solves/sec", "value": value, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: 10^5 van Genuchten (alpha, n, Ks) sweep, horizontal imbibition, "
                               "b=0.7-1e-7, i=0.025, itol=1e-3", "problems_per_gpu": args.problems,
                   "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": "port",
                         "sample": f"first {sample} problems of the C2 sweep per step, {threads} threads; CPU "
                                   "restatement of the reference algorithm, not the reference's JAX executable "
                                   "(jax/diffrax not installable here)"},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from frontx_b200 import _dist, _lib

    rank, world, local = _dist.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py measures the CUDA path; no CUDA device is visible (no CPU fallback)")
    dev = torch.device("cuda", local)
    n = args.problems
    L = _lib.lib()

    models = c2_workload(n, seed=rank)
    ids, par = models.to_device(dev)
    b_t = torch.full((n,), B_BOUNDARY, dtype=torch.float64, device=dev)
    i_t = torch.full((n,), I_INITIAL, dtype=torch.float64, device=dev)
    summary = torch.empty((n, _lib.NSUM), dtype=torch.float64, device=dev)
    counters = torch.zeros((n, _lib.NCNT), dtype=torch.int32, device=dev)
    gathered = torch.empty((world * n, _lib.NSUM), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    opt = _lib.default_options(itol=ITOL)
    stream = torch.cuda.current_stream(dev)

    def step() -> None:
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
        ms = ctypes.c_float()
        _lib.check(L.fx_last_solve_kernel_ms(VAN_GENUCHTEN, ctypes.byref(ms)))
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
    h_ids = torch.from_numpy(models.model_id).pin_memory()
    h_par = torch.from_numpy(models.params).pin_memory()
    h_b = torch.full((n,), B_BOUNDARY, dtype=torch.float64).pin_memory()
    h_i = torch.full((n,), I_INITIAL, dtype=torch.float64).pin_memory()
    h_sum = torch.empty((n, _lib.NSUM), dtype=torch.float64).pin_memory()
    h_cnt = torch.zeros((n, _lib.NCNT), dtype=torch.int32).pin_memory()

    def e2e_step() -> None:
        _lib.check(L.fx_solve_batch_host(n, h_ids.data_ptr(), h_par.data_ptr(), h_b.data_ptr(), h_i.data_ptr(),
                                         opt, h_sum.data_ptr(), h_cnt.data_ptr(), 0, None))

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

    # ---- roofline of the dominant kernel: shoot_bisect_kernel<van_genuchten> ----
    cnt = counters.cpu().numpy().astype(np.float64)
    flops = float((cnt[:, 5] * F_RHS + cnt[:, 6] * F_JAC + cnt[:, 3] * F_ATTEMPT).sum())
    k_ms = float(np.mean(kernel_ms))
    peak = ctypes.c_double()
    _lib.check(L.fx_measure_fp64_peak(ctypes.byref(peak), None))
    achieved = flops / (k_ms * 1e-3)
    roofline = {
        "bound": "fp64", "kernel": "fx::shoot_bisect_kernel<van_genuchten>",
        "achieved": achieved / 1e12, "peak": peak.value / 1e12, "unit": "TFLOP/s",
        "frac": achieved / peak.value, "traffic": ncu_dram_bytes_per_launch(n),
        "ncu": ncu_pipe_utilisation(n),
        "kernel_ms": k_ms, "algorithmic_flops_per_launch": flops,
        "algorithmic_flops_per_solve": flops / n,
        "hbm_gbs": n * (BYTES_IN + BYTES_OUT) / (k_ms * 1e-3) / 1e9,
        "peak_source": "DFMA micro-benchmark run in this process (MEASURED_PEAKS.json has no FP64 entry); "
                       "weights: add/mul=1, FMA=2, div/sqrt=4, exp/log=8",
    }
    ok_frac = float((cnt[:, 0] == 0).mean())
    shots = {"mean": float(cnt[:, 1].mean()), "min": int(cnt[:, 1].min()), "max": int(cnt[:, 1].max())}

    # ---- the other kernel of the path, for the record (not the metric): the inverse scan on config C5 ----
    inverse_c5 = None
    if rank == 0 and world == 1:
        try:
            inverse_c5 = inverse_c5_probe(dev, stream)
        except Exception as exc:  # the metric line must not depend on it
            inverse_c5 = {"error": str(exc)[:200]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import fxo

        fxo.build()
        threads = fxo.num_threads()
        probe = min(64 * threads, n)
        rate, _, _ = cpu_arm(models, probe, threads)
        sample = int(min(n, max(probe, rate * 15.0)))  # ~15 s of CPU work
        rate, dt, ref = cpu_arm(models, sample, threads)
        agree = float(np.mean(np.abs(ref["summary"][:, 0] / summary[:sample, 0].cpu().numpy() - 1) < 1e-9))
        cpu_baseline = {"value": rate, "unit": "solves/s", "cores": threads, "kind": "port",
                        "sample": f"first {sample} problems of the C2 sweep, {threads} threads, {dt:.1f} s; CPU "
                                  "restatement of the reference algorithm (oracle/), not the reference's JAX "
                                  "executable (jax/diffrax not installable here)",
                        "d_dob_agreement_1e-9": agree}

    if rank == 0:
        line = {
            "metric": "fp64 Boltzmann shooting solves/sec", "value": value, "unit": "solves/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "C2: 10^5 van Genuchten (alpha, n, Ks) sweep, horizontal imbibition, "
                                   "b=0.7-1e-7, i=0.025, itol=1e-3", "problems_per_gpu": n,
                       "parallelism": f"problems sharded over {world} GPU(s), no data-path collective; "
                                      "all-gather of 64-byte summaries" if world > 1 else "1 GPU",
                       "l2": "256 MB buffer written between timed iterations (L2 flush)",
                       "success_fraction": ok_frac, "shots": shots},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": n * BYTES_IN,
                    "d2h_bytes_per_step": n * BYTES_OUT},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "inverse_c5": inverse_c5,
        }
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
    ap.add_argument("--problems", type=int, default=100_000, help="problems per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
