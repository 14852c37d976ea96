"""Developer experiment: build scheduling variants of the shooting kernel (results never depend on the
knobs, see fx_solve.cuh) and time them on C2.

    python scripts/exp_variants.py build             # here (nvcc cross-compiles), into build_variants/
    python scripts/exp_variants.py run [names...]    # on the GPU box: each variant in its own process
"""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "build_variants"
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC", "-shared"]
VARIANTS = {
    "cur": [],
    "ctas6": ["-DFX_SOLVE_CTAS=6"],
    "ctas4": ["-DFX_SOLVE_CTAS=4"],
    "nap500": ["-DFX_SPARSE_NAP=500", "-DFX_SPARSE_POLL_MASK=15u"],
    "nofrozen": ["-DFX_FROZEN_FF=0"],
}


def build(names):
    OUT.mkdir(exist_ok=True)
    procs = []
    for name in names:
        out = OUT / f"lib_{name}.so"
        cmd = ["nvcc", *FLAGS, *VARIANTS[name], "-o", str(out), "fx_api.cu"]
        procs.append((name, subprocess.Popen(cmd, cwd=str(ROOT / "frontx_b200" / "csrc"))))
    for name, p in procs:
        if p.wait() != 0:
            raise SystemExit(f"build of {name} failed")
    print("built", names)


def run_one():
    import numpy as np
    import torch

    sys.path.insert(0, str(ROOT))
    import bench
    from frontx_b200 import _lib

    _lib.LIB_PATH = Path(os.environ["FX_LIB"]).resolve()
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev)
    out = []
    for tag, n in (("C2", 100_000), ("C3", 1_000_000), ("C3", 4_000_000)):
        models = bench.c2_workload(n, seed=0) if tag == "C2" else bench.c3_rows(np.arange(n))
        ids, par = models.to_device(dev)
        b = torch.full((n,), bench.B_BOUNDARY, dtype=torch.float64, device=dev)
        i = torch.full((n,), bench.I_INITIAL, dtype=torch.float64, device=dev)
        summary = torch.empty((n, _lib.NSUM), dtype=torch.float64, device=dev)
        counters = torch.zeros((n, _lib.NCNT), dtype=torch.int32, device=dev)
        opt = _lib.default_options(itol=bench.ITOL, model_mask=models.mask())
        best = 1e9
        for rep in range(4 if n == 100_000 else 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            _lib.check(L.fx_solve_batch(n, ids.data_ptr(), par.data_ptr(), b.data_ptr(), i.data_ptr(), opt,
                                        summary.data_ptr(), counters.data_ptr(), 0, None, stream.cuda_stream))
            e1.record(stream)
            torch.cuda.synchronize()
            if rep or n > 100_000:
                best = min(best, e0.elapsed_time(e1))
        h = hashlib.sha256(summary.cpu().numpy().tobytes() + counters.cpu().numpy().tobytes()).hexdigest()[:16]
        c = counters.cpu().numpy()
        out.append(f"{tag} n={n}: {best:8.2f} ms {n / best / 1e3:6.3f} Msolves/s hash {h} fail {int((c[:, 0] != 0).sum())}")
        del ids, par, b, i, summary, counters, models
        torch.cuda.empty_cache()
    print(os.environ.get("FX_VARIANT", "?"), " | ".join(out), flush=True)


def run(names):
    for name in names:
        env = dict(os.environ, FX_LIB=str(OUT / f"lib_{name}.so"), FX_VARIANT=name)
        subprocess.run([sys.executable, __file__, "one"], env=env, check=False)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "build"
    names = sys.argv[2:] or list(VARIANTS)
    if what == "build":
        build(names)
    elif what == "one":
        run_one()
    else:
        run(names)
