"""Developer experiment: build variants of the inverse scan kernel (fx_inverse.cuh knobs) and time them on C5
and on the shapes round 1 found weak (odd npts, a few very long profiles).

    python scripts/exp_inv_variants.py build          # here, into build_variants/
    python scripts/exp_inv_variants.py run [names]    # on the GPU box
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
VARIANTS = {f"inv_lag{a}_{b}": [f"-DFX_INV_LAG={a}", f"-DFX_INV_LAG_SORP={b}"]
            for a, b in ((1, 1), (2, 2), (3, 2), (4, 3), (5, 4), (6, 5), (7, 6), (9, 9))}
VARIANTS["inv_r1like"] = ["-DFX_INV_FASTDIV=0", "-DFX_INV_TILESIGN=0"]


def build(names):
    OUT.mkdir(exist_ok=True)
    procs = [(n, subprocess.Popen(["nvcc", *FLAGS, *VARIANTS[n], "-o", str(OUT / f"lib_{n}.so"), "fx_api.cu"],
                                  cwd=str(ROOT / "frontx_b200" / "csrc"))) for n in names]
    for n, p in procs:
        if p.wait() != 0:
            raise SystemExit(f"build of {n} failed")
    print("built", names)


def run_one():
    import torch

    sys.path.insert(0, str(ROOT))
    from frontx_b200 import _lib

    _lib.LIB_PATH = Path(os.environ["FX_LIB"]).resolve()
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev)
    peak = 6527.5
    out = []
    for batch, npts in ((1000, 1_000_000), (1000, 999_999), (4, 100_000_000), (100_000, 10_000)):
        gen = torch.Generator(device=dev).manual_seed(3)
        a = 0.5 + 1.5 * torch.rand((batch, 1), dtype=torch.float64, device=dev, generator=gen)
        o = (torch.linspace(0, 20, npts, dtype=torch.float64, device=dev).unsqueeze(0) / a).contiguous()
        theta = torch.exp(-a * o)
        D = torch.empty_like(o)
        sorp = torch.empty((batch, 4), dtype=torch.float64, device=dev)
        status = torch.empty((batch,), dtype=torch.int32, device=dev)
        best = {}
        for with_sorp in (False, True):
            best[with_sorp] = 1e9
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                _lib.check(L.fx_inverse_batch(batch, npts, o.data_ptr(), theta.data_ptr(), D.data_ptr(),
                                              sorp.data_ptr() if with_sorp else None, status.data_ptr(), None, None,
                                              stream.cuda_stream))
                e1.record(stream)
                torch.cuda.synchronize()
                best[with_sorp] = min(best[with_sorp], e0.elapsed_time(e1))
        h = hashlib.sha256(D[: min(batch, 8)].cpu().numpy().tobytes() + sorp[:, :3].cpu().numpy().tobytes()).hexdigest()[:10]
        f = lambda ms: batch * npts * 24 / 1e6 / ms / peak * 100  # noqa: E731
        out.append(f"{batch}x{npts}: {best[False]:.2f} ms {f(best[False]):.1f}% | sorp {best[True]:.2f} ms "
                   f"{f(best[True]):.1f}% ok {int((status == 0).sum())} {h}")
        del o, theta, D
        torch.cuda.empty_cache()
    print(os.environ.get("FX_VARIANT", "?"), " || ".join(out), flush=True)


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
