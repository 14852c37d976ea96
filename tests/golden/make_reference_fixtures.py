"""Writes tests/golden/reference_fixtures.npz FROM THE REFERENCE ITSELF: run this wherever gerlero/frontx and
its dependencies (jax, diffrax, optimistix, equinox, interpax) import -- it cannot run in the build container or
on the GPU box (no jax wheels, no network), which is why the forward parity is "unpinned" today.  The moment
the file exists, tests/test_reference_fixtures.py pins the CPU oracle (and, with -m gpu, the CUDA path) against
the reference's own numbers:

    pip install frontx            # or: pip install -e /path/to/frontx
    python tests/golden/make_reference_fixtures.py [--n 1000]

Per case it stores what `frontx.solve` returns: d_dob, oi, theta(oi), sorptivity, the status, and theta(o),
d theta/do(o) on 128 points of [0, oi]  (src/frontx/_forward.py:25-57, src/frontx/_boltzmann.py:158-161).
Cases: the six parameter sets of the reference's tests (tests/test_solve.py:47-113), Philip's exact case (:38-42),
the C1 power law, and the first --n problems of BASELINE configs C2 and C3 as bench.py generates them.
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

NPTS = 128


def reference_model(model_id, p):
    """(model id, 12 doubles) of include/frontx_b200.h -> the reference's model object / callable."""
    import jax.numpy as jnp
    from frontx.models import BrooksAndCorey, LETd, LETxs, VanGenuchten

    if model_id == 0:
        return lambda theta: p[0] + 0 * theta
    if model_id == 1:
        return lambda theta: p[0] * theta ** p[1] + p[2]
    if model_id == 2:
        return BrooksAndCorey(n=p[0], l=p[1], Ks=p[2], alpha=p[3], theta_range=(p[4], p[5]))
    if model_id == 3:
        return VanGenuchten(n=p[0], l=p[2], Ks=p[3], alpha=p[4], theta_range=(p[5], p[6]))
    if model_id == 4:
        return LETxs(Lw=p[0], Ew=p[1], Tw=p[2], Ls=p[3], Es=p[4], Ts=p[5], Ks=p[6], alpha=p[7], theta_range=(p[8], p[9]))
    if model_id == 5:
        return LETd(L=p[0], E=p[1], T=p[2], Dwt=p[3], theta_range=(p[4], p[5]))
    if model_id == 6:
        return lambda theta: p[0] + p[1] * jnp.log(theta)
    raise ValueError(model_id)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000, help="problems of each of C2 and C3")
    args = ap.parse_args()
    try:
        import jax

        jax.config.update("jax_enable_x64", True)
        import frontx
    except ImportError as e:
        raise SystemExit(f"this script needs the reference itself (jax, diffrax, optimistix, interpax, frontx): {e}") from None

    import bench
    from tests.conftest import B_PAPER, I_PAPER, paper_models

    ids, par, bs, is_, names = [], [], [], [], []
    for name, m in paper_models().items():
        ids.append(m.model_id), par.append(m.params()), bs.append(B_PAPER), is_.append(I_PAPER), names.append(name)
    from frontx_b200 import D as FD

    for name, m, b, i in (("philip", FD.philip_exact(), 1.0, 0.0), ("power_law k=4 (C1)", FD.power_law(k=4), 1.0, 0.1)):
        ids.append(m.model_id), par.append(m.params()), bs.append(b), is_.append(i), names.append(name)
    c2 = bench.c2_workload(100_000, seed=0)
    c3 = bench.c3_rows(np.arange(args.n))
    for tag, mb in (("C2", c2), ("C3", c3)):
        for j in range(args.n):
            ids.append(int(mb.model_id[j])), par.append(mb.params[j]), bs.append(bench.B_BOUNDARY)
            is_.append(bench.I_INITIAL), names.append(f"{tag}[{j}]")
    n = len(ids)
    out = {k: np.full(n, np.nan) for k in ("d_dob", "oi", "theta_oi", "sorptivity")}
    status = np.zeros(n, dtype=np.int32)
    theta = np.full((n, NPTS), np.nan)
    dtheta = np.full((n, NPTS), np.nan)
    for j in range(n):
        sol = frontx.solve(reference_model(ids[j], par[j]), b=bs[j], i=is_[j], throw=False)
        status[j] = 0 if sol.result == frontx.RESULTS.successful else 1
        out["d_dob"][j], out["oi"][j], out["theta_oi"][j] = float(sol.d_dob), float(sol.oi), float(sol.i)
        out["sorptivity"][j] = float(sol.sorptivity())
        o = np.linspace(0.0, float(sol.oi), NPTS)
        theta[j], dtheta[j] = np.asarray(sol(o)), np.asarray(sol.d_do(o))
        if j % 100 == 0:
            print(j, names[j], out["d_dob"][j], flush=True)
    import diffrax
    import optimistix

    np.savez_compressed(Path(__file__).with_name("reference_fixtures.npz"), model_id=np.array(ids, dtype=np.int32),
                        params=np.stack(par), b=np.array(bs), i=np.array(is_), names=np.array(names), status=status,
                        theta=theta, dtheta_do=dtheta,
                        versions=np.array([f"frontx {frontx.__version__ if hasattr(frontx, '__version__') else '?'}",
                                           f"jax {jax.__version__}", f"diffrax {diffrax.__version__}",
                                           f"optimistix {optimistix.__version__}"]), **out)
    print("wrote", n, "cases")


if __name__ == "__main__":
    main()
