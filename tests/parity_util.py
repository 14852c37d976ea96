"""Helpers shared by the parity tests: profile-level comparison of two solved batches and the
conditioning control that explains (rather than tolerates) what lies outside the 1e-6 criterion.

The shooting problem is ill-conditioned at the reference's own test point b = theta_s - 1e-7: moving ONE
input of the literal oracle by ONE ulp moves 6-30 % of its own van Genuchten profiles by more than 1e-6.
Two implementations of the reference that round D(theta) differently -- the reference on two backends
included -- therefore cannot agree to 1e-6 on those problems.  The tests (a) compare theta(o) on 128
points and the sorptivity per problem, (b) measure the oracle's deviation from ITSELF under +-1 ulp
perturbations of b and of two model parameters, and (c) assert that every problem outside 1e-6 either
took a different discrete decision somewhere (step accept/reject, chord convergence, bisection path) or
is one the oracle itself does not reproduce to 1e-7 under those perturbations.
"""

from __future__ import annotations

import numpy as np

SEQ = [0, 1, 2, 3, 4, 6, 7]  # status, shots, expansions, attempts, rejected, Jacobians, knots (not n_rhs)
NPTS = 128


def as_record(sol) -> dict:
    """A frontx_b200.BatchSolution as the dict the oracle wrappers return (numpy)."""
    return {"summary": sol.summary.cpu().numpy(), "counters": sol.counters.cpu().numpy(),
            "knots": None if sol.knots is None else sol.knots.cpu().numpy()}


def _k5(k: np.ndarray) -> np.ndarray:
    """[nk,4] kernel/twin knots (o, theta, theta', theta'') -> the oracle's [nk,5] (o, theta, theta', F0, F1)."""
    return k if k.shape[1] == 5 else np.concatenate([k[:, :3], k[:, 2:3], k[:, 3:4]], axis=1)


def profile_deviation(fxo, A: dict, B: dict, npts: int = NPTS) -> np.ndarray:
    """max over npts points of [0, min(oi_A, oi_B)] of |theta_A(o) - theta_B(o)| / |theta_B(o)|, per problem."""
    n = A["summary"].shape[0]
    dev = np.zeros(n)
    cap_a, cap_b = A["knots"].shape[1], B["knots"].shape[1]
    for j in range(n):
        nk_a, nk_b = int(A["counters"][j, 7]), int(B["counters"][j, 7])
        if nk_a > cap_a or nk_b > cap_b or nk_a < 1 or nk_b < 1:
            dev[j] = np.inf
            continue
        o = np.linspace(0.0, min(A["summary"][j, 1], B["summary"][j, 1]), npts)
        a, _ = fxo.evaluate(_k5(A["knots"][j, :nk_a]), o)
        b, _ = fxo.evaluate(_k5(B["knots"][j, :nk_b]), o)
        dev[j] = np.max(np.abs(a - b) / np.abs(b))
    return dev


def same_decisions(A: dict, B: dict) -> np.ndarray:
    return (A["counters"][:, SEQ] == B["counters"][:, SEQ]).all(axis=1)


def ulp_perturbations(ids: np.ndarray, par: np.ndarray, b):
    """+-1 ulp of b and of two parameters of each model (a scale of D and a shape exponent)."""
    cols = {0: (0, 0), 1: (0, 1), 2: (2, 0), 3: (3, 4), 4: (6, 0), 5: (3, 0), 6: (0, 1)}  # model id -> columns
    out = []
    for sgn in (np.inf, -np.inf):
        out.append((f"b{'+' if sgn > 0 else '-'}1ulp", par, np.nextafter(np.asarray(b, dtype=np.float64), sgn)))
    for which in (0, 1):
        for sgn in (np.inf, -np.inf):
            p2 = par.copy()
            for mid, cc in cols.items():
                rows = ids == mid
                p2[rows, cc[which]] = np.nextafter(p2[rows, cc[which]], sgn)
            out.append((f"param{which}{'+' if sgn > 0 else '-'}1ulp", p2, b))
    return out


def conditioning(fxo, ids, par, b, i, base: dict, k_max: int, **kw):
    """The literal oracle against itself under 1-ulp input perturbations: [n_perturbations, n] deviations."""
    devs, names = [], []
    for name, p2, b2 in ulp_perturbations(ids, par, b):
        other = fxo.solve_batch(ids, p2, b2, i, k_max=k_max, **kw)
        devs.append(profile_deviation(fxo, other, base))
        names.append(name)
    return names, np.array(devs)


def report(label: str, dev, rel_S, same, names, self_dev) -> str:
    pct = lambda v: "/".join(f"{np.percentile(v, q):.1e}" for q in (50, 90, 99))  # noqa: E731
    lines = [f"{label}: same decisions {same.mean():.4f}; theta(o) on {NPTS} pts within 1e-6 {(dev < 1e-6).mean():.4f} "
             f"(p50/p90/p99 {pct(dev)}); sorptivity within 1e-6 {(rel_S < 1e-6).mean():.4f}"]
    for name, d in zip(names, self_dev):
        lines.append(f"    oracle vs itself, {name}: within 1e-6 {(d < 1e-6).mean():.4f} (p50/p90/p99 {pct(d)})")
    return "\n".join(lines)
