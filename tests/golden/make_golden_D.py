"""Generate tests/golden/diffusivity_golden.json.

The reference (gerlero/frontx) cannot be imported here (no jax), so the golden
values of D, dD/dtheta, d2D/dtheta2 are computed with mpmath at 60 digits
directly from the reference's model DEFINITIONS -- h(theta), kr(theta),
D = Ks*kr*dh/dtheta -- exactly as written in /root/reference/src/frontx/models.py
(:62-66 LETd, :88-100 D=K/C with C=1/(dh/dtheta), :160-174 Brooks and Corey,
:241-253 van Genuchten, :298-315 LETxs), with derivatives by mpmath's
high-precision numerical differentiation.  Parameter sets and evaluation points
are those of the reference's own test, tests/test_solve.py:47-138
(theta = 0.5, theta = i = 0.025) plus the boundary value b = 0.7 - 1e-7 (:127).

Run:  python tests/golden/make_golden_D.py
"""

from __future__ import annotations

import json
from pathlib import Path

import mpmath as mp

mp.mp.dps = 60

PAPER_SETS = {
    "LETd#1": ("letd", {"L": "0.004569", "E": "12930", "T": "1.505", "Dwt": "4.660e-4",
                        "theta_range": ("0.019852", "0.7")}),
    "LETd#2": ("letd", {"Dwt": "1.004e-3", "L": "1.356", "E": "10010", "T": "1.224",
                        "theta_range": ("0.00625", "0.7")}),
    "VG(n)": ("van_genuchten", {"n": "8.093", "l": "2.344", "Ks": "2.079e-6",
                                "theta_range": ("0.004943", "0.7")}),
    "VG(m)": ("van_genuchten", {"m": "0.8861", "l": "2.331", "Ks": "2.105e-6",
                                "theta_range": ("0.005623", "0.7")}),
    "BC": ("brooks_and_corey", {"n": "0.2837", "l": "4.795", "Ks": "3.983e-6",
                                "theta_range": ("2.378e-5", "0.7")}),
    "LETxs": ("letxs", {"Lw": "1.651", "Ew": "230.5", "Tw": "0.9115", "Ls": "0.517", "Es": "493.6",
                        "Ts": "0.3806", "Ks": "8.900e-3", "theta_range": ("0.01176", "0.7")}),
}


def f64(x: str) -> mp.mpf:
    """The fp64 value the reference would hold for this literal."""
    return mp.mpf(float(x))


def model_fn(kind: str, kw: dict):
    tr, ts = (f64(v) for v in kw["theta_range"])
    alpha = f64(kw.get("alpha", "1"))

    def Se(th):
        return (th - tr) / (ts - tr)

    if kind == "letd":
        L, E, T, Dwt = f64(kw["L"]), f64(kw["E"]), f64(kw["T"]), f64(kw["Dwt"])
        return lambda th: Dwt * Se(th) ** L / (Se(th) ** L + E * (1 - Se(th)) ** T)
    Ks = f64(kw["Ks"])
    if kind == "brooks_and_corey":
        n, l = f64(kw["n"]), f64(kw["l"])
        h = lambda th: -1 / (alpha * Se(th) ** (1 / n))  # noqa: E731
        kr = lambda th: Se(th) ** (2 / n + l + 2)  # noqa: E731
    elif kind == "van_genuchten":
        l = f64(kw["l"])
        if "n" in kw:
            n = f64(kw["n"])
            m = mp.mpf(1 - 1 / float(kw["n"]))  # reference computes m in fp64 (models.py:239)
        else:
            m = f64(kw["m"])
            n = mp.mpf(1 / (1 - float(kw["m"])))  # models.py:227
        h = lambda th: -((1 / (Se(th) ** (1 / m)) - 1) ** (1 / n)) / alpha  # noqa: E731
        kr = lambda th: Se(th) ** l * (1 - (1 - Se(th) ** (1 / m)) ** m) ** 2  # noqa: E731
    elif kind == "letxs":
        Lw, Ew, Tw = f64(kw["Lw"]), f64(kw["Ew"]), f64(kw["Tw"])
        Ls, Es, Ts = f64(kw["Ls"]), f64(kw["Es"]), f64(kw["Ts"])
        kr = lambda th: Se(th) ** Lw / (Se(th) ** Lw + Ew * (1 - Se(th)) ** Tw)  # noqa: E731
        h = lambda th: -((1 - Se(th)) ** Ls / ((1 - Se(th)) ** Ls + Es * Se(th) ** Ts)) / alpha  # noqa: E731
    else:
        raise ValueError(kind)
    return lambda th: Ks * kr(th) / (1 / mp.diff(h, th))  # D = K/C, C = 1/(dh/dtheta)


def main() -> None:
    b = mp.mpf(0.7 - 1e-7)
    points = {"0.5": mp.mpf(0.5), "i=0.025": mp.mpf(0.025), "b=0.7-1e-7": b}
    out = {"_doc": "mpmath 60-digit D, D', D'' from the reference model definitions; see make_golden_D.py",
           "sets": {}}
    for name, (kind, kw) in PAPER_SETS.items():
        D = model_fn(kind, kw)
        rows = {}
        for pname, th in points.items():
            d0 = D(th)
            d1 = mp.diff(D, th, 1)
            d2 = mp.diff(D, th, 2)
            rows[pname] = {"theta": float(th), "D": mp.nstr(d0, 20), "D1": mp.nstr(d1, 20), "D2": mp.nstr(d2, 20)}
        out["sets"][name] = {"kind": kind, "kwargs": kw, "values": rows}
    path = Path(__file__).with_name("diffusivity_golden.json")
    path.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
