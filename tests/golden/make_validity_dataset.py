"""Writes tests/golden/validity_dataset.json: the 141-sample experimental profile the reference ships and
fits its models to (/root/reference/src/frontx/examples/data/validity/{__init__.py:11-22, processed.npz};
Gerlero et al. 2022).  It is the data set behind the six parameter sets of the reference's tests
(tests/test_solve.py:47-113) and the natural input of `fit` / `ScaledSolution.fitting_data`; the fit-path
parity tests (tests/test_fit.py) use it as their golden INPUT vector.  Run here (needs /root/reference):

    python tests/golden/make_validity_dataset.py
"""
import json
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/src/frontx/examples/data/validity/processed.npz")
theta_s, theta_i = 0.7, 0.025  # __init__.py:11-13

f = np.load(SRC)
out = {
    "source": "gerlero/frontx src/frontx/examples/data/validity/processed.npz (o, I, std); theta = theta_i + "
              "(theta_s - theta_i) I, std scaled likewise (__init__.py:20-22)",
    "theta_s": theta_s, "theta_i": theta_i,
    "o": [float(v) for v in f["o"]],
    "theta": [float(v) for v in theta_i + (theta_s - theta_i) * f["I"]],
    "std": [float(v) for v in (theta_s - theta_i) * f["std"]],
}
Path(__file__).with_name("validity_dataset.json").write_text(json.dumps(out))
print(len(out["o"]), "samples")
