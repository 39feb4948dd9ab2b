"""Writes tests/golden/clock_optimum.json: the reference's known-answer value for the clock-constrained optimum of
test/test_likelihood_clock.cpp (71.2657, lines 80-93 and 121) next to the parameter point at which the oracle attains it.

The optimum of a likelihood surface does not depend on the optimiser that finds it, so minimising the ORACLE's -lnL over
the clock parametrisation (TotalHeight, HeightP<id>: RHomogeneousClockTreeLikelihood.cpp:121-179) and T92's kappa / theta
must land on the reference's value; the argmin is stored so that the CUDA path can be evaluated at it without an optimiser.
(The other "final" values of the reference tests -- 65.7229 in test_likelihood.cpp, 71.2657 for the unconstrained fit of the
clock test -- are end points of the reference's optimiser on surfaces where the oracle finds better optima (64.9261, 71.0564):
they pin the optimiser, not the likelihood, and are not used.)

    python tests/golden/make_clock_optimum.py
"""
import json
import pathlib
import sys

import numpy as np
from scipy.optimize import minimize

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import cases  # noqa: E402
from oracle import ref_models as rm  # noqa: E402
from oracle import ref_tree as rt  # noqa: E402

TREE = "(((A:0.01, B:0.01):0.02,C:0.03):0.01,D:0.04);"
SEQS = {"A": "AAATGGCTGTGCACGTC", "B": "AACTGGATCTGCATGTC", "C": "ATCTGGACGTGCACGTG", "D": "CAACGGGAGTGCGCCTA"}


def clock_case():
    r, p = rm.constant_rate()
    return cases.case_from_alignment(TREE, SEQS, rm.t92(3.0, 0.5), r, p, check_rooted=False)


def minus_lnl(c, total_height, height_p, kappa, theta):
    m = rm.t92(kappa, theta)
    c.root_freqs = np.asarray(m.freq)
    return -cases.oracle_eval(c, brlen=rt.clock_branch_lengths(c.flat, total_height, height_p), model=m).lnl


def optimise(c):
    """from the reference test's starting point: the tree's own heights, kappa = 3, theta = 0.5"""
    h0, hp0 = rt.clock_parameters(c.flat)
    ids = sorted(hp0)

    def f(x):
        return minus_lnl(c, x[0], dict(zip(ids, x[1:1 + len(ids)])), x[-2], x[-1])
    x = np.array([h0] + [hp0[i] for i in ids] + [3.0, 0.5])
    bounds = [(1e-6, 1e4)] + [(1e-9, 1 - 1e-9)] * len(ids) + [(1e-4, 1e3), (1e-3, 0.999)]
    for _ in range(4):
        x = minimize(f, x, method="L-BFGS-B", bounds=bounds, options={"maxiter": 3000, "ftol": 1e-15, "gtol": 1e-9}).x
    return ids, x, f(x)


if __name__ == "__main__":
    c = clock_case()
    ids, x, _ = optimise(c)

    def f(y):
        return minus_lnl(c, y[0], dict(zip(ids, y[1:1 + len(ids)])), y[-2], y[-1])
    for _ in range(3):
        x = minimize(f, x, method="Nelder-Mead", options={"xatol": 1e-12, "fatol": 1e-14, "maxiter": 20000, "maxfev": 40000}).x
    doc = {
        "_comment": "reference value transcribed from test/test_likelihood_clock.cpp:121 (fitModelHClock final value, tolerance "
                    "0.001 at :91-92); argmin and oracle value written by tests/golden/make_clock_optimum.py",
        "tree": TREE, "sequences": SEQS, "model": "T92", "rates": "Constant",
        "reference_minus_lnl": 71.2657, "reference_tolerance": 0.001,
        "argmin": {"TotalHeight": float("%.12g" % x[0]), "HeightP": {str(i): float("%.12g" % v) for i, v in zip(ids, x[1:1 + len(ids)])},
                   "kappa": float("%.12g" % x[-2]), "theta": float("%.12g" % x[-1])},
    }
    a = doc["argmin"]
    doc["oracle_minus_lnl"] = minus_lnl(c, a["TotalHeight"], {int(k): v for k, v in a["HeightP"].items()}, a["kappa"], a["theta"])
    (HERE / "clock_optimum.json").write_text(json.dumps(doc, indent=1) + "\n")
    print(json.dumps(doc, indent=1))
