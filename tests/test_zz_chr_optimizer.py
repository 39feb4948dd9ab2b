"""SURVEY 8f-1 end to end: the multi-start flow of ChromosomeNumberMng::runChromEvol (App/ChromosomeNumberMng.cpp:264-288) through the
shim on the device -- ChromosomeNumberOptimizer with every point's Brent probe in one device call per step, then the joint ML and
marginal reconstructions at the best point -- with the printed parameter points re-evaluated by the oracle.  (The line-search
machinery itself is checked on CPU: tests/test_cpp_shim.py::test_batched_brent_line_searches_host_side.)"""
import subprocess

import numpy as np
import pytest

import cases
from oracle import ref_likelihood as rl
from oracle import ref_models as rm
from oracle import ref_tree as rt
from test_cpp_shim import compile_cpp


def _case():
    flat = rt.FlatTree(rt.parse_newick("(((a:0.3,b:0.2):0.4,c:0.5):0.1,(d:0.3,e:0.6):0.2);"), check_rooted=False)
    counts = {"a": 7, "b": 8, "c": 14, "d": 9, "e": None}
    ch = cases.Case()
    ch.flat, ch.rates, ch.probs = flat, np.ones(1), np.ones(1)
    ch.table, ch.N, ch.weights = np.vstack([np.eye(30), np.ones((1, 30))]), 1, np.ones(1, np.uint32)
    ch.codes_by_leaf = {lid: np.array([30 if counts[flat.nodes[lid].name] is None else counts[flat.nodes[lid].name] - 1], np.uint8)
                        for lid in flat.leaf_ids}
    return ch


def _oracle_at(ch, p, want_d1=False):
    m = rm.chromosome(1, 30, gain=p[0], loss=p[1], dupl=p[2], demi=p[3])
    ch.model, ch.root_freqs = m, m.freq
    return cases.oracle_eval(ch, weighted_root=True, want_d1=want_d1), m


def test_oracle_line_searches_improve_the_chromosome_likelihood():
    """CPU: what the device run is compared with -- the oracle at the starting points (values the batch test also uses) and one
    bounded line search, which must not make things worse."""
    from scipy.optimize import minimize_scalar
    ch = _case()
    p = [0.7, 0.4, 0.2, 0.1]
    v0 = -_oracle_at(ch, p)[0].lnl
    assert abs(v0 - 9.542596633515) < 1e-9
    r = minimize_scalar(lambda g: -_oracle_at(ch, [g] + p[1:])[0].lnl, bounds=(1e-3, 3.0), method="bounded", options={"xatol": 1e-3})
    assert r.fun <= v0


@pytest.mark.gpu
def test_multi_start_chromosome_optimisation_and_reconstruction_through_the_shim(built_lib):
    """Starting values follow the one parity rule of cases.chromosome_value_parity: 1e-9 on the same exponentiation route; point 5
    (gain 5, demi 1; cond(V) = 1.9e14) takes the series on the device and the eigen route in the oracle, where the device value is
    5.6e-10 from the exact one (mpmath expm) and the oracle's 1.4e-8."""
    exe = compile_cpp("test_chr_optimizer", built_lib)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    vals = {f[0]: float(f[1]) for f in (line.split() for line in r.stdout.splitlines()) if len(f) == 2}
    ch = _case()
    starts = [(0.7, 0.4, 0.2, 0.1), (1.1, 0.4, 0.2, 0.05), (0.2, 1.3, 0.6, 0.3), (2.0, 2.0, 0.01, 0.4), (0.05, 0.05, 0.9, 0.02), (5.0, 0.5, 0.1, 1.0)]
    routes = []
    for k, p in enumerate(starts):
        _, m = _oracle_at(ch, p)
        routes.append(cases.chromosome_value_parity(ch, m, -vals["OPT_START_%d" % k], int(vals["OPT_START_NONSINGULAR_%d" % k]))[0])
    assert routes[:5] == ["eigen"] * 5          # point 5 (cond(V) = 1.9e14) is the one where the two sides pick different routes
    order = [int(vals["OPT_ORDER_%d" % r_]) for r_ in range(6)]
    assert sorted(order) == list(range(6))
    best = order[0]
    # the search never makes a point worse, the best point improved, and it is the best of all
    for k in range(6):
        assert vals["OPT_FINAL_%d" % k] <= vals["OPT_START_%d" % k] + 1e-9, k
    assert vals["OPT_BEST"] == vals["OPT_FINAL_%d" % best]
    assert vals["OPT_BEST"] < min(vals["OPT_START_%d" % k] for k in range(6)) - 0.01
    assert vals["OPT_BEST"] <= min(vals["OPT_FINAL_%d" % k] for k in range(6)) + 1e-12
    # batching: six point evaluations per device call
    assert vals["OPT_POINT_EVALS"] >= 5 * vals["OPT_BATCH_EVALS"]
    # a single likelihood built on the optimised model gives the batch's value
    assert abs(vals["OPT_BEST_SINGLE"] - vals["OPT_BEST"]) <= 1e-8 * vals["OPT_BEST"]
    for n in range(ch.flat.n_nodes):
        assert 0 <= int(vals["OPT_ML_%d" % n]) < 30 and 0 <= int(vals["OPT_MARG_%d" % n]) < 30
        assert 0.0 < vals["OPT_MARGP_%d" % n] <= 1.0 + 1e-12
    # the oracle at the returned parameters: to 1e-9 where both sides exponentiate through a well-conditioned eigensystem; on the
    # edge of the search box the generator is (nearly) defective, one side may fall back to the reference's Taylor rule (tolerance
    # 1e-4 on P, ChromosomeSubstitutionModel.cpp:852-899) and the values agree to that rule's accuracy only
    p = [vals["OPT_PARAM_%d_%s" % (best, n)] for n in ("gain", "loss", "dupl", "demi")]
    res, m = _oracle_at(ch, p, want_d1=True)
    cases.chromosome_value_parity(ch, m, -vals["OPT_BEST_SINGLE"], int(vals["OPT_BEST_NONSINGULAR"]))
    strict = bool(m.nonsingular) and int(vals["OPT_BEST_NONSINGULAR"]) == 1
    if strict:
        flat = ch.flat
        _, ml_root = rl.ml_joint_reconstruction(flat, ch.codes_by_leaf, ch.table, res.P, res.root_freqs)
        assert abs(vals["OPT_ML_BEST_LNL"] - np.log(ml_root[0, 0].max())) <= 1e-8
        for n in range(flat.n_nodes):
            post, _ = rl.marginal_posteriors(flat, res, res.P, n, ch.probs)
            assert abs(vals["OPT_MARGP_%d" % n] - post[0].max()) <= 1e-8, n
