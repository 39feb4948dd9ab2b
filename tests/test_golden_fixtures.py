"""Committed golden fixtures (tests/golden/).

reference_vectors.json: the known-answer values of the reference's own tests (test/test_likelihood.cpp,
test/test_likelihood_clock.cpp) -- the oracle (CPU) and the CUDA path through the C ABI (GPU) must both reproduce them.
oracle_vectors.json: seeded outputs of the oracle for every model family of the benchmark configs (regression fixtures,
written by tests/golden/make_oracle_vectors.py; parity for LG08 / YN98 / Chromosome is unpinned against the reference, which
holds no known-answer test for them) -- the CPU suite keeps the oracle on them, the GPU suite compares the CUDA path with the
stored numbers without executing the oracle.
"""
import importlib.util
import json
import pathlib

import numpy as np
import pytest

import cases
from oracle import ref_models as rm
from oracle import ref_patterns as rp

GOLD = pathlib.Path(__file__).resolve().parent / "golden"
REF = json.loads((GOLD / "reference_vectors.json").read_text())
ORA = json.loads((GOLD / "oracle_vectors.json").read_text())

_spec = importlib.util.spec_from_file_location("make_oracle_vectors", GOLD / "make_oracle_vectors.py")
_mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mk)


def ref_case(v):
    assert v["model"]["name"] == "T92"
    m = rm.t92(v["model"]["kappa"], v["model"]["theta"])
    r, p = rm.gamma_rates(v["rates"]["ncat"], v["rates"]["alpha"]) if v["rates"]["name"] == "Gamma" else rm.constant_rate()
    return cases.case_from_alignment(v["tree"], v["sequences"], m, r, p, check_rooted=v["check_rooted"])


@pytest.mark.parametrize("v", REF["vectors"], ids=[v["id"] for v in REF["vectors"]])
def test_oracle_reproduces_reference_vectors(v):
    res = cases.oracle_eval(ref_case(v))
    digits = len(repr(v["minus_lnl"]).split(".")[1])
    tol = 1e-9 if digits > 8 else 0.5 * 10.0 ** (-digits)      # 94.3957 is quoted to four decimals
    assert abs(-res.lnl - v["minus_lnl"]) <= tol
    assert abs(-res.lnl - v["minus_lnl"]) <= v["tolerance"]     # the reference test's own bound


def test_gamma_means_and_pattern_order_fixtures():
    g = REF["gamma_class_means"]
    r, p = rm.gamma_rates(4, 1.0)
    np.testing.assert_allclose(r, g["values"], atol=g["atol"])
    c = ref_case(REF["vectors"][0])
    assert [u.decode() for u in c.patterns] == REF["pattern_order"]["patterns"]
    assert int(c.weights.sum()) == REF["pattern_order"]["n_sites"]


@pytest.mark.parametrize("v", ORA["vectors"], ids=[v["spec"]["id"] for v in ORA["vectors"]])
def test_oracle_reproduces_its_fixtures(v):
    c, kw = _mk.build(v["spec"])
    res = cases.oracle_eval(c, **kw)
    assert c.N == v["n_patterns"]
    assert abs(res.lnl - v["lnl"]) <= 1e-12 * abs(v["lnl"])
    np.testing.assert_allclose(res.site_lnl[:16], v["site_lnl"], rtol=1e-12)
    if "d1" in v:
        np.testing.assert_allclose(res.d1, v["d1"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(res.d2, v["d2"], rtol=1e-10, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("v", REF["vectors"], ids=[v["id"] for v in REF["vectors"]])
def test_cuda_path_reproduces_reference_vectors(v):
    from bpp_phyl_b200 import capi
    c = ref_case(v)
    for flags in (0, capi.FLAG_R_SEMANTICS, capi.FLAG_FORCE_GENERIC):
        with cases.make_engine(c, flags=flags) as e:
            lnl, _, _ = e.eval(capi.EVAL_LNL)
        assert abs(-lnl[0] - v["minus_lnl"]) <= v["tolerance"]
        assert abs(-lnl[0] - v["minus_lnl"]) <= (1e-9 if v["minus_lnl"] != 94.3957 else 5e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("v", ORA["vectors"], ids=[v["spec"]["id"] for v in ORA["vectors"]])
def test_cuda_path_reproduces_oracle_fixtures(v):
    """the stored numbers, not a live oracle run"""
    from bpp_phyl_b200 import capi
    spec = v["spec"]
    c, kw = _mk.build(spec)
    flags = (capi.FLAG_KEEP_CLVS if spec["derivs"] else 0) | (capi.FLAG_NH_DERIV if spec.get("nh_form") else 0) | \
        (capi.FLAG_WEIGHTED_ROOT if spec.get("weighted_root") else 0)
    with cases.make_engine(c, flags=flags) as e:
        lnl, d1, d2 = e.eval(7 if spec["derivs"] else 1)
        site = e.site_lnl()
        rf = e.root_freqs() if spec.get("weighted_root") else None
    assert abs(lnl[0] - v["lnl"]) <= 1e-9 * abs(v["lnl"])
    np.testing.assert_allclose(site[:16], v["site_lnl"], rtol=1e-11, atol=1e-11)
    if spec["derivs"]:
        nb = c.flat.n_nodes - 1
        np.testing.assert_allclose(-d1[0, :nb], v["d1"], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(-d2[0, :nb], v["d2"], rtol=1e-8, atol=1e-7)
    if rf is not None:
        np.testing.assert_allclose(rf, v["root_freqs"], rtol=1e-9, atol=1e-15)


# ---- the clock-constrained optimum of test/test_likelihood_clock.cpp (a second known-answer value, at other parameters) ----
CLOCK = json.loads((GOLD / "clock_optimum.json").read_text())
_cspec = importlib.util.spec_from_file_location("make_clock_optimum", GOLD / "make_clock_optimum.py")
_ck = importlib.util.module_from_spec(_cspec)
_cspec.loader.exec_module(_ck)


def test_clock_parametrisation_round_trip():
    """TotalHeight / HeightP<id> of the test tree (RHomogeneousClockTreeLikelihood.cpp:121-157) and back (:161-179): the tree is
    ultrametric, so the clock branch lengths are its own and the clock likelihood starts at the unconstrained 94.3957."""
    from oracle import ref_tree as rt
    c = _ck.clock_case()
    h, hp = rt.clock_parameters(c.flat)
    assert abs(h - 0.04) < 1e-15 and set(hp) == {2, 4}
    assert abs(hp[4] - 0.75) < 1e-12 and abs(hp[2] - 1.0 / 3.0) < 1e-12
    np.testing.assert_allclose(rt.clock_branch_lengths(c.flat, h, hp), c.flat.brlen, rtol=0, atol=1e-15)
    assert abs(_ck.minus_lnl(c, h, hp, 3.0, 0.5) - 94.3957) < 5e-5


def test_oracle_reaches_the_reference_clock_optimum():
    """Minimising the ORACLE's -lnL under the clock constraint, from the reference test's starting point, lands on the value the
    reference test demands (71.2657 +- 0.001, test/test_likelihood_clock.cpp:91-92,121) -- in fact on all its printed digits; the
    stored argmin reproduces the stored oracle value."""
    c = _ck.clock_case()
    ids, x, fx = _ck.optimise(c)
    assert abs(fx - CLOCK["reference_minus_lnl"]) <= 1e-4
    a = CLOCK["argmin"]
    v = _ck.minus_lnl(c, a["TotalHeight"], {int(k): w for k, w in a["HeightP"].items()}, a["kappa"], a["theta"])
    assert abs(v - CLOCK["oracle_minus_lnl"]) <= 1e-12 * v
    assert abs(v - CLOCK["reference_minus_lnl"]) <= 1e-4 and fx >= v - 1e-9


@pytest.mark.gpu
def test_cuda_path_at_the_reference_clock_optimum():
    """the CUDA path evaluated at the stored argmin: the reference's known-answer optimum within the reference's tolerance (and
    to its last printed digit), the stored oracle value to 1e-9 relative"""
    from bpp_phyl_b200 import capi
    from oracle import ref_tree as rt
    c = _ck.clock_case()
    a = CLOCK["argmin"]
    bl = rt.clock_branch_lengths(c.flat, a["TotalHeight"], {int(k): w for k, w in a["HeightP"].items()})
    m = rm.t92(a["kappa"], a["theta"])
    c.model, c.root_freqs = m, np.asarray(m.freq)
    for flags in (0, capi.FLAG_R_SEMANTICS, capi.FLAG_FORCE_GENERIC):
        with cases.make_engine(c, flags=flags) as e:
            e.set_branch_lengths(0, bl)
            lnl, _, _ = e.eval(capi.EVAL_LNL)
        assert abs(-lnl[0] - CLOCK["reference_minus_lnl"]) <= 1e-4
        assert abs(-lnl[0] - CLOCK["oracle_minus_lnl"]) <= 1e-9 * CLOCK["oracle_minus_lnl"]
