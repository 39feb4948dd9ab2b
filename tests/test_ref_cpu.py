"""The C++ CPU restatement (what bench.py times as cpu_baseline) against the numpy oracle and the golden value."""
import numpy as np

import cases
from oracle import ref_cpu
from oracle import ref_models as rm
from test_oracle_golden import GOLD1, SEQS1, TREE1


def test_cpp_port_reproduces_the_reference_golden_value():
    r, p = rm.gamma_rates(4, 1.0)
    c = cases.case_from_alignment(TREE1, SEQS1, rm.t92(3.0, 0.5), r, p)
    for scaled in (False, True):
        out = ref_cpu.eval_case(c, scaled=scaled)
        assert abs(-out["lnl"] - GOLD1) < 1e-9


def test_cpp_port_matches_numpy_oracle_with_derivatives_and_threads():
    r, p = rm.gamma_rates(4, 0.5)
    c = cases.make_case(40, 300, rm.gtr(1.2, 0.8, 0.6, 1.5, 0.9, (.3, .2, .25, .25)), r, p, seed=9)
    res = cases.oracle_eval(c, want_d1=True, want_d2=True)
    for nt in (1, 3):
        out = ref_cpu.eval_case(c, want=7, nthreads=nt, site=True)
        assert abs(out["lnl"] - res.lnl) < 1e-10 * abs(res.lnl)
        nb = c.flat.n_nodes - 1
        np.testing.assert_allclose(out["d1"][:nb], res.d1, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(out["d2"][:nb], res.d2, rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(out["site_lnl"], res.site_lnl, rtol=1e-12)


def test_cpp_port_protein_and_underflow():
    r, p = rm.gamma_rates(4, 0.7)
    c = cases.make_case(12, 60, rm.lg08(), r, p, seed=10)
    assert abs(ref_cpu.eval_case(c)["lnl"] - cases.oracle_eval(c).lnl) < 1e-9
    big = cases.make_case(700, 3, rm.gtr(), r, p, seed=4, random_tips=True, mean_brlen=0.5)
    assert ref_cpu.eval_case(big, scaled=False)["lnl"] == -np.inf
    assert abs(ref_cpu.eval_case(big, scaled=True)["lnl"] - cases.oracle_eval(big).lnl) < 1e-9 * 1e4


def test_cpp_port_chromosome_block_form_and_weighted_root():
    """conjugate eigen-pairs (block form), the Chromosome clamp and the fork's weighted root frequencies"""
    r, p = rm.constant_rate()
    m = rm.chromosome(1, 50, gain=1.02, loss=1.90, dupl=0.14, demi=0.95)
    assert m.nonsingular and not m.diagonalizable
    c = cases.make_case(15, 1, m, r, p, seed=12, rooted=True, mean_brlen=0.2, compress=False)
    res = cases.oracle_eval(c, weighted_root=True)
    out = ref_cpu.eval_case(c, weighted_root=True)
    assert abs(out["lnl"] - res.lnl) < 1e-10 * abs(res.lnl)


def test_blocked_driver_equals_the_single_call():
    """refcpu_eval_blocks (arrays allocated once, re-pointed block by block: what bench.py's full-size parity leg runs) returns
    the same lnL as one refcpu_eval over everything, for block sizes that do and do not divide the input."""
    import cases
    from oracle import ref_models as rm
    r, p = rm.gamma_rates(4, 0.5)
    c = cases.make_case(20, 700, rm.gtr(1.2, 0.8, 0.6, 1.5, 0.9, (.3, .2, .25, .25)), r, p, seed=3, compress=False)
    ref = ref_cpu.eval_case(c, nthreads=2)["lnl"]
    off, ch = c.flat.csr()
    codes = np.stack([c.codes_by_leaf[l] for l in range(c.flat.n_nodes) if c.flat.is_leaf[l]])
    m = c.model
    for blk, th in ((700, 1), (128, 4), (96, 3), (333, 8)):
        lnl, sec = ref_cpu.eval_blocks(4, 4, c.N, blk, off, ch, c.flat.root, codes, c.table, c.weights, c.rates, c.probs, m.V, m.Vinv,
                                       m.ev_re, m.rate, c.flat.brlen, c.root_freqs, nthreads=th)
        assert abs(lnl - ref) <= 1e-13 * abs(ref) and sec > 0
