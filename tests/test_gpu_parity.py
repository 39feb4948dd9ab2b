"""Parity of the CUDA path (through the C ABI) against the oracle.  Needs a B200: -m gpu."""
import numpy as np
import pytest

import cases
from oracle import ref_models as rm
from test_oracle_golden import GOLD1, GOLD2, SEQS1, SEQS2, TREE1, TREE2

pytestmark = pytest.mark.gpu

REL = 1e-9      # north-star tolerance: log L within 1e-9 relative in FP64


def _capi():
    from bpp_phyl_b200 import capi
    return capi


def gtr():
    return rm.gtr(1.2, 0.8, 0.6, 1.5, 0.9, (.3, .2, .25, .25))


def check_value(c, flags=0, **okw):
    capi = _capi()
    res = cases.oracle_eval(c, **okw)
    with cases.make_engine(c, flags=flags) as e:
        lnl, _, _ = e.eval(capi.EVAL_LNL)
        site = e.site_lnl()
        st = e.stats()
    assert abs(lnl[0] - res.lnl) <= REL * abs(res.lnl), (lnl[0], res.lnl)
    np.testing.assert_allclose(site, res.site_lnl, rtol=1e-11, atol=1e-11)
    return st


@pytest.mark.parametrize("flags", [0, 2, 16, 1])
def test_config1_golden_value_through_c_abi(flags):
    """test/test_likelihood.cpp:108 through every kernel family (walk, R semantics, generic, keep)."""
    r, p = rm.gamma_rates(4, 1.0)
    c = cases.case_from_alignment(TREE1, SEQS1, rm.t92(3.0, 0.5), r, p)
    with cases.make_engine(c, flags=flags) as e:
        lnl, _, _ = e.eval()
    assert abs(-lnl[0] - GOLD1) < 1e-9


def test_clock_golden_value_rooted():
    r, p = rm.constant_rate()
    c = cases.case_from_alignment(TREE2, SEQS2, rm.t92(3.0, 0.5), r, p, check_rooted=False)
    with cases.make_engine(c) as e:
        lnl, _, _ = e.eval()
    assert abs(-lnl[0] - GOLD2) < 1e-4


@pytest.mark.parametrize("ncat", [1, 2, 4, 8])
@pytest.mark.parametrize("flags", [0, 1, 16])
def test_dna_random_trees(ncat, flags):
    r, p = rm.gamma_rates(ncat, 0.5)
    c = cases.make_case(64, 700, gtr(), r, p, seed=11 + ncat, ambiguity=0.02)
    st = check_value(c, flags=flags)
    assert st["path"] == (3 if flags & 16 else 1)


@pytest.mark.parametrize("pt,pipe", [(1, 0), (2, 0), (4, 0)])
@pytest.mark.parametrize("flags", [0, 1])
def test_dna_walk_kernel_variants(monkeypatch, pt, pipe, flags):
    """every instantiation of the DNA walk (patterns per thread, software-pipelined event kernel), with and
    without streaming the CLVs out, on ragged sizes and a multifurcating root"""
    monkeypatch.setenv("BPPGPU_WALK4_PT", str(pt))
    monkeypatch.setenv("BPPGPU_WALK4_PIPE", str(pipe))
    r, p = rm.gamma_rates(4, 0.5)
    for ntaxa, nsites, seed in ((64, 700, 5), (9, 130, 6), (300, 65, 7)):
        c = cases.make_case(ntaxa, nsites, gtr(), r, p, seed=seed, ambiguity=0.02, mean_brlen=0.3 if ntaxa == 300 else 0.05,
                            random_tips=ntaxa == 300)
        st = check_value(c, flags=flags)
        assert st["path"] == 1
    if flags & 1:
        capi = _capi()
        c = cases.make_case(30, 200, gtr(), r, p, seed=41)
        res = cases.oracle_eval(c, want_d1=True, want_d2=True)
        with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS) as e:
            lnl, d1, d2 = e.eval(7)
            nb = c.flat.n_nodes - 1
            np.testing.assert_allclose(-d1[0, :nb], res.d1, rtol=1e-8, atol=1e-8)
            np.testing.assert_allclose(-d2[0, :nb], res.d2, rtol=1e-8, atol=1e-7)


@pytest.mark.parametrize("ncat", [1, 2, 4, 8])
def test_dna_chunk_streamed_walk_equals_register_walk(monkeypatch, ncat):
    """walk4c_kernel (class-uniform warps, TMA-streamed table chunks, register slot) against walk4_kernel (BPPGPU_WALK4C=0):
    same arithmetic in the same order, so the per-site log-likelihoods agree to the last bits (2 ulp: the compiler may contract
    log(L) - E ln2 differently in the two kernels); trees deep enough to need several
    chunks and shared-memory stack slots, ambiguity codes (wide tip tables), ragged pattern counts, a star tree."""
    capi = _capi()
    r, p = rm.gamma_rates(ncat, 0.5) if ncat > 1 else rm.constant_rate()
    todo = [cases.make_case(ntaxa, nsites, gtr(), r, p, seed=seed, ambiguity=amb, random_tips=rt_, mean_brlen=bl)
            for ntaxa, nsites, seed, amb, rt_, bl in ((200, 333, 71, 0.05, False, 0.05), (700, 130, 72, 0.0, True, 0.5),
                                                       (5, 1, 73, 0.3, False, 0.1), (33, 64, 74, 0.0, False, 0.05))]
    todo.append(cases.case_from_alignment("(A:0.1,B:0.2,C:0.3,D:0.05,E:0.01);",
                                          {"A": "ACGTNACG", "B": "ACGTRACC", "C": "AGGT-TCG", "D": "ACCTYACG", "E": "TCGTAACG"},
                                          gtr(), r, p))
    for c in todo:
        out = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("BPPGPU_WALK4C", mode)
            for pt in ("1", "2", "4"):
                monkeypatch.setenv("BPPGPU_WALK4_PT", pt)
                with cases.make_engine(c) as e:
                    lnl, _, _ = e.eval(capi.EVAL_LNL)
                    out[mode, pt] = (lnl[0], e.site_lnl().copy())
        res = cases.oracle_eval(c)
        for k, (lnl, site) in out.items():
            np.testing.assert_allclose(site, out["0", "2"][1], rtol=5e-16, atol=0, err_msg=str(k))
            assert abs(lnl - res.lnl) <= REL * abs(res.lnl), k


@pytest.mark.parametrize("flags,env,path", [(0, None, 4), (1, None, 4), (16, None, 3), (0, "walk", 2)])
def test_protein_random_trees(monkeypatch, flags, env, path):
    """S = 20: tensor-core node kernels (default), the generic kernels, and the register walk (BPPGPU_PATH=walk)"""
    if env:
        monkeypatch.setenv("BPPGPU_PATH", env)
    r, p = rm.gamma_rates(4, 0.7)
    c = cases.make_case(40, 300, rm.lg08(), r, p, seed=21, ambiguity=0.02)
    st = check_value(c, flags=flags)
    assert st["path"] == path


def test_codon_generic():
    r, p = rm.constant_rate()
    c = cases.make_case(12, 150, rm.yn98(2.0, 0.3), r, p, seed=31, mean_brlen=0.1)
    check_value(c)


@pytest.mark.parametrize("which,ncat", [("lg08", 4), ("lg08", 1), ("yn98", 1), ("yn98", 2)])
def test_dmma_tensor_core_path(monkeypatch, which, ncat):
    """FP64 tensor-core (mma.sync DMMA) node / upper / derivative kernels: S = 20 (forced) and S = 64 (default)."""
    capi = _capi()
    monkeypatch.setenv("BPPGPU_PATH", "dmma")
    m = rm.lg08() if which == "lg08" else rm.yn98(2.0, 0.3)
    r, p = rm.gamma_rates(ncat, 0.7) if ncat > 1 else rm.constant_rate()
    for ntaxa, nsites, seed, nh in ((14, 150, 61, False), (9, 37, 62, True)):
        c = cases.make_case(ntaxa, nsites, m, r, p, seed=seed, mean_brlen=0.1, ambiguity=0.03, rooted=nh)
        res = cases.oracle_eval(c, want_d1=True, want_d2=True, nh_form=nh)
        flags = capi.FLAG_KEEP_CLVS | (capi.FLAG_NH_DERIV if nh else 0)
        with cases.make_engine(c, flags=flags) as e:
            lnl, d1, d2 = e.eval(7)
            assert e.stats()["path"] == 4
            nb = c.flat.n_nodes - 1
            assert abs(lnl[0] - res.lnl) <= REL * abs(res.lnl)
            np.testing.assert_allclose(e.site_lnl(), res.site_lnl, rtol=1e-11, atol=1e-11)
            np.testing.assert_allclose(-d1[0, :nb], res.d1, rtol=1e-8, atol=1e-8)
            np.testing.assert_allclose(-d2[0, :nb], res.d2, rtol=1e-8, atol=1e-7)
            for nid in range(c.flat.n_nodes - 1):
                if c.flat.is_leaf[nid]:
                    continue
                clv, ex = e.clv(nid, 0)
                np.testing.assert_array_equal(ex, res.lexp[nid])
                np.testing.assert_allclose(clv, res.lower[nid], rtol=1e-10, atol=1e-14 * res.lower[nid].max())
                clv, ex = e.clv(nid, 1)
                got = np.ldexp(clv, -ex[:, :, None].astype(np.int64))
                exp = np.ldexp(res.upper[nid], -res.uexp[nid][:, :, None])
                np.testing.assert_allclose(got, exp, rtol=1e-9, atol=1e-13 * exp.max())


def _check_derivs_and_uppers(c, e, res, want=7, uppers=True):
    capi = _capi()
    lnl, d1, d2 = e.eval(want)
    nb = c.flat.n_nodes - 1
    assert abs(lnl[0] - res.lnl) <= REL * abs(res.lnl)
    np.testing.assert_allclose(-d1[0, :nb], res.d1, rtol=1e-8, atol=1e-8)
    if want & capi.EVAL_D2:
        np.testing.assert_allclose(-d2[0, :nb], res.d2, rtol=1e-8, atol=1e-7)
    if not uppers:
        return
    for nid in range(nb):
        clv, ex = e.clv(nid, 1)
        got = np.ldexp(clv, -ex[:, :, None].astype(np.int64))
        exp = np.ldexp(res.upper[nid], -res.uexp[nid][:, :, None])
        np.testing.assert_allclose(got, exp, rtol=1e-9, atol=1e-13 * exp.max())


@pytest.mark.parametrize("family", ["1", "0"])
@pytest.mark.parametrize("want", [7, 3])
def test_family_kernel_vs_per_branch(monkeypatch, family, want):
    """S = 20 prefix + derivative pass: the per-father fused kernel (default) and the per-branch kernels it replaces,
    value+d1 only (5 column blocks) and value+d1+d2 (8), rooted NH form and unrooted DR form, ragged pattern counts."""
    capi = _capi()
    monkeypatch.setenv("BPPGPU_FAMILY", family)
    r, p = rm.gamma_rates(4, 0.7)
    for ntaxa, nsites, seed, nh in ((14, 150, 71, False), (9, 37, 72, True), (6, 3, 73, False)):
        c = cases.make_case(ntaxa, nsites, rm.lg08(), r, p, seed=seed, mean_brlen=0.1, ambiguity=0.03, rooted=nh,
                            compress=False)
        res = cases.oracle_eval(c, want_d1=True, want_d2=True, nh_form=nh)
        with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS | (capi.FLAG_NH_DERIV if nh else 0)) as e:
            assert e.stats()["path"] == 4
            _check_derivs_and_uppers(c, e, res, want)
            launches = e.stats()["kernel_launches"]
        # one launch per father instead of three per branch
        if family == "1":
            assert launches < 3 * (c.flat.n_nodes - 1)


def test_family_kernel_chunks_multifurcation_and_underflow(monkeypatch):
    """Few CTAs (several 512-pattern accumulation chunks per CTA), a father with 4 sons (falls back to the per-branch
    kernels inside the same pass), a 3-son father below the root, and upper CLVs that need rescaling."""
    capi = _capi()
    r, p = rm.gamma_rates(2, 0.5)
    monkeypatch.setenv("BPPGPU_FAMILY_GRID", "2")
    c = cases.make_case(8, 1300, rm.lg08(), r, p, seed=74, random_tips=True, compress=False)
    res = cases.oracle_eval(c, want_d1=True, want_d2=True)
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS) as e:
        _check_derivs_and_uppers(c, e, res)
    monkeypatch.delenv("BPPGPU_FAMILY_GRID")
    aa = "ARNDCQEGHILKMFPSTWYV"
    rng = np.random.default_rng(75)
    names = list("ABCDEFGHIJ")
    seqs = {n: "".join(rng.choice(list(aa), size=23)) for n in names}
    nwk = "((A:0.1,B:0.2,C:0.3,D:0.05):0.1,(E:0.01,F:0.2,(G:0.1,H:0.1):0.05):0.2,I:0.3,J:0.1);"
    c = cases.case_from_alignment(nwk, seqs, rm.lg08(), r, p, states=aa, aliases={ch: [ch] for ch in aa})
    res = cases.oracle_eval(c, want_d1=True, want_d2=True)
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS) as e:
        assert e.stats()["path"] == 4
        _check_derivs_and_uppers(c, e, res)
    # long branches, i.i.d. tips, 300 taxa: lower and upper rows are rescaled many times
    c = cases.make_case(300, 24, rm.lg08(), r, p, seed=76, random_tips=True, mean_brlen=0.6, compress=False)
    res = cases.oracle_eval(c, want_d1=True, want_d2=True)
    assert res.SR_exp.max() > 256
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS) as e:
        _check_derivs_and_uppers(c, e, res, uppers=False)


@pytest.mark.parametrize("S,ncat,ntaxa,nsites", [(20, 4, 60, 700), (20, 3, 33, 150), (64, 1, 40, 300)])
def test_level_batched_launches_equal_one_launch_per_node(monkeypatch, S, ncat, ntaxa, nsites):
    """dmma_prune_level_kernel (one launch per tree level and kind of sons, the default) and dmma_family_level_kernel (opt-in) do
    the arithmetic of the per-node / per-father launches row by row: lnL, per-site lnL, every lower CLV with its exponents and the
    derivatives are bit-identical, with far fewer launches.  Includes a multifurcation below the root and ragged pattern counts."""
    capi = _capi()
    r, p = rm.gamma_rates(ncat, 0.6) if ncat > 1 else rm.constant_rate()
    model = rm.lg08() if S == 20 else rm.yn98(2.0, 0.3)
    c = cases.make_case(ntaxa, nsites, model, r, p, seed=91, mean_brlen=0.2, ambiguity=0.02, compress=False)
    derivs = S == 20
    got = {}
    for mode in ("per_node", "levels", "levels+deriv"):
        if mode == "levels+deriv" and not derivs:
            continue
        monkeypatch.setenv("BPPGPU_LEVEL_BATCH", "0" if mode == "per_node" else "1")
        monkeypatch.setenv("BPPGPU_LEVEL_BATCH_DERIV", "1" if mode == "levels+deriv" else "0")
        with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS) as e:
            assert e.stats()["path"] == 4
            lnl, d1, d2 = e.eval(capi.EVAL_LNL | ((capi.EVAL_D1 | capi.EVAL_D2) if derivs else 0))
            launches = e.stats()["kernel_launches"]
            site = e.site_lnl()
            clvs = [e.clv(nid, 0) for nid in range(c.flat.n_nodes) if not c.flat.is_leaf[nid]]
        got[mode] = (lnl.copy(), None if d1 is None else d1.copy(), None if d2 is None else d2.copy(), site.copy(), clvs, launches)
    ref = got["per_node"]
    for mode, g in got.items():
        if mode == "per_node":
            continue
        assert g[0][0] == ref[0][0]
        np.testing.assert_array_equal(g[3], ref[3])
        for (a, ea), (b, eb) in zip(g[4], ref[4]):
            np.testing.assert_array_equal(a, b)
            np.testing.assert_array_equal(ea, eb)
        if derivs and mode == "levels":
            np.testing.assert_array_equal(g[1], ref[1])
            np.testing.assert_array_equal(g[2], ref[2])
        elif derivs:   # the per-CTA partial sums of w dL are cut at other pattern boundaries: same terms, another association
            np.testing.assert_allclose(g[1], ref[1], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(g[2], ref[2], rtol=1e-12, atol=1e-12)
        assert g[5] < ref[5]


def test_s20_three_classes_two_byte_codes_with_derivatives():
    """fragment-order S = 20 kernels with an odd class count (class-major rows: crow = N) and 2-byte tip codes, d1 + d2"""
    capi = _capi()
    rng = np.random.default_rng(85)
    r, p = rm.gamma_rates(3, 0.9)
    c = cases.make_case(13, 75, rm.lg08(), r, p, seed=85, compress=False, mean_brlen=0.15)
    extra = rng.dirichlet(np.ones(20), size=300)
    c.table = np.vstack([c.table, extra])
    for lid in c.codes_by_leaf:
        codes = c.codes_by_leaf[lid].astype(np.uint16)
        mask = rng.random(c.N) < 0.3
        codes[mask] = rng.integers(22, 322, size=int(mask.sum()))
        c.codes_by_leaf[lid] = codes
    c.code_dtype = np.uint16
    res = cases.oracle_eval(c, want_d1=True, want_d2=True)
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS) as e:
        assert e.stats()["path"] == 4
        _check_derivs_and_uppers(c, e, res)
        for nid in range(c.flat.n_nodes):
            if not c.flat.is_leaf[nid]:
                clv, ex = e.clv(nid, 0)
                np.testing.assert_array_equal(ex, res.lexp[nid])
                np.testing.assert_allclose(clv, res.lower[nid], rtol=1e-10, atol=1e-14 * res.lower[nid].max())


def test_s20_fallbacks_eight_classes_and_two_points():
    """The S = 20 fragment-order kernels hold the operands of <= 4 rate classes and serve single-point engines; 8 classes
    and 2-point engines must take the per-node / per-branch tensor-core kernels and agree with the oracle all the same."""
    capi = _capi()
    r, p = rm.gamma_rates(8, 0.8)
    c = cases.make_case(11, 90, rm.lg08(), r, p, seed=81, mean_brlen=0.1, ambiguity=0.02, compress=False)
    res = cases.oracle_eval(c, want_d1=True, want_d2=True)
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS) as e:
        assert e.stats()["path"] == 4
        _check_derivs_and_uppers(c, e, res)
    r, p = rm.gamma_rates(4, 0.8)
    c = cases.make_case(9, 70, rm.lg08(), r, p, seed=82, mean_brlen=0.1, compress=False)
    res = cases.oracle_eval(c)
    bl2 = c.flat.brlen * 1.7
    res2 = cases.oracle_eval(c, brlen=bl2)
    with cases.make_engine(c, n_points=2) as e:
        e.set_branch_lengths(1, bl2)
        lnl, _, _ = e.eval(capi.EVAL_LNL)
        assert e.stats()["path"] == 4
    assert abs(lnl[0] - res.lnl) <= REL * abs(res.lnl)
    assert abs(lnl[1] - res2.lnl) <= REL * abs(res2.lnl)


@pytest.mark.parametrize("mk,ncat,rooted", [(gtr, 4, False), (rm.lg08, 4, False), (rm.lg08, 2, True),
                                            (lambda: rm.yn98(2.0, 0.3), 1, False)])
def test_likelihood_at_node_and_posteriors(mk, ncat, rooted):
    """DRTreeLikelihood::computeLikelihoodAtNode and DRTreeLikelihoodTools' posterior probabilities per state
    and rate class, computed from the device-resident lower / upper arrays, for every node (leaves, internal nodes, root)."""
    capi = _capi()
    r, p = rm.gamma_rates(ncat, 0.6) if ncat > 1 else rm.constant_rate()
    c = cases.make_case(12, 60, mk(), r, p, seed=91, rooted=rooted, ambiguity=0.05, compress=False, mean_brlen=0.2)
    res = cases.oracle_eval(c, want_d1=True)
    from oracle import ref_likelihood as rl
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS) as e:
        e.eval(capi.EVAL_LNL | capi.EVAL_D1)
        for nid in range(c.flat.n_nodes):
            la, ex, post = e.node_posteriors(nid)
            exp_post = rl.posterior_probabilities(c.flat, res, res.P, nid, c.probs, c.codes_by_leaf, c.table)
            np.testing.assert_allclose(post, exp_post, rtol=1e-9, atol=1e-14)
            np.testing.assert_allclose(post.sum(axis=(1, 2)), 1.0, rtol=1e-12)
            A, E = rl.likelihood_at_node(c.flat, res, res.P, nid)
            got = np.ldexp(la, -ex[:, :, None].astype(np.int64))
            want = np.ldexp(A, -E[:, :, None])
            np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-13 * want.max())
            # every node sees all the data: sum_x full[i][c][x] is the class likelihood of the site, whatever the node
            site_c = np.einsum("icx->ic", got)
            root_c = np.einsum("icx->ic", np.ldexp(res.lower[c.flat.root] * res.root_freqs, -res.lexp[c.flat.root][:, :, None]))
            np.testing.assert_allclose(site_c, root_c, rtol=1e-8)


def test_dmma_underflow_scaling():
    r, p = rm.constant_rate()
    c = cases.make_case(150, 40, rm.yn98(2.0, 0.3), r, p, seed=63, mean_brlen=0.8)     # simulated: no stop codons
    res = cases.oracle_eval(c)
    assert res.SR_exp.max() > 256 and np.isfinite(res.lnl)
    st = check_value(c)
    assert st["path"] == 4


def test_underflow_scaling_large_tree():
    """i.i.d. tips on a long-branched 700-taxon tree: the unscaled reference arithmetic gives -inf."""
    r, p = rm.gamma_rates(4, 0.5)
    c = cases.make_case(700, 300, gtr(), r, p, seed=4, random_tips=True, mean_brlen=0.5)
    check_value(c)
    check_value(c, flags=16)


def test_ragged_pattern_counts_and_multifurcation():
    r, p = rm.gamma_rates(4, 0.5)
    for n in (1, 2, 63, 65, 257):
        c = cases.make_case(9, n, gtr(), r, p, seed=100 + n, random_tips=True, compress=False)
        check_value(c)
    # star tree: root with 5 sons
    c = cases.case_from_alignment("(A:0.1,B:0.2,C:0.3,D:0.05,E:0.01);",
                                  {"A": "ACGTNACG", "B": "ACGTRACC", "C": "AGGT-TCG", "D": "ACCTYACG", "E": "TCGTAACG"},
                                  gtr(), r, p)
    check_value(c)
    check_value(c, flags=16)


def test_empty_pattern_list():
    r, p = rm.gamma_rates(4, 0.5)
    c = cases.make_case(5, 0, gtr(), r, p, seed=1, random_tips=True, compress=False)
    with cases.make_engine(c) as e:
        lnl, _, _ = e.eval()
    assert lnl[0] == 0.0


@pytest.mark.parametrize("mk,ncat,ntaxa,nsites,flags", [
    (gtr, 4, 30, 200, 0), (gtr, 4, 30, 200, 16), (rm.lg08, 4, 16, 80, 0), (rm.lg08, 2, 16, 80, 16),
    (lambda: rm.yn98(2.0, 0.3), 1, 8, 40, 0)])
def test_branch_derivatives(mk, ncat, ntaxa, nsites, flags):
    capi = _capi()
    r, p = rm.gamma_rates(ncat, 0.6) if ncat > 1 else rm.constant_rate()
    c = cases.make_case(ntaxa, nsites, mk(), r, p, seed=41)
    res = cases.oracle_eval(c, want_d1=True, want_d2=True)
    with cases.make_engine(c, flags=flags | capi.FLAG_KEEP_CLVS) as e:
        lnl, d1, d2 = e.eval(capi.EVAL_LNL | capi.EVAL_D1 | capi.EVAL_D2)
        nb = c.flat.n_nodes - 1
        assert abs(lnl[0] - res.lnl) <= REL * abs(res.lnl)
        # the ABI returns derivatives of +lnL; the oracle (like getFirstOrderDerivative) of -lnL
        np.testing.assert_allclose(-d1[0, :nb], res.d1, rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(-d2[0, :nb], res.d2, rtol=1e-8, atol=1e-7)
        # device-resident CLVs (getLikelihoodData consumers)
        for nid in range(c.flat.n_nodes):
            if c.flat.is_leaf[nid]:
                continue
            clv, ex = e.clv(nid, 0)
            np.testing.assert_array_equal(ex, res.lexp[nid])
            np.testing.assert_allclose(clv, res.lower[nid], rtol=1e-11, atol=1e-14 * res.lower[nid].max())
        for nid in range(nb):
            clv, ex = e.clv(nid, 1)
            got = np.ldexp(clv, -ex[:, :, None].astype(np.int64))
            exp = np.ldexp(res.upper[nid], -res.uexp[nid][:, :, None])
            np.testing.assert_allclose(got, exp, rtol=1e-10, atol=1e-13 * exp.max())
        # pxy_/dpxy_/d2pxy_ tables
        for nid in (0, nb - 1):
            np.testing.assert_allclose(e.transition_probabilities(nid, capi.WANT_P), res.P[nid], rtol=0, atol=1e-13)
            np.testing.assert_allclose(e.transition_probabilities(nid, capi.WANT_DP), res.dP[nid], rtol=0, atol=1e-11)
            np.testing.assert_allclose(e.transition_probabilities(nid, capi.WANT_D2P), res.d2P[nid], rtol=0, atol=1e-9)


@pytest.mark.parametrize("mk,ncat,flags", [(gtr, 4, 1), (rm.lg08, 4, 1), (rm.lg08, 3, 1 | 8), (lambda: rm.yn98(2.0, 0.3), 1, 1)])
def test_per_site_derivative_arrays(mk, ncat, flags):
    """bppgpu_get_site_derivatives = DRASDRTreeLikelihoodData::getDLikelihoodArray / getD2LikelihoodArray: (dL_i/dt)/L_i and
    (d2L_i/dt2)/L_i of every branch, per pattern, against the oracle's arrays (both slab layouts, tips and internal nodes, the NH
    form), and their weighted sums against the evaluation's own d1 / d2."""
    capi = _capi()
    r, p = rm.gamma_rates(ncat, 0.6) if ncat > 1 else rm.constant_rate()
    c = cases.make_case(10, 60, mk(), r, p, seed=123, ambiguity=0.03)
    res = cases.oracle_eval(c, want_d1=True, want_d2=True, nh_form=bool(flags & 8))
    w = c.weights.astype(float)
    with cases.make_engine(c, flags=flags) as e:
        lnl, d1, d2 = e.eval(7)
        for n in range(c.flat.n_nodes - 1):
            a, b = e.site_derivatives(n)
            np.testing.assert_allclose(a, res.dL[n], rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(b, res.d2L[n], rtol=1e-8, atol=1e-9)
            assert abs(np.sum(w * a) - d1[0, n]) <= 1e-9 * max(1.0, abs(d1[0, n]))
            assert abs(np.sum(w * (b - a * a)) - d2[0, n]) <= 1e-8 * max(1.0, abs(d2[0, n]))
        with pytest.raises(capi.BppGpuError):
            e.site_derivatives(c.flat.root)


def test_nh_derivative_form():
    capi = _capi()
    r, p = rm.gamma_rates(4, 0.6)
    c = cases.make_case(20, 100, gtr(), r, p, seed=43, rooted=True)
    res = cases.oracle_eval(c, want_d1=True, want_d2=True, nh_form=True)
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS | capi.FLAG_NH_DERIV) as e:
        lnl, d1, d2 = e.eval(7)
    nb = c.flat.n_nodes - 1
    np.testing.assert_allclose(-d1[0, :nb], res.d1, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(-d2[0, :nb], res.d2, rtol=1e-8, atol=1e-7)


def _nh_models(kind):
    if kind == "dna":
        return [rm.t92(3.0, th) for th in (0.2, 0.5, 0.8)] + [gtr()]
    if kind == "protein":
        lg = rm.lg08()
        rng = np.random.default_rng(8)
        other = []
        exch = lg.Q / lg.freq[None, :]
        np.fill_diagonal(exch, 0.0)
        for _ in range(2):        # LG exchangeabilities with other equilibrium frequencies (LG08+F style)
            f = rng.dirichlet(np.ones(20) * 5)
            other.append(rm._reversible("LG08F", exch, f))
        return [lg] + other
    return [rm.yn98(2.0, 0.3), rm.yn98(1.0, 1.7), rm.yn98(4.0, 0.05)]


@pytest.mark.parametrize("kind,ncat,ntaxa,nsites", [("dna", 4, 14, 150), ("protein", 4, 10, 60), ("protein", 3, 10, 60), ("codon", 1, 8, 30),
                                                    ("codon", 2, 6, 20)])
def test_nonhomogeneous_model_set_per_branch_models(kind, ncat, ntaxa, nsites):
    """SubstitutionModelSet semantics (AbstractNonHomogeneousTreeLikelihood.cpp:394-468): every branch takes P, dP, d2P from the
    model of its own node (bppgpu_set_branch_models), rooted tree, free root frequencies; value, per-site values, NH-form
    derivatives and the P tables against the oracle."""
    capi = _capi()
    models = _nh_models(kind)
    r, p = rm.gamma_rates(ncat, 0.8) if ncat > 1 else rm.constant_rate()
    c = cases.make_case(ntaxa, nsites, models[0], r, p, seed=61, rooted=True, ambiguity=0.03, mean_brlen=0.15)
    rng = np.random.default_rng(12)
    nn = c.flat.n_nodes
    slots = rng.integers(len(models), size=nn).astype(np.int32)
    slots[:len(models)] = np.arange(len(models))             # every model used at least once
    c.root_freqs = rng.dirichlet(np.ones(models[0].size) * 3)
    res = cases.oracle_eval_nh(c, models, slots, want_d1=True, want_d2=True)
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS | capi.FLAG_NH_DERIV, n_models=len(models)) as e:
        holders = [cases.to_model_desc(m) for m in models]
        e._model_holders = holders
        for k, md in enumerate(holders):
            e.set_model(k, md)
        e.set_branch_models(0, slots)
        lnl, d1, d2 = e.eval(7)
        assert abs(lnl[0] - res.lnl) <= REL * abs(res.lnl)
        np.testing.assert_allclose(e.site_lnl(), res.site_lnl, rtol=1e-11, atol=1e-11)
        nb = nn - 1
        np.testing.assert_allclose(-d1[0, :nb], res.d1, rtol=1e-8, atol=1e-7)
        np.testing.assert_allclose(-d2[0, :nb], res.d2, rtol=1e-8, atol=1e-6)
        for nid in (0, nb // 2, nb - 1):
            np.testing.assert_allclose(e.transition_probabilities(nid), res.P[nid], rtol=0, atol=1e-12)
        # value-only evaluation after the derivative one, and back to a homogeneous assignment
        lnl2, _, _ = e.eval(capi.EVAL_LNL)
        assert lnl2[0] == lnl[0]
        e.set_branch_models(0, np.zeros(nn, np.int32))
        c0 = cases.oracle_eval_nh(c, models, np.zeros(nn, np.int64))
        lnl3, _, _ = e.eval(capi.EVAL_LNL)
        assert abs(lnl3[0] - c0.lnl) <= REL * abs(c0.lnl)


def _generic_singular(S):
    """a pure-birth chain: upper bidiagonal generator with equal rates -> one Jordan block, not diagonalisable"""
    Q = np.zeros((S, S))
    for i in range(S - 1):
        Q[i, i + 1] = 1.0
    Q = rm._set_diagonal(Q)
    m = rm.update_matrices(rm.Model("birth", Q, np.full(S, 1.0 / S), reversible=False, scalable=False))
    return m


@pytest.mark.parametrize("name", ["t92", "gtr", "lg08", "yn98", "chr_eigen", "chr_complex", "chr_complex50", "chr_real200",
                                  "chr_complex200", "chr_singular", "chr_singular200", "chr_singular52_exact", "generic_singular40", "nonrev4"])
def test_pt_batch_interface(name):
    """Interface 1: getPij_t / getdPij_dt / getd2Pij_dt2 for a batch of t."""
    capi = _capi()
    if name == "nonrev4":
        Q = np.array([[0, .9, .05, .05], [.05, 0, .9, .05], [.05, .05, 0, .9], [.9, .05, .05, 0.]])
        Q = rm._set_diagonal(Q)
        m = rm.update_matrices(rm.Model("NR", Q, np.full(4, .25), reversible=False))
        assert not m.diagonalizable and m.nonsingular          # complex pair -> block form
    else:
        m = {"t92": lambda: rm.t92(3.0, 0.5), "gtr": gtr, "lg08": rm.lg08, "yn98": lambda: rm.yn98(2.0, 0.3),
             "chr_eigen": lambda: rm.chromosome(1, 40, gain=0.7, loss=0.4, dupl=0.2, demi=rm.DEMI_EQUAL_DUPL),
             "chr_complex": lambda: rm.chromosome(1, 25, gain=1.5, loss=0.1, dupl=0.9, demi=0.4, gain_r=0.05),
             "chr_complex50": lambda: rm.chromosome(1, 50, gain=1.02, loss=1.90, dupl=0.14, demi=0.95),
             "chr_real200": lambda: rm.chromosome(1, 104, gain=1.0, loss=1.0, dupl=0.01),     # real spectrum, S % 8 == 0
             "chr_complex200": lambda: rm.chromosome(1, 200, gain=0.08, loss=1.06, dupl=0.46, demi=0.06),
             "chr_singular": lambda: rm.chromosome(1, 20, gain=0.5, loss=0.0, dupl=0.0),
             # the series route on the FP64 tensor cores (S >= 32: mat_mul_dmma): the reference's adaptive Taylor rule at S = 200,
             # S not a multiple of 8, and the generic 30-term rule of a non-Chromosome singular generator
             "chr_singular200": lambda: rm.chromosome(1, 200, gain=0.5, loss=0.0, dupl=0.1),
             "chr_singular52_exact": lambda: rm.chromosome(1, 52, gain=0.5, loss=0.0, dupl=0.0),
             "generic_singular40": lambda: _generic_singular(40)}[name]()
        if name.startswith("chr_singular") or name == "generic_singular40":
            assert not m.nonsingular
        if name in ("chr_complex50", "chr_complex200"):
            assert m.nonsingular and not m.diagonalizable       # conjugate pairs -> block form on the tensor cores
    ts = np.array([0.0, 1e-6, 0.013, 0.2, 1.0, 4.5])
    P, dP, d2P = capi.pt_batch(cases.to_model_desc(m), ts, 7)
    for k, t in enumerate(ts):
        # absolute tolerances scale with the conditioning of the eigenvector basis (V.f(L).V^-1 in FP64)
        kap = np.linalg.cond(m.V) if m.nonsingular else 1.0
        q = max(1.0, np.abs(m.Q).max() * m.rate)
        np.testing.assert_allclose(P[k], rm.pij_t(m, t), rtol=0, atol=2e-15 * kap + 2e-13, err_msg="P t=%g" % t)
        np.testing.assert_allclose(dP[k], rm.dpij_dt(m, t), rtol=1e-10, atol=(2e-15 * kap + 1e-12) * q, err_msg="dP t=%g" % t)
        np.testing.assert_allclose(d2P[k], rm.d2pij_dt2(m, t), rtol=1e-10, atol=(2e-15 * kap + 1e-12) * q * q * 10, err_msg="d2P t=%g" % t)


@pytest.mark.parametrize("S,npts,ntaxa", [(30, 5, 25), (144, 3, 12)])
def test_chromosome_weighted_root_batched_points(S, npts, ntaxa):
    """ChromEvol shape: one character, C = 1, weighted root frequencies, many parameter points.  S = 144 takes the
    stacked tensor-core P(t) kernel (row blocks of a point's matrices stacked along M), conjugate eigen-pairs included."""
    capi = _capi()
    rng = np.random.default_rng(7)
    r, p = rm.constant_rate()
    pts = []
    while len(pts) < npts:
        m = rm.chromosome(1, S, gain=rng.uniform(0.2, 2), loss=rng.uniform(0.2, 2), dupl=rng.uniform(0.05, 1),
                          demi=rm.DEMI_EQUAL_DUPL if S == 30 else rng.uniform(0.05, 1))
        if m.nonsingular and np.linalg.cond(m.V) < 1e6:
            pts.append(m)
    c = cases.make_case(ntaxa, 1, pts[0], r, p, seed=51, rooted=True, mean_brlen=0.2 if S == 30 else 0.05, compress=False)
    off, ch = c.flat.csr()
    e = capi.Engine(S, 1, 1, off, ch, c.flat.root, c.table, n_points=len(pts), n_models=len(pts),
                    flags=capi.FLAG_WEIGHTED_ROOT)
    for lid, codes in c.codes_by_leaf.items():
        e.set_tip_codes(lid, codes)
    e.set_pattern_weights(c.weights)
    e.set_rates(r, p)
    holders = [cases.to_model_desc(m) for m in pts]
    for k, h in enumerate(holders):
        e.set_model(k, h)
        e.set_branch_lengths(k, c.flat.brlen)
    lnl, _, _ = e.eval()
    for k, m in enumerate(pts):
        res = cases.oracle_eval(c, model=m, weighted_root=True)
        assert abs(lnl[k] - res.lnl) <= REL * abs(res.lnl)
        # entries near 1e-14 differ between the routes: the table route clamps P < 0 to 1e-20 entry by entry, the factored one cannot
        np.testing.assert_allclose(e.root_freqs(k), res.root_freqs, rtol=1e-9, atol=1e-12)
    e.close()


@pytest.mark.parametrize("S,ntaxa", [(30, 25), (52, 40), (200, 60)])
def test_factored_points_route_against_the_table_route_and_the_oracle(monkeypatch, S, ntaxa):
    """chr_level_kernel (P(t) applied as V T(t) V^-1 x on the FP64 tensor cores, no tables) against the table route
    (BPPGPU_POINTS_FACTORED=0) and the oracle: real and complex spectra, S not a multiple of 8, an unknown count ('X': dense leaf
    column), the guard sending an ill-conditioned and a singular point to the table route inside the same batch."""
    capi = _capi()
    rng = np.random.default_rng(S)
    r, p = rm.constant_rate()
    pts = []
    while len(pts) < 5:
        m = rm.chromosome(1, S, gain=rng.uniform(0.2, 2), loss=rng.uniform(0.2, 2), dupl=rng.uniform(0.05, 1), demi=rng.uniform(0.05, 1))
        if m.nonsingular and np.linalg.cond(m.V) < 1e7:
            pts.append(m)
    pts.append(rm.chromosome(1, S, gain=0.5, loss=0.0, dupl=0.0))                   # singular generator: Taylor route
    assert not pts[-1].nonsingular
    c = cases.make_case(ntaxa, 1, pts[0], r, p, seed=77, rooted=True, mean_brlen=0.1, compress=False)
    unknown = S                                                                       # the all-ones row of cases' code table
    some_leaf = sorted(c.codes_by_leaf)[3]
    c.codes_by_leaf[some_leaf] = np.array([unknown], c.code_dtype)
    off, ch = c.flat.csr()
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("BPPGPU_POINTS_FACTORED", mode)
        e = capi.Engine(S, 1, 1, off, ch, c.flat.root, c.table, n_points=len(pts), n_models=len(pts), flags=capi.FLAG_WEIGHTED_ROOT,
                        code_bytes=np.dtype(c.code_dtype).itemsize)
        for lid, codes in c.codes_by_leaf.items():
            e.set_tip_codes(lid, codes)
        e.set_pattern_weights(c.weights)
        e.set_rates(r, p)
        holders = [cases.to_model_desc(m) for m in pts]
        for k, h in enumerate(holders):
            e.set_model(k, h)
            e.set_branch_lengths(k, c.flat.brlen * (1.0 + 0.1 * k))
        lnl, _, _ = e.eval()
        st = e.stats()
        out[mode] = (lnl.copy(), np.array([e.root_freqs(k) for k in range(len(pts))]), st["factored_points"], st["table_points"])
        e.close()
    np.testing.assert_allclose(out["1"][0], out["0"][0], rtol=1e-9)
    np.testing.assert_allclose(out["1"][1], out["0"][1], rtol=1e-8, atol=1e-12)
    assert out["1"][2] >= 1 and out["1"][3] >= 1 and out["1"][2] + out["1"][3] == len(pts)   # both routes inside one batch
    assert out["0"][2] == 0
    for k, m in enumerate(pts):
        res = cases.oracle_eval(c, model=m, weighted_root=True, brlen=c.flat.brlen * (1.0 + 0.1 * k))
        assert abs(out["1"][0][k] - res.lnl) <= REL * abs(res.lnl), k


@pytest.mark.parametrize("S,ntaxa", [(52, 40), (200, 60)])
def test_slab_streamed_chromosome_kernels_equal_the_fragment_loaded_ones(monkeypatch, S, ntaxa):
    """chr_level_slab_kernel / chr_chain_slab_kernel (A through the TMA slab ring, 7 consumer warps) accumulate over k in the
    order of chr_level_kernel / chr_chain_kernel (A fragments from L2): log L and the root frequencies in use are bit-identical."""
    capi = _capi()
    rng = np.random.default_rng(S + 1)
    r, p = rm.constant_rate()
    pts = []
    while len(pts) < 6:
        m = rm.chromosome(1, S, gain=rng.uniform(0.2, 2), loss=rng.uniform(0.2, 2), dupl=rng.uniform(0.05, 1), demi=rng.uniform(0.05, 1))
        if m.nonsingular and np.linalg.cond(m.V) < 1e7:
            pts.append(m)
    c = cases.make_case(ntaxa, 1, pts[0], r, p, seed=78, rooted=True, mean_brlen=0.1, compress=False)
    off, ch = c.flat.csr()
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("BPPGPU_CHR_SLAB", mode)
        e = capi.Engine(S, 1, 1, off, ch, c.flat.root, c.table, n_points=len(pts), n_models=len(pts), flags=capi.FLAG_WEIGHTED_ROOT,
                        code_bytes=np.dtype(c.code_dtype).itemsize)
        for lid, codes in c.codes_by_leaf.items():
            e.set_tip_codes(lid, codes)
        e.set_pattern_weights(c.weights)
        e.set_rates(r, p)
        holders = [cases.to_model_desc(m) for m in pts]
        for k, h in enumerate(holders):
            e.set_model(k, h)
            e.set_branch_lengths(k, c.flat.brlen * (1.0 + 0.07 * k))
        lnl, _, _ = e.eval()
        assert e.stats()["factored_points"] == len(pts)
        out[mode] = (lnl.copy(), np.array([e.root_freqs(k) for k in range(len(pts))]))
        e.close()
    np.testing.assert_array_equal(out["1"][0], out["0"][0])
    np.testing.assert_array_equal(out["1"][1], out["0"][1])


def test_batched_points_path_general_shapes():
    """n_points > 1 on a small pattern set: per-point models AND branch lengths, several patterns and classes,
    ambiguity codes, rescaling, device-resident CLVs of a point."""
    capi = _capi()
    rng = np.random.default_rng(17)
    r, p = rm.gamma_rates(2, 0.8)
    base = rm.lg08()
    c = cases.make_case(110, 9, base, r, p, seed=71, mean_brlen=0.6, ambiguity=0.05, compress=False)
    npts = 7
    off, ch = c.flat.csr()
    models = [rm.lg08()] + [rm._reversible("R%d" % k, np.triu(rng.gamma(0.6, 1.0, (20, 20)), 1) + np.triu(rng.gamma(0.6, 1.0, (20, 20)), 1).T
                                           if False else (lambda a: (a + a.T) / 2)(rng.gamma(0.6, 1.0, (20, 20)) + 1e-3),
                                           rng.dirichlet(np.full(20, 5.0))) for k in range(npts - 1)]
    brl = [c.flat.brlen * rng.uniform(0.5, 1.5, size=c.flat.n_nodes) for _ in range(npts)]
    e = capi.Engine(20, 2, c.N, off, ch, c.flat.root, c.table, n_points=npts, n_models=npts)
    for lid, codes in c.codes_by_leaf.items():
        e.set_tip_codes(lid, codes)
    e.set_pattern_weights(c.weights)
    e.set_rates(r, p)
    holders = [cases.to_model_desc(m) for m in models]
    for k in range(npts):
        e.set_model(k, holders[k])
        e.set_branch_lengths(k, brl[k])
        e.set_root_freqs(k, models[k].freq)
    lnl, _, _ = e.eval()
    assert e.stats()["path"] == 5
    for k in range(npts):
        cc = cases.Case()
        cc.__dict__.update(c.__dict__)
        cc.root_freqs = models[k].freq
        res = cases.oracle_eval(cc, model=models[k], brlen=brl[k])
        assert abs(lnl[k] - res.lnl) <= REL * abs(res.lnl), k
        np.testing.assert_allclose(e.site_lnl(k), res.site_lnl, rtol=1e-11)
        assert res.SR_exp.max() > 0          # rescaling really happened on this tree
    clv, ex = e.clv(c.flat.root, 0, point=npts - 1)
    np.testing.assert_array_equal(ex, res.lexp[c.flat.root])
    np.testing.assert_allclose(clv, res.lower[c.flat.root], rtol=1e-10, atol=1e-14 * res.lower[c.flat.root].max())
    with pytest.raises(capi.BppGpuError):
        e.eval(7)                            # no branch derivatives on this path
    e.close()


def test_odd_class_counts_two_byte_codes_and_impossible_sites():
    capi = _capi()
    # C = 3 (not a power of two) -> generic kernels for DNA
    r, p = rm.gamma_rates(3, 0.9)
    c = cases.make_case(20, 120, gtr(), r, p, seed=81, ambiguity=0.05)
    st = check_value(c)
    assert st["path"] == 3
    # C = 8 on the DNA walk
    r8, p8 = rm.gamma_rates(8, 0.4)
    c = cases.make_case(33, 257, gtr(), r8, p8, seed=82)
    assert check_value(c)["path"] == 1
    # more than 256 distinct tip codes -> 2-byte codes (composite states carry probabilities, like chromosome counts)
    rng = np.random.default_rng(83)
    r4, p4 = rm.gamma_rates(4, 0.7)
    c = cases.make_case(10, 90, rm.lg08(), r4, p4, seed=83)
    extra = rng.dirichlet(np.ones(20), size=300)
    c.table = np.vstack([c.table, extra])
    for lid in c.codes_by_leaf:
        codes = c.codes_by_leaf[lid].astype(np.uint16)
        mask = rng.random(c.N) < 0.3
        codes[mask] = rng.integers(22, 322, size=int(mask.sum()))
        c.codes_by_leaf[lid] = codes
    c.code_dtype = np.uint16
    st = check_value(c)
    assert st["path"] == 4
    # a site that is impossible under the model (all-zero tip vector): log L = -inf like the reference
    # ("Likelihood will be 0 for site", DRASRTreeLikelihoodData.cpp:305-306), the other sites are unaffected
    c = cases.make_case(8, 30, gtr(), r4, p4, seed=84, compress=False)
    c.table = np.vstack([c.table, np.zeros((1, 4))])
    first = c.flat.leaf_ids[0]
    c.codes_by_leaf[first] = c.codes_by_leaf[first].copy()
    c.codes_by_leaf[first][5] = c.table.shape[0] - 1
    res = cases.oracle_eval(c)
    with cases.make_engine(c) as e:
        lnl, _, _ = e.eval()
        site = e.site_lnl()
    assert lnl[0] == -np.inf and res.lnl == -np.inf
    assert site[5] == -np.inf
    keep = np.arange(c.N) != 5
    np.testing.assert_allclose(site[keep], res.site_lnl[keep], rtol=1e-11)


def test_error_conventions():
    capi = _capi()
    r, p = rm.gamma_rates(4, 0.5)
    c = cases.make_case(6, 10, gtr(), r, p, seed=1)
    off, ch = c.flat.csr()
    e = capi.Engine(4, 4, c.N, off, ch, c.flat.root, c.table)
    with pytest.raises(capi.BppGpuError) as ei:
        e.eval()
    assert ei.value.code == capi.E_STATE
    with pytest.raises(capi.BppGpuError) as ei:
        e.set_tip_codes(c.flat.root, np.zeros(c.N, np.uint8))
    assert ei.value.code == capi.E_INVALID
    e.close()
    with pytest.raises(capi.BppGpuError):
        capi.Engine(4, 4, 10, np.array([0, 0, 0, 2], np.int32), np.array([0, 0], np.int32), 2, c.table)


def test_node_posteriors_error_conventions():
    """bppgpu_get_node_posteriors: needs kept CLVs, needs the prefix pass for non-root nodes (the root works after a value-only
    evaluation), rejects bad node ids; a leaf's posteriors need no upper array."""
    capi = _capi()
    r, p = rm.gamma_rates(4, 0.5)
    c = cases.make_case(6, 10, gtr(), r, p, seed=3)
    with cases.make_engine(c) as e:                      # no KEEP_CLVS
        e.eval()
        with pytest.raises(capi.BppGpuError) as ei:
            e.node_posteriors(c.flat.root)
        assert ei.value.code == capi.E_STATE
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS) as e:
        e.eval(capi.EVAL_LNL)
        la, ex, post = e.node_posteriors(c.flat.root)    # root: lower array and root frequencies only
        np.testing.assert_allclose(post.sum(axis=(1, 2)), 1.0, rtol=1e-12)
        inner = next(n for n in range(c.flat.n_nodes - 1) if not c.flat.is_leaf[n])
        with pytest.raises(capi.BppGpuError) as ei:
            e.node_posteriors(inner)
        assert ei.value.code == capi.E_STATE
        _, _, post = e.node_posteriors(c.flat.leaf_ids[0], full=False)
        np.testing.assert_allclose(post.sum(axis=(1, 2)), 1.0, rtol=1e-12)
        with pytest.raises(capi.BppGpuError) as ei:
            e.node_posteriors(c.flat.n_nodes)
        assert ei.value.code == capi.E_INVALID


@pytest.mark.parametrize("mk,ncat,rooted,nsites", [
    (gtr, 4, True, 70), (rm.lg08, 4, False, 40), (rm.lg08, 2, True, 40), (lambda: rm.yn98(2.0, 0.3), 1, True, 12),
    (lambda: rm.chromosome(1, 40, gain=0.7, loss=0.4, dupl=0.2, demi=rm.DEMI_EQUAL_DUPL), 1, True, 1),
    (lambda: rm.chromosome(1, 25, gain=1.5, loss=0.1, dupl=0.9, demi=0.4, gain_r=0.05), 2, True, 3)])
def test_marginal_nonrev_ancestral_posteriors(mk, ncat, rooted, nsites):
    """MarginalNonRevAncestralStateReconstruction (fork): node posteriors and (node, father) joint posteriors of every node
    from the device-resident arrays against the oracle (itself held to the reference's S-pass form and to brute-force
    enumeration on CPU); non-stationary root frequencies so that the non-reversible case is exercised."""
    capi = _capi()
    from oracle import ref_likelihood as rl
    r, p = rm.gamma_rates(ncat, 0.6) if ncat > 1 else rm.constant_rate()
    m = mk()
    c = cases.make_case(9, nsites, m, r, p, seed=23, rooted=rooted, ambiguity=0.05, compress=False,
                        mean_brlen=0.2 if m.size <= 64 else 0.05)
    rng = np.random.default_rng(5)
    c.root_freqs = rng.dirichlet(np.ones(m.size))
    res = cases.oracle_eval(c, want_d1=True)
    flat = c.flat
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS) as e:
        lnl, _, _ = e.eval(capi.EVAL_LNL | capi.EVAL_D1)
        assert abs(lnl[0] - res.lnl) <= REL * abs(res.lnl)
        posts = {}
        for nid in range(flat.n_nodes):
            is_root = nid == flat.root
            post, joint = e.marginal_posteriors(nid, joint=not is_root)
            want_post, want_joint = rl.marginal_posteriors(flat, res, res.P, nid, c.probs)
            np.testing.assert_allclose(post, want_post, rtol=1e-9, atol=1e-14)
            np.testing.assert_allclose(post.sum(axis=1), 1.0, rtol=1e-9)
            posts[nid] = post
            if not is_root:
                np.testing.assert_allclose(joint, want_joint, rtol=1e-9, atol=1e-14)
                post_only, none = e.marginal_posteriors(nid, joint=False)
                assert none is None
                np.testing.assert_array_equal(post_only, post)
        # marginalising the node's state out of the joint table gives the father's posterior
        for nid in range(flat.n_nodes - 1):
            _, joint = e.marginal_posteriors(nid)
            np.testing.assert_allclose(joint.sum(axis=1), posts[int(flat.parent[nid])], rtol=1e-8, atol=1e-13)


@pytest.mark.parametrize("mk,ncat,rooted,ntaxa,nsites,amb", [
    (gtr, 1, True, 12, 80, 0.05), (gtr, 4, False, 10, 50, 0.0), (rm.lg08, 1, True, 9, 40, 0.03),
    (lambda: rm.chromosome(1, 40, gain=0.7, loss=0.4, dupl=0.2, demi=rm.DEMI_EQUAL_DUPL), 1, True, 14, 2, 0.0),
    (lambda: rm.chromosome(1, 70, gain=1.5, loss=0.1, dupl=0.9, demi=0.4, gain_r=0.05), 1, True, 9, 1, 0.0)])
def test_ml_joint_ancestral_reconstruction(mk, ncat, rooted, ntaxa, nsites, amb):
    """MLAncestralStateReconstruction (fork): the device's max-product pass and trace back against the oracle (held to brute-force
    enumeration on CPU).  Ties are compared through the value: the joint likelihood of the returned assignment and, where the
    maximiser is unique, the states themselves."""
    capi = _capi()
    from oracle import ref_likelihood as rl
    r, p = rm.gamma_rates(ncat, 0.6) if ncat > 1 else rm.constant_rate()
    m = mk()
    c = cases.make_case(ntaxa, nsites, m, r, p, seed=37, rooted=rooted, ambiguity=amb, compress=False,
                        mean_brlen=0.2 if m.size <= 20 else 0.05)
    c.root_freqs = np.random.default_rng(6).dirichlet(np.ones(m.size) * 2)
    res = cases.oracle_eval(c)
    want, Lroot = rl.ml_joint_reconstruction(c.flat, c.codes_by_leaf, c.table, res.P, c.root_freqs)
    with cases.make_engine(c) as e:
        e.eval(capi.EVAL_LNL)
        got, best = e.ml_ancestral_states()
    np.testing.assert_allclose(best, np.log(Lroot[:, 0, :].max(axis=1)), rtol=1e-10, atol=1e-10)
    np.testing.assert_array_equal(got, want)


def test_ml_joint_reconstruction_does_not_underflow_on_a_large_tree():
    """700 taxa, long branches: the reference's unscaled arrays are all zero here (its reconstruction degenerates to state 0
    everywhere); the device rows carry exponents, so the leaves come back as observed and the joint likelihood is finite."""
    capi = _capi()
    r, p = rm.constant_rate()
    c = cases.make_case(700, 3, gtr(), r, p, seed=4, random_tips=True, mean_brlen=0.5, compress=False)
    with cases.make_engine(c) as e:
        lnl, _, _ = e.eval(capi.EVAL_LNL)
        got, best = e.ml_ancestral_states()
    assert np.all(np.isfinite(best)) and np.all(best < -708)          # below the smallest double: the reference returns zeros
    for l in c.flat.leaf_ids:
        np.testing.assert_array_equal(got[l], c.codes_by_leaf[l])


def test_marginal_posteriors_error_conventions():
    """bppgpu_get_marginal_posteriors: needs kept CLVs and an evaluation, the prefix pass for non-root nodes, no joint table at
    the root, valid node ids."""
    capi = _capi()
    r, p = rm.gamma_rates(2, 0.5)
    c = cases.make_case(6, 20, gtr(), r, p, seed=3, rooted=True)
    with cases.make_engine(c) as e:
        e.eval(capi.EVAL_LNL)
        with pytest.raises(capi.BppGpuError):
            e.marginal_posteriors(c.flat.root, joint=False)
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS) as e:
        with pytest.raises(capi.BppGpuError):
            e.marginal_posteriors(c.flat.root, joint=False)          # no evaluation yet
        e.eval(capi.EVAL_LNL)
        post, _ = e.marginal_posteriors(c.flat.root, joint=False)    # the root needs the value pass only
        np.testing.assert_allclose(post.sum(axis=1), 1.0, rtol=1e-12)
        with pytest.raises(capi.BppGpuError):
            e.marginal_posteriors(c.flat.root, joint=True)
        with pytest.raises(capi.BppGpuError):
            e.marginal_posteriors(0, joint=False)                    # no prefix pass yet
        with pytest.raises(capi.BppGpuError):
            e.marginal_posteriors(c.flat.n_nodes, joint=False)


@pytest.mark.parametrize("mk,ncat", [(gtr, 4), (rm.lg08, 3), (lambda: rm.yn98(2.0, 0.3), 1)])
def test_root_reparametrisation_derivatives(mk, ncat):
    """BrLenRoot / RootPosition (reparametrizeRoot): first and second derivatives rebuilt at the root on the device, against
    the oracle's restatement of DRNonHomogeneousTreeLikelihood.cpp:445-478 / :576-867 (itself checked by finite differences)."""
    capi = _capi()
    from oracle import ref_likelihood as rl
    r, p = rm.gamma_rates(ncat, 0.7) if ncat > 1 else rm.constant_rate()
    m = mk()
    c = cases.make_case(9, 50, m, r, p, seed=95, rooted=True, ambiguity=0.04, mean_brlen=0.2, compress=False)
    rng = np.random.default_rng(96)
    c.root_freqs = rng.dirichlet(np.ones(m.size))            # non-stationary root: the root position matters
    res = cases.oracle_eval(c, want_d1=True, want_d2=True, nh_form=True)
    g = rl.root_reparam_derivatives(c.flat, res, res.P, res.dP, res.d2P, c.probs, c.weights)
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS | capi.FLAG_NH_DERIV) as e:
        e.eval(7)
        out = e.root_reparam_derivatives()
        np.testing.assert_allclose(-out, [g["d1_len"], g["d1_pos"], g["d2_len"], g["d2_pos"]], rtol=1e-8, atol=1e-8)
    with cases.make_engine(c, flags=capi.FLAG_KEEP_CLVS) as e:
        e.eval(capi.EVAL_LNL | capi.EVAL_D1)
        with pytest.raises(capi.BppGpuError) as ei:          # d2P is not resident
            e.root_reparam_derivatives()
        assert ei.value.code == capi.E_STATE
