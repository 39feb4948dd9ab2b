"""Site-pattern compression (bit-exact integer work): the product's host routine against the oracle's restatement of
SitePatterns::SitePatterns (SitePatterns.cpp:52-106): CPU only, no device is needed for that entry point.  GPU part:
bppgpu_site_patterns_device (radix sort of the columns on the device + tip-code extraction) returns the same arrays bit for bit."""
import numpy as np
import pytest

from oracle import ref_patterns as rp


def _cols(seqs):
    return np.array([[ord(s[i]) for s in seqs] for i in range(len(seqs[0]))], np.uint8).reshape(len(seqs[0]), len(seqs))


def check(cols, built_lib):
    from bpp_phyl_b200 import capi
    keys = [bytes(c) for c in cols]
    uniq, w, idx = rp.site_patterns(keys)
    ps, w2, idx2 = capi.site_patterns(cols)
    assert [keys[i] for i in ps] == uniq                  # same patterns in the same (lexicographic) order
    np.testing.assert_array_equal(w2, w)
    np.testing.assert_array_equal(idx2, idx)
    assert int(w2.sum()) == len(cols)


def test_reference_alignment_patterns(built_lib):
    seqs = ["AAATGGCTGTGCACGTC", "GACTGGATCTGCACGTC", "CTCTGGATGTGCACGTG", "AAATGGCGGTGCGCCTA"]
    cols = _cols(seqs)
    check(cols, built_lib)
    from bpp_phyl_b200 import capi
    ps, w, idx = capi.site_patterns(cols)
    assert ["".join(chr(x) for x in cols[i]) for i in ps] == ["AAAG", "AATA", "ACCA", "AGCA", "CAAC", "CCCC", "CCGA",
                                                               "GCGG", "GGGC", "GGGG", "TTTG", "TTTT"]


@pytest.mark.parametrize("n,ntaxa,nstates,seed", [(1, 3, 4, 0), (2, 1, 2, 1), (500, 5, 2, 2), (5000, 7, 4, 3),
                                                   (3000, 40, 20, 4), (257, 300, 4, 5)])
def test_random_alignments(built_lib, n, ntaxa, nstates, seed):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTRYKMSWBDHVN-?XQZ", np.uint8)[:nstates]
    cols = alphabet[rng.integers(nstates, size=(n, ntaxa))]
    check(cols, built_lib)


def test_empty_and_all_identical(built_lib):
    from bpp_phyl_b200 import capi
    ps, w, idx = capi.site_patterns(np.zeros((0, 4), np.uint8))
    assert len(ps) == 0 and len(w) == 0
    cols = np.full((100, 6), ord("A"), np.uint8)
    ps, w, idx = capi.site_patterns(cols)
    assert list(w) == [100] and list(ps) == [0] and not idx.any()


def test_device_entry_has_no_cpu_fallback(built_lib):
    """Without a GPU the device routine refuses (E_CUDA) instead of silently running the host one."""
    import torch
    from bpp_phyl_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.BppGpuError) as ei:
        capi.site_patterns_device(np.zeros((4, 3), np.uint8))
    assert ei.value.code == capi.E_CUDA


def check_device(cols, code_bytes=1):
    from bpp_phyl_b200 import capi
    raw = cols.view(np.uint8) if cols.dtype == np.uint16 else cols
    ps, w, idx = capi.site_patterns(raw)                        # the host routine (itself held to the oracle above)
    ps2, w2, idx2, tips = capi.site_patterns_device(cols, code_bytes=code_bytes)
    np.testing.assert_array_equal(ps2, ps)
    np.testing.assert_array_equal(w2, w)
    np.testing.assert_array_equal(idx2, idx)
    np.testing.assert_array_equal(tips, cols[ps].T)             # row l = bppgpu_set_tip_codes input of leaf l
    return len(ps)


@pytest.mark.gpu
@pytest.mark.parametrize("n,ntaxa,nstates,seed", [(1, 3, 4, 0), (2, 1, 2, 1), (500, 5, 2, 2), (5000, 7, 4, 3), (3000, 40, 20, 4),
                                                   (257, 300, 4, 5), (4000, 8, 2, 6), (2500, 16, 3, 7), (1000, 33, 4, 8),
                                                   (70000, 64, 4, 9)])
def test_device_patterns_equal_host_patterns(built_lib, n, ntaxa, nstates, seed):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTRYKMSWBDHVN-?XQZ", np.uint8)[:nstates]
    cols = alphabet[rng.integers(nstates, size=(n, ntaxa))]
    check_device(cols)
    # realistic redundancy: few distinct columns, many repeats, bytes above 127 (unsigned comparison like memcmp)
    base = rng.integers(0, 256, size=(max(2, n // 50), ntaxa)).astype(np.uint8)
    cols = base[rng.integers(len(base), size=n)]
    assert check_device(cols) <= len(base)


@pytest.mark.gpu
def test_device_patterns_reference_alignment_two_byte_codes_and_edges(built_lib):
    from bpp_phyl_b200 import capi
    seqs = ["AAATGGCTGTGCACGTC", "GACTGGATCTGCACGTC", "CTCTGGATGTGCACGTG", "AAATGGCGGTGCGCCTA"]
    cols = _cols(seqs)
    ps, w, idx, tips = capi.site_patterns_device(cols)
    assert ["".join(chr(x) for x in cols[i]) for i in ps] == ["AAAG", "AATA", "ACCA", "AGCA", "CAAC", "CCCC", "CCGA",
                                                               "GCGG", "GGGC", "GGGG", "TTTG", "TTTT"]
    assert int(w.sum()) == 17
    rng = np.random.default_rng(3)
    cols16 = rng.integers(0, 700, size=(3000, 21)).astype(np.uint16)          # chromosome-style counts above 255
    cols16[rng.integers(3000, size=1500)] = cols16[rng.integers(3000, size=1500)]
    check_device(cols16)
    ps, w, idx, tips = capi.site_patterns_device(np.zeros((0, 4), np.uint8))
    assert len(ps) == 0 and len(w) == 0
    allsame = np.full((1000, 9), ord("A"), np.uint8)
    ps, w, idx, tips = capi.site_patterns_device(allsame)
    assert list(w) == [1000] and list(ps) == [0] and not idx.any() and tips.shape == (9, 1)


@pytest.mark.gpu
def test_device_patterns_feed_the_engine(built_lib):
    """Ingestion straight to tip codes: the device routine's rows go to bppgpu_set_tip_codes unchanged and the evaluation
    equals the one fed by the oracle's compression."""
    import cases
    from bpp_phyl_b200 import capi
    from oracle import ref_models as rm
    r, p = rm.gamma_rates(4, 0.5)
    m = rm.gtr(1.2, 0.8, 0.6, 1.5, 0.9, (.3, .2, .25, .25))
    c = cases.make_case(12, 3000, m, r, p, seed=5, compress=False)                # one column per site
    cols = np.stack([c.codes_by_leaf[l] for l in c.flat.leaf_ids], axis=1)      # [site][leaf]
    ps, w, idx, tips = capi.site_patterns_device(cols)
    c2 = cases.make_case(12, 3000, m, r, p, seed=5, compress=True)
    assert len(ps) == c2.N
    np.testing.assert_array_equal(w, c2.weights)
    res = cases.oracle_eval(c2)
    c.N, c.weights = len(ps), w
    c.codes_by_leaf = {l: np.ascontiguousarray(tips[k]) for k, l in enumerate(c.flat.leaf_ids)}
    with cases.make_engine(c) as e:
        lnl, _, _ = e.eval(capi.EVAL_LNL)
    assert abs(lnl[0] - res.lnl) <= 1e-9 * abs(res.lnl)


@pytest.mark.gpu
@pytest.mark.parametrize("n,ntaxa,nstates,seed", [(1, 3, 4, 0), (500, 5, 2, 2), (3000, 40, 20, 4), (257, 300, 4, 5), (70000, 64, 4, 9)])
def test_device_patterns_plain_radix_variant(built_lib, monkeypatch, n, ntaxa, nstates, seed):
    """The two device algorithms -- "dedup" (default: merge identical columns by hash first, sort only the unique ones) and "radix"
    (sort every column) -- are bit-identical to the host routine; this test pins the non-default one."""
    monkeypatch.setenv("BPPGPU_PATTERNS_ALGO", "radix")
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTRYKMSWBDHVN-?XQZ", np.uint8)[:nstates]
    check_device(alphabet[rng.integers(nstates, size=(n, ntaxa))])
    base = rng.integers(0, 256, size=(max(2, n // 50), ntaxa)).astype(np.uint8)
    check_device(base[rng.integers(len(base), size=n)])
    check_device(np.full((1000, 9), ord("A"), np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("where", ["pinned", "device"])
def test_device_patterns_on_pinned_and_device_resident_buffers(built_lib, where):
    """The alignment and every output buffer may be pinned host memory or device memory (include/bppgpu.h): same bits as the
    host routine.  (torch only allocates the buffers.)"""
    import torch
    from bpp_phyl_b200 import capi
    rng = np.random.default_rng(11)
    n, taxa = 600_000, 24                                  # > 4 staging chunks, so the pageable path would take the staged copy
    base = rng.integers(0, 4, size=(n // 6, taxa)).astype(np.uint8)
    cols = np.ascontiguousarray(base[rng.integers(len(base), size=n)])
    ps, w, idx = capi.site_patterns(cols)
    make = (lambda t: t.pin_memory()) if where == "pinned" else (lambda t: t.cuda())
    c = make(torch.from_numpy(cols))
    o_ps, o_ix = make(torch.zeros(n, dtype=torch.int64)), make(torch.zeros(n, dtype=torch.int64))
    o_w, o_tip = make(torch.zeros(n, dtype=torch.int32)), make(torch.zeros(n * taxa, dtype=torch.uint8))
    k = capi.site_patterns_device_raw(c.data_ptr(), n, taxa, o_ps.data_ptr(), o_w.data_ptr(), o_ix.data_ptr(), o_tip.data_ptr())
    torch.cuda.synchronize()
    assert k == len(ps)
    np.testing.assert_array_equal(o_ps[:k].cpu().numpy(), ps)
    np.testing.assert_array_equal(o_w[:k].cpu().numpy().view(np.uint32), w)
    np.testing.assert_array_equal(o_ix.cpu().numpy(), idx)
    np.testing.assert_array_equal(o_tip[:k * taxa].cpu().numpy().reshape(taxa, k), cols[ps].T)


@pytest.mark.parametrize("ntaxa,nsites,nstates,seed,width", [(4, 17, 4, 0, 1), (9, 300, 2, 1, 1), (25, 800, 4, 2, 1), (12, 200, 4, 3, 3),
                                                         (6, 0, 4, 4, 1), (2, 50, 3, 5, 1)])
def test_recursive_subtree_patterns_are_bit_exact(built_lib, ntaxa, nsites, nstates, seed, width):
    """SURVEY 8 a4 through the C ABI (host integer routine): bppgpu_subtree_patterns against the oracle's restatement of
    DRASRTreeLikelihoodData::initLikelihoodsWithPatterns (:218-332) -- array length of every node, patternLinks_[father][son] for
    every branch, root indices and weights, all IDENTICAL; container order differs from the tree's leaf order on purpose;
    codon-width elements, an empty alignment and a two-leaf tree included."""
    from bpp_phyl_b200 import capi
    from oracle import ref_tree as rt
    rng = np.random.default_rng(seed)
    root = rt.random_tree(ntaxa, rng, rooted=ntaxa < 3 or bool(seed % 2))
    flat = rt.FlatTree(root, check_rooted=False)
    names = list(flat.leaf_names)
    rng.shuffle(names)                                              # container order != tree order
    alphabet = "ACGT"[:nstates]
    base = ["".join(alphabet[k] for k in rng.integers(nstates, size=width * max(1, nsites // 6))) for _ in names]
    pick = rng.integers(max(1, nsites // 6), size=nsites)            # many repeated columns
    seqs = {nm: "".join(b[width * i:width * (i + 1)] for i in pick) for nm, b in zip(names, base)}
    rec = rp.recursive_patterns(flat, seqs, width=width)
    cols = np.array([[seqs[nm][width * i:width * (i + 1)].encode() for nm in names] for i in range(nsites)],
                    dtype="S%d" % width).reshape(nsites, len(names))
    off, ch = flat.csr()
    leaf_seq = np.full(flat.n_nodes, -1, np.int32)
    for lid, nm in zip(flat.leaf_ids, flat.leaf_names):
        leaf_seq[lid] = names.index(nm)
    npat, links, rl, rw = capi.subtree_patterns(cols, off, ch, flat.root, leaf_seq)
    for n in range(flat.n_nodes):
        assert npat[n] == rec["n"][n], n
    for f, sons in rec["links"].items():
        for s_, idx in sons.items():
            np.testing.assert_array_equal(links[s_], idx)
    np.testing.assert_array_equal(rl, rec["root_links"])
    np.testing.assert_array_equal(rw, rec["weights"])
