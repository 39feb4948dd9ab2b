"""Site-pattern compression (bit-exact integer work): the product's host routine against the oracle's restatement of
SitePatterns::SitePatterns (SitePatterns.cpp:52-106).  CPU only: no device is needed for this entry point."""
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
