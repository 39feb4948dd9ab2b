"""CPU oracle for the bpp-phyl tree-likelihood hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference
(anatshafir1/bpp-phyl, Bio++ bpp-phyl 2.4.1 + ChromEvol fork) for the one hot
path this repository accelerates: P(t) construction, Felsenstein pruning,
root reduction and branch-length derivatives.  It exists so the CUDA path can
be checked; nothing in the product (``bpp_phyl_b200``) imports it, links it or
falls back to it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it.

Why a restatement: the reference cannot be compiled here (it needs bpp-core
2.4.1 and the author's forked bpp-seq, neither vendored nor installed; no
network), so ``oracle/_ref`` does not exist.

Parity pinning status
---------------------
* DNA (T92/HKY85 + Gamma4 / constant rate): PINNED against the reference's own
  known-answer tests, ``test/test_likelihood.cpp:108`` (85.030942031997312824)
  and ``test/test_likelihood_clock.cpp:115`` (94.3957), and against its R-vs-DR
  derivative identity (``test/test_likelihood.cpp:129-135``).
  See tests/test_oracle_golden.py.
* Protein (LG08), codon (YN98/GY94) and Chromosome likelihood VALUES:
  "parity unpinned" -- the reference holds no known-answer test for them
  (SURVEY.md section 4).  The restatement is validated structurally instead
  (row sums, detailed balance, brute-force enumeration over internal states
  on small trees, scipy ``expm``, finite differences).

Modules
-------
ref_models      generators / eigensystems / P(t) family / rate categories
ref_tree        Newick, unrooting, post-order ``BrLen`` indexing
ref_patterns    SitePatterns + recursive per-subtree compression
ref_likelihood  R (single) and DR (double) recursion, root reduction, d1/d2
ref_cpu.cpp     C++ twin keeping the reference's nested-vector layout and loop
                nest; this is what is timed as the CPU baseline
"""
