"""Oracle (TEST INFRASTRUCTURE): alignment -> site patterns (bit-exact integer work).

Restates (paths relative to /root/reference/src/Bpp/Phyl/):
* SitePatterns::SitePatterns                 SitePatterns.cpp:52-106 (+ SortableSite::operator< SitePatterns.h:94)
* PatternTools::getSequenceSubset            PatternTools.cpp:59-70
* DRASRTreeLikelihoodData::initLikelihoodsWithPatterns  Likelihood/DRASRTreeLikelihoodData.cpp:218-332
* DRASDRTreeLikelihoodData::initLikelihoods  Likelihood/DRASDRTreeLikelihoodData.cpp:51-206
* AbstractTransitionModel::getInitValue      Model/AbstractSubstitutionModel.cpp:98-112

A column's sort key is ``Site::toString()`` (bpp-seq): the concatenation of the
per-sequence character strings in container order, compared as byte strings.
"""
from __future__ import annotations

import numpy as np

DNA_STATES = "ACGT"
PROTEIN_STATES = "ARNDCQEGHILKMFPSTWYV"      # Bio++ ProteicAlphabet order

# bpp-seq DNA alphabet aliases (IUPAC); value = set of resolved states
DNA_ALIASES = {
    "A": "A", "C": "C", "G": "G", "T": "T", "U": "T",
    "M": "AC", "R": "AG", "W": "AT", "S": "CG", "Y": "CT", "K": "GT",
    "V": "ACG", "H": "ACT", "D": "AGT", "B": "CGT",
    "N": "ACGT", "X": "ACGT", "O": "ACGT", "0": "ACGT", "?": "ACGT", "-": "ACGT",
}
PROTEIN_ALIASES = {c: c for c in PROTEIN_STATES}
PROTEIN_ALIASES.update({"B": "ND", "Z": "QE", "J": "IL", "X": PROTEIN_STATES, "O": PROTEIN_STATES,
                        "0": PROTEIN_STATES, "?": PROTEIN_STATES, "-": PROTEIN_STATES})


def site_patterns(columns):
    """SitePatterns::SitePatterns.  ``columns`` = list of per-site byte strings.
    Returns (unique_cols sorted lexicographically, weights u32, indices site->pattern).
    std::sort is not stable, but equal keys are identical columns, so the result
    (patterns, weights, indices) does not depend on the order among equals."""
    n = len(columns)
    if n == 0:
        return [], np.zeros(0, np.uint32), np.zeros(0, np.int64)
    order = sorted(range(n), key=lambda i: columns[i])
    uniq = [columns[order[0]]]
    weights = [1]
    indices = np.zeros(n, np.int64)
    indices[order[0]] = 0
    for k in range(1, n):
        c = columns[order[k]]
        if c == uniq[-1]:
            weights[-1] += 1
        else:
            uniq.append(c)
            weights.append(1)
        indices[order[k]] = len(uniq) - 1
    return uniq, np.array(weights, np.uint32), indices


def columns_from_sequences(seqs, width=1):
    """Column strings in container order; ``width`` chars per state (3 for codons)."""
    n = len(seqs[0]) // width
    return [b"".join(s[i * width:(i + 1) * width].encode() for s in seqs) for i in range(n)]


def global_patterns(seq_by_name, leaf_names, width=1):
    """DR layout (DRASDRTreeLikelihoodData::initLikelihoods :51-72): sequences re-ordered
    to the tree's leaf order (getSequenceSubset), then one global compression."""
    seqs = [seq_by_name[n] for n in leaf_names]
    cols = columns_from_sequences(seqs, width)
    uniq, w, idx = site_patterns(cols)
    return uniq, w, idx


def recursive_patterns(flat, seq_by_name, width=1):
    """R layout (DRASRTreeLikelihoodData::initLikelihoodsWithPatterns :218-332).

    Returns dict with, per node id: ``cols`` (that subtree's unique columns over its
    own leaves), and per internal node ``links[son] = indices`` (father pattern ->
    son pattern, :323-325); plus root ``weights`` and ``root_links`` (site -> root pattern).
    """
    from .ref_tree import leaves
    out = {"cols": {}, "links": {}, "n": {}}

    def rec(node, names, cols):
        # names: sequence order of the incoming container; cols: its columns
        lv = [l.name for l in leaves(node)]
        pos = [names.index(nm) for nm in lv]                       # getSequenceSubset
        sub = [b"".join(c[p * width:(p + 1) * width] for p in pos) for c in cols]
        uniq, w, idx = site_patterns(sub)
        out["cols"][node.id] = uniq
        out["n"][node.id] = len(uniq)
        if node.sons:
            out["links"][node.id] = {}
            for s in node.sons:
                _, _, sidx = rec(s, lv, uniq)
                out["links"][node.id][s.id] = sidx
        return uniq, w, idx

    root = flat.root_node
    names0 = list(seq_by_name.keys())
    seqs0 = [seq_by_name[n] for n in names0]
    uniq, w, idx = rec(root, names0, columns_from_sequences(seqs0, width))
    out["weights"] = w
    out["root_links"] = idx
    return out


def init_value_table(states: str, aliases: dict):
    """Code table for tips: code k <-> character; table[k][s] = getInitValue(s, code)
    (1 if model state s is in alphabet->getAlias(code), :98-112).  Returns
    (chars, table[n_codes][S])."""
    chars = list(aliases.keys())
    S = len(states)
    tab = np.zeros((len(chars), S))
    for k, ch in enumerate(chars):
        for r in aliases[ch]:
            tab[k, states.index(r)] = 1.0
    return chars, tab


def encode_columns(uniq_cols, chars, width=1):
    """[n_tips][N] code array from unique column strings (tip order = column order)."""
    N = len(uniq_cols)
    if N == 0:
        return np.zeros((0, 0), np.uint8)
    ntip = len(uniq_cols[0]) // width
    lut = {c.encode() if isinstance(c, str) else c: k for k, c in enumerate(chars)}
    codes = np.zeros((ntip, N), np.uint16 if len(chars) > 256 else np.uint8)
    for i, col in enumerate(uniq_cols):
        for t in range(ntip):
            codes[t, i] = lut[col[t * width:(t + 1) * width]]
    return codes
