"""Shared builders for the parity tests: a *case* = tree + model + rate classes + tip codes,
expressed once for the oracle (numpy arrays) and once for the C ABI (an Engine).

Test infrastructure: uses oracle/ (allowed in tests/, smoke() and bench.py's cpu_baseline only).
"""
from __future__ import annotations

import numpy as np

from oracle import ref_likelihood as rl
from oracle import ref_models as rm
from oracle import ref_patterns as rp
from oracle import ref_tree as rt


class Case:
    pass


def model_flags(m: rm.Model):
    from bpp_phyl_b200 import capi
    f = 0
    if m.diagonalizable:
        f |= capi.MODEL_DIAGONALIZABLE
    if m.nonsingular:
        f |= capi.MODEL_NONSINGULAR
    if m.chromosome:
        f |= capi.MODEL_CLAMP01 | capi.MODEL_CHR_DERIV | capi.MODEL_CHR_TAYLOR
    return f


def to_model_desc(m: rm.Model, extra_flags=0):
    from bpp_phyl_b200 import capi
    return capi.model_desc(m.size, model_flags(m) | extra_flags, rate=m.rate, V=m.V, Vinv=m.Vinv,
                           ev_re=m.ev_re, ev_im=m.ev_im, Q=m.Q)


def simulate_tips(flat: rt.FlatTree, m: rm.Model, rates, n_sites, rng, root_freqs=None):
    """Sample states down the tree (one rate class per site); returns [n_leaves(pre-order)][n_sites] states."""
    S = m.size
    pi = m.freq if root_freqs is None else root_freqs
    pi = np.asarray(pi, float)
    pi = pi / pi.sum()
    cls = rng.integers(len(rates), size=n_sites)
    st = {flat.root: rng.choice(S, size=n_sites, p=pi)}
    Pc = {}
    for nid in range(flat.n_nodes - 2, -1, -1):       # fathers have larger post-order ids
        f = int(flat.parent[nid])
        out = np.empty(n_sites, np.int64)
        for c, r in enumerate(rates):
            P = np.clip(rm.pij_t(m, flat.brlen[nid] * r), 0, None)
            P = P / P.sum(axis=1, keepdims=True)
            cum = np.cumsum(P, axis=1)
            sel = np.where(cls == c)[0]
            u = rng.random(len(sel))
            out[sel] = (u[:, None] > cum[st[f][sel]]).sum(axis=1).clip(0, S - 1)
        st[nid] = out
    return np.stack([st[l] for l in flat.leaf_ids])


def make_case(n_taxa, n_sites, model: rm.Model, rates, probs, seed, rooted=False, mean_brlen=0.05,
              random_tips=False, ambiguity=0.0, compress=True):
    """Random tree + simulated (or i.i.d.) tip states, globally compressed (DR layout)."""
    rng = np.random.default_rng(seed)
    root = rt.random_tree(n_taxa, rng, mean_brlen=mean_brlen, rooted=rooted)
    flat = rt.FlatTree(root, check_rooted=not rooted)
    S = model.size
    if random_tips:
        tips = rng.integers(S, size=(n_taxa, n_sites))
    else:
        tips = simulate_tips(flat, model, rates, n_sites, rng)
    # code table: S plain states + one fully ambiguous code + (S>=4) one two-state ambiguity
    table = np.eye(S)
    table = np.vstack([table, np.ones((1, S))])
    amb2 = np.zeros((1, S))
    amb2[0, :2] = 1.0
    table = np.vstack([table, amb2])
    if ambiguity > 0:
        mask = rng.random(tips.shape) < ambiguity
        tips = np.where(mask, rng.integers(S, S + 2, size=tips.shape), tips)
    code_dtype = np.uint8 if table.shape[0] <= 256 else np.uint16
    cols = np.ascontiguousarray(tips.T.astype(code_dtype))             # [site][leaf]
    if compress:
        keys = [c.tobytes() for c in cols]
        uniq, w, idx = rp.site_patterns(keys)
        pat = np.frombuffer(b"".join(uniq), dtype=code_dtype).reshape(len(uniq), n_taxa)
    else:
        pat, w, idx = cols, np.ones(len(cols), np.uint32), np.arange(len(cols))
    c = Case()
    c.flat, c.model, c.rates, c.probs = flat, model, np.asarray(rates, float), np.asarray(probs, float)
    c.table = table
    c.N = pat.shape[0]
    c.weights = w
    c.site_index = idx
    c.codes_by_leaf = {lid: np.ascontiguousarray(pat[:, k]) for k, lid in enumerate(flat.leaf_ids)}
    c.code_dtype = code_dtype
    c.root_freqs = np.asarray(model.freq, float)
    return c


def case_from_alignment(newick, seqs, model, rates, probs, check_rooted=True, states=rp.DNA_STATES,
                        aliases=rp.DNA_ALIASES):
    """The reference tests' way: Newick string + named sequences (test/test_likelihood.cpp:91-103)."""
    flat = rt.FlatTree(rt.parse_newick(newick), check_rooted=check_rooted)
    uniq, w, idx = rp.global_patterns(seqs, flat.leaf_names)
    chars, table = rp.init_value_table(states, aliases)
    codes = rp.encode_columns(uniq, chars)
    c = Case()
    c.flat, c.model, c.rates, c.probs = flat, model, np.asarray(rates, float), np.asarray(probs, float)
    c.table = table
    c.N = len(uniq)
    c.weights = w
    c.site_index = idx
    c.codes_by_leaf = {lid: codes[k] for k, lid in enumerate(flat.leaf_ids)}
    c.code_dtype = codes.dtype
    c.root_freqs = np.asarray(model.freq, float)
    c.patterns = uniq
    return c


def oracle_eval(c: Case, want_d1=False, want_d2=False, scaled=True, nh_form=False, weighted_root=False,
                brlen=None, model=None):
    m = c.model if model is None else model
    bl = c.flat.brlen if brlen is None else brlen
    P, dP, d2P = rm.transition_tables(m, bl, c.rates, want_d1 or want_d2, want_d2)
    res = rl.dr_eval(c.flat, c.codes_by_leaf, c.table, P, len(c.rates), c.root_freqs, c.probs,
                     c.weights.astype(float), dP=dP, d2P=d2P, scaled=scaled, nh_form=nh_form,
                     weighted_root=weighted_root)
    res.P, res.dP, res.d2P = P, dP, d2P
    return res


def oracle_eval_nh(c: Case, models, slot_of_node, want_d1=False, want_d2=False, nh_form=True, scaled=True):
    """Non-homogeneous evaluation: the branch above node n uses models[slot_of_node[n]]
    (AbstractNonHomogeneousTreeLikelihood::computeTransitionProbabilitiesForNode, .cpp:410-468: the node's own model from the
    SubstitutionModelSet); root frequencies are c.root_freqs (the set's root FrequencySet)."""
    tabs = [rm.transition_tables(m, c.flat.brlen, c.rates, want_d1 or want_d2, want_d2) for m in models]
    nb = len(c.flat.brlen)
    pick = lambda k: None if tabs[0][k] is None else np.stack([tabs[slot_of_node[n]][k][n] for n in range(nb)])
    P, dP, d2P = pick(0), pick(1), pick(2)
    res = rl.dr_eval(c.flat, c.codes_by_leaf, c.table, P, len(c.rates), c.root_freqs, c.probs, c.weights.astype(float),
                     dP=dP, d2P=d2P, scaled=scaled, nh_form=nh_form)
    res.P, res.dP, res.d2P = P, dP, d2P
    return res


def make_engine(c: Case, flags=0, n_points=1, n_models=1, device=0):
    from bpp_phyl_b200 import capi
    off, ch = c.flat.csr()
    e = capi.Engine(c.model.size, len(c.rates), c.N, off, ch, c.flat.root, c.table, n_points=n_points,
                    n_models=n_models, code_bytes=np.dtype(c.code_dtype).itemsize, device=device, flags=flags)
    for lid, codes in c.codes_by_leaf.items():
        e.set_tip_codes(lid, codes)
    e.set_pattern_weights(c.weights)
    e.set_rates(c.rates, c.probs)
    md = to_model_desc(c.model)
    e._model_holders = [md]
    e.set_model(0, md)
    for p in range(n_points):
        e.set_branch_lengths(p, c.flat.brlen)
        e.set_root_freqs(p, c.root_freqs)
    return e


# ---- the one rule for value parity on (nearly) defective chromosome generators --------------------------------------------
# Both the reference and this repo pick between two exponentiation routes per model (ChromosomeSubstitutionModel.cpp:686-767):
# V exp(D t) V^-1 when the eigensystem passes the reference's checks, else the truncated Taylor series (:852-899).  Close to a
# defective Q the eigenvector matrix is ill conditioned, the checks depend on rounding inside the eigen-solver (bpp-core's JAMA port
# there, LAPACK in the oracle, a Hessenberg-QR in the shim -- which additionally refuses an eigensystem that does not reproduce Q to
# 1e-8) and the two sides may take DIFFERENT routes.  The rule:
#   (1) same route on both sides            -> 1e-9 relative, as everywhere else;
#   (2) different routes                    -> the oracle is re-run on the device's route and must agree to 1e-9 (same algorithm),
#       and the device must be at least as close as the oracle's own route to the exact value (mpmath expm, 50 digits):
#       |device - exact| <= max(1e-9 |exact|, |oracle_own_route - exact|).
# Measured on the point that prompted it (gain 5, loss .5, dupl .1, demi 1, S = 30, cond(V) = 1.9e14): eigen route 1.4e-8 from
# exact (its P is off by 1e-3), series route 5.6e-10 from exact.
def exact_expm_tables(Q, brlen, rates, dps=50):
    import mpmath as mp
    old = mp.mp.dps
    mp.mp.dps = dps
    try:
        Qm = mp.matrix(np.asarray(Q, float).tolist())
        S = Qm.rows
        out = np.zeros((len(brlen), len(rates), S, S))
        for n, t in enumerate(brlen):
            for c, r in enumerate(rates):
                E = mp.expm(Qm * (mp.mpf(float(t)) * mp.mpf(float(r))))
                out[n, c] = [[float(E[i, j]) for j in range(S)] for i in range(S)]
        return out
    finally:
        mp.mp.dps = old


def chromosome_value_parity(c: Case, model: rm.Model, device_lnl, device_nonsingular, weighted_root=True):
    """Apply the rule above; returns (route_of_the_device, rel. distance device<->oracle on that route)."""
    import copy
    own = oracle_eval(c, weighted_root=weighted_root, model=model)
    if bool(model.nonsingular) == bool(device_nonsingular):
        rel = abs(device_lnl - own.lnl) / abs(own.lnl)
        assert rel <= 1e-9, ("same route", rel)
        return ("eigen" if model.nonsingular else "series"), rel
    forced = copy.copy(model)
    forced.nonsingular = bool(device_nonsingular)
    forced.diagonalizable = bool(device_nonsingular) and model.diagonalizable
    same = oracle_eval(c, weighted_root=weighted_root, model=forced)
    rel = abs(device_lnl - same.lnl) / abs(same.lnl)
    assert rel <= 1e-9, ("device route", rel)
    Pex = exact_expm_tables(model.Q, c.flat.brlen, c.rates)
    exact = rl.dr_eval(c.flat, c.codes_by_leaf, c.table, Pex, len(c.rates), c.root_freqs, c.probs, c.weights.astype(float),
                       weighted_root=weighted_root).lnl
    assert abs(device_lnl - exact) <= max(1e-9 * abs(exact), abs(own.lnl - exact)), (device_lnl, own.lnl, exact)
    return ("eigen" if device_nonsingular else "series"), rel
