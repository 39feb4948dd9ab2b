"""Oracle (TEST INFRASTRUCTURE): pruning, root reduction, branch derivatives.

Vectorised NumPy restatement of (paths relative to
/root/reference/src/Bpp/Phyl/Likelihood/):

* RHomogeneousTreeLikelihood::computeSubtreeLikelihood     RHomogeneousTreeLikelihood.cpp:802-863
* ... getLogLikelihood / ForASite / ForARateClass           :162-216
* ... computeTreeDLikelihood / computeDownSubtreeDLikelihood (+D2)  :365-541, :615-791
* DRHomogeneousTreeLikelihood::computeSubtreeLikelihoodPostfix / Prefix  DRHomogeneousTreeLikelihood.cpp:483-649
* ... computeRootLikelihood :653-719, getLogLikelihood :170-186
* ... computeTreeDLikelihoodAtNode / D2 :287-326, :373-411; reductions :340-368, :425-454
* DRNonHomogeneousTreeLikelihood (per-branch models, NH derivative form :370-413,
  weighted root frequencies :927-962)

Arrays are indexed [pattern][class][state] like the reference's VVVdouble.
``scaled=True`` adds the per-(pattern, class) power-of-two rescaling the GPU path uses
(the reference has none, SURVEY.md finding 5); it is exact, so wherever the
unscaled value is finite both agree to rounding of the final log.
"""
from __future__ import annotations

import numpy as np

SCALE_THRESHOLD_EXP = -256   # rescale a pattern when max CLV < 2^-256 (same rule as the CUDA path)


def _contract(P, L):
    """sum_y P[c][x][y] * L[i][c][y] -> [i][c][x]   (:851-856)."""
    return np.einsum("cxy,icy->icx", P, L)


def _contract_T(P, L):
    """root-side: sum_y P[c][y][x] * L[i][c][y]   (DRHomogeneousTreeLikelihood.cpp:934-940)."""
    return np.einsum("cyx,icy->icx", P, L)


def _rescale(A, E):
    """Power-of-two rescale per (pattern, class) row: A[i][c] *= 2^k, E[i][c] += k when
    0 < max_x A[i][c][x] < 2^-256, k chosen so that the max lands in [0.5, 1)."""
    m = A.max(axis=2)
    need = (m >= 2.0 ** -1022) & (m < 2.0 ** SCALE_THRESHOLD_EXP)
    if need.any():
        _, ex = np.frexp(m[need])           # m = f * 2^ex, f in [0.5,1)
        k = -ex
        A[need] = np.ldexp(A[need], k[:, None])
        E[need] += k
    return A, E


def _align(root_clv, root_e):
    """Bring the C rows of each pattern to the pattern's smallest exponent: returns (clv', Emin[N])."""
    emin = root_e.min(axis=1)
    return np.ldexp(root_clv, -(root_e - emin[:, None])[:, :, None]), emin


def leaf_array(codes_row, table, C):
    """Tip CLV [N][C][S] from codes (DRASRTreeLikelihoodData.cpp:277-303)."""
    L = table[codes_row.astype(np.int64)]            # [N][S]
    return np.repeat(L[:, None, :], C, axis=1)


# ----------------------------------------------------------------------------
# single recursion (R classes)
# ----------------------------------------------------------------------------
def prune(flat, tip_codes, table, P, C, links=None, n_per_node=None, scaled=False, keep=False):
    """computeSubtreeLikelihood.  tip_codes: dict leaf id -> code vector over that leaf's
    own pattern list (recursive mode) or over the global list.  P[b][c][x][y] indexed by
    node id.  links[father][son] = index vector (None -> identity).
    Returns (root CLV [N][C][S], scale exponents [N], {node: (clv, exp)} if keep)."""
    clv = {}
    exps = {}
    for nid in range(flat.n_nodes):                     # post-order ids
        if flat.is_leaf[nid]:
            clv[nid] = leaf_array(tip_codes[nid], table, C)
            exps[nid] = np.zeros(clv[nid].shape[:2], np.int64)
            continue
        A = None
        E = None
        for s in flat.children[nid]:
            Ls, Es = clv[s], exps[s]
            if links is not None:
                idx = links[nid][s]
                Ls, Es = Ls[idx], Es[idx]
            t = _contract(P[s], Ls)
            A = t if A is None else A * t
            E = Es.copy() if E is None else E + Es
            if not keep:
                pass
        if scaled:
            A, E = _rescale(A, E)
        clv[nid], exps[nid] = A, E
    root = flat.root
    if keep:
        return clv[root], exps[root], (clv, exps)
    return clv[root], exps[root], None


def site_likelihoods_R(root_clv, root_freqs, probs):
    """getLikelihoodForASiteForARateClass (:205-216, drops negative terms) and
    getLogLikelihoodForASite (:192-201, drops non-positive class terms)."""
    t = root_clv * root_freqs[None, None, :]
    t = np.where(t > 0, t, 0.0)
    lc = t.sum(axis=2) * probs[None, :]
    lc = np.where(lc > 0, lc, 0.0)
    return lc.sum(axis=1)


def loglik_R(root_clv, root_exp, root_freqs, probs, root_links):
    """getLogLikelihood (:162-176): per SITE (duplicates included) through
    rootPatternLinks_, sorted, summed from the largest."""
    al, emin = _align(root_clv, root_exp)
    with np.errstate(divide="ignore"):
        lp = np.log(site_likelihoods_R(al, root_freqs, probs)) - emin * np.log(2.0)
    la = np.sort(lp[root_links])
    return float(np.sum(la[::-1])), lp


# ----------------------------------------------------------------------------
# double recursion (DR classes), global patterns
# ----------------------------------------------------------------------------
class DRResult:
    pass


def dr_eval(flat, tip_codes, table, P, C, root_freqs, probs, weights,
            dP=None, d2P=None, scaled=False, nh_form=False, weighted_root=False):
    """computeTreeLikelihood (DRHomogeneousTreeLikelihood.cpp:474-479) + derivatives.

    lower[n]  = likelihood of n's subtree given the state at n           (postfix :483-539)
    upper[n]  = likelihood of everything else given the state at n's FATHER,
                root frequencies folded in at the root's sons             (prefix :543-649)
    """
    N = len(weights)
    nn = flat.n_nodes
    lower, lexp = {}, {}
    for nid in range(nn):
        if flat.is_leaf[nid]:
            lower[nid] = leaf_array(tip_codes[nid], table, C)
            lexp[nid] = np.zeros((N, C), np.int64)
        else:
            A, E = None, np.zeros((N, C), np.int64)
            for s in flat.children[nid]:
                t = _contract(P[s], lower[s])
                A = t if A is None else A * t
                E = E + lexp[s]
            if scaled:
                A, E = _rescale(A, E)
            lower[nid], lexp[nid] = A, E
    root = flat.root
    res = DRResult()
    res.lower, res.lexp = lower, lexp
    root_clv, root_e = lower[root], lexp[root]

    if weighted_root:
        # setWeightedRootFreq (DRNonHomogeneousTreeLikelihood.cpp:927-962):
        # pi_x = sum_i sum_c p_c L_root[i][c][x] / sum_x(...)
        # (scale exponents: per-pattern; with N=1 they cancel in the normalisation)
        tot = np.einsum("icx,c->x", np.ldexp(root_clv, -(root_e - root_e.min())[:, :, None]), probs)
        root_freqs = tot / tot.sum()
    res.root_freqs = root_freqs

    # computeRootLikelihood (:653-719)
    aligned, emin = _align(root_clv, root_e)
    S_ic = np.einsum("icx,x->ic", aligned, root_freqs)
    SR = S_ic @ probs
    SR = np.where(SR < 0, 0.0, SR)
    res.SR, res.SR_exp = SR, emin
    with np.errstate(divide="ignore"):
        lp = np.log(SR) - emin * np.log(2.0)
    res.site_lnl = lp
    la = np.sort(weights * lp)                    # getLogLikelihood (:170-186)
    res.lnl = float(np.sum(la[::-1]))

    if dP is None and d2P is None:
        return res

    # prefix pass
    upper, uexp = {}, {}
    order = list(range(nn - 1, -1, -1))           # fathers before sons
    for nid in order:
        if nid == root:
            continue
        f = int(flat.parent[nid])
        A = None
        E = np.zeros((N, C), np.int64)
        for b in flat.children[f]:
            if b == nid:
                continue
            t = _contract(P[b], lower[b])
            A = t if A is None else A * t
            E = E + lexp[b]
        if f != root:
            t = _contract_T(P[f], upper[f])
            A = t if A is None else A * t
            E = E + uexp[f]
        else:
            A = A * root_freqs[None, None, :]
        if scaled:
            A, E = _rescale(A, E)
        upper[nid], uexp[nid] = A, E
    res.upper, res.uexp = upper, uexp

    nb = nn - 1
    d1 = np.zeros(nb)
    d2 = np.zeros(nb)
    res.dL, res.d2L = {}, {}
    for nid in range(nb):
        U, D = upper[nid], lower[nid]
        sh = (emin[:, None] - uexp[nid] - lexp[nid])   # [N][C]; true = stored * 2^-(e); ratio needs 2^(eR-eU-eD)
        if nh_form:
            # DRNonHomogeneousTreeLikelihood.cpp:370-413: larray = full conditional at the
            # father INCLUDING this son; divide by (P.lower), 0 where the denominator is 0
            den = _contract(P[nid], D)
            full = U * den
            num1 = _contract(dP[nid], D) if dP is not None else None
            num2 = _contract(d2P[nid], D) if d2P is not None else None
            with np.errstate(divide="ignore", invalid="ignore"):
                if num1 is not None:
                    q = np.where(den == 0, 0.0, full * num1 / den)
                    dLi = np.einsum("ic,c->i", np.ldexp(q.sum(axis=2), sh), probs) / SR
                if num2 is not None:
                    q = np.where(den == 0, 0.0, full * num2 / den)
                    d2Li = np.einsum("ic,c->i", np.ldexp(q.sum(axis=2), sh), probs) / SR
        else:
            if dP is not None:
                dLi = np.einsum("ic,c->i", np.ldexp(np.einsum("icx,icx->ic", U, _contract(dP[nid], D)), sh), probs) / SR
            if d2P is not None:
                d2Li = np.einsum("ic,c->i", np.ldexp(np.einsum("icx,icx->ic", U, _contract(d2P[nid], D)), sh), probs) / SR
        if dP is not None:
            res.dL[nid] = dLi
            d1[nid] = -float(np.sum(weights * dLi))                     # :340-368
        if d2P is not None and dP is not None:
            res.d2L[nid] = d2Li
            d2[nid] = -float(np.sum(weights * (d2Li - dLi ** 2)))       # :425-454
    res.d1, res.d2 = d1, d2
    return res


# ----------------------------------------------------------------------------
# Consumers of the DR arrays: computeLikelihoodAtNode + posterior probabilities
def likelihood_at_node(flat, res, P, nid):
    """DRTreeLikelihood::computeLikelihoodAtNode(nodeId, VVVdouble&)
    (DRHomogeneousTreeLikelihood::computeLikelihoodAtNode_, Likelihood/DRHomogeneousTreeLikelihood.cpp:723-815):
    the conditional likelihood of ALL the data given the state at `nid`,
        leaf / subtree part  x  sum_y P_nid[y][x] upper_nid[y]      (P transposed, :919-945),
    times the root frequencies at the root (:797-813).  ``res`` is a dr_eval result with the
    prefix pass done.  Returns (array [N][C][S], exponents [N][C]): true = array * 2^-exp."""
    L, E = res.lower[nid], res.lexp[nid]
    if nid == flat.root:
        return L * res.root_freqs[None, None, :], E.copy()
    return L * _contract_T(P[nid], res.upper[nid]), E + res.uexp[nid]


def posterior_probabilities(flat, res, P, nid, probs, tip_codes=None, table=None):
    """DRTreeLikelihoodTools::getPosteriorProbabilitiesForEachStateForEachRate
    (Likelihood/DRTreeLikelihoodTools.cpp:46-119), [N][C][S].
    Internal node: computeLikelihoodAtNode / sum over (c, x) -- as in the reference the class
    probabilities are NOT applied (they cancel for the equiprobable classes of the Gamma law).
    Leaf: leaf likelihoods[x] * p_c / sum_x leaf likelihoods (the rest of the tree is ignored, :58-79)."""
    if flat.is_leaf[nid]:
        la = table[np.asarray(tip_codes[nid], dtype=np.int64)]          # [N][S]
        return la[:, None, :] * np.asarray(probs)[None, :, None] / la.sum(axis=1)[:, None, None]
    A, E = likelihood_at_node(flat, res, P, nid)
    al = np.ldexp(A, -(E - E.min(axis=1, keepdims=True))[:, :, None])
    return al / al.sum(axis=(1, 2))[:, None, None]


def marginal_posteriors(flat, res, P, nid, probs):
    """MarginalNonRevAncestralStateReconstruction (fork): posterior probability of every state at a node, and the joint
    posterior of (node state x, father state y), per distinct site.

    The reference loops over the S root states, re-runs the prefix pass conditional on each
    (DRNonHomogeneousTreeLikelihood::computeLikelihoodPrefixConditionalOnRoot, .cpp:1026-1162) and adds
        FatherTerm_y(r) P[c][y][x] lower[x] p_c pi_r / L_site     (getJointLikelihoodFatherNode, MarginalNonRev...cpp:52-88)
    over r (computePosteriorProbabilitiesOfNodesForEachStatePerSite, :10-48).  The conditional prefix arrays are linear in
    the indicator of the root state, so their pi_r-weighted sum over r is the ordinary prefix array (root frequencies folded
    in at the root's sons, DRHomogeneousTreeLikelihood.cpp:625-640) and ONE pass suffices:
        joint[i][x][y] = sum_c p_c upper[i][c][y] P[c][y][x] sub[i][c][x] / L_i ,   post[i][x] = sum_y joint[i][x][y]
    (root: post[i][x] = sum_c p_c pi_x lower[i][c][x] / L_i, :112-122).  ``sub`` is the node's lower array (leaf likelihoods
    for a leaf).  ``marginal_posteriors_by_root_state`` below is the literal S-pass form; tests hold the two equal.
    Returns (post [N][S], joint [N][S][S] or None at the root)."""
    probs = np.asarray(probs, float)
    sub, esub = res.lower[nid], res.lexp[nid]
    if nid == flat.root:
        sh = res.SR_exp[:, None] - esub
        post = np.einsum("icx,c->ix", np.ldexp(sub * res.root_freqs[None, None, :], sh[:, :, None]), probs) / res.SR[:, None]
        return post, None
    U, eU = res.upper[nid], res.uexp[nid]
    sh = res.SR_exp[:, None] - esub - eU
    t = np.einsum("icy,cyx,icx->icxy", U, P[nid], sub)          # U[i][c][y] P[c][y][x] sub[i][c][x]
    joint = np.einsum("icxy,c->ixy", np.ldexp(t, sh[:, :, None, None]), probs) / res.SR[:, None, None]
    return joint.sum(axis=2), joint


def marginal_posteriors_by_root_state(flat, res, P, probs):
    """The reference's own S-pass algorithm, statement by statement, on UNSCALED arrays (small cases only):
    for every root state r: computeLikelihoodPrefixConditionalOnRoot(root, r) (DRNonHomogeneousTreeLikelihood.cpp:1026-1117;
    at a root son only father state r survives and NO root frequency is applied, computeLikelihoodRootSonConditionalOnState
    :1119-1162), then getPosteriorProbabilitiesOfNodesForEachRootStatePerSite (MarginalNonRev...cpp:91-136) adds
    getJointLikelihoodFatherNode = FatherTerm_y P[c][y][x] lower[x] r_c pi_r / l_i (:52-88) into jointProbabilities_ and
    postProbNode_.  ``res`` must come from dr_eval(scaled=False).  Returns ({node: post [N][S]}, {node: joint [N][S][S]})."""
    probs = np.asarray(probs, float)
    nn, root = flat.n_nodes, flat.root
    lower = res.lower
    N, C, S = lower[root].shape
    l_site = res.SR * np.ldexp(1.0, -res.SR_exp)            # getRootRateSiteLikelihoodArray
    post = {n: np.zeros((N, S)) for n in range(nn)}
    joint = {n: np.zeros((N, S, S)) for n in range(nn) if n != root}
    for r in range(S):
        cond = {}
        for nid in range(nn - 1, -1, -1):                    # fathers before sons (pre-order recursion of the reference)
            if nid == root:
                continue
            f = int(flat.parent[nid])
            A = np.ones((N, C, S))                           # resetLikelihoodArray
            if f != root:
                for b in flat.children[f]:
                    if b != nid:
                        A = A * _contract(P[b], lower[b])
                A = A * _contract_T(P[f], cond[f])           # computeLikelihoodFromArrays, root-side overload
            else:
                ind = np.zeros(S)
                ind[r] = 1.0                                 # "if (x == initState)" (:1147), else the product is 0
                for b in flat.children[f]:
                    if b != nid:
                        A = A * (_contract(P[b], lower[b]) * ind[None, None, :])
            cond[nid] = A
        pi_r = res.root_freqs[r]
        for nid in range(nn):
            if nid == root:                                  # :104-115
                post[nid][:, r] += pi_r * np.einsum("ic,c->i", lower[root][:, :, r], probs) / l_site
                continue
            # joint[i][x][y] += sum_c cond[i][c][y] P[c][y][x] lower[i][c][x] r_c pi_r / l_i
            j = np.einsum("icy,cyx,icx,c->ixy", cond[nid], P[nid], lower[nid], probs) * pi_r / l_site[:, None, None]
            joint[nid] += j
            post[nid] += j.sum(axis=2)
    return post, joint


def ml_joint_reconstruction(flat, tip_codes, table, P, root_freqs):
    """MLAncestralStateReconstruction (fork; Likelihood/MLAncestralStateReconstruction.cpp:6-188): joint ML reconstruction by
    Pupko's max-product recursion, arrays indexed by the FATHER's state.
      leaf with observed state s (first state whose init value is 1; DRASRTreeLikelihoodData.cpp:283-292):
          L[i][c][x] = P[c][x][s], anc[i][x] = s   (no such state: L = 1, anc = 0)
      internal (fillLikelihoodsArrays :101-134):  L[i][c][x] = max_y P[c][x][y] prod_sons L_son[i][c][y], anc[i][x] = first y
          with a strictly larger (positive) value; the table of the LAST class survives
      root (fillRootLikelihoodsArrays :66-98):    L[i][c][x] = pi_x prod_sons L_son[i][c][x]
      trace back (getAllAncestralStatesRec :143-186): root = first maximum of class 0, node = anc[node][i][father's state].
    Unscaled, like the reference.  Returns (states [n_nodes][N] int, L_root [N][C][S])."""
    nn = flat.n_nodes
    C, S = P.shape[1], P.shape[2]
    N = len(next(iter(tip_codes.values())))
    L, anc = {}, {}
    for nid in range(nn):
        if flat.is_leaf[nid]:
            t = table[np.asarray(tip_codes[nid], dtype=np.int64)]              # [N][S]
            has = (t == 1.0).any(axis=1)
            s = np.argmax(t == 1.0, axis=1)
            A = np.transpose(P[nid][:, :, s], (2, 0, 1)).copy()                 # [N][C][x] = P[c][x][s_i]
            A[~has] = 1.0
            L[nid] = A
            anc[nid] = np.repeat(np.where(has, s, 0)[:, None], S, axis=1)
            continue
        prod = np.ones((N, C, S))
        for son in flat.children[nid]:
            prod = prod * L[son]
        if nid == flat.root:
            L[nid] = prod * np.asarray(root_freqs)[None, None, :]
            continue
        cand = prod[:, :, None, :] * P[nid][None, :, :, :]                     # [i][c][x][y]
        L[nid] = cand.max(axis=3)
        anc[nid] = np.argmax(cand[:, C - 1], axis=2)                            # first maximum; all-zero row -> 0
    states = np.zeros((nn, N), np.int64)
    states[flat.root] = np.argmax(L[flat.root][:, 0, :], axis=1)
    for nid in range(nn - 2, -1, -1):                                           # fathers (larger ids) first
        f = int(flat.parent[nid])
        states[nid] = anc[nid][np.arange(N), states[f]]
    return states, L[flat.root]


def root_reparam_derivatives(flat, res, P, dP, d2P, probs, weights):
    """Derivatives of -lnL with respect to ``BrLenRoot`` (l1 + l2) and ``RootPosition`` (l1 / (l1 + l2)), the
    re-parametrisation of the two root branches of a rooted tree
    (AbstractNonHomogeneousTreeLikelihood::initBranchLengthsParameters, .cpp:386-389; applyParameters :319-330).
    First order: DRNonHomogeneousTreeLikelihood::getFirstOrderDerivative (.cpp:445-478) -- combinations of the two
    branches' derivatives.  Second order: getSecondOrderDerivative (.cpp:576-867) -- the cross term
    dP_1 L_1 . dP_2 L_2 cannot be deduced from the per-branch values and is rebuilt at the root:
        dl  = pos dl1 l2 + (1 - pos) dl2 l1,   d2l = pos^2 d2l1 l2 + (1 - pos)^2 d2l2 l1 + 2 pos (1 - pos) dl1 dl2   (:662-663)
    per root state, times the other root sons, the root frequencies and the class probabilities, then
    -sum_i w_i (d2l_i / SR_i - (dl_i / SR_i)^2)  (:705-720).  Returns dict d1_len, d1_pos, d2_len, d2_pos."""
    root = flat.root
    sons = list(flat.children[root])
    r1, r2 = sons[0], sons[1]
    l1b, l2b = flat.brlen[r1], flat.brlen[r2]
    length, pos = l1b + l2b, l1b / (l1b + l2b)
    L1, L2 = res.lower[r1], res.lower[r2]
    e = res.lexp[r1] + res.lexp[r2]
    others = np.ones_like(_contract(P[r1], L1))
    for s in sons[2:]:
        others = others * _contract(P[s], res.lower[s])
        e = e + res.lexp[s]
    l1, l2 = _contract(P[r1], L1), _contract(P[r2], L2)
    dl1, dl2 = _contract(dP[r1], L1), _contract(dP[r2], L2)
    d2l1, d2l2 = _contract(d2P[r1], L1), _contract(d2P[r2], L2)
    sh = (res.SR_exp[:, None] - e)                       # [N][C]

    def site(v):                                        # sum_c p_c sum_x pi_x others v / SR
        t = np.einsum("icx,x->ic", others * v, res.root_freqs)
        return np.einsum("ic,c->i", np.ldexp(t, sh), probs) / res.SR

    d_len = site(pos * dl1 * l2 + (1 - pos) * dl2 * l1)
    d_pos = site(length * (dl1 * l2 - dl2 * l1))
    d2_len = site(pos * pos * d2l1 * l2 + (1 - pos) ** 2 * d2l2 * l1 + 2 * pos * (1 - pos) * dl1 * dl2)
    d2_pos = site(length * length * (d2l1 * l2 + d2l2 * l1 - 2 * dl1 * dl2))
    w = np.asarray(weights, float)
    return {"d1_len": -float(np.sum(w * d_len)), "d1_pos": -float(np.sum(w * d_pos)),
            "d2_len": -float(np.sum(w * (d2_len - d_len ** 2))), "d2_pos": -float(np.sum(w * (d2_pos - d_pos ** 2))),
            "length": length, "pos": pos, "root1": r1, "root2": r2}


# ----------------------------------------------------------------------------
# R-class derivatives (single recursion re-pruned along the path to the root)
# ----------------------------------------------------------------------------
def r_derivative(flat, tip_codes, table, P, dP, C, root_freqs, probs, weights, branch, order=1, d2P=None):
    """computeTreeDLikelihood / computeDownSubtreeDLikelihood
    (RHomogeneousTreeLikelihood.cpp:365-541; D2 :615-791) on the global pattern list:
    substitute dP (d2P) on ``branch`` and re-prune the path to the root; then
    getFirstOrderDerivative (:346-361) / getSecondOrderDerivative (:596-611)."""
    N = len(weights)
    clv = {}
    for nid in range(flat.n_nodes):
        if flat.is_leaf[nid]:
            clv[nid] = leaf_array(tip_codes[nid], table, C)
        else:
            A = None
            for s in flat.children[nid]:
                t = _contract(P[s], clv[s])
                A = t if A is None else A * t
            clv[nid] = A
    L = site_like_plain(clv[flat.root], root_freqs, probs)

    def along(M):
        path = set()
        n = branch
        while n != flat.root:
            path.add(n)
            n = int(flat.parent[n])
        d = {}
        for nid in range(flat.n_nodes):
            if flat.is_leaf[nid]:
                continue
            hit = [s for s in flat.children[nid] if s in path]
            if not hit:
                continue
            s0 = hit[0]
            A = None
            for s in flat.children[nid]:
                if s == s0:
                    t = _contract(M[s], clv[s]) if s == branch else _contract(P[s], d[s])
                else:
                    t = _contract(P[s], clv[s])
                A = t if A is None else A * t
            d[nid] = A
        return site_like_plain(d[flat.root], root_freqs, probs)

    dL = along(dP)
    if order == 1:
        return -float(np.sum(weights * dL / L))
    d2L = along(d2P)
    return -float(np.sum(weights * (d2L / L - (dL / L) ** 2)))


def site_like_plain(root_clv, root_freqs, probs):
    return np.einsum("icx,x,c->i", root_clv, root_freqs, probs)


# ----------------------------------------------------------------------------
# brute force over all internal-state assignments (structural check, tiny trees)
# ----------------------------------------------------------------------------
def brute_force_site(flat, tip_states_probs, P_c, root_freqs):
    """Sum over every assignment of states to internal nodes for ONE pattern and ONE
    rate class.  tip_states_probs: leaf id -> vector[S]; P_c[node] = S x S."""
    import itertools
    S = len(root_freqs)
    internal = [i for i in range(flat.n_nodes) if not flat.is_leaf[i]]
    total = 0.0
    for assign in itertools.product(range(S), repeat=len(internal)):
        st = dict(zip(internal, assign))
        p = root_freqs[st[flat.root]]
        for nid in range(flat.n_nodes):
            if nid == flat.root:
                continue
            x = st[int(flat.parent[nid])]
            if flat.is_leaf[nid]:
                p *= float(P_c[nid][x] @ tip_states_probs[nid])
            else:
                p *= P_c[nid][x, st[nid]]
            if p == 0:
                break
        total += p
    return total
