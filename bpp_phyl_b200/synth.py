"""Synthetic inputs of the BASELINE.json shapes for bench.py: random trees in the flattened
post-order layout the C ABI takes, model eigensystems, Gamma rate classes, and tip states simulated down the
tree on the GPU with torch (plumbing) from P(t) tables produced by the product's own bppgpu_pt_batch.

Not a port of the reference's model classes (those live in C++ under host/); this only has to produce
well-formed inputs deterministically from a seed.
"""
from __future__ import annotations

import numpy as np


class Tree:
    """Flattened topology: node ids are post-order positions (root last), CSR children, branch lengths."""

    def __init__(self, child_off, children, brlen):
        self.child_off = np.asarray(child_off, np.int32)
        self.children = np.asarray(children, np.int32)
        self.brlen = np.asarray(brlen, np.float64)
        self.nn = len(self.child_off) - 1
        self.root = self.nn - 1
        self.is_leaf = np.diff(self.child_off) == 0
        self.n_leaves = int(self.is_leaf.sum())
        self.n_internal = self.nn - self.n_leaves
        self.parent = np.full(self.nn, -1, np.int32)
        for n in range(self.nn):
            self.parent[self.children[self.child_off[n]:self.child_off[n + 1]]] = n
        self.leaf_nodes = np.where(self.is_leaf)[0]           # leaf slot -> node id (increasing id)

    def sons(self, n):
        return self.children[self.child_off[n]:self.child_off[n + 1]]


def random_tree(n_taxa, rng, mean_brlen=0.05, rooted=False):
    """Random binary topology by sequential random attachment, Exp(mean) lengths clamped to [1e-6, 1e4]
    (AbstractHomogeneousTreeLikelihood.cpp:305-337).  Unrooted: 3-son root; rooted: 2-son root."""
    sons = {0: []}          # node -> list of sons (temporary ids)
    parent = {}
    nxt = 1
    edges = []
    start = 2 if rooted or n_taxa < 3 else 3
    for _ in range(start):
        sons[0].append(nxt)
        parent[nxt] = 0
        sons[nxt] = []
        edges.append(nxt)
        nxt += 1
    for _ in range(start, n_taxa):
        e = edges[int(rng.integers(len(edges)))]
        f = parent[e]
        mid, leaf = nxt, nxt + 1
        nxt += 2
        sons[f][sons[f].index(e)] = mid
        parent[mid] = f
        sons[mid] = [e, leaf]
        parent[e] = mid
        parent[leaf] = mid
        sons[leaf] = []
        edges.extend([mid, leaf])
    # post-order renumbering
    order = []
    st = [(0, 0)]
    while st:
        n, i = st.pop()
        if i < len(sons[n]):
            st.append((n, i + 1))
            st.append((sons[n][i], 0))
        else:
            order.append(n)
    new = {old: k for k, old in enumerate(order)}
    child_off = [0]
    children = []
    for old in order:
        children.extend(new[s] for s in sons[old])
        child_off.append(len(children))
    brlen = np.clip(rng.exponential(mean_brlen, size=len(order)), 1e-6, 1e4)
    brlen[-1] = 0.0
    return Tree(child_off, children, brlen)


def gamma_rates(ncat, alpha):
    """Equiprobable Gamma(alpha, alpha) classes, class value = class mean (bpp-core GammaDiscreteDistribution)."""
    from scipy import special
    if ncat == 1:
        return np.ones(1), np.ones(1)
    q = special.gammaincinv(alpha, np.arange(1, ncat) / ncat) / alpha
    b = np.concatenate([[0.0], q, [np.inf]])
    return ncat * np.diff(special.gammainc(alpha + 1.0, b * alpha)), np.full(ncat, 1.0 / ncat)


def reversible_eigensystem(exch, pi):
    """Q = exch * pi (off-diagonal), normalised to one substitution per unit time; eigensystem through the
    symmetric similarity pi^1/2 Q pi^-1/2.  Returns dict(Q, V, Vinv, ev, pi)."""
    pi = np.asarray(pi, float)
    pi = pi / pi.sum()
    Q = np.asarray(exch, float) * pi[None, :]
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(axis=1))
    Q = Q / -(np.diag(Q) @ pi)
    s = np.sqrt(pi)
    B = (s[:, None] * Q) / s[None, :]
    w, U = np.linalg.eigh((B + B.T) / 2)
    w[np.argmax(w)] = 0.0
    return {"Q": Q, "V": U / s[:, None], "Vinv": U.T * s[None, :], "ev": w, "pi": pi}


def gtr(a=1.2, b=0.8, c=0.6, d=1.5, e=0.9, pi=(.3, .2, .25, .25)):
    """GTR exchangeabilities as in Model/Nucleotide/GTR.cpp:84-124 (AC=d AG=1 AT=b CG=e CT=a GT=c)."""
    ex = np.zeros((4, 4))
    ex[0, 1] = ex[1, 0] = d
    ex[0, 2] = ex[2, 0] = 1.0
    ex[0, 3] = ex[3, 0] = b
    ex[1, 2] = ex[2, 1] = e
    ex[1, 3] = ex[3, 1] = a
    ex[2, 3] = ex[3, 2] = c
    return reversible_eigensystem(ex, pi)


def random_reversible(S, rng):
    """A random reversible model with S states (stands in for LG08-shaped inputs where only the shape matters)."""
    ex = rng.gamma(0.6, 1.0, size=(S, S)) + 1e-3
    ex = (ex + ex.T) / 2
    pi = rng.dirichlet(np.full(S, 5.0))
    return reversible_eigensystem(ex, pi)


def chromosome_generator(n_states, gain, loss, dupl, demi):
    """ChromEvol generator on counts 1..n_states (Model/ChromosomeSubstitutionModel.cpp:431-577, constant rates, no base
    number): +1 gain, -1 loss, x2 duplication, x1.5 demi-duplication (split between floor/ceil for odd counts), overflow to
    the maximum state.  Not normalised."""
    n = n_states
    Q = np.zeros((n, n))
    for i in range(1, n + 1):
        r = i - 1
        if i + 1 <= n:
            Q[r, r + 1] += gain
        if i - 1 >= 1:
            Q[r, r - 1] += loss
        if 2 * i <= n:
            Q[r, 2 * i - 1] += dupl
        elif i != n:
            Q[r, n - 1] += dupl
        if i != n:
            if i % 2 == 0 and int(i * 1.5) <= n:
                Q[r, int(i * 1.5) - 1] += demi
            elif i % 2 != 0 and int(np.ceil(i * 1.5)) <= n:
                if i == 1:
                    Q[r, int(np.ceil(i * 1.5)) - 1] += demi
                else:
                    Q[r, int(np.ceil(i * 1.5)) - 1] += demi / 2
                    Q[r, int(np.floor(i * 1.5)) - 1] += demi / 2
            else:
                Q[r, n - 1] += demi
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(axis=1))
    return Q


def real_block_form(w, U):
    """Complex eigenpairs -> the real form bpp-core's EigenValue<double> presents: eigenvalues (re, im), real V with
    A V = V D, D block diagonal [[re, im], [-im, re]] per conjugate pair, the +im member first."""
    n = len(w)
    re, im, V = np.zeros(n), np.zeros(n), np.zeros((n, n))
    used = np.zeros(n, bool)
    k = 0
    for i in np.argsort(-w.real, kind="stable"):
        if used[i]:
            continue
        if abs(w[i].imag) <= 1e-13 * max(1.0, abs(w[i])):
            re[k], V[:, k] = w[i].real, U[:, i].real
            used[i] = True
            k += 1
        else:
            cand = [j for j in range(n) if not used[j] and j != i]
            j = min(cand, key=lambda jj: abs(w[jj] - np.conj(w[i])))
            ip = i if w[i].imag > 0 else j
            re[k] = re[k + 1] = w[ip].real
            im[k], im[k + 1] = w[ip].imag, -w[ip].imag
            V[:, k], V[:, k + 1] = U[:, ip].real, U[:, ip].imag
            used[i] = used[j] = True
            k += 2
    return re, im, V


def chromosome_eigensystem(args):
    """Host eigendecomposition of one parameter point (what ChromosomeSubstitutionModel::updateEigenMatrices does per
    likelihood object: Model/ChromosomeSubstitutionModel.cpp:589-802) and the reference's choice of route for it (:686-767):
    the eigen form V exp(L t) V^-1 when V can be inverted and exactly one eigenvalue is null, the Taylor series otherwise
    ("route": "eigen" | "series"; like oracle/ref_models.py, an inverse with non-finite entries or cond(V) > 1e15 stands for
    MatrixTools::inv's ZeroDivisionException).  Every point is returned -- nothing is redrawn; "resid" = max |V D V^-1 - Q| / max |Q|
    says how well the eigen form a point travels with reproduces its generator (the reference never checks)."""
    n, gain, loss, dupl, demi = args
    Q = chromosome_generator(n, gain, loss, dupl, demi)
    base = {"Q": Q, "pi": np.full(n, 1.0 / n), "params": (gain, loss, dupl, demi), "route": "series", "resid": np.inf}
    with np.errstate(all="ignore"):
        w, U = np.linalg.eig(Q)
        re, im, V = real_block_form(w, U)
        try:
            Vinv = np.linalg.inv(V)
        except np.linalg.LinAlgError:
            return base
        if not np.all(np.isfinite(Vinv)) or not np.all(np.isfinite(V)) or np.linalg.cond(V) > 1e15:
            return base
        null = (np.abs(re) < 1e-6) & (np.abs(im) < 1e-6)          # NumConstants::SMALL()
        if null.sum() != 1:
            return base
        D = np.diag(re)
        for k in range(n - 1):
            if im[k] > 0:
                D[k, k + 1], D[k + 1, k] = im[k], -im[k]
        resid = float(np.abs(V @ D @ Vinv - Q).max() / np.abs(Q).max())
    re[np.argmax(null)] = 0.0
    return dict(base, V=V, Vinv=Vinv, ev=re, ev_im=im, route="eigen", resid=resid)


def chromosome_model_desc(es):
    from . import capi
    S = len(es["Q"])
    chr_flags = capi.MODEL_CLAMP01 | capi.MODEL_CHR_DERIV | capi.MODEL_CHR_TAYLOR
    if es.get("route", "eigen") == "series":                     # isNonSingular_ false: the reference's Taylor + squaring route
        return capi.model_desc(S, chr_flags, rate=1.0, Q=es["Q"])
    return capi.model_desc(S, chr_flags | capi.MODEL_NONSINGULAR | (0 if np.any(es["ev_im"]) else capi.MODEL_DIAGONALIZABLE),
                           rate=1.0, V=es["V"], Vinv=es["Vinv"], ev_re=es["ev"], ev_im=es["ev_im"], Q=es["Q"])


def _single_thread_blas():
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass


def chromosome_points(n_states, n_points, seed, workers=None, well_conditioned_only=False):
    """`n_points` parameter points drawn uniformly from the box gain, loss ~ U(0,2); dupl, demi ~ U(0,1), eigensystems computed
    in parallel.  Every drawn point is kept with the route the reference would take for it (chromosome_eigensystem);
    `well_conditioned_only` restores the round-1 behaviour (points whose eigen form misses Q by more than 1e-9 are redrawn)."""
    import concurrent.futures as cf
    import os
    rng = np.random.default_rng(seed)
    out = []
    with cf.ProcessPoolExecutor(max_workers=workers or (os.cpu_count() or 1), initializer=_single_thread_blas) as ex:
        while len(out) < n_points:
            need = n_points - len(out)
            args = [(n_states, rng.uniform(0, 2), rng.uniform(0, 2), rng.uniform(0, 1), rng.uniform(0, 1))
                    for _ in range((int(need * 2.6) + 8) if well_conditioned_only else need)]
            for es in ex.map(chromosome_eigensystem, args, chunksize=8):
                if len(out) < n_points and (not well_conditioned_only or (es["route"] == "eigen" and es["resid"] <= 1e-9)):
                    out.append(es)
    return out


def simulate_single_character(tree: Tree, P, root_state, seed):
    """One character down the tree: P[node] is the S x S transition matrix of the branch above node."""
    rng = np.random.default_rng(seed)
    st = {tree.root: root_state}
    codes = np.zeros((tree.n_leaves, 1), np.uint8)
    slot = {int(n): k for k, n in enumerate(tree.leaf_nodes)}
    for node in range(tree.nn - 2, -1, -1):
        row = np.clip(P[node][st[int(tree.parent[node])]], 0, None)
        st[node] = int(rng.choice(len(row), p=row / row.sum()))
        if tree.is_leaf[node]:
            codes[slot[node], 0] = st[node]
    return codes


def model_desc(es):
    from . import capi
    S = len(es["ev"])
    return capi.model_desc(S, es.get("flags", capi.MODEL_DIAGONALIZABLE | capi.MODEL_NONSINGULAR), rate=es.get("rate", 1.0),
                           V=es["V"], Vinv=es["Vinv"], ev_re=es["ev"], Q=es["Q"])


def host_pt(es, t):
    """P(t) = V exp(ev rate t) V^-1 on the host (real spectra), for callers that must not touch the device."""
    t = np.asarray(t, float)
    return np.einsum("ik,tk,kj->tij", es["V"], np.exp(np.outer(t * es.get("rate", 1.0), es["ev"])), es["Vinv"])


def simulate_tip_codes(tree: Tree, es, rates, n_sites, seed, device="cuda:0", chunk=1 << 20):
    """Tip states [n_leaves][n_sites] uint8 (leaf slots in increasing node id), simulated down the tree with torch on `device`.
    On a GPU, P(t) comes from bppgpu_pt_batch (the product's interface 1); with device="cpu" from host_pt (no library call)."""
    import torch
    S = len(es["ev"])
    C = len(rates)
    t = (tree.brlen[:, None] * np.asarray(rates)[None, :]).ravel()
    if str(device) == "cpu":
        P = host_pt(es, t)
    else:
        from . import capi
        P, _, _ = capi.pt_batch(model_desc(es), t, capi.WANT_P, device=int(str(device).split(":")[-1]) if ":" in str(device) else 0)
    P = np.clip(P.reshape(tree.nn, C, S, S), 0, None)
    cum = torch.tensor(np.cumsum(P / P.sum(-1, keepdims=True), axis=-1), device=device)     # [nn][C][S][S]
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((tree.n_leaves, n_sites), dtype=torch.uint8, device="cpu")
    if n_sites and str(device) != "cpu":
        out = out.pin_memory()
    leaf_slot = {int(n): k for k, n in enumerate(tree.leaf_nodes)}
    picum = torch.tensor(np.cumsum(es["pi"]), device=device)
    for s0 in range(0, n_sites, chunk):
        n = min(chunk, n_sites - s0)
        cls = torch.randint(C, (n,), device=device, generator=g)
        st = {tree.root: torch.searchsorted(picum, torch.rand(n, device=device, generator=g, dtype=torch.float64)).clamp_(max=S - 1).to(torch.uint8)}
        for node in range(tree.nn - 2, -1, -1):
            f = int(tree.parent[node])
            rows = cum[node][cls, st[f].long()]                                    # [n][S]
            u = torch.rand(n, 1, device=device, generator=g, dtype=torch.float64)
            child = (u > rows).sum(dim=1).clamp_(max=S - 1)
            if tree.is_leaf[node]:
                out[leaf_slot[node], s0:s0 + n] = child.to(torch.uint8).cpu()
            else:
                st[node] = child.to(torch.uint8)
            # free fathers whose sons are all done
        del st
    return out.numpy()
