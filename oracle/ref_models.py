"""Oracle (TEST INFRASTRUCTURE): substitution-model side of the hot path.

NumPy/SciPy restatement of the reference's generator construction, eigen
bookkeeping and the getPij_t / getdPij_dt / getd2Pij_dt2 family.  Every
function cites the reference file:line it follows (paths relative to
/root/reference/src/Bpp/Phyl/).

Third-party arithmetic that is NOT under /root/reference (bpp-core 2.4.1:
MatrixTools::mult/inv/Taylor/pow, EigenValue<double>, GammaDiscreteDistribution,
NumConstants) is restated from its published behaviour; only the product
V.f(Lambda).V^-1 matters for parity, so LAPACK's eigenvector scaling/order is
used in place of JAMA's.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy import special

# NumConstants (bpp-core; values from upstream Bio++): used by updateMatrices
TINY = 1e-12
SMALL = 1e-6
VERY_TINY = 1e-20
MILLI = 1e-3


# --------------------------------------------------------------------------
# Rate categories  (Model/RateDistribution/GammaDiscreteRateDistribution.h:52-58)
# --------------------------------------------------------------------------
def gamma_rates(ncat: int, alpha: float):
    """Equiprobable classes of Gamma(alpha, beta=alpha); class value = class MEAN.

    bpp-core GammaDiscreteDistribution::discretize; verified numerically in
    SURVEY.md appendix C (K=4, alpha=1 -> .13695378 .47675186 1. 2.38629436).
    """
    beta = alpha
    if ncat == 1:
        return np.array([1.0]), np.array([1.0])
    q = special.gammaincinv(alpha, np.arange(1, ncat) / ncat) / beta  # class bounds
    bounds = np.concatenate([[0.0], q, [np.inf]])
    cdf1 = special.gammainc(alpha + 1.0, bounds * beta)  # I(beta*q; alpha+1)
    rates = ncat * (alpha / beta) * np.diff(cdf1)
    return rates, np.full(ncat, 1.0 / ncat)


def constant_rate():
    """Model/RateDistribution/ConstantRateDistribution.h"""
    return np.array([1.0]), np.array([1.0])


# --------------------------------------------------------------------------
# Model container mirroring the accessors of SubstitutionModel.h:215-525
# --------------------------------------------------------------------------
@dataclass
class Model:
    name: str
    Q: np.ndarray                      # generator_, row = from, col = to
    freq: np.ndarray                   # freq_
    ev_re: np.ndarray = None           # eigenValues_
    ev_im: np.ndarray = None           # iEigenValues_
    V: np.ndarray = None               # rightEigenVectors_ (columns)
    Vinv: np.ndarray = None            # leftEigenVectors_ (rows)
    diagonalizable: bool = True        # isDiagonalizable_
    nonsingular: bool = True           # isNonSingular_
    rate: float = 1.0                  # rate_
    scalable: bool = True              # isScalable_
    reversible: bool = False
    chromosome: bool = False           # ChromosomeSubstitutionModel P(t) semantics
    pow_gen: list = field(default_factory=list)   # vPowGen_/vPowExp_
    first_norm: float = 0.0            # firstNormQ_ (chromosome)

    @property
    def size(self):
        return self.Q.shape[0]


def _set_diagonal(Q):
    """AbstractSubstitutionModel::setDiagonal (Model/AbstractSubstitutionModel.cpp:666-680)."""
    Q = Q.copy()
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(axis=1))
    return Q


def _real_block_eig(A):
    """Real eigen-form of a real matrix as JAMA's EigenValue<double> returns it:
    real eigenvalues d, imaginary parts e, and a REAL matrix V such that
    A.V = V.D with D block diagonal (2x2 blocks [[re, im], [-im, re]] for
    conjugate pairs, the +im member first).  Only V.f(D).V^-1 matters.
    """
    w, U = np.linalg.eig(A)
    n = A.shape[0]
    d = np.zeros(n)
    e = np.zeros(n)
    V = np.zeros((n, n))
    used = np.zeros(n, bool)
    k = 0
    order = list(range(n))
    for i in order:
        if used[i]:
            continue
        if abs(w[i].imag) <= 1e-14 * max(1.0, abs(w[i])):
            d[k] = w[i].real
            V[:, k] = U[:, i].real
            used[i] = True
            k += 1
        else:
            # find conjugate partner
            j = min((jj for jj in range(n) if not used[jj] and jj != i),
                    key=lambda jj: abs(w[jj] - np.conj(w[i])))
            ip, im_ = (i, j) if w[i].imag > 0 else (j, i)
            lam = w[ip]
            u = U[:, ip]
            # A (x + i y) = (a + i b)(x + i y): A x = a x - b y ; A y = b x + a y
            d[k] = d[k + 1] = lam.real
            e[k], e[k + 1] = lam.imag, -lam.imag
            V[:, k] = u.real
            V[:, k + 1] = u.imag
            used[i] = used[j] = True
            k += 2
    return d, e, V


def update_matrices(m: Model, compute_freq: bool = True) -> Model:
    """AbstractSubstitutionModel::updateMatrices (Model/AbstractSubstitutionModel.cpp:175-421).

    Strips all-zero ("stop") lines, eigen-decomposes the rest, inverts, decides
    diagonalisable / non-singular, pins the ~0 eigenvalue to exactly 0, takes the
    equilibrium frequencies from its left vector, normalises if scalable.
    """
    Q = m.Q
    n = Q.shape[0]
    vnull = np.array([abs(Q[i, i]) < TINY and np.all(np.abs(Q[:, i]) < TINY) for i in range(n)])  # :189-208
    ok = np.where(~vnull)[0]
    nstop = int(vnull.sum())
    d, e, Vk = _real_block_eig(Q[np.ix_(ok, ok)])
    if nstop:
        ev_re = np.concatenate([d, np.zeros(nstop)])
        ev_im = np.concatenate([e, np.zeros(nstop)])
        V = np.zeros((n, n))
        V[np.ix_(ok, np.arange(len(ok)))] = Vk
        gi = 0
        for i in range(n):
            if vnull[i]:
                gi += 1
                V[i, len(ok) + gi - 1] = 1.0       # :253-262
    else:
        ev_re, ev_im, V = d, e, Vk
    m.ev_re, m.ev_im, m.V = ev_re, ev_im, V
    try:
        m.Vinv = np.linalg.inv(V)
        if not np.all(np.isfinite(m.Vinv)) or np.linalg.cond(V) > 1e15:
            raise np.linalg.LinAlgError
        m.diagonalizable = True
        if not m.reversible and np.any(np.abs(ev_im) > TINY):
            m.diagonalizable = False                  # :293-303
        nullev = []
        fact = 0.1
        while not nullev and fact < 1000:             # :307-316
            fact *= 10
            nullev = [i for i in range(n - nstop)
                      if abs(ev_re[i]) < fact * SMALL and abs(ev_im[i]) < SMALL]
        m.nonsingular = len(nullev) == 1
        nulleigen = None
        if not m.nonsingular:
            first = int(np.where(~vnull)[0][0])
            for c in nullev:                          # :326-352
                val = V[first, c]
                if all(abs((V[i, c] - val) / val) <= SMALL for i in ok[1:]):
                    m.nonsingular = True
                    nulleigen = c
                    break
        else:
            nulleigen = nullev[0]
        if m.nonsingular:
            ev_re[nulleigen] = 0.0
            ev_im[nulleigen] = 0.0
            if compute_freq:
                f = m.Vinv[nulleigen, :].copy()
                m.freq = f / f.sum()                   # :362-370
        else:
            m.diagonalizable = False
    except np.linalg.LinAlgError:
        m.nonsingular = False
        m.diagonalizable = False
    if not m.nonsingular:                              # :386-410
        mn = np.min(np.diag(m.Q))
        _set_scale(m, -1.0 / mn)
        if compute_freq:
            T = np.linalg.matrix_power(np.eye(n) + m.Q, 256)
            m.freq = T[0, :].copy()
    _normalize(m)                                      # :413-415
    if not m.nonsingular:
        m.pow_gen = _taylor_powers(m.Q, 30)            # :417-418
    return m


def _taylor_powers(Q, n):
    """MatrixTools::Taylor(A, p, vO): vO[i] = A^i, i = 0..p-1 (bpp-core)."""
    out = [np.eye(Q.shape[0])]
    for _ in range(1, n):
        out.append(out[-1] @ Q)
    return out


def _set_scale(m: Model, s: float):
    """AbstractSubstitutionModel::setScale (:655-663)."""
    if m.scalable:
        m.Q = m.Q * s
        if m.ev_re is not None:
            m.ev_re = m.ev_re * s
            m.ev_im = m.ev_im * s


def _normalize(m: Model):
    """normalize/getScale (:645-652, :684-688): -sum_i pi_i Q_ii = 1."""
    if m.scalable:
        _set_scale(m, 1.0 / (-float(np.dot(np.diag(m.Q), m.freq))))


def _reversible(name, exch, freq):
    """AbstractReversibleSubstitutionModel::updateMatrices (:694-703)."""
    freq = np.asarray(freq, float)
    Q = exch * freq[None, :]
    Q = _set_diagonal(Q)
    m = Model(name, Q, freq.copy(), reversible=True)
    _normalize(m)
    return update_matrices(m, compute_freq=True)


# ---- nucleotide models ----------------------------------------------------
def gtr(a=1., b=1., c=1., d=1., e=1., pi=(.25, .25, .25, .25)):
    """GTR::updateMatrices (Model/Nucleotide/GTR.cpp:84-124): exchangeabilities
    AC=d AG=1 AT=b CG=e CT=a GT=c (then normalised)."""
    ex = np.zeros((4, 4))
    ex[0, 1] = ex[1, 0] = d
    ex[0, 2] = ex[2, 0] = 1.0
    ex[0, 3] = ex[3, 0] = b
    ex[1, 2] = ex[2, 1] = e
    ex[1, 3] = ex[3, 1] = a
    ex[2, 3] = ex[3, 2] = c
    return _reversible("GTR", ex, pi)


def hky85(kappa=1., pi=(.25, .25, .25, .25)):
    """HKY85::updateMatrices (Model/Nucleotide/HKY85.cpp:80-191); the generic
    eigen path is used instead of the closed form (identical P, T92.cpp:355-386
    checked to 4e-17 in SURVEY.md appendix C)."""
    ex = np.ones((4, 4))
    ex[0, 2] = ex[2, 0] = kappa
    ex[1, 3] = ex[3, 1] = kappa
    np.fill_diagonal(ex, 0.0)
    return _reversible("HKY85", ex, pi)


def t92(kappa=1., theta=0.5):
    """T92::updateMatrices (Model/Nucleotide/T92.cpp:81-187)."""
    m = hky85(kappa, ((1 - theta) / 2, theta / 2, theta / 2, (1 - theta) / 2))
    m.name = "T92"
    return m


def k80(kappa=1.):
    """K80::updateMatrices (Model/Nucleotide/K80.cpp:64-95)."""
    m = hky85(kappa)
    m.name = "K80"
    return m


def jc69():
    m = hky85(1.0)
    m.name = "JC69"
    return m


# ---- protein ----------------------------------------------------------------
def lg08():
    """LG08::LG08 (Model/Protein/LG08.cpp:53-62) + data tables."""
    from .lg08_data import LG08_LOWER, LG08_FREQ
    ex = np.zeros((20, 20))
    for i in range(1, 20):
        for j in range(i):
            ex[i, j] = ex[j, i] = LG08_LOWER[i - 1][j]
    return _reversible("LG08", ex, LG08_FREQ)


# ---- codon ------------------------------------------------------------------
_TCAG_AA = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"


def standard_genetic_code():
    """Amino-acid letter per codon index 16*n1+4*n2+n3 with A,C,G,T = 0..3
    (bpp-seq CodonAlphabet order).  '*' = stop."""
    tcag = "TCAG"
    aa = [None] * 64
    k = 0
    for a in tcag:
        for b in tcag:
            for c in tcag:
                idx = 16 * "ACGT".index(a) + 4 * "ACGT".index(b) + "ACGT".index(c)
                aa[idx] = _TCAG_AA[k]
                k += 1
    return "".join(aa)


def f3x4(pos_freqs):
    """Codon frequencies from 3 position-specific nucleotide frequency vectors with
    stop mass moved onto non-stop single-nucleotide neighbours proportionally to
    freq^2 ("quadratic", FrequencySet/CodonFrequencySet.cpp:473-530).  Benchmarks
    may pass an explicit 64-vector instead."""
    aa = standard_genetic_code()
    pf = np.asarray(pos_freqs, float)
    f = np.array([pf[0, i // 16] * pf[1, (i // 4) % 4] * pf[2, i % 4] for i in range(64)])
    stops = [i for i in range(64) if aa[i] == "*"]
    out = f.copy()
    for s in stops:
        nb = []
        for pos, mul in ((0, 16), (1, 4), (2, 1)):
            cur = (s // mul) % 4
            for n in range(4):
                if n != cur:
                    j = s + (n - cur) * mul
                    if aa[j] != "*":
                        nb.append(j)
        w = np.array([f[j] ** 2 for j in nb])
        w = w / w.sum()
        for j, wj in zip(nb, w):
            out[j] += f[s] * wj
        out[s] = 0.0
    return out / out.sum()


def yn98(kappa=1., omega=1., codon_freq=None):
    """YN98 (Model/Codon/YN98.cpp:51-77) = CodonDistanceFrequenciesSubstitutionModel
    over K80: AbstractWordSubstitutionModel::fillBasicGenerator (:366-398) puts the
    K80 rate (after K80's own scaling 1/(kappa+2), Nucleotide/K80.cpp:64-95) x 1/3
    on every single-nucleotide change; completeMatrices
    (Codon/AbstractCodonSubstitutionModel.cpp:174-191) zeroes stop rows/cols and
    multiplies by omega-or-1 (AbstractCodonDistanceSubstitutionModel.cpp:80-88) and
    by the target codon frequency (AbstractCodonFrequenciesSubstitutionModel.cpp:82-85);
    then setDiagonal and AbstractSubstitutionModel::updateMatrices with freq_ fixed
    to the codon frequency set (computeFrequencies(false))."""
    aa = standard_genetic_code()
    if codon_freq is None:
        codon_freq = np.array([0.0 if aa[i] == "*" else 1.0 for i in range(64)])
        codon_freq /= codon_freq.sum()
    codon_freq = np.asarray(codon_freq, float)
    ts = {(0, 2), (2, 0), (1, 3), (3, 1)}
    Q = np.zeros((64, 64))
    for i in range(64):
        for j in range(64):
            if i == j:
                continue
            di = [(i // m) % 4 for m in (16, 4, 1)]
            dj = [(j // m) % 4 for m in (16, 4, 1)]
            diff = [p for p in range(3) if di[p] != dj[p]]
            if len(diff) != 1:
                continue
            p = diff[0]
            r = (kappa if (di[p], dj[p]) in ts else 1.0) / (kappa + 2.0)
            r *= 1.0 / 3.0
            if aa[i] == "*" or aa[j] == "*":
                r = 0.0
            else:
                r *= (1.0 if aa[i] == aa[j] else omega) * codon_freq[j]
            Q[i, j] = r
    Q = _set_diagonal(Q)
    m = Model("YN98", Q, codon_freq.copy(), reversible=False)
    return update_matrices(m, compute_freq=False)


# Grantham (1974) distances = bpp-seq GranthamAAChemicalDistance::getIndex (symmetric mode), the index GY94 is built on
# (Model/Codon/GY94.h:45,88).  bpp-seq is not under /root/reference: the published table is restated (Grantham's order, upper
# triangle by rows) and checked in tests against Grantham's own formula from composition / polarity / volume.
GRANTHAM_ORDER = "SRLPTAVGIFYCHQNKDEMW"
_GRANTHAM_UPPER = """
110 145 74 58 99 124 56 142 155 144 112 89 68 46 121 65 80 135 177
102 103 71 112 96 125 97 97 77 180 29 43 86 26 96 54 91 101
98 92 96 32 138 5 22 36 198 99 113 153 107 172 138 15 61
38 27 68 42 95 114 110 169 77 76 91 103 108 93 87 147
58 69 59 89 103 92 149 47 42 65 78 85 65 81 128
64 60 94 113 112 195 86 91 111 106 126 107 84 148
109 29 50 55 192 84 96 133 97 152 121 21 88
135 153 147 159 98 87 80 127 94 98 127 184
21 33 198 94 109 149 102 168 134 10 61
22 205 100 116 158 102 177 140 28 40
194 83 99 143 85 160 122 36 37
174 154 139 202 154 170 196 215
24 68 32 81 40 87 115
46 53 61 29 101 130
94 23 42 142 174
101 56 95 110
45 160 181
126 152
67
"""
# composition, polarity, molecular volume (Grantham 1974, table 1)
GRANTHAM_PROPERTIES = {"S": (1.42, 9.2, 32), "R": (0.65, 10.5, 124), "L": (0, 4.9, 111), "P": (0.39, 8.0, 32.5), "T": (0.71, 8.6, 61),
                       "A": (0, 8.1, 31), "V": (0, 5.9, 84), "G": (0.74, 9.0, 3), "I": (0, 5.2, 111), "F": (0, 5.2, 132),
                       "Y": (0.20, 6.2, 136), "C": (2.75, 5.5, 55), "H": (0.58, 10.4, 96), "Q": (0.89, 10.5, 85),
                       "N": (1.33, 11.6, 56), "K": (0.33, 11.3, 119), "D": (1.38, 13.0, 54), "E": (0.92, 12.3, 83),
                       "M": (0, 5.7, 105), "W": (0.13, 5.4, 170)}


def grantham_matrix():
    """dict (a, b) -> distance, both orders, zero diagonal"""
    rows = [list(map(float, line.split())) for line in _GRANTHAM_UPPER.strip().splitlines()]
    d = {}
    for i, row in enumerate(rows):
        for k, v in enumerate(row):
            a, b = GRANTHAM_ORDER[i], GRANTHAM_ORDER[i + 1 + k]
            d[a, b] = d[b, a] = v
    for a in GRANTHAM_ORDER:
        d[a, a] = 0.0
    return d


def gy94(kappa=1., V=10000., codon_freq=None):
    """GY94 (Model/Codon/GY94.cpp:49-71) = CodonDistanceFrequenciesSubstitutionModel over K80 with the Grantham index:
    like yn98() with the non-synonymous factor exp(-d(aa_i, aa_j) / V) (beta = gamma = 1,
    AbstractCodonDistanceSubstitutionModel.cpp:80-88; V is the model's `alpha`, GY94.cpp:66)."""
    aa = standard_genetic_code()
    dist = grantham_matrix()
    if codon_freq is None:
        codon_freq = np.array([0.0 if aa[i] == "*" else 1.0 for i in range(64)])
        codon_freq /= codon_freq.sum()
    codon_freq = np.asarray(codon_freq, float)
    ts = {(0, 2), (2, 0), (1, 3), (3, 1)}
    Q = np.zeros((64, 64))
    for i in range(64):
        for j in range(64):
            if i == j:
                continue
            di = [(i // m) % 4 for m in (16, 4, 1)]
            dj = [(j // m) % 4 for m in (16, 4, 1)]
            diff = [p for p in range(3) if di[p] != dj[p]]
            if len(diff) != 1 or aa[i] == "*" or aa[j] == "*":
                continue
            p = diff[0]
            r = (kappa if (di[p], dj[p]) in ts else 1.0) / (kappa + 2.0) / 3.0
            r *= (1.0 if aa[i] == aa[j] else math.exp(-dist[aa[i], aa[j]] / V)) * codon_freq[j]
            Q[i, j] = r
    Q = _set_diagonal(Q)
    m = Model("GY94", Q, codon_freq.copy(), reversible=False)
    return update_matrices(m, compute_freq=False)


def _simple_distribution_probs(thetas):
    """SimpleDiscreteDistribution's simplex parameters (bpp-core): p_1 = theta1, p_2 = (1 - theta1) theta2, ...,
    the last class takes the rest (YNGP_M2.cpp / RELAX.cpp: "theta1 = p0, theta2 = p1 / (p1 + p2)")."""
    probs, rest = [], 1.0
    for th in thetas:
        probs.append(rest * th)
        rest *= 1.0 - th
    return np.array(probs + [rest])


def _omega_mixture(kappa, omegas, probs, codon_freq=None):
    """MixtureOfASubstitutionModel over YN98's omega + the synonymous-rate homogenisation of the YNGP / RELAX wrappers
    (YNGP_M2::updateMatrices, Model/Codon/YNGP_M2.cpp:134-146; RELAX.cpp:212-218): sub-model k gets the relative rate
    1 / Q_k(synfrom, synto) -- the first synonymous pair with a non-zero rate, AAG -> AAA in the 64-state order -- and
    MixtureOfASubstitutionModel::setVRates normalises the rates to mean 1 under the class probabilities."""
    models = [yn98(kappa, w, codon_freq) for w in omegas]
    rates = np.array([1.0 / m.Q[2, 0] for m in models])
    rates = rates / float(np.dot(probs, rates))
    for m, r in zip(models, rates):
        m.rate = float(r)
    return models, np.asarray(probs, float)


def yngp_m2(kappa=1., omega0=0.5, omega2=2., theta1=0.333333, theta2=0.5, codon_freq=None):
    """YNGP_M2 (Model/Codon/YNGP_M2.cpp:52-146): three YN98 with omega in {omega0 < 1, 1, omega2 > 1}."""
    return _omega_mixture(kappa, [omega0, 1.0, omega2], _simple_distribution_probs([theta1, theta2]), codon_freq)


def relax(kappa=1., p=0.5, omega1=1., omega2=2., k=1., theta1=0.333333, theta2=0.5, codon_freq=None):
    """RELAX (fork; Model/Codon/RELAX.cpp:52-218): omegas ((p omega1)^k, omega1^k, omega2^k), the first two floored at 0.001 and
    the last capped at 999 (:176-205)."""
    w0 = max((p * omega1) ** k, 0.001)
    w1 = max(omega1 ** k, 0.001)
    w2 = min(omega2 ** k, 999.0)
    return _omega_mixture(kappa, [w0, w1, w2], _simple_distribution_probs([theta1, theta2]), codon_freq)


# ---- chromosome number (fork) -------------------------------------------------
IGNORE_PARAM = -999.0      # Model/ChromosomeSubstitutionModel.h:15-23
DEMI_EQUAL_DUPL = -2.0


def chromosome(min_chr, max_chr, gain=0.0, loss=0.0, dupl=0.0, demi=IGNORE_PARAM,
               gain_r=IGNORE_PARAM, loss_r=IGNORE_PARAM, dupl_r=IGNORE_PARAM,
               base_num=IGNORE_PARAM, base_num_r=IGNORE_PARAM, max_chr_range=0,
               rate_change="LINEAR"):
    """ChromosomeSubstitutionModel::updateMatrices
    (Model/ChromosomeSubstitutionModel.cpp:431-469, updateQWith* :471-577,
    getRate :504-526) followed by updateEigenMatrices (:589-802).  Not
    normalised (isScalable_=false, :58)."""
    n = max_chr - min_chr + 1
    Q = np.zeros((n, n))

    def rate(i, const, lin):
        # getRate (:504-526); callers skip when both are IgnoreParam (:529-557)
        if const == IGNORE_PARAM and lin == IGNORE_PARAM:
            return 0.0
        total = lin if const == IGNORE_PARAM else const
        if lin == IGNORE_PARAM:
            return total
        if rate_change == "LINEAR":
            return total + lin * (i - 1)
        return total * math.exp(lin * (i - 1))

    if demi == DEMI_EQUAL_DUPL:
        demi_v = dupl
    else:
        demi_v = demi
    for i in range(min_chr, max_chr + 1):
        r = i - min_chr
        if i + 1 <= max_chr:
            Q[r, r + 1] += rate(i, gain, gain_r)
        if i - 1 >= min_chr:
            Q[r, r - 1] += rate(i, loss, loss_r)
        if 2 * i <= max_chr:
            Q[r, 2 * i - min_chr] += rate(i, dupl, dupl_r)
        elif i != max_chr:
            Q[r, max_chr - min_chr] += rate(i, dupl, dupl_r)
        # demi-polyploidy (:533-560)
        if demi_v != IGNORE_PARAM and i != max_chr:
            if i % 2 == 0 and int(i * 1.5) <= max_chr:
                Q[r, int(i * 1.5) - min_chr] += demi_v
            elif i % 2 != 0 and math.ceil(i * 1.5) <= max_chr:
                if i == 1:
                    Q[r, math.ceil(i * 1.5) - min_chr] += demi_v
                else:
                    Q[r, math.ceil(i * 1.5) - min_chr] += demi_v / 2
                    Q[r, math.floor(i * 1.5) - min_chr] += demi_v / 2
            else:
                Q[r, max_chr - min_chr] += demi_v
        if i < max_chr and base_num != IGNORE_PARAM:
            # base-number transitions (:562-577)
            for j in range(i + 1, max_chr + 1):
                if j == max_chr:
                    if j - i <= max_chr_range:
                        Q[r, j - min_chr] += base_num_r
                elif (j - i) % int(base_num) == 0 and j - i <= max_chr_range:
                    Q[r, j - min_chr] += base_num_r
    Q = _set_diagonal(Q)
    m = Model("Chromosome", Q, np.full(n, 1.0 / n), reversible=False, scalable=False, chromosome=True)
    update_matrices_chromosome(m)
    return m


def update_matrices_chromosome(m: Model):
    """ChromosomeSubstitutionModel::updateEigenMatrices (:589-802): like the generic
    one but null lines are rows with |Q_ii| < TINY only (:604-613), the zero-eigen
    search has no tolerance ladder (:708-711) and compares eigenvector entries
    absolutely (:735), no frequencies, no normalisation, and 30 powers of Q are
    ALWAYS tabulated (:787-799)."""
    Q = m.Q
    n = Q.shape[0]
    vnull = np.abs(np.diag(Q)) < TINY
    ok = np.where(~vnull)[0]
    nstop = int(vnull.sum())
    d, e, Vk = _real_block_eig(Q[np.ix_(ok, ok)])
    if nstop:
        ev_re = np.concatenate([d, np.zeros(nstop)])
        ev_im = np.concatenate([e, np.zeros(nstop)])
        V = np.zeros((n, n))
        V[np.ix_(ok, np.arange(len(ok)))] = Vk
        gi = 0
        for i in range(n):
            if vnull[i]:
                gi += 1
                V[i, len(ok) + gi - 1] = 1.0
    else:
        ev_re, ev_im, V = d, e, Vk
    m.ev_re, m.ev_im, m.V = ev_re, ev_im, V
    try:
        m.Vinv = np.linalg.inv(V)
        if not np.all(np.isfinite(m.Vinv)) or np.linalg.cond(V) > 1e15:
            raise np.linalg.LinAlgError
        m.diagonalizable = not np.any(np.abs(ev_im) > TINY)
        nullev = [i for i in range(n - nstop) if abs(ev_re[i]) < SMALL and abs(ev_im[i]) < SMALL]
        m.nonsingular = len(nullev) == 1
        nulleigen = nullev[0] if m.nonsingular else None
        if not m.nonsingular:
            first = int(ok[0])
            for c in nullev:
                val = V[first, c]
                if all(abs(V[i, c] - val) <= SMALL for i in ok[1:]):
                    m.nonsingular = True
                    nulleigen = c
                    break
        if m.nonsingular:
            ev_re[nulleigen] = 0.0
            ev_im[nulleigen] = 0.0
        else:
            m.diagonalizable = False
    except np.linalg.LinAlgError:
        m.nonsingular = False
        m.diagonalizable = False
    m.pow_gen = _taylor_powers(Q, 30)
    m.first_norm = float(np.abs(Q).sum())      # getFirstNorm (:921-930)
    return m


# --------------------------------------------------------------------------
# P(t) family
# --------------------------------------------------------------------------
def _block_exp_factor(m: Model, l: float, order: int):
    """Tridiagonal factor T(l) with V.T.Vinv = d^order/dt^order exp(Q t) up to the
    rate factor, following the complex-pair block form of
    AbstractSubstitutionModel::getPij_t (:438-468), getdPij_dt (:507-537),
    getd2Pij_dt2 (:578-612)."""
    n = m.size
    T = np.zeros((n, n))
    i = 0
    while i < n:
        a, b = m.ev_re[i], m.ev_im[i]
        ex = math.exp(a * l)
        if b != 0.0:
            s, c = math.sin(b * l), math.cos(b * l)
            if order == 0:
                dia, up = ex * c, ex * s
            elif order == 1:
                dia = m.rate * (a * c - b * s) * ex
                up = m.rate * (a * s + b * c) * ex
            else:
                dia = m.rate ** 2 * ((a * a - b * b) * c - 2 * a * b * s) * ex
                up = m.rate ** 2 * ((a * a - b * b) * s + 2 * a * b * c) * ex
            T[i, i] = T[i + 1, i + 1] = dia
            T[i, i + 1] = up
            T[i + 1, i] = -up
            i += 2
        else:
            T[i, i] = ex * (1.0 if order == 0 else (m.rate * a) ** order)
            i += 1
    return T


def _taylor_expm(m: Model, t: float):
    """Singular branch of getPij_t (:470-492): sum_{k<30} v^k/k! Q^k, v = rate*t
    halved m times until <= 0.5, result squared m times."""
    n = m.size
    P = np.eye(n)
    s = 1.0
    v = m.rate * t
    k = 0
    while v > 0.5:
        k += 1
        v /= 2
    for i in range(1, len(m.pow_gen)):
        s *= v / i
        P = P + s * m.pow_gen[i]
    for _ in range(k):
        P = P @ P
    return P


def pij_t(m: Model, t: float):
    """getPij_t: Model/AbstractSubstitutionModel.cpp:426-495;
    Chromosome override Model/ChromosomeSubstitutionModel.cpp:808-920."""
    if m.chromosome:
        return _chr_pij_t(m, t, from_deriv=False)
    n = m.size
    if t == 0:
        return np.eye(n)
    if m.nonsingular:
        if m.diagonalizable:
            return (m.V * np.exp(m.ev_re * (m.rate * t))[None, :]) @ m.Vinv   # :436
        return m.V @ _block_exp_factor(m, m.rate * t, 0) @ m.Vinv            # :438-468
    return _taylor_expm(m, t)


def dpij_dt(m: Model, t: float):
    """getdPij_dt: :499-566;  Chromosome: P.Q.rate (:966-982)."""
    if m.chromosome:
        return _chr_pij_t(m, t, True) @ m.Q * m.rate
    if m.nonsingular:
        if m.diagonalizable:
            lam = m.rate * m.ev_re
            return (m.V * (lam * np.exp(lam * t))[None, :]) @ m.Vinv          # :505
        return m.V @ _block_exp_factor(m, m.rate * t, 1) @ m.Vinv
    # singular: rate * Q * P  (:539-563: derivative of the Taylor series, squared back)
    return m.rate * (m.Q @ _taylor_expm(m, t))


def d2pij_dt2(m: Model, t: float):
    """getd2Pij_dt2: :570-641;  Chromosome: Q^2.P.rate^2 (:986-1001)."""
    if m.chromosome:
        return (m.pow_gen[2] @ _chr_pij_t(m, t, True)) * m.rate ** 2
    if m.nonsingular:
        if m.diagonalizable:
            lam = m.rate * m.ev_re
            return (m.V * (lam * lam * np.exp(lam * t))[None, :]) @ m.Vinv    # :576
        return m.V @ _block_exp_factor(m, m.rate * t, 2) @ m.Vinv
    return m.rate ** 2 * (m.Q @ m.Q @ _taylor_expm(m, t))


def _chr_pij_t(m: Model, t: float, from_deriv: bool, epsilon: float = 1e-4):
    """ChromosomeSubstitutionModel::getPij_t (:808-920): eigen / block form as the
    generic model; otherwise the adaptive Taylor variant (L1-norm scaling until
    <= 0.5, terms added until all |dP_ij| <= epsilon and P within
    [-epsilon, 1+epsilon], :852-899, :934-946).  Post-clamp P<0 -> 1e-20,
    P>1 -> 1 unless called from a derivative (:903-916)."""
    n = m.size
    if t == 0:
        P = np.eye(n)
    elif m.nonsingular:
        if m.diagonalizable:
            P = (m.V * np.exp(m.ev_re * (m.rate * t))[None, :]) @ m.Vinv
        else:
            P = m.V @ _block_exp_factor(m, m.rate * t, 0) @ m.Vinv
    else:
        v = m.rate * t
        norm = v * m.first_norm
        k = 0
        while norm > 0.5:
            k += 1
            v /= 2
            norm /= 2
        pw = list(m.pow_gen)
        prev = np.eye(n)
        P = prev
        it = 2
        while True:
            while len(pw) <= it:
                pw.append(pw[-1] @ m.Q)
            P = np.eye(n)
            s = 1.0
            for i in range(1, it + 1):         # calculateExp_Qt (:948-962)
                s *= v / i
                P = P + s * pw[i]
            if it > 2:
                ok = np.all(np.abs(P - prev) <= epsilon) and np.all(P + epsilon >= 0) and np.all(P <= 1 + epsilon)
                if ok:
                    break
            prev = P
            if it > 250:
                raise RuntimeError("ChromosomeSubstitutionModel: Taylor series did not reach convergence!")
            it += 1
        for _ in range(k):
            P = P @ P
    if not from_deriv:
        P = P.copy()
        P[P < 0] = VERY_TINY
        P[P > 1] = 1.0
    return P


def transition_tables(m: Model, brlens, rates, want_d1=False, want_d2=False):
    """computeAllTransitionProbabilities / computeTransitionProbabilitiesForNode
    (Likelihood/AbstractHomogeneousTreeLikelihood.cpp:341-414): pxy[b][c] =
    P(l_b r_c), dpxy = r_c P'(l_b r_c), d2pxy = r_c^2 P''(l_b r_c)."""
    B, C, S = len(brlens), len(rates), m.size
    P = np.empty((B, C, S, S))
    dP = np.empty((B, C, S, S)) if want_d1 else None
    d2P = np.empty((B, C, S, S)) if want_d2 else None
    for b, l in enumerate(brlens):
        for c, r in enumerate(rates):
            P[b, c] = pij_t(m, l * r)
            if want_d1:
                dP[b, c] = r * dpij_dt(m, l * r)
            if want_d2:
                d2P[b, c] = r * r * d2pij_dt2(m, l * r)
    return P, dP, d2P
