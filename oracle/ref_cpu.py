"""ctypes binding of oracle/ref_cpu.cpp (TEST / BASELINE INFRASTRUCTURE).

``eval_case`` runs the C++ restatement (reference loop nest and nested-vector layout) and returns lnL,
derivatives and the best wall time of ``reps`` evaluations; bench.py times it as the CPU baseline.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
SO = HERE / "_build" / "libref_cpu.so"


def build(force=False):
    src = HERE / "ref_cpu.cpp"
    if force or not SO.exists() or SO.stat().st_mtime < src.stat().st_mtime:
        SO.parent.mkdir(exist_ok=True)
        # -O2 -g = the reference's default RelWithDebInfo (CMakeLists.txt:14-18)
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++11", "-fPIC", "-shared", "-pthread", "-o", str(SO), str(src)])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def eval_raw(S, Ccat, N, child_off, children, root, codes, code_table, weights, rates, probs, V, Vinv, ev,
             model_rate, brlen, rootfreq, scaled=True, want=1, nthreads=1, reps=1, site=False, ev_im=None, chr_clamp=False,
             weighted_root=False):
    """codes: [n_leaves][N] rows in increasing leaf node id."""
    f64 = lambda a: np.ascontiguousarray(a, np.float64)
    child_off = np.ascontiguousarray(child_off, np.int32)
    children = np.ascontiguousarray(children, np.int32)
    codes = np.ascontiguousarray(codes)
    assert codes.dtype in (np.uint8, np.uint16)
    code_table, rates, probs, V, Vinv, ev, brlen, rootfreq = map(f64, (code_table, rates, probs, V, Vinv, ev, brlen, rootfreq))
    weights = np.ascontiguousarray(weights, np.uint32)
    nn = len(child_off) - 1
    if ev_im is not None:
        ev_im = f64(ev_im)
        if not np.any(ev_im):
            ev_im = None
    lnl = C.c_double(0)
    sec = C.c_double(0)
    tot = C.c_double(0)
    d1 = np.zeros(nn) if want & 6 else None
    d2 = np.zeros(nn) if want & 4 else None
    sl = np.zeros(N) if site else None
    rc = lib().refcpu_eval(C.c_int(S), C.c_int(Ccat), C.c_long(N), C.c_int(nn), C.c_int(root), _p(child_off, C.c_int),
                           _p(children, C.c_int), codes.ctypes.data_as(C.c_void_p), C.c_int(codes.dtype.itemsize),
                           C.c_int(code_table.shape[0]), _p(code_table), _p(weights, C.c_uint), _p(rates), _p(probs),
                           _p(V), _p(Vinv), _p(ev), _p(ev_im), C.c_int(int(chr_clamp)), C.c_int(int(weighted_root)),
                           C.c_double(model_rate), _p(brlen), _p(rootfreq),
                           C.c_int(int(scaled)), C.c_int(want), C.c_int(nthreads), C.c_int(reps), C.byref(lnl),
                           _p(d1), _p(d2), _p(sl), C.byref(sec), C.byref(tot))
    assert rc == 0
    return {"lnl": lnl.value, "d1": d1, "d2": d2, "site_lnl": sl, "seconds": sec.value, "total_seconds": tot.value}


def eval_blocks(S, Ccat, N, block, child_off, children, root, codes, code_table, weights, rates, probs, V, Vinv, ev, model_rate,
                brlen, rootfreq, scaled=True, nthreads=1, ev_im=None, chr_clamp=False):
    """Value-only evaluation of a large alignment in blocks of `block` patterns (arrays allocated once): (lnL, seconds)."""
    f64 = lambda a: np.ascontiguousarray(a, np.float64)
    child_off = np.ascontiguousarray(child_off, np.int32)
    children = np.ascontiguousarray(children, np.int32)
    codes = np.ascontiguousarray(codes)
    assert codes.dtype in (np.uint8, np.uint16) and codes.shape[1] == N
    code_table, rates, probs, V, Vinv, ev, brlen, rootfreq = map(f64, (code_table, rates, probs, V, Vinv, ev, brlen, rootfreq))
    weights = np.ascontiguousarray(weights, np.uint32)
    nn = len(child_off) - 1
    if ev_im is not None:
        ev_im = f64(ev_im)
        if not np.any(ev_im):
            ev_im = None
    lnl, tot = C.c_double(0), C.c_double(0)
    rc = lib().refcpu_eval_blocks(C.c_int(S), C.c_int(Ccat), C.c_long(N), C.c_long(block), C.c_int(nn), C.c_int(root),
                                  _p(child_off, C.c_int), _p(children, C.c_int), codes.ctypes.data_as(C.c_void_p),
                                  C.c_int(codes.dtype.itemsize), C.c_int(code_table.shape[0]), _p(code_table), _p(weights, C.c_uint),
                                  _p(rates), _p(probs), _p(V), _p(Vinv), _p(ev), _p(ev_im), C.c_int(int(chr_clamp)),
                                  C.c_double(model_rate), _p(brlen), _p(rootfreq), C.c_int(int(scaled)), C.c_int(nthreads),
                                  C.byref(lnl), C.byref(tot))
    assert rc == 0, rc
    return lnl.value, tot.value


def eval_case(case, **kw):
    """``case`` as built by tests/cases.py (diagonalisable models only)."""
    flat, m = case.flat, case.model
    assert m.nonsingular
    kw.setdefault("ev_im", m.ev_im)
    kw.setdefault("chr_clamp", m.chromosome)
    off, ch = flat.csr()
    leaf_ids = [i for i in range(flat.n_nodes) if flat.is_leaf[i]]
    codes = np.stack([case.codes_by_leaf[l] for l in leaf_ids]) if case.N else np.zeros((len(leaf_ids), 0), np.uint8)
    return eval_raw(m.size, len(case.rates), case.N, off, ch, flat.root, codes, case.table, case.weights, case.rates,
                    case.probs, m.V, m.Vinv, m.ev_re, m.rate, flat.brlen, case.root_freqs, **kw)
