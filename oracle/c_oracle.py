"""ctypes loader for oracle/build/liboracle.so (the plain-C restatement, swrt_oracle.c).

TEST INFRASTRUCTURE ONLY (see swrt_oracle.c header): used by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
SO = _HERE / "build" / "liboracle.so"
_dp = C.POINTER(C.c_double)
_lib = None


def build(force=False):
    src = _HERE / "swrt_oracle.c"
    if SO.exists() and not force and SO.stat().st_mtime >= src.stat().st_mtime:
        return SO
    SO.parent.mkdir(exist_ok=True)
    subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-o", str(SO), str(src), "-lm"])
    return SO


def lib():
    global _lib
    if _lib is None:
        if not SO.exists():
            build()
        _lib = C.CDLL(str(SO))
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def _cm(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64)).ravel(order="F").copy()


def _table(arrs):
    return (_dp * len(arrs))(*[_p(a) for a in arrs])


def num_threads():
    return int(lib().orc_num_threads())


def set_threads(n):
    lib().orc_set_threads(C.c_int(int(n)))


def interpolate(x, y, F, dx, dy, bump=1e-13):
    x = _f(x).ravel(); y = _f(y).ravel()
    nx, ny = np.asarray(F).shape
    Ff = _cm(F); out = np.empty(x.size)
    lib().orc_interpolate(_p(x), _p(y), C.c_int64(x.size), _p(Ff), C.c_int(nx), C.c_int(ny), C.c_double(dx),
                          C.c_double(dy), C.c_double(bump), _p(out))
    return out


def interpolate6(x, y, grids, dx, bump=1e-13):
    x = _f(x).ravel(); y = _f(y).ravel()
    nx = np.asarray(grids[0]).shape[0]
    gs = [_cm(g) for g in grids]
    outs = [np.empty(x.size) for _ in grids]
    lib().orc_interpolate6(_p(x), _p(y), C.c_int64(x.size), _table(gs), C.c_int(len(gs)), C.c_int(nx), C.c_double(dx),
                           C.c_double(bump), _table(outs))
    return np.stack(outs)


def leapfrog_lagrange(x, y, k, l, grids, dx, f, gH, dt, nsteps, bump=1e-13):
    x, y, k, l = (_f(a).copy() for a in (x, y, k, l))
    nx = np.asarray(grids[0]).shape[0]
    gs = [_cm(g) for g in grids[:6]]
    lib().orc_leapfrog_lagrange(_p(x), _p(y), _p(k), _p(l), C.c_int64(x.size), _table(gs), C.c_int(nx), C.c_double(dx),
                                C.c_double(bump), C.c_double(f), C.c_double(gH), C.c_double(dt), C.c_int(nsteps))
    return x, y, k, l


def leapfrog_lagrange2(x, y, k, l, grids1, grids2, dx, f, gH, dt, nsteps, alpha0, dalpha, bump=1e-13, prepared=None):
    """two-frame leapfrog (interpolate_U.m blend per plane, alpha_j = alpha0 + j*dalpha); ``prepared`` = the
    column-major copies a previous call returned (so a timing loop does not re-copy 12 planes per call)"""
    x, y, k, l = (_f(a).copy() for a in (x, y, k, l))
    nx = np.asarray(grids1[0]).shape[0]
    g1, g2 = prepared if prepared is not None else ([_cm(g) for g in grids1[:6]], [_cm(g) for g in grids2[:6]])
    lib().orc_leapfrog_lagrange2(_p(x), _p(y), _p(k), _p(l), C.c_int64(x.size), _table(g1), _table(g2), C.c_int(nx), C.c_double(dx),
                                 C.c_double(bump), C.c_double(f), C.c_double(gH), C.c_double(dt), C.c_int(nsteps),
                                 C.c_double(alpha0), C.c_double(dalpha))
    return x, y, k, l


def prepare_grids(grids):
    return [_cm(g) for g in grids[:6]]


def _planes(planes_k):
    res = [_cm(np.asarray(p).real) for p in planes_k]
    ims = [_cm(np.asarray(p).imag) for p in planes_k]
    return res, ims


def spectral_eval(x, y, planes_k, dx, nx, precise=True):
    x = _f(x).ravel(); y = _f(y).ravel()
    res, ims = _planes(planes_k)
    outs = [np.empty(x.size) for _ in planes_k]
    lib().orc_spectral_eval(_p(x), _p(y), C.c_int64(x.size), _table(res), _table(ims), C.c_int(len(res)), C.c_int(nx),
                            C.c_double(dx), C.c_int(int(precise)), _table(outs))
    return np.stack(outs)


def leapfrog_spectral(x, y, k, l, planes_k, dx, nx, f, gH, dt, nsteps, precise=True):
    x, y, k, l = (_f(a).copy() for a in (x, y, k, l))
    res, ims = _planes(planes_k[:6])
    lib().orc_leapfrog_spectral(_p(x), _p(y), _p(k), _p(l), C.c_int64(x.size), _table(res), _table(ims), C.c_int(nx),
                                C.c_double(dx), C.c_int(int(precise)), C.c_double(f), C.c_double(gH), C.c_double(dt),
                                C.c_int(nsteps))
    return x, y, k, l


def rk4_lagrange(x, y, k, l, a, grids, dx, f, C0, dt, nsteps, xka, bump=1e-13):
    x, y, k, l = (_f(v).copy() for v in (x, y, k, l))
    a = _f(a).copy() if a is not None else np.ones_like(x)
    nx = np.asarray(grids[0]).shape[0]
    gs = [_cm(g) for g in grids]
    if len(gs) == 6:
        gs.append(np.ones_like(gs[0]))
    lib().orc_rk4_lagrange(_p(x), _p(y), _p(k), _p(l), _p(a), C.c_int64(x.size), _table(gs), C.c_int(nx), C.c_double(dx),
                           C.c_double(bump), C.c_double(f), C.c_double(C0), C.c_double(dt), C.c_int(nsteps), C.c_int(int(xka)))
    return x, y, k, l, a


def rhs(k, l, e6, f, Cg):
    k = _f(k); l = _f(l)
    es = [_f(e) for e in e6]
    outs = [np.empty(k.size) for _ in range(4)]
    lib().orc_rhs(_p(k), _p(l), C.c_int64(k.size), _table(es), C.c_double(f), C.c_double(Cg), *[_p(o) for o in outs])
    return tuple(outs)


def histcounts(w, edges):
    w = _f(w).ravel(); edges = _f(edges)
    counts = np.zeros(edges.size - 1, dtype=np.uint64)
    lib().orc_histcounts(_p(w), C.c_int64(w.size), _p(edges), C.c_int(edges.size), counts.ctypes.data_as(C.POINTER(C.c_uint64)))
    return counts
