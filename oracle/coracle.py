"""ORACLE (test infrastructure, not product): ctypes access to oracle/liboracle.so (plain-C restatement,
cv_oracle.c) and oracle/librefcpu.so (pass-structured C++14/OpenMP port, ref_cpu.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
u8p = C.POINTER(C.c_uint8)
f64p = C.POINTER(C.c_double)


class Params(C.Structure):
    _fields_ = [("mu", C.c_double), ("nu", C.c_double), ("dt", C.c_double), ("eps", C.c_double),
                ("lambda1", C.c_double * 3), ("lambda2", C.c_double * 3)]


def build():
    """make -C oracle (gcc -O2 -ffp-contract=off; g++ with the reference's flags)."""
    srcs = [os.path.join(HERE, f) for f in ("cv_oracle.c", "ref_cpu.cpp", "Makefile")]
    libs = [os.path.join(HERE, f) for f in ("liboracle.so", "librefcpu.so")]
    if all(os.path.exists(l) for l in libs) and min(map(os.path.getmtime, libs)) >= max(map(os.path.getmtime, srcs)):
        return
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run(["make", "-C", HERE, "-s"], check=True, env=env)


_o = None
_r = None


def oracle():
    global _o
    if _o is None:
        build()
        o = C.CDLL(os.path.join(HERE, "liboracle.so"))
        o.cvo_heaviside.restype = C.c_double
        o.cvo_heaviside.argtypes = [C.c_double, C.c_double]
        o.cvo_delta.restype = C.c_double
        o.cvo_delta.argtypes = [C.c_double, C.c_double]
        o.cvo_levelset_checkerboard.argtypes = [C.c_int, C.c_int, f64p]
        o.cvo_levelset_rect.argtypes = [C.c_int] * 6 + [f64p]
        o.cvo_levelset_circ.argtypes = [C.c_int] * 5 + [f64p]
        o.cvo_region_variance.restype = C.c_double
        o.cvo_region_variance.argtypes = [u8p, f64p, C.c_int, C.c_int, C.c_int, C.c_double]
        o.cvo_curvature.argtypes = [f64p, C.c_int, C.c_int, f64p]
        o.cvo_pm_num_steps.restype = C.c_int
        o.cvo_pm_num_steps.argtypes = [C.c_double, C.c_double]
        o.cvo_pm_evolve.argtypes = [f64p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        o.cvo_perona_malik.restype = C.c_int
        o.cvo_perona_malik.argtypes = [C.POINTER(u8p), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                       C.POINTER(u8p)]
        o.cvo_stop_condition.restype = C.c_double
        o.cvo_stop_condition.argtypes = [C.POINTER(u8p), C.c_int, C.c_int, C.c_int, C.c_double]
        o.cvo_csv_step_ex.restype = C.c_double
        o.cvo_csv_step_ex.argtypes = [C.POINTER(u8p), C.c_int, C.c_int, C.c_int, f64p, C.POINTER(Params), f64p, f64p,
                                      f64p, f64p]
        o.cvo_csv_run.restype = C.c_int
        o.cvo_csv_run.argtypes = [C.POINTER(u8p), C.c_int, C.c_int, C.c_int, f64p, C.POINTER(Params), C.c_double,
                                  C.c_int, f64p]
        o.cvo_mask.argtypes = [f64p, C.c_int, C.c_int, C.c_int, u8p]
        o.cvo_delta_map.argtypes = [f64p, C.c_size_t, C.c_double]
        _o = o
    return _o


def refcpu():
    global _r
    if _r is None:
        build()
        r = C.CDLL(os.path.join(HERE, "librefcpu.so"))
        r.refcpu_perona_malik.restype = C.c_int
        r.refcpu_perona_malik.argtypes = [C.POINTER(u8p), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                          C.POINTER(u8p)]
        r.refcpu_stop_condition.restype = C.c_double
        r.refcpu_stop_condition.argtypes = [C.POINTER(u8p), C.c_int, C.c_int, C.c_int, C.c_double]
        r.refcpu_csv_run.restype = C.c_int
        r.refcpu_csv_run.argtypes = [C.POINTER(u8p), C.c_int, C.c_int, C.c_int, f64p, C.POINTER(Params), C.c_double,
                                     C.c_int, f64p]
        _r = r
    return _r


def _pl(channels):
    arrs = [np.ascontiguousarray(c, dtype=np.uint8) for c in channels]
    return arrs, (u8p * len(arrs))(*[a.ctypes.data_as(u8p) for a in arrs])


def _f(a):
    return a.ctypes.data_as(f64p)


def params(mu=0.5, nu=0.0, dt=1.0, eps=1.0, lambda1=None, lambda2=None, nch=3):
    p = Params()
    p.mu, p.nu, p.dt, p.eps = mu, nu, dt, eps
    for k in range(3):
        p.lambda1[k] = lambda1[k] if lambda1 is not None and k < len(lambda1) else 1.0
        p.lambda2[k] = lambda2[k] if lambda2 is not None and k < len(lambda2) else 1.0
    return p


def levelset_checkerboard(h, w):
    u = np.empty((h, w))
    oracle().cvo_levelset_checkerboard(h, w, _f(u))
    return u


def levelset_rect(h, w, x, y, rw, rh):
    u = np.empty((h, w))
    oracle().cvo_levelset_rect(h, w, x, y, rw, rh, _f(u))
    return u


def levelset_circ(h, w, cx, cy, r):
    u = np.empty((h, w))
    oracle().cvo_levelset_circ(h, w, cx, cy, r, _f(u))
    return u


def region_variance(img, u, inside, eps=1.0):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    u = np.ascontiguousarray(u, dtype=np.float64)
    h, w = u.shape
    return oracle().cvo_region_variance(img.ctypes.data_as(u8p), _f(u), h, w, int(inside), eps)


def curvature(u):
    u = np.ascontiguousarray(u, dtype=np.float64)
    k = np.empty_like(u)
    oracle().cvo_curvature(_f(u), u.shape[0], u.shape[1], _f(k))
    return k


def pm_evolve(I, K, L, nsteps):
    I = np.array(I, dtype=np.float64, order="C", copy=True)
    oracle().cvo_pm_evolve(_f(I), I.shape[0], I.shape[1], K, L, nsteps)
    return I


def perona_malik(channels, K, L, T, impl="oracle"):
    arrs, ptrs = _pl(channels)
    h, w = arrs[0].shape
    outs = [np.empty((h, w), dtype=np.uint8) for _ in arrs]
    optrs = (u8p * len(arrs))(*[a.ctypes.data_as(u8p) for a in outs])
    fn = oracle().cvo_perona_malik if impl == "oracle" else refcpu().refcpu_perona_malik
    n = fn(ptrs, len(arrs), h, w, K, L, T, optrs)
    return outs, n


def stop_condition(channels, tol):
    arrs, ptrs = _pl(channels)
    h, w = arrs[0].shape
    return oracle().cvo_stop_condition(ptrs, len(arrs), h, w, tol)


def csv_step(channels, u, p, c1=None, c2=None):
    """One step; returns (u_new, norm, c1_used, c2_used)."""
    arrs, ptrs = _pl(channels)
    h, w = arrs[0].shape
    u = np.array(u, dtype=np.float64, order="C", copy=True)
    c1o, c2o = np.zeros(3), np.zeros(3)
    c1i = _f(np.ascontiguousarray(c1, dtype=np.float64)) if c1 is not None else None
    c2i = _f(np.ascontiguousarray(c2, dtype=np.float64)) if c2 is not None else None
    nrm = oracle().cvo_csv_step_ex(ptrs, len(arrs), h, w, _f(u), C.byref(p), c1i, c2i, _f(c1o), _f(c2o))
    return u, nrm, c1o[:len(arrs)], c2o[:len(arrs)]


def csv_run(channels, u, p, tol, max_steps, impl="oracle"):
    arrs, ptrs = _pl(channels)
    h, w = arrs[0].shape
    u = np.array(u, dtype=np.float64, order="C", copy=True)
    nrm = C.c_double(0.0)
    fn = oracle().cvo_csv_run if impl == "oracle" else refcpu().refcpu_csv_run
    steps = fn(ptrs, len(arrs), h, w, _f(u), C.byref(p), tol, max_steps, C.byref(nrm))
    return u, steps, nrm.value


def mask(u, invert=False):
    u = np.ascontiguousarray(u, dtype=np.float64)
    m = np.empty(u.shape, dtype=np.uint8)
    oracle().cvo_mask(_f(u), u.shape[0], u.shape[1], int(bool(invert)), m.ctypes.data_as(u8p))
    return m


def delta_map(data, eps=1.0):
    d = np.array(data, dtype=np.float64, order="C", copy=True)
    oracle().cvo_delta_map(_f(d), d.size, eps)
    return d
