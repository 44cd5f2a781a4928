"""ctypes binding of oracle/gp_oracle.c (plain-C oracle).  TEST INFRASTRUCTURE ONLY.

``build()`` compiles libgp_oracle.so next to the source with gcc; ``use_openblas()`` hands the
OpenBLAS bundled with SciPy to the C oracle (dpotrf / dtrtrs) so that the CPU *baseline timing*
runs the reference's own LAPACK path (Julia -> OpenBLAS dpotrf/dtrtrs).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

from .gp_oracle import Op

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgp_oracle.so")
_lib = None
_blas = None


class COp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("col", C.c_int32), ("theta_slot", C.c_int32), ("var_slot", C.c_int32),
                ("value", C.c_double), ("var", C.c_double)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-fopenmp", "-std=c11", "-ffp-contract=off", "-shared",
                               "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.gpo_lml.restype = C.c_double
    return _lib


def use_openblas(threads: int = 1) -> str | None:
    """Route the C oracle's Cholesky / triangular solve through SciPy's bundled OpenBLAS."""
    global _blas
    import scipy
    cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
    if not cands:
        return None
    _blas = C.CDLL(cands[0], mode=C.RTLD_GLOBAL)
    _blas.scipy_openblas_set_num_threads(C.c_int(threads))
    lib().gpo_set_lapack(C.cast(_blas.scipy_dpotrf_, C.c_void_p), C.cast(_blas.scipy_dtrtrs_, C.c_void_p))
    try:
        _blas.scipy_openblas_get_config.restype = C.c_char_p
        return _blas.scipy_openblas_get_config().decode()
    except Exception:
        return os.path.basename(cands[0])


def use_plain_c() -> None:
    lib().gpo_set_lapack(None, None)


def _ops(ops):
    arr = (COp * len(ops))()
    for i, o in enumerate(ops):
        arr[i] = COp(o.kind, o.col, o.theta_slot, o.var_slot, o.value, o.var)
    return arr


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def cov(ops, X, theta, sigma2, jitter=0.0, mode=0):
    X = _f(np.asarray(X, dtype=np.float64).reshape(len(X), -1))
    n, d = X.shape
    th = _f(np.atleast_1d(theta))
    K = np.empty((n, n), order="F")
    rc = lib().gpo_cov(_ops(ops), len(ops), n, d, _p(X), _p(th), C.c_double(sigma2), C.c_double(jitter), mode, _p(K))
    if rc:
        raise ValueError("malformed program")
    return K


def lml(ops, X, y, theta, sigma2, jitter=0.0, mode=0):
    X = _f(np.asarray(X, dtype=np.float64).reshape(len(X), -1))
    n, d = X.shape
    y, th = _f(y), _f(np.atleast_1d(theta))
    info = C.c_int(0)
    v = lib().gpo_lml(_ops(ops), len(ops), n, d, _p(X), _p(y), _p(th), C.c_double(sigma2), C.c_double(jitter), mode,
                      C.byref(info))
    return v, info.value


def lml_batched(ops, X, Y, Theta, sigma2, jitter=0.0, mode=0, threads=0, x_batched=False):
    """Theta: (B, p) row per item. Y: (n,) shared or (B, n). X: (n, d) shared or (B, n, d). sigma2 scalar or (B,)."""
    Theta = np.ascontiguousarray(np.atleast_2d(np.asarray(Theta, dtype=np.float64)))
    B, p = Theta.shape
    if x_batched:
        Xb = np.asarray(X, dtype=np.float64)
        n, d = Xb.shape[1], Xb.shape[2]
        Xf = np.ascontiguousarray(np.transpose(Xb, (0, 2, 1)))  # item-major, each item col-major n x d
        xs = n * d
    else:
        Xf = _f(np.asarray(X, dtype=np.float64).reshape(len(X), -1))
        n, d = Xf.shape
        xs = 0
    Y = np.asarray(Y, dtype=np.float64)
    ys = 0 if Y.ndim == 1 else n
    Yc = np.ascontiguousarray(Y)
    s2 = np.ascontiguousarray(np.atleast_1d(np.asarray(sigma2, dtype=np.float64)))
    ss = 0 if s2.size == 1 else 1
    out = np.empty(B)
    info = np.zeros(B, dtype=np.int32)
    lib().gpo_lml_batched(_ops(ops), len(ops), n, d, _p(Xf), C.c_long(xs), _p(Yc), C.c_long(ys), _p(Theta), p,
                          _p(s2), C.c_long(ss), C.c_double(jitter), mode, B, threads, _p(out),
                          info.ctypes.data_as(C.POINTER(C.c_int)))
    return out, info


def posterior(ops, X, y, theta, sigma2, jitter=0.0, mode=0):
    X = _f(np.asarray(X, dtype=np.float64).reshape(len(X), -1))
    n, d = X.shape
    y, th = _f(y), _f(np.atleast_1d(theta))
    U = np.empty((n, n), order="F")
    alpha = np.empty(n)
    rc = lib().gpo_posterior(_ops(ops), len(ops), n, d, _p(X), _p(y), _p(th), C.c_double(sigma2), C.c_double(jitter),
                             mode, _p(U), _p(alpha))
    if rc:
        raise np.linalg.LinAlgError(f"not positive definite at pivot {rc}")
    return np.triu(U), alpha


def mean_and_var(ops, X, U, alpha, Xs, theta, mode=0):
    X = _f(np.asarray(X, dtype=np.float64).reshape(len(X), -1))
    Xs = _f(np.asarray(Xs, dtype=np.float64).reshape(len(Xs), -1))
    n, d = X.shape
    m = Xs.shape[0]
    U, alpha, th = _f(U), _f(alpha), _f(np.atleast_1d(theta))
    mean, var = np.empty(m), np.empty(m)
    lib().gpo_mean_var(_ops(ops), len(ops), n, d, _p(X), _p(U), _p(alpha), _p(th), m, _p(Xs), mode, _p(mean), _p(var))
    return mean, var


def sample(ops, X, theta, sigma2, Z, jitter=0.0, mode=0):
    X = _f(np.asarray(X, dtype=np.float64).reshape(len(X), -1))
    n, d = X.shape
    Z = _f(np.asarray(Z, dtype=np.float64).reshape(n, -1))
    S = Z.shape[1]
    th = _f(np.atleast_1d(theta))
    out = np.empty((n, S), order="F")
    rc = lib().gpo_sample(_ops(ops), len(ops), n, d, _p(X), _p(th), C.c_double(sigma2), C.c_double(jitter), mode,
                          _p(Z), S, _p(out))
    if rc:
        raise np.linalg.LinAlgError(f"not positive definite at pivot {rc}")
    return out


def chol_logdet(A):
    A = np.array(A, dtype=np.float64, order="F")
    ld = C.c_double(0.0)
    rc = lib().gpo_chol_logdet(A.shape[0], _p(A), C.byref(ld))
    return np.triu(A), ld.value, rc


__all__ = ["Op", "build", "lib", "use_openblas", "use_plain_c", "cov", "lml", "lml_batched", "posterior",
           "mean_and_var", "sample", "chol_logdet"]
