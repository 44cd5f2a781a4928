"""ctypes binding of libgaplac_b200.so — the stand-in for the Julia `ccall` layer (julia/GaPLACB200.jl).

Layout rules are the ones `ccall` imposes: column-major Float64 arrays, Cint sizes, caller-owned buffers,
no ownership transfer.  There is no CPU fallback: if the shared library is missing, or no CUDA device is
present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GAPLAC_B200_LIB", os.path.join(_HERE, "libgaplac_b200.so"))  # env: experiment builds

GPL_OK, GPL_ERR_ARG, GPL_ERR_CUDA, GPL_ERR_LIMIT, GPL_ERR_NOTPD = 0, -1, -2, -3, -4
SQEXP, OU, LINEAR, CAT, CONSTANT, NOISE, ADD, MUL = range(8)

# every symbol include/gaplac_b200.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = [
    "gpl_init", "gpl_destroy", "gpl_last_error", "gpl_abi_version", "gpl_launch_count", "gpl_set_option",
    "gpl_device_info", "gpl_program_create", "gpl_program_destroy", "gpl_program_n_theta", "gpl_program_n_cols",
    "gpl_cov", "gpl_cov_dev", "gpl_cross_cov", "gpl_lml_batched", "gpl_lml_batched_dev", "gpl_posterior_fit",
    "gpl_posterior_free", "gpl_posterior_logpdf", "gpl_posterior_alpha", "gpl_posterior_factor",
    "gpl_posterior_mean_var", "gpl_sample", "gpl_chol_logdet", "gpl_chol_logdet_dev", "gpl_lml_large",
    "gpl_predict_batched", "gpl_last_timing", "gpl_set_stream", "gpl_release_workspace", "gpl_mcmc_nuts", "gpl_multi_init", "gpl_multi_destroy",
    "gpl_multi_device_count", "gpl_multi_context", "gpl_multi_last_error", "gpl_multi_lml_batched", "gpl_multi_mcmc_nuts",
]


class GplOp(C.Structure):
    """struct gpl_op (32 bytes)."""
    _fields_ = [("kind", C.c_int32), ("col", C.c_int32), ("theta_slot", C.c_int32), ("var_slot", C.c_int32),
                ("value", C.c_double), ("var", C.c_double)]


class GplTiming(C.Structure):
    """struct gpl_timing."""
    _fields_ = [("n_phases", C.c_int32), ("launches", C.c_int32 * 8), ("ms", C.c_double * 8)]


class GaplacError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"gaplac_b200 status {status}: {message}")
        self.status = status
        self.message = message


class PosDefException(GaplacError):
    """Mirror of Julia's LinearAlgebra.PosDefException(info) raised by `cholesky` [upstream]."""

    def __init__(self, status: int, message: str):
        super().__init__(status, message)
        import re
        m = re.search(r"pivot (\d+)", message)
        self.info = int(m.group(1)) if m else -1


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p


def load() -> C.CDLL:
    """Load the shared library (building is `python -m gaplac_b200.build`; never done implicitly here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(f"{LIB_PATH} not found: build it with `python -m gaplac_b200.build` "
                      "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.gpl_last_error.restype = C.c_char_p
    lib.gpl_last_error.argtypes = [_vp]
    lib.gpl_launch_count.restype = C.c_uint64
    lib.gpl_launch_count.argtypes = [_vp]
    lib.gpl_init.argtypes = [C.c_int, C.POINTER(_vp)]
    lib.gpl_destroy.argtypes = [_vp]
    lib.gpl_set_option.argtypes = [_vp, C.c_char_p, C.c_int]
    lib.gpl_device_info.argtypes = [_vp, C.c_char_p, C.c_int, _ip, _ip]
    lib.gpl_last_timing.argtypes = [_vp, C.POINTER(GplTiming)]
    lib.gpl_set_stream.argtypes = [_vp, _vp]
    lib.gpl_release_workspace.argtypes = [_vp]
    lib.gpl_multi_init.argtypes = [_ip, C.c_int, C.POINTER(_vp)]
    lib.gpl_multi_destroy.argtypes = [_vp]
    lib.gpl_multi_device_count.argtypes = [_vp]
    lib.gpl_multi_context.argtypes = [_vp, C.c_int]
    lib.gpl_multi_context.restype = _vp
    lib.gpl_multi_last_error.argtypes = [_vp]
    lib.gpl_multi_last_error.restype = C.c_char_p
    lib.gpl_multi_mcmc_nuts.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int,
                                        C.c_double, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                        C.POINTER(C.c_longlong)]
    lib.gpl_mcmc_nuts.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int,
                                  C.c_double, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  C.POINTER(C.c_longlong)]
    lib.gpl_program_create.argtypes = [_vp, C.POINTER(GplOp), C.c_int, C.POINTER(_vp)]
    lib.gpl_program_destroy.argtypes = [_vp]
    lib.gpl_program_n_theta.argtypes = [_vp]
    lib.gpl_program_n_cols.argtypes = [_vp]
    lib.gpl_cov.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_double, C.c_double, _vp]
    lib.gpl_cov_dev.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_double, C.c_double, _vp, _vp]
    lib.gpl_cross_cov.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp]
    lml_args = [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_double,
                C.c_int, _vp, _vp, _vp, _vp]
    lib.gpl_lml_batched.argtypes = lml_args
    lib.gpl_multi_lml_batched.argtypes = lml_args
    lib.gpl_lml_batched_dev.argtypes = lml_args + [_vp]
    lib.gpl_posterior_fit.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_double, C.c_double,
                                      C.POINTER(_vp)]
    lib.gpl_posterior_free.argtypes = [_vp]
    lib.gpl_posterior_logpdf.argtypes = [_vp, _dp]
    lib.gpl_posterior_alpha.argtypes = [_vp, _vp]
    lib.gpl_posterior_factor.argtypes = [_vp, _vp]
    lib.gpl_posterior_mean_var.argtypes = [_vp, C.c_int, _vp, _vp, _vp]
    lib.gpl_predict_batched.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_double,
                                        C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]
    lib.gpl_sample.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_double, C.c_double, _vp, C.c_int, _vp]
    lib.gpl_chol_logdet.argtypes = [_vp, C.c_int, _vp, C.c_int, _dp, _ip]
    lib.gpl_chol_logdet_dev.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp]
    lib.gpl_lml_large.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_double, C.c_double, _dp, _dp, _ip]
    _lib = lib
    return lib


def _check(ctx, rc: int) -> None:
    if rc == GPL_OK:
        return
    msg = load().gpl_last_error(ctx)
    msg = msg.decode() if msg else ""
    if rc == GPL_ERR_NOTPD:
        raise PosDefException(rc, msg)
    raise GaplacError(rc, msg)


def _fa(a, ndim=None) -> np.ndarray:
    """Column-major float64 view/copy (what a Julia Array{Float64} is)."""
    a = np.asarray(a, dtype=np.float64)
    if ndim == 2 and a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


def ops_array(ops):
    """ops: iterable of objects/tuples with (kind, col, theta_slot, var_slot, value, var)."""
    arr = (GplOp * len(ops))()
    for i, o in enumerate(ops):
        if isinstance(o, (tuple, list)):
            arr[i] = GplOp(*o)
        else:
            arr[i] = GplOp(o.kind, o.col, o.theta_slot, o.var_slot, o.value, o.var)
    return arr


class Program:
    """Compiled kernel-program handle (gpl_prog)."""

    def __init__(self, ctx: "Context", ops):
        self.ctx = ctx
        self.ops = list(ops)
        h = _vp()
        arr = ops_array(self.ops)
        _check(ctx.h, load().gpl_program_create(ctx.h, arr, len(self.ops), C.byref(h)))
        self.h = h
        self.n_theta = load().gpl_program_n_theta(h)
        self.n_cols = load().gpl_program_n_cols(h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                load().gpl_program_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Posterior:
    """gpl_post handle: Cholesky factor and alpha resident in HBM."""

    def __init__(self, ctx: "Context", h, n: int, d: int):
        self.ctx, self.h, self.n, self.d = ctx, h, n, d

    def logpdf(self) -> float:
        v = C.c_double()
        _check(self.ctx.h, load().gpl_posterior_logpdf(self.h, C.byref(v)))
        return v.value

    def alpha(self) -> np.ndarray:
        out = np.empty(self.n)
        _check(self.ctx.h, load().gpl_posterior_alpha(self.h, _ptr(out)))
        return out

    def factor(self) -> np.ndarray:
        out = np.empty((self.n, self.n), order="F")
        _check(self.ctx.h, load().gpl_posterior_factor(self.h, _ptr(out)))
        return out

    def mean_and_var(self, Xs, want_var: bool = True):
        Xs = _fa(Xs, 2)
        if Xs.shape[1] != self.d:
            raise GaplacError(GPL_ERR_ARG, f"test points have {Xs.shape[1]} columns, model has {self.d}")
        m = Xs.shape[0]
        mean = np.empty(m)
        var = np.empty(m) if want_var else None
        _check(self.ctx.h, load().gpl_posterior_mean_var(self.h, m, _ptr(Xs), _ptr(mean),
                                                         _ptr(var) if want_var else None))
        return (mean, var) if want_var else mean

    def free(self):
        if self.h:
            load().gpl_posterior_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _lml_batched_call(fn, owner, prog, X, Y, Theta, sigma2, jitter, grad):
    """Argument marshalling shared by Context.lml_batched and MultiContext.lml_batched (same C signature)."""
    Theta = np.ascontiguousarray(np.atleast_2d(np.asarray(Theta, dtype=np.float64)))  # (B,p) C-order == p x B col-major
    B, p = Theta.shape
    Xa = np.asarray(X, dtype=np.float64)
    x_batched = Xa.ndim == 3
    if x_batched:
        n, d = Xa.shape[1], Xa.shape[2]
        Xf = np.ascontiguousarray(np.transpose(Xa, (0, 2, 1)))  # per item: column-major n x d
    else:
        Xf = _fa(Xa, 2)
        n, d = Xf.shape
    Ya = np.asarray(Y, dtype=np.float64)
    y_batched = Ya.ndim == 2
    Yc = np.ascontiguousarray(Ya)  # (B, n) C-order == n x B col-major
    s2 = np.ascontiguousarray(np.atleast_1d(np.asarray(sigma2, dtype=np.float64)))
    s_batched = s2.size > 1
    lml = np.empty(B)
    info = np.zeros(B, dtype=np.int32)
    dth = np.empty((B, p)) if grad else None
    dy = np.empty((B, n)) if grad else None
    owner._check(fn(owner.h, prog.h, n, d, _ptr(Xf), int(x_batched), _ptr(Yc), int(y_batched), _ptr(Theta), p, _ptr(s2),
                    int(s_batched), jitter, B, _ptr(lml), _ptr(dth) if grad else None, _ptr(dy) if grad else None,
                    info.ctypes.data_as(_vp)))
    return (lml, info, dth, dy) if grad else (lml, info)


class _NoCtx:
    h = None


class MultiContext:
    """gpl_multi: one context per device behind one handle; batched calls shard their independent items over the devices
    and gather into the caller's buffers (one call, as a Julia host would make one `ccall`)."""

    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*[int(x) for x in devices])
        h = _vp()
        rc = load().gpl_multi_init(devs, len(devices), C.byref(h))
        if rc != GPL_OK:
            msg = load().gpl_multi_last_error(None)
            raise GaplacError(rc, msg.decode() if msg else "")
        self.h = h
        self.devices = list(devices)

    def _check(self, rc: int) -> None:
        if rc != GPL_OK:
            msg = load().gpl_multi_last_error(self.h)
            raise GaplacError(rc, msg.decode() if msg else "")

    def program(self, ops) -> Program:
        return Program(_NoCtx, ops)      # programs are device-independent

    def set_option(self, key: str, value: int) -> None:
        for r in range(len(self.devices)):
            _check(None, load().gpl_set_option(load().gpl_multi_context(self.h, r), key.encode(), value))

    def launch_count(self) -> int:
        return sum(int(load().gpl_launch_count(load().gpl_multi_context(self.h, r))) for r in range(len(self.devices)))

    def lml_batched(self, prog: Program, X, Y, Theta, sigma2, jitter: float = 0.0, grad: bool = False):
        return _lml_batched_call(load().gpl_multi_lml_batched, self, prog, X, Y, Theta, sigma2, jitter, grad)

    def close(self):
        if self.h:
            load().gpl_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """gpl_ctx: one CUDA device, one stream, a grow-only workspace pool."""

    def _check(self, rc: int) -> None:
        _check(self.h, rc)

    def __init__(self, device: int = -1):
        h = _vp()
        rc = load().gpl_init(device, C.byref(h))
        if rc != GPL_OK:
            msg = load().gpl_last_error(None)
            raise GaplacError(rc, msg.decode() if msg else "")
        self.h = h

    def close(self):
        if self.h:
            load().gpl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- facts ------------------------------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(load().gpl_launch_count(self.h))

    def set_option(self, key: str, value: int) -> None:
        _check(self.h, load().gpl_set_option(self.h, key.encode(), value))

    def device_info(self):
        name = C.create_string_buffer(128)
        sm, clk = C.c_int(), C.c_int()
        _check(self.h, load().gpl_device_info(self.h, name, 128, C.byref(sm), C.byref(clk)))
        return name.value.decode(), sm.value, clk.value

    def set_stream(self, stream: int = 0) -> None:
        """Host entry points run on this CUDA stream (raw cudaStream_t, e.g. torch.cuda.Stream.cuda_stream); 0 restores."""
        _check(self.h, load().gpl_set_stream(self.h, stream or None))

    def release_workspace(self) -> None:
        """Free the grow-only workspace (re-grown on demand)."""
        _check(self.h, load().gpl_release_workspace(self.h))

    def last_timing(self):
        """(ms[7], launches[7]) per phase of the last call made with option profile_events = 1 (gpl_last_timing)."""
        t = GplTiming()
        _check(self.h, load().gpl_last_timing(self.h, C.byref(t)))
        return np.array(t.ms[: t.n_phases]), np.array(t.launches[: t.n_phases])

    def program(self, ops) -> Program:
        return Program(self, ops)

    # -- covariance ---------------------------------------------------------------------------------------------
    def cov(self, prog: Program, X, theta, sigma2: float, jitter: float = 0.0) -> np.ndarray:
        X = _fa(X, 2)
        n, d = X.shape
        th = _fa(np.atleast_1d(theta))
        K = np.empty((n, n), order="F")
        _check(self.h, load().gpl_cov(self.h, prog.h, n, d, _ptr(X), _ptr(th), th.size, sigma2, jitter, _ptr(K)))
        return K

    def cov_dev(self, prog: Program, n: int, d: int, dX: int, dtheta: int, p: int, sigma2: float, jitter: float, dK: int,
                stream: int = 0) -> None:
        """Device-pointer entry: K (n x n, column-major) into dK; asynchronous on `stream`."""
        _check(self.h, load().gpl_cov_dev(self.h, prog.h, n, d, dX, dtheta or None, p, sigma2, jitter, dK, stream or None))

    def cross_cov(self, prog: Program, X, Xs, theta) -> np.ndarray:
        X, Xs = _fa(X, 2), _fa(Xs, 2)
        n, d = X.shape
        m = Xs.shape[0]
        th = _fa(np.atleast_1d(theta))
        K = np.empty((n, m), order="F")
        _check(self.h, load().gpl_cross_cov(self.h, prog.h, n, m, d, _ptr(X), _ptr(Xs), _ptr(th), th.size, _ptr(K)))
        return K

    # -- batched lml ----------------------------------------------------------------------------------------------
    def lml_batched(self, prog: Program, X, Y, Theta, sigma2, jitter: float = 0.0, grad: bool = False):
        """X: (n, d) shared or (B, n, d); Y: (n,) shared or (B, n); Theta: (B, p); sigma2: scalar or (B,).
        Returns (lml[B], info[B]) or, with grad, (lml, info, dtheta[B, p], dy[B, n])."""
        return _lml_batched_call(load().gpl_lml_batched, self, prog, X, Y, Theta, sigma2, jitter, grad)

    def lml_batched_dev(self, prog: Program, n: int, d: int, dX: int, x_batched: bool, dY: int, y_batched: bool,
                        dTheta: int, p: int, dsigma2: int, sigma2_batched: bool, jitter: float, B: int, dlml: int,
                        ddtheta: int = 0, ddy: int = 0, dinfo: int = 0, stream: int = 0) -> None:
        """Device-pointer entry (integers are raw device addresses, e.g. torch.Tensor.data_ptr()); asynchronous."""
        _check(self.h, load().gpl_lml_batched_dev(self.h, prog.h, n, d, dX, int(x_batched), dY, int(y_batched), dTheta,
                                                  p, dsigma2, int(sigma2_batched), jitter, B, dlml, ddtheta or None,
                                                  ddy or None, dinfo or None, stream or None))

    # -- posterior / sampling -----------------------------------------------------------------------------------------
    def posterior_fit(self, prog: Program, X, y, theta, sigma2: float, jitter: float = 0.0) -> Posterior:
        X = _fa(X, 2)
        n, d = X.shape
        y, th = _fa(y), _fa(np.atleast_1d(theta))
        h = _vp()
        _check(self.h, load().gpl_posterior_fit(self.h, prog.h, n, d, _ptr(X), _ptr(y), _ptr(th), th.size, sigma2,
                                                jitter, C.byref(h)))
        return Posterior(self, h, n, d)

    def predict_batched(self, prog: Program, X, y, Theta, sigma2, Xs, jitter: float = 0.0, want_var: bool = True):
        """One posterior per row of Theta (B, p) over shared (X, y); predictions at Xs (m, d).
        Returns (mean[B, m], var[B, m] or None, lml[B], info[B])."""
        X = _fa(X, 2)
        n, d = X.shape
        y = _fa(y)
        Theta = np.ascontiguousarray(np.atleast_2d(np.asarray(Theta, dtype=np.float64)))  # (B, p) C-order == p x B col-major
        B, p = Theta.shape
        s2 = np.ascontiguousarray(np.atleast_1d(np.asarray(sigma2, dtype=np.float64)))
        if s2.size not in (1, B):
            raise ValueError("sigma2 must be a scalar or have one value per row of Theta")
        Xs = _fa(Xs, 2)
        m = Xs.shape[0]
        if Xs.shape[1] != d:
            raise ValueError("Xs must have the same number of columns as X")
        mean = np.empty((B, m))
        var = np.empty((B, m)) if want_var else None
        lml = np.empty(B)
        info = np.zeros(B, dtype=np.int32)
        _check(self.h, load().gpl_predict_batched(self.h, prog.h, n, d, _ptr(X), _ptr(y), _ptr(Theta), p, _ptr(s2),
                                                  int(s2.size == B and B > 1), jitter, B, m, _ptr(Xs), _ptr(mean),
                                                  _ptr(var) if want_var else None, _ptr(lml), _ptr(info)))
        return mean, var, lml, info

    def sample(self, prog: Program, X, theta, sigma2: float, Z, jitter: float = 0.0) -> np.ndarray:
        X = _fa(X, 2)
        n, d = X.shape
        Z = _fa(np.asarray(Z, dtype=np.float64).reshape(n, -1))
        S = Z.shape[1]
        th = _fa(np.atleast_1d(theta))
        out = np.empty((n, S), order="F")
        _check(self.h, load().gpl_sample(self.h, prog.h, n, d, _ptr(X), _ptr(th), th.size, sigma2, jitter, _ptr(Z), S,
                                         _ptr(out)))
        return out

    # -- large n ----------------------------------------------------------------------------------------------------------
    def chol_logdet(self, A, want_factor: bool = True):
        A = np.array(A, dtype=np.float64, order="F")
        n = A.shape[0]
        ld, info = C.c_double(), C.c_int()
        _check(self.h, load().gpl_chol_logdet(self.h, n, _ptr(A), int(want_factor), C.byref(ld), C.byref(info)))
        return (A if want_factor else None), ld.value, info.value

    def chol_logdet_dev(self, n: int, dA: int, want_factor: bool, dlogdet: int, dinfo: int, stream: int = 0) -> None:
        _check(self.h, load().gpl_chol_logdet_dev(self.h, n, dA, int(want_factor), dlogdet, dinfo, stream or None))

    def lml_large(self, prog: Program, X, y, theta, sigma2: float, jitter: float = 0.0):
        X = _fa(X, 2)
        n, d = X.shape
        y, th = _fa(y), _fa(np.atleast_1d(theta))
        lml, ld, info = C.c_double(), C.c_double(), C.c_int()
        _check(self.h, load().gpl_lml_large(self.h, prog.h, n, d, _ptr(X), _ptr(y), _ptr(th), th.size, sigma2, jitter,
                                            C.byref(lml), C.byref(ld), C.byref(info)))
        return lml.value, ld.value, info.value
