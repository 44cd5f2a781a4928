"""Host build of gaplac_b200/csrc/fastexp.h (the table exp of the covariance kernels) against mpmath.

The device code is the same source (GPL_HD functions); the GPU side is covered by the covariance parity tests
(entries <= 1e-14 relative to the oracle's libm exp)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r"""
#include "%s/gaplac_b200/csrc/fastexp.h"
static const double TAB[64] = {GPL_EXP_TABLE_VALUES};
extern "C" void fast_exp_array(const double *x, double *out, int n) {
    for (int i = 0; i < n; ++i) out[i] = gpl::fast_exp(x[i], TAB);
}
extern "C" void fast_exp_vec8(const double *x, double *out) {
    double v[8];
    for (int i = 0; i < 8; ++i) v[i] = x[i];
    gpl::fast_exp_vec<8>(v, TAB);
    for (int i = 0; i < 8; ++i) out[i] = v[i];
}
"""


@pytest.fixture(scope="module")
def lib():
    d = tempfile.mkdtemp(prefix="fastexp_")
    src = os.path.join(d, "fe.cpp")
    with open(src, "w") as f:
        f.write(SRC % ROOT)
    so = os.path.join(d, "libfe.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    lib.fast_exp_array.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.fast_exp_vec8.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    return lib


def _call(lib, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib.fast_exp_array(x.ctypes.data, out.ctypes.data, x.size)
    return out


def test_fast_exp_within_1p5_ulp_of_mpmath(lib):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.uniform(0, 700, 4000), -np.exp(rng.uniform(-40, 3, 2000)), rng.uniform(0, 700, 500),
                        [0.0, -0.0, -1e-300, -700.0, 700.0, -np.log(2) / 128, -np.log(2) / 64]])
    got = _call(lib, x)
    worst = 0.0
    for xi, gi in zip(x, got):
        ref = mp.exp(mp.mpf(float(xi)))
        ulp = float(abs(mp.mpf(float(gi)) - ref) / ref) / 2.0 ** -52
        worst = max(worst, ulp)
    assert worst <= 1.5, worst


def test_fast_exp_general_path_for_out_of_range_and_nan(lib):
    x = np.array([-800.0, -745.0, -1e4, 710.0, np.nan, -np.inf, np.inf])
    got = _call(lib, x)
    with np.errstate(over="ignore"):
        ref = np.exp(x)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.allclose(got[ok], ref[ok], rtol=1e-15, atol=0.0)


def test_fast_exp_vec_group_matches_scalar_including_a_slow_path_member(lib):
    for x in (np.linspace(-30.0, -0.5, 8), np.array([-1.0, -2.0, -900.0, -3.0, -4.0, np.nan, -5.0, -6.0])):
        out = np.empty(8)
        xc = np.ascontiguousarray(x)
        lib.fast_exp_vec8(xc.ctypes.data, out.ctypes.data)
        ref = np.exp(x)
        ok = ~np.isnan(ref)
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        assert np.allclose(out[ok], ref[ok], rtol=4e-16, atol=0.0)


def test_fast_exp_vec_deep_underflow_group_is_flushed_branch_free(lib):
    """A group with arguments below -700 but no NaN / overflow takes the clamped branch-free path: exact zeros below
    -708 (the true values are below 3.3e-308), the usual <= 1.5 ulp elsewhere - also between -700 and -708."""
    x = np.array([-0.25, -699.0, -703.5, -707.9, -708.5, -745.0, -5000.0, -np.inf])
    out = np.empty(8)
    lib.fast_exp_vec8(np.ascontiguousarray(x).ctypes.data, out.ctypes.data)
    ref = np.exp(x)
    assert np.all(out[4:] == 0.0) and np.all(ref[4:] < 3.4e-308)
    assert np.allclose(out[:4], ref[:4], rtol=4e-16, atol=0.0)
