// exp(x) for the covariance construction: 64-entry table + degree-5 polynomial, ~10 FP64-pipe operations instead of
// the ~21 of the general-purpose exp() (which, at two exp per entry, made the covariance tiles cost as much FP64
// pipe time as the factorisation's GEMMs; profiles/ncu_lml_r01_v2_*).
//
//   x = (64 m + j) ln2/64 + r,  |r| <= ln2/128   ->   exp(x) = 2^m * T[j] * (1 + p(r)),  T[j] = 2^(j/64) correctly rounded
//
// Truncation error r^6/720 <= 3.6e-17; total error <= 1 ulp on [-700, 0] (tests/test_fastexp.py checks the host
// build of this same code against mpmath).  Outside [-700, 700], and for NaN, it defers to exp().
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define GPL_HD __host__ __device__ __forceinline__
#else
#define GPL_HD static inline
#endif

namespace gpl {

#define GPL_EXP_TABLE_VALUES                                                                                          \
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0, 0x1.0b5586cf9890fp+0,      \
        0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0, 0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0,  \
        0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0, 0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0,  \
        0x1.2d285a6e4030bp+0, 0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,  \
        0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0, 0x1.4bfdad5362a27p+0,  \
        0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0, 0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0,  \
        0x1.6247eb03a5585p+0, 0x1.6623882552225p+0, 0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0,  \
        0x1.75feb564267c9p+0, 0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,  \
        0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0, 0x1.9c49182a3f090p+0,  \
        0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0, 0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0,  \
        0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0, 0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0,  \
        0x1.d072d4a07897cp+0, 0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,  \
        0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0

// branch-free core: valid for |x| <= 700 (no range check, no special values)
GPL_HD double fast_exp_core(double x, const double *__restrict__ tab) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer in the low bits
    const double t = fma(x, 0x1.71547652b82fep+6, MAGIC);
    const double tn = t - MAGIC;
    double r = fma(tn, -0x1.62e42ff000000p-7, x);
    r = fma(tn, 0x1.718432a1b0e26p-41, r);
    int64_t tb;
#ifdef __CUDA_ARCH__
    tb = __double_as_longlong(t);
#else
    memcpy(&tb, &t, 8);
#endif
    const int n = (int)(uint32_t)tb;  // low 32 bits hold the (two's complement) integer
    const int j = n & 63, m = n >> 6;
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    q = q * r;
    const double p = fma(q, r, r);
    const double T = tab[j];
    const double v = fma(T, p, T);
    int64_t vb;
#ifdef __CUDA_ARCH__
    // m << 20 added to the high word: (n & ~63) * 2^14 is the same value (two instructions: LOP3 + IMAD)
    (void)vb;
    (void)m;
    return __hiloint2double(__double2hiint(v) + (n & ~63) * 16384, __double2loint(v));
#else
    memcpy(&vb, &v, 8);
    vb += (int64_t)m << 52;
    double out;
    memcpy(&out, &vb, 8);
    return out;
#endif
}

#ifdef __CUDACC__
static __device__ __noinline__ double exp_general(double x) { return exp(x); }  // out of line: keeps the hot code small
#endif

GPL_HD double fast_exp(double x, const double *__restrict__ tab) {
#ifdef __CUDA_ARCH__
    if (!(fabs(x) <= 700.0)) return exp_general(x);  // NaN, huge, deep underflow: the general path
#else
    if (!(fabs(x) <= 700.0)) return exp(x);
#endif
    return fast_exp_core(x, tab);
}

// N independent evaluations with ONE range check for the group: the N dependency chains interleave (the scalar
// version's per-call branch kept them serial, which left the covariance tiles latency-bound: profiles/ README).
template <int N>
GPL_HD void fast_exp_vec(double (&x)[N], const double *__restrict__ tab) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) ok = ok && (fabs(x[i]) <= 700.0);  // false for NaN
    if (ok) {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = fast_exp_core(x[i], tab);
        return;
    }
    // Deep underflow somewhere in the group (far-apart points under a short length scale: most entries of a wide
    // SqExp covariance are exp(-thousands)): still branch-free - clamp the argument to -708 and flush to zero below it
    // (results under 3.3e-308 read as 0: absolute error below the smallest normal number).  Measured on the n = 8192
    // covariance build (x ~ U(-50, 50), l = 1): the general exp() below took 55 % of that kernel's instructions.
    bool okneg = true;
#pragma unroll
    for (int i = 0; i < N; ++i) okneg = okneg && (x[i] <= 700.0);  // false for NaN and overflow
    if (okneg) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const bool tiny = x[i] < -708.0;
            const double v = fast_exp_core(tiny ? -708.0 : x[i], tab);
            x[i] = tiny ? 0.0 : v;
        }
    } else {
        // rare (deep underflow / NaN somewhere in the group): the general exp for every entry.  Unrolled with static
        // indices so that x[] stays in registers (a rolled loop here put the whole array in local memory).
#pragma unroll
        for (int i = 0; i < N; ++i) {
#ifdef __CUDA_ARCH__
            x[i] = exp_general(x[i]);
#else
            x[i] = exp(x[i]);
#endif
        }
    }
}

}  // namespace gpl
