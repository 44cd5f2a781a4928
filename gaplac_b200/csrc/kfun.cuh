// Device-side evaluation of the compiled kernel-program (program.h) on register blocks of matrix entries.
//
// Replaces kernelmatrix(k, RowVecs(X)) [upstream KernelFunctions 0.10.38] reached from the FiniteGP call
// sites (CLI/src/mcmc.jl:35, CLI/src/select.jl:43,47, CLI/src/sample.jl:25, src/plotting.jl:6): leaves per
// src/abstractgp_translations.jl:8-15 and src/gp_parts.jl:11-13.  Nothing is materialised per node: each
// thread evaluates sum_t coef_t prod_f leaf_f for its own R x C block of entries, in registers.
#pragma once
#include "fastexp.h"
#include "program.h"

namespace gpl {

// per-item scalars derived from theta once per CTA (shared memory)
struct ItemScalars {
    double a[GPL_MAX_FACTORS];   // SQEXP: -1/(2 l^2); OU: -1/l; LINEAR: c; PARAM: theta
    double da[GPL_MAX_FACTORS];  // derivative scale: SQEXP 1/l^3 (dk = k d^2 / l^3); OU 1/l^2 (dk = k |d| / l^2);
                                 // PARAM: the term's coefficient without this factor
    double tc[GPL_MAX_TERMS];    // per term: coef * prod of its F_PARAM factors
    double etab[64];             // 2^(j/64) for fast_exp (fastexp.h), copied from constant memory once per CTA
};

// Separable form of 1-D OU leaves for a 64 x 64 block whose rows all lie at or above its columns along the leaf's input
// column (inputs sorted by that column, block below the diagonal):  exp(-|x_i - x_j| / l) = exp(-(x_i - c) / l) *
// exp(-(c - x_j) / l)  for any c between the two groups - 128 exponentials per block instead of 4096, one multiply per
// entry.  With c = the block's smallest row coordinate both exponents are <= 0 (no overflow).  u, v are filled by the
// kernel that owns the block (lk_below_kernel); leaf[] lists the factor indices they stand for.
struct SepCtx {
    int n_sep;
    int leaf[2];
    double u[2][64], v[2][64];
};

static __constant__ double c_exptab[64] = {GPL_EXP_TABLE_VALUES};

__device__ __forceinline__ void prepare_item_scalars(const DevProgram &P, const double *__restrict__ theta,
                                                     ItemScalars *S, int tid) {
    if (tid < P.n_factors) {
        const DevFactor f = P.f[tid];
        double h = f.slot >= 0 ? theta[f.slot] : f.value;
        double a = h, da = 1.0;
        if (f.kind == F_SQEXP) {
            a = -0.5 / (h * h);
            da = 1.0 / (h * h * h);
        } else if (f.kind == F_OU) {
            a = -1.0 / h;
            da = 1.0 / (h * h);
        }
        // a length scale must be > 0 (ScaleTransform throws otherwise [upstream]): poison the item instead of letting a
        // negative or zero proposal produce a plausible-looking kernel -> NaN covariance -> info = 1, lml = -Inf
        if ((f.kind == F_SQEXP || f.kind == F_OU) && !(h > 0.0)) a = da = __longlong_as_double(0x7ff8000000000000LL);
        if (f.kind == F_PARAM) {  // d term / d theta_f = (coef * the term's other per-item scalars) * product of its leaves
            int t = 0;
            while (tid >= P.term_begin[t + 1]) ++t;
            da = P.coef[t];
            for (int g = P.term_begin[t]; g < P.leaf_begin[t]; ++g)
                if (g != tid) da *= P.f[g].slot >= 0 ? theta[P.f[g].slot] : P.f[g].value;
        }
        S->a[tid] = a;
        S->da[tid] = da;
    }
    if (tid >= 64 && tid < 128) S->etab[tid - 64] = c_exptab[tid - 64];
    if (tid >= 32 && tid < 32 + P.n_terms) {
        const int t = tid - 32;
        double tc = P.coef[t];
        for (int f = P.term_begin[t]; f < P.leaf_begin[t]; ++f) tc *= P.f[f].slot >= 0 ? theta[P.f[f].slot] : P.f[f].value;
        S->tc[t] = tc;
    }
}

// True (block-uniform) iff the 64 x 64 tile (rows ti*64.., columns tj*64.. of the same observation set X) is zero for every
// hyperparameter value: every term but the noise has a Cat factor none of whose row categories occurs among the columns
// (rows / columns of different subjects under Cat(:subject) * k).  K and dK/dtheta vanish on such a tile, so its
// covariance code can be skipped outright.  Guarded so that skipping cannot hide a non-finite value (0 * NaN): all item
// scalars and the tile's inputs must be finite.  scratch: 64 doubles of shared memory; contains block barriers.
__device__ __forceinline__ bool tile_cat_dead(const DevProgram &P, const ItemScalars &S, const double *__restrict__ X, int n, int d,
                                              int ti, int tj, double *scratch, int tid) {
    const int loc = tid & 63, idx = (tid < 64 ? ti : tj) * 64 + loc;
    int bad = 0;
    if (tid < 128 && idx < n)
        for (int c = 0; c < d; ++c) bad |= !isfinite(X[(size_t)c * n + idx]);
    if (tid < P.n_factors) bad |= !isfinite(S.a[tid]) | !isfinite(S.da[tid]);
    if (tid < P.n_terms) bad |= !isfinite(S.tc[tid]);
    if (__syncthreads_or(bad)) return false;
    for (int t = 0; t < P.n_terms; ++t) {
        if (P.has_noise >> t & 1) continue;
        bool term_dead = false;
        for (int f = P.leaf_begin[t]; f < P.term_begin[t + 1] && !term_dead; ++f) {
            if (P.f[f].kind != F_CAT) continue;
            const double *xc = X + (size_t)P.f[f].col * n;
            if (tid < 64) scratch[tid] = (tj * 64 + tid < n) ? xc[tj * 64 + tid] : NAN;  // padding columns match nothing
            __syncthreads();
            int match = 0;
            if (tid < 64 && ti * 64 + tid < n) {
                const double xi = xc[ti * 64 + tid];
#pragma unroll 8
                for (int c = 0; c < 64; ++c) match |= (xi == scratch[c]);
            }
            term_dead = !__syncthreads_or(match);
        }
        if (!term_dead) return false;
    }
    return P.n_terms > 0;
}

// Evaluate the program on the block rows gi[0..R) x cols gj[0..C).
//   Xa: column-major, leading dimension lda, na valid rows (row indices); Xb likewise for column indices.
//   SAME: Xa and Xb are the same observation set (K(X,X)): Noise = [gi == gj], diag_add goes on gi == gj,
//         and indices >= na are the identity padding of the tiled factorisation (1 on the diagonal, 0 off it).
//   !SAME: cross-covariance K(X, X*): Noise = 0, out-of-range entries = 0.
// One leaf factor on an R x C block: k[r * C + c] = leaf(x_i[col], x_j[col]).
template <int R, int C, bool SAME>
__device__ __forceinline__ void leaf_block(const DevProgram &P, const ItemScalars &S, int f,
                                           const double *__restrict__ Xa, int lda, const int (&ci)[R],
                                           const double *__restrict__ Xb, int ldb, const int (&cj)[C],
                                           const int (&gi)[R], const int (&gj)[C], double (&k)[R * C],
                                           const SepCtx *sep = nullptr) {
    const int kind = P.f[f].kind;
    if (kind == F_NOISE) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) k[r * C + c] = (SAME && gi[r] == gj[c]) ? 1.0 : 0.0;
        return;
    }
    const double a = S.a[f];
    const int col = P.f[f].col;
    double xi[R], xj[C];
#pragma unroll
    for (int r = 0; r < R; ++r) xi[r] = Xa[(size_t)col * lda + ci[r]];
#pragma unroll
    for (int c = 0; c < C; ++c) xj[c] = Xb[(size_t)col * ldb + cj[c]];
    if (kind == F_SQEXP) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double d = xi[r] - xj[c];
                k[r * C + c] = a * (d * d);
            }
        fast_exp_vec<R * C>(k, S.etab);
    } else if (kind == F_OU) {
        if (!SAME && sep) {  // block-uniform: the separable form, when this leaf has one (SepCtx above)
            for (int q = 0; q < sep->n_sep; ++q)
                if (sep->leaf[q] == f) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int c = 0; c < C; ++c) k[r * C + c] = sep->u[q][gi[r] & 63] * sep->v[q][gj[c] & 63];
                    return;
                }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) k[r * C + c] = a * fabs(xi[r] - xj[c]);
        fast_exp_vec<R * C>(k, S.etab);
    } else if (kind == F_LINEAR) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) k[r * C + c] = fma(xi[r], xj[c], a);
    } else {  // F_CAT
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) k[r * C + c] = (xi[r] == xj[c]) ? 1.0 : 0.0;
    }
}

// out = sum_t tc_t prod_{leaves of t}.  The first leaf of a term is evaluated straight into the product and the
// coefficient is applied by the closing FMA, so a single-leaf term (the common case) costs its leaf plus one FMA per
// entry; keeping one accumulating array live across a factor loop instead cost 16-30 register moves per factor.
template <int R, int C, bool SAME>
__device__ __forceinline__ void eval_block(const DevProgram &P, const ItemScalars &S, const double *__restrict__ Xa,
                                           int lda, int na, const int (&gi)[R], const double *__restrict__ Xb, int ldb,
                                           int nb, const int (&gj)[C], double diag_add, double (&out)[R][C],
                                           const SepCtx *sep = nullptr) {
    int ci[R], cj[C];
#pragma unroll
    for (int r = 0; r < R; ++r) ci[r] = gi[r] < na ? gi[r] : na - 1;
#pragma unroll
    for (int c = 0; c < C; ++c) cj[c] = gj[c] < nb ? gj[c] : nb - 1;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) out[r][c] = 0.0;

    for (int t = 0; t < P.n_terms; ++t) {
        if (!SAME && ((P.has_noise >> t) & 1)) continue;  // Noise terms vanish on cross-covariances
        const double tc = S.tc[t];
        const int f0 = P.leaf_begin[t], f1 = P.term_begin[t + 1];
        if (f0 == f1) {  // a term without leaves: a per-item constant
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int c = 0; c < C; ++c) out[r][c] += tc;
            continue;
        }
        double k[R * C];
        leaf_block<R, C, SAME>(P, S, f0, Xa, lda, ci, Xb, ldb, cj, gi, gj, k, sep);
        for (int f = f0 + 1; f < f1; ++f) {
            double e[R * C];
            leaf_block<R, C, SAME>(P, S, f, Xa, lda, ci, Xb, ldb, cj, gi, gj, e, sep);
#pragma unroll
            for (int q = 0; q < R * C; ++q) k[q] *= e[q];
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) out[r][c] = fma(tc, k[r * C + c], out[r][c]);
    }
    if (SAME) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const bool inr = gi[r] < na && gj[c] < nb;
                if (gi[r] == gj[c]) out[r][c] = inr ? out[r][c] + diag_add : 1.0;
                else if (!inr) out[r][c] = 0.0;
            }
    } else if (gi[R - 1] >= na || gj[C - 1] >= nb) {  // indices ascend within a block: only edge blocks need masking
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (gi[r] >= na || gj[c] >= nb) out[r][c] = 0.0;
    }
}

// The 2 x 16 accumulator block of one thread (tile.cuh: rows gi[0..1], columns cbase + 8 (cc/2) + 2 t + cc%2) as four
// 2 x 4 quarters.  ROLLED = true keeps one copy of the evaluation code (small kernels, instruction-cache bound fused
// kernel) at the price of select-copies into the dynamically indexed quarter; ROLLED = false unrolls the quarters
// (static register indices, no copies) for the throughput kernels of the lockstep schedule.
// hmax: quarters h >= hmax (16 columns each) are not evaluated and read as 0 (the part of a diagonal tile that lies
// above the diagonal for this warp's rows).
template <bool SAME, bool ROLLED = true>
__device__ __forceinline__ void eval_block_acc(const DevProgram &P, const ItemScalars &S, const double *__restrict__ Xa,
                                               int lda, int na, const int (&gi)[2], const double *__restrict__ Xb,
                                               int ldb, int nb, int cbase, int t, double diag_add,
                                               double (&out)[2][16], int hmax = 4) {
    if (ROLLED) {
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
            int gjh[4];
            double o[2][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) gjh[c] = cbase + 16 * h + 8 * (c >> 1) + 2 * t + (c & 1);
            if (h < hmax) {
                eval_block<2, 4, SAME>(P, S, Xa, lda, na, gi, Xb, ldb, nb, gjh, diag_add, o);
            } else {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[r][c] = 0.0;
            }
#pragma unroll
            for (int hh = 0; hh < 4; ++hh)
                if (hh == h) {
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) out[r][4 * hh + c] = o[r][c];
                }
        }
    } else {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            int gjh[4];
            double o[2][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) gjh[c] = cbase + 16 * h + 8 * (c >> 1) + 2 * t + (c & 1);
            if (h < hmax) {
                eval_block<2, 4, SAME>(P, S, Xa, lda, na, gi, Xb, ldb, nb, gjh, diag_add, o);
            } else {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[r][c] = 0.0;
            }
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) out[r][4 * h + c] = o[r][c];
        }
    }
}

// Same result as eval_block_acc, built for the throughput kernels: the quarters run through ONE rolled copy of the
// evaluation code per pass (the four unrolled copies made a 38 KB loop body that the instruction cache could not hold:
// 20 % of the j = 0 launch stalled on instruction fetch), and each quarter parks its 8 values in a thread-private
// column of a shared-memory slot (NTHREADS x 8 doubles = 8 KiB, conflict-free, no barrier needed) from where they are
// read back into statically indexed accumulator registers.  NSLOT slots -> 4 / NSLOT passes.
template <bool SAME, int NSLOT, int CW = 4>
__device__ __forceinline__ void eval_block_acc_scr(const DevProgram &P, const ItemScalars &S,
                                                   const double *__restrict__ Xa, int lda, int na, const int (&gi)[2],
                                                   const double *__restrict__ Xb, int ldb, int nb, int cbase, int t,
                                                   double diag_add, double *const (&slot)[NSLOT], int tid,
                                                   double (&out)[2][16], int hmax = 4, const SepCtx *sep = nullptr) {
    // CW = 4: the code runs per quarter (2 x 4 entries, 16 columns of the tile); CW = 8: per half (2 x 8 entries), which
    // halves the per-step interpretation / load / bookkeeping instructions and doubles the independent exp chains.
    static_assert(NSLOT == 1 || NSLOT == 2 || NSLOT == 4, "slots");
    static_assert(CW == 4 || CW == 8, "entries per row and step");
    constexpr int STEPS = 16 / CW;          // steps per tile
    constexpr int SPS = CW / 4;             // slots per step (a slot parks 8 values per thread)
    constexpr int STEPS_PER_PASS = NSLOT / SPS;
    static_assert(STEPS_PER_PASS >= 1, "a step must fit the slots");
    const int smax = (hmax + SPS - 1) / SPS;
#pragma unroll
    for (int pass = 0; pass < STEPS / STEPS_PER_PASS; ++pass) {
#pragma unroll 1
        for (int ss = 0; ss < STEPS_PER_PASS; ++ss) {
            const int st = pass * STEPS_PER_PASS + ss;
            int gjh[CW];
            double o[2][CW];
#pragma unroll
            for (int c = 0; c < CW; ++c) gjh[c] = cbase + 4 * CW * st + 8 * (c >> 1) + 2 * t + (c & 1);
            if (st < smax) {
                eval_block<2, CW, SAME>(P, S, Xa, lda, na, gi, Xb, ldb, nb, gjh, diag_add, o, sep);
            } else {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < CW; ++c) o[r][c] = 0.0;
            }
#pragma unroll
            for (int u = 0; u < SPS; ++u) {  // slot ss * SPS + u takes columns 4u..4u+3 of the step
                double *dst = slot[u];
#pragma unroll
                for (int q = 1; q < STEPS_PER_PASS; ++q) dst = (ss == q) ? slot[q * SPS + u] : dst;
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) dst[(r * 4 + c) * 128 + tid] = o[r][4 * u + c];
            }
        }
#pragma unroll
        for (int hh = 0; hh < NSLOT; ++hh)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) out[r][4 * (pass * NSLOT + hh) + c] = slot[hh][(r * 4 + c) * 128 + tid];
    }
}

// Contract a weight block with dK/dtheta_s for every slot s:  g[s] += sum_rc w[r][c] dK[r][c]/dtheta_s, for the 2 x 4
// entries rows gi x columns gj of K(X, X).  Every leaf of a term is evaluated once (fast_exp, 8 wide); a term
// T = tc * prod_leaves k contributes
//   per-item scalar factor f (variance, Constant):  da_f * sum w prod_leaves k
//   SqExp / OU length scale:                         tc * da_f * sum w (prod_leaves k) d^2   (resp. |d|)
//   Linear offset c:                                 tc * sum w prod_{leaves != f} k
// gsum: shared-memory accumulators, one row of GPL_MAX_THETA doubles PER WARP (the caller passes its warp's row and
// adds the rows up in a fixed order afterwards): lane 0 adds the warp total, so the summation order - and with it every
// bit of the gradient - is the same from run to run (shared atomics were not).
__device__ __forceinline__ void grad_reduce_add(double part, double *dst) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) *dst += part;
}
__device__ __forceinline__ void contract_grad_quarter(const DevProgram &P, const ItemScalars &S,
                                                      const double *__restrict__ X, int ldx, int n, const int (&gi)[2],
                                                      const int (&gj)[4], const double (&w)[8], double *gsum,
                                                      const SepCtx *sep = nullptr) {
    // sep != nullptr: a block below the diagonal on sorted inputs - cross form of the leaves (Noise = 0 there anyway) with
    // the separable OU factors
    int ci[2], cj[4];
#pragma unroll
    for (int r = 0; r < 2; ++r) ci[r] = gi[r] < n ? gi[r] : n - 1;
#pragma unroll
    for (int c = 0; c < 4; ++c) cj[c] = gj[c] < n ? gj[c] : n - 1;
    double wm[8];  // weights, zero outside the matrix
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) wm[r * 4 + c] = (gi[r] < n && gj[c] < n) ? w[r * 4 + c] : 0.0;
    for (int t = 0; t < P.n_terms; ++t) {
        const int p0 = P.term_begin[t], f0 = P.leaf_begin[t], f1 = P.term_begin[t + 1];
        bool any = false;
        for (int f = p0; f < f1; ++f) any = any || P.f[f].slot >= 0;
        if (!any) continue;
        double wl[8];  // w * product of the term's leaves
#pragma unroll
        for (int e = 0; e < 8; ++e) wl[e] = wm[e];
        for (int f = f0; f < f1; ++f) {
            double k[8];
            if (sep) leaf_block<2, 4, false>(P, S, f, X, ldx, ci, X, ldx, cj, gi, gj, k, sep);
            else leaf_block<2, 4, true>(P, S, f, X, ldx, ci, X, ldx, cj, gi, gj, k);
#pragma unroll
            for (int e = 0; e < 8; ++e) wl[e] *= k[e];
        }
        double sl = 0.0;
#pragma unroll
        for (int e = 0; e < 8; ++e) sl += wl[e];
        for (int f = p0; f < f0; ++f)  // per-item scalar factors
            if (P.f[f].slot >= 0) grad_reduce_add(S.da[f] * sl, &gsum[P.f[f].slot]);
        const double tc = S.tc[t];
        for (int f = f0; f < f1; ++f) {
            const int slot = P.f[f].slot, kind = P.f[f].kind;
            if (slot < 0) continue;
            const int col = P.f[f].col;
            double part = 0.0;
            if (kind == F_SQEXP || kind == F_OU) {
                double xi[2], xj[4];
#pragma unroll
                for (int r = 0; r < 2; ++r) xi[r] = X[(size_t)col * ldx + ci[r]];
#pragma unroll
                for (int c = 0; c < 4; ++c) xj[c] = X[(size_t)col * ldx + cj[c]];
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double d = xi[r] - xj[c];
                        part = fma(wl[r * 4 + c], kind == F_SQEXP ? d * d : fabs(d), part);
                    }
                part *= tc * S.da[f];
            } else if (kind == F_LINEAR) {  // dk/dc = 1: the product of the other leaves
                double wo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) wo[e] = wm[e];
                for (int g = f0; g < f1; ++g) {
                    if (g == f) continue;
                    double k[8];
                    leaf_block<2, 4, true>(P, S, g, X, ldx, ci, X, ldx, cj, gi, gj, k);
#pragma unroll
                    for (int e = 0; e < 8; ++e) wo[e] *= k[e];
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) part += wo[e];
                part *= tc;
            }
            grad_reduce_add(part, &gsum[slot]);
        }
    }
}

// The 2 x 16 block of one thread (rows gi, columns cbase + col_of): the weights are parked in thread-private columns of
// a 32 KiB shared-memory scratch so that one rolled copy of the quarter code serves all four quarters.
__device__ __forceinline__ void contract_grad_block(const DevProgram &P, const ItemScalars &S,
                                                    const double *__restrict__ X, int ldx, int n, const int (&gi)[2],
                                                    int cbase, int t, const double (&w)[2][16], double *scratch, int tid,
                                                    double *gsum, int hmax = 4,  // quarters h >= hmax carry zero weights: skipped
                                                    const SepCtx *sep = nullptr) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) scratch[(r * 16 + cc) * 128 + tid] = w[r][cc];
#pragma unroll 1
    for (int h = 0; h < hmax; ++h) {
        int gjh[4];
        double wq[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) gjh[c] = cbase + 16 * h + 8 * (c >> 1) + 2 * t + (c & 1);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) wq[r * 4 + c] = scratch[(r * 16 + 4 * h + c) * 128 + tid];
        contract_grad_quarter(P, S, X, ldx, n, gi, gjh, wq, gsum, sep);
    }
}

}  // namespace gpl
