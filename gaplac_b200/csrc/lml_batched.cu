// Batched fused log-marginal-likelihood kernel: covariance construction + Cholesky + triangular solve +
// logdet (+ analytic gradient) for many independent GPs, one CTA per GP at a time.
//
// Replaces, per batch item, the reference's logpdf(FiniteGP, y) [upstream AbstractGPs 0.5.12] =
// kernelmatrix -> + sigma2 I -> cholesky(Symmetric) (dpotrf) -> U' \ y (dtrtrs) -> logdet, called at
// CLI/src/select.jl:49-50 and once per leapfrog step at CLI/src/mcmc.jl:35; and posterior(FiniteGP, y)
// (CLI/src/select.jl:51-52, src/plotting.jl:8) when `keep` is set.
//
// Algorithm (per item): left-looking blocked Cholesky on 64 x 64 tiles.  K never exists in memory: tile (i, j)
// is generated in registers from the input columns and the hyperparameters, updated with the previously
// factored tiles streamed (double-buffered cp.async stages of 16 columns) from the CTA's private workspace, then
// factored (diagonal tile, tile_potrf_inv) or multiplied by the inverse of the diagonal tile (tiles below).
// All contractions run on the FP64 tensor-core path (DMMA, tile.cuh).  The forward solve z = L^-1 y and logdet
// ride along with the diagonal tiles.  The gradient phase forms M = L^-1 and K^-1 tile by tile and contracts
// (K^-1 - alpha alpha') with dK/dtheta generated on the fly.
#include "kernels.h"
#include "kfun.cuh"
#include "tile.cuh"

namespace gpl {

namespace {

#ifndef GPL_LML_CTAS_PER_SM
#define GPL_LML_CTAS_PER_SM 4
#endif
#ifndef GPL_LML_ZMAX
#define GPL_LML_ZMAX 1024  // rows of z kept in shared memory (8 KiB)
#endif
#ifndef GPL_LML_STREAM_TRSM
#define GPL_LML_STREAM_TRSM 1  // 1: stream L_jj through the load pipeline; 0: re-load the whole tile, then solve
#endif
// Optional phase timers (build with -DGPL_LML_PROFILE): thread 0 of every CTA accumulates clock64() deltas per phase
// into prm.dtheta (reused as a raw buffer of gridDim.x * 8 doubles).  tools/phase_profile.py reads them.
#ifdef GPL_LML_PROFILE
#define PH_DECL long long ph_t0 = clock64(), ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PH_MARK(k)                         \
    do {                                   \
        const long long ph_t1 = clock64(); \
        ph_acc[k] += ph_t1 - ph_t0;        \
        ph_t0 = ph_t1;                     \
    } while (0)
#define PH_DUMP                                                                                         \
    do {                                                                                                \
        if (tid == 0 && prm.dtheta)                                                                     \
            for (int k = 0; k < 8; ++k) prm.dtheta[(size_t)blockIdx.x * 8 + k] = (double)ph_acc[k];      \
    } while (0)
#else
#define PH_DECL
#define PH_MARK(k)
#define PH_DUMP
#endif
constexpr int LKC = 16;        // columns per pipeline stage
constexpr int LCH = LKC * TS;  // doubles per operand stage: two stages x (row + column operand) = the 32 KiB buffer S
static_assert(4 * LCH == TILE_ELEMS, "two stages of both operands fill the staging buffer exactly");

constexpr double LOG2PI = 1.8378770664093454835606594728112;

// Shared memory of one CTA (~48 KiB: four CTAs per SM).  The 32 KiB buffer S holds, in turn: the two 16-column
// pipeline stages of the row operand (S[0], S[LCH]) and of the column operand (S[2 LCH], S[3 LCH]) during the
// update loop; the factorisation scratch of the diagonal tile; the factor L_jj of the diagonal tile, re-loaded
// from the workspace (L2) for the triangular solve of every tile below it; one whole tile for the single-buffered
// alpha / gradient phases.  Only the gradient kernel carries a second whole-tile buffer (Bt).
template <bool GRAD>
struct __align__(16) LmlSmem {
    double S[TILE_ELEMS];
    // Gradient kernel: a second whole-tile buffer for the K^-1 phases.  Until those start it hosts D and zs (below), which
    // keeps the kernel at three CTAs per SM (it ran at two: 12 % of the warp slots, DMMA pipe 41 % busy).
    double Bt[GRAD ? TILE_ELEMS : 2];
    double Dbuf[GRAD ? 2 : DSIZE];         // inverses of the four 16 x 16 diagonal blocks of the current diagonal tile
    double zbuf[GRAD ? 2 : GPL_LML_ZMAX];  // z = L^-1 y of the current item (global workspace instead when n is larger)
    ItemScalars sc;
    double rsbuf[16];
    double pivbuf[TS];
    double ybuf[TS];
    double tmp16[16];
    double L16s[256];
    double gsum[NWARPS * GPL_MAX_THETA];  // one row per warp (deterministic summation order)
    double red[NWARPS];
    double logdet;
    int item;
    int info;
};

__device__ __forceinline__ void block_indices(const TMap &tm, int i, int j, int (&gi)[2], int (&gj)[NCC]) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) gi[mb] = i * TS + row_of(tm, mb);
#pragma unroll
    for (int cc = 0; cc < NCC; ++cc) gj[cc] = j * TS + col_of(tm, cc);
}

// acc (+/-)= sum_{k0 <= k < k1} A_k B_k' over whole tiles, both operands streamed in 16-column chunks through a 4-stage
// cp.async ring (ring = 64 KiB: stage = 8 KiB of A + 8 KiB of B; three stages in flight, one barrier per stage).
// srcA(k) / srcB(k) give the tiles; `same`: B_k is A_k (one copy is loaded).
template <bool SUB, class FA, class FB>
__device__ __forceinline__ void ring_tile_mma(double (&acc)[2][NCC], double *ring, FA srcA, FB srcB, int k0, int k1,
                                              bool same, const TMap &tm, int tid) {
    constexpr int RCH = 16 * TS, NS = 4;  // doubles per operand chunk, stages
    const int Q = 4 * (k1 - k0);
    auto issue = [&](int s) {
        if (s < Q) {
            const int k = k0 + (s >> 2), c = s & 3;
            double *dst = ring + (s % NS) * 2 * RCH;
            block_load_async<RCH * 8>(dst, srcA(k) + c * RCH, tid);
            if (!same) block_load_async<RCH * 8>(dst + RCH, srcB(k) + c * RCH, tid);
        }
        cp_async_commit();
    };
    __syncthreads();  // the ring is free
#pragma unroll
    for (int s = 0; s < NS - 1; ++s) issue(s);
    for (int s = 0; s < Q; ++s) {
        cp_async_wait<NS - 2>();
        __syncthreads();
        issue(s + NS - 1);
        const double *a = ring + (s % NS) * 2 * RCH;
        tile_mma<SUB>(acc, a, same ? a : a + RCH, tm, 0, 16);
    }
    cp_async_wait<0>();
}

}  // namespace

template <bool GRAD>
__device__ __forceinline__ void lml_batched_body(const LmlParams &prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LmlSmem<GRAD> &sm = *reinterpret_cast<LmlSmem<GRAD> *>(smem_raw);
    const DevProgram &P = prm.prog;
    const int tid = threadIdx.x, warp = tid >> 5;
    const TMap tm = thread_map(tid);
    const int n = prm.n, nt = prm.nt;
    const long long ntri = tri_index(nt, 0);

    double *wsL = prm.ws + (size_t)blockIdx.x * prm.ws_stride;  // lower tiles of L
    double *wsW = wsL + ntri * TILE_ELEMS;                      // inverses of the diagonal tiles
    double *wsM = wsW + (size_t)nt * TILE_ELEMS;                // (L^-1)' tiles (gradient only)
    double *wsZg = prm.vec + (size_t)blockIdx.x * 2 * nt * TS;  // z = L^-1 y (global copy: alpha phase, large n)
    double *const smD = GRAD ? sm.Bt : sm.Dbuf;
    double *const smZ = GRAD ? sm.Bt + DSIZE : sm.zbuf;
    static_assert(DSIZE + GPL_LML_ZMAX <= TILE_ELEMS, "D and zs fit the second tile buffer");
    double *wsZ = (nt * TS <= GPL_LML_ZMAX) ? smZ : wsZg;
    double *wsAl = wsZg + (size_t)nt * TS;                      // alpha = K^-1 y

    PH_DECL;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            sm.item = prm.counter ? (int)atomicAdd(prm.counter, 1u) : (int)blockIdx.x;
            sm.logdet = 0.0;
            sm.info = 0;
        }
        if (tid < NWARPS * GPL_MAX_THETA) sm.gsum[tid] = 0.0;
        __syncthreads();
        const int b = sm.item;
        if (b >= prm.B) {
            PH_DUMP;
            break;
        }
        const double *X = prm.X + (size_t)b * prm.x_stride;
        const double *Y = prm.Y + (size_t)b * prm.y_stride;
        const double *theta = prm.Theta + (size_t)b * prm.p;
        const double diag_add = prm.sigma2[(size_t)b * prm.sigma2_stride] + prm.jitter;
        prepare_item_scalars(P, theta, &sm.sc, tid);

        // ------------------------------------------------------------------ factorisation ------------
        for (int j = 0; j < nt; ++j) {
            for (int i = j; i < nt; ++i) {
                const bool diag = (i == j);
                const int Q = (TS / LKC) * j;   // update steps: LKC columns of L_ik / L_jk each, k = 0..j-1
                const int QT = (diag || !GPL_LML_STREAM_TRSM) ? Q : Q + 3;  // + 3 steps streaming L_jj for the solve
                const double *srcA = wsL + tri_index(i, 0) * TILE_ELEMS;  // tiles (i, 0..j-1) are contiguous
                const double *srcB = wsL + tri_index(j, 0) * TILE_ELEMS;  // tiles (j, 0..j-1), then L_jj itself
                // one load pipeline for the whole tile: step s < Q stages 16 columns of both operands of the
                // update; step Q + q stages columns 16q.. of L_jj (they follow tile (j, j-1) in the workspace)
                auto issue = [&](int s) {
                    const int st = (s & 1) * LCH;
                    if (s < Q) block_load_async<LCH * 8>(sm.S + st, srcA + (size_t)s * LCH, tid);
                    if (!diag) block_load_async<LCH * 8>(sm.S + 2 * LCH + st, srcB + (size_t)s * LCH, tid);
                    cp_async_commit();
                };
                (void)QT;
                // all readers of S (previous tile) are done: start the first loads, then generate the covariance
                // tile while they are in flight
                PH_MARK(7);
                __syncthreads();
                PH_MARK(0);
                if (QT > 0) issue(0);
                double acc[2][NCC];
                {
                    int gi[2];
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb) gi[mb] = i * TS + row_of(tm, mb);
                    // a diagonal tile is only needed on and below the diagonal: warp w (rows 16w..) skips the 16-column
                    // quarters h > w here and the n-blocks nb > 2w + 1 in its update loop
                    eval_block_acc<true>(P, sm.sc, X, n, n, gi, X, n, n, j * TS, tm.t, diag_add, acc, diag ? warp + 1 : 4);
                }
                PH_MARK(1);
                double ytmp = 0.0;
                if (diag && tid < TS) ytmp = (j * TS + tid < n) ? Y[j * TS + tid] : 0.0;
                for (int q = 0; q < Q; ++q) {
                    cp_async_wait<0>();
                    __syncthreads();  // step q landed for everyone; everyone finished step q-1
                    if (q + 1 < QT) issue(q + 1);
                    // 16 whole columns starting at a multiple of 4 are themselves in tile format (swizzle uses c & 3)
                    const double *a = sm.S + (q & 1) * LCH;
                    const double *bt = a + 2 * LCH;
                    if (diag) tile_mma<true, 0xFF, true>(acc, a, a, tm, 0, LKC, 2 * warp + 2);
                    else tile_mma<true>(acc, a, bt, tm, 0, LKC);
                    if (diag && tid < TS) ytmp -= tile_row_dot(a, wsZ + q * LKC, tid, 0, LKC);
                }
                PH_MARK(2);
                if (diag) {
                    __syncthreads();  // S becomes the factorisation scratch
                    const int fail = tile_potrf(acc, tm, sm.S, sm.L16s, smD, sm.rsbuf, sm.pivbuf, tid);
                    PH_MARK(3);
                    if (tid == 0 && fail >= 0 && sm.info == 0) sm.info = j * TS + fail + 1;
                    acc_to_tile(wsL + tri_index(j, j) * TILE_ELEMS, acc, tm);
                    __syncthreads();  // scratch no longer read
                    acc_to_tile(sm.S, acc, tm);  // L_jj in shared memory for the forward solve (and the inverse)
                    if (tid < TS) sm.ybuf[tid] = ytmp;
                    tile_forward_solve(sm.S, smD, sm.ybuf, sm.tmp16, tid);  // z_j = L_jj^-1 (y_j - sum_k L_jk z_k)
                    if (tid < TS) {
                        wsZ[j * TS + tid] = sm.ybuf[tid];
                        if (wsZ != wsZg && (prm.want_grad || prm.keep)) wsZg[j * TS + tid] = sm.ybuf[tid];
                    }
                    if (tid < 32) {
                        double lg = log(sm.pivbuf[tid]) + log(sm.pivbuf[tid + 32]);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
                        if (tid == 0) sm.logdet += lg;
                    }
                    if (prm.want_grad || prm.keep) {
                        // full inverse W_jj = L_jj^-1 for the alpha / gradient / prediction phases: X L' = I gives W'
                        double e[2][NCC];
#pragma unroll
                        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                            for (int cc = 0; cc < NCC; ++cc) e[mb][cc] = (row_of(tm, mb) == col_of(tm, cc)) ? 1.0 : 0.0;
                        tile_trsm_ld(e, sm.S, smD, tm);
                        acc_to_tile_t(wsW + (size_t)j * TILE_ELEMS, e, tm);
                    }
                    PH_MARK(4);
                } else {
                    // L_ij = T_ij L_jj^-T, right-looking over the four 16-column panels: the row operands come from
                    // the accumulator registers, the block inverses from D (still there from the diagonal tile), and
                    // the columns of L_jj arrive through the same pipeline (steps Q, Q+1, Q+2; Q is even)
#if !GPL_LML_STREAM_TRSM
                    __syncthreads();
                    tile_load_async(sm.S, wsL + tri_index(j, j) * TILE_ELEMS, tid);
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncthreads();
                    tile_trsm_ld(acc, sm.S, smD, tm);
#else
                    trsm_rl_solve<0>(acc, smD, tm);
                    cp_async_wait<0>();
                    __syncthreads();
                    issue(Q + 1);
                    trsm_rl_update<0>(acc, sm.S + 2 * LCH, tm);
                    trsm_rl_solve<1>(acc, smD, tm);
                    cp_async_wait<0>();
                    __syncthreads();
                    issue(Q + 2);
                    trsm_rl_update<1>(acc, sm.S + 3 * LCH, tm);
                    trsm_rl_solve<2>(acc, smD, tm);
                    cp_async_wait<0>();
                    __syncthreads();
                    trsm_rl_update<2>(acc, sm.S + 2 * LCH, tm);
                    trsm_rl_solve<3>(acc, smD, tm);
#endif
                    PH_MARK(5);
                    acc_to_tile(wsL + tri_index(i, j) * TILE_ELEMS, acc, tm);
                    PH_MARK(6);
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------------ lml ----------------------
        double q = 0.0;
        for (int t = tid; t < nt * TS; t += NTHREADS) {
            const double z = wsZ[t];
            q = fma(z, z, q);
        }
        q = block_sum(q, sm.red, tid);
        const int info = sm.info;
        if (tid == 0) {
            const double val = -0.5 * ((double)n * LOG2PI + sm.logdet + q);
            prm.lml[b] = info ? -INFINITY : val;
            if (prm.info) prm.info[b] = info;
        }
        if (!prm.want_grad && !prm.keep) continue;

        // ------------------------------------------------------------------ alpha = L^-T z ------------
        // backward substitution by tiles: alpha_j = W_jj' (z_j - sum_{i>j} L_ij' alpha_i)
        for (int j = nt - 1; j >= 0; --j) {
            double rj = 0.0;
            if (tid < TS) rj = wsZ[j * TS + tid];
            for (int i = nt - 1; i > j; --i) {
                __syncthreads();
                tile_load_async(sm.S, wsL + tri_index(i, j) * TILE_ELEMS, tid);
                cp_async_commit();
                cp_async_wait<0>();
                if (tid < TS) sm.ybuf[tid] = wsAl[i * TS + tid];
                __syncthreads();
                if (tid < TS) rj -= tile_col_dot(sm.S, sm.ybuf, tid);
            }
            __syncthreads();
            tile_load_async(sm.S, wsW + (size_t)j * TILE_ELEMS, tid);
            cp_async_commit();
            cp_async_wait<0>();
            if (tid < TS) sm.ybuf[tid] = rj;
            __syncthreads();
            if (tid < TS) {
                const double a = tile_col_dot(sm.S, sm.ybuf, tid);  // W upper part is zero
                wsAl[j * TS + tid] = a;
                if (prm.dy && j * TS + tid < n) prm.dy[(size_t)b * n + j * TS + tid] = info ? NAN : -a;
            }
        }
        __syncthreads();
        if (!GRAD || !prm.want_grad) continue;

        if constexpr (GRAD) {
        // ------------------------------------------------------------------ M = L^-1, stored as M' tiles ----
        // M_jj = W_jj ; M_ij = -W_ii * sum_{k=j}^{i-1} L_ik M_kj  (i > j).  wsM holds the TRANSPOSE of every tile of
        // M, so that M_kj can be the column operand of C += A B' (B'(k, n) = M_kj(k, n)).
        for (int j = 0; j < nt; ++j) {
            for (int i = j; i < nt; ++i) {
                double acc[2][NCC];
                if (i == j) {
                    __syncthreads();
                    tile_load_async(sm.S, wsW + (size_t)j * TILE_ELEMS, tid);
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncthreads();
                    acc_from_tile(acc, sm.S, tm);
                    acc_to_tile_t(wsM + tri_index(j, j) * TILE_ELEMS, acc, tm);
                    continue;
                }
                acc_zero(acc);
                ring_tile_mma<false>(  // S += L_ik M_kj
                    acc, sm.S, [&](int k) { return wsL + tri_index(i, k) * TILE_ELEMS; },
                    [&](int k) { return wsM + tri_index(k, j) * TILE_ELEMS; }, j, i, false, tm, tid);
                // M_ij = -W_ii S : row operand W_ii, column operand S' (transposed store of the accumulator)
                __syncthreads();
                tile_load_async(sm.S, wsW + (size_t)i * TILE_ELEMS, tid);
                cp_async_commit();
                acc_to_tile_t(sm.Bt, acc, tm);
                cp_async_wait<0>();
                __syncthreads();
                double mij[2][NCC];
                acc_zero(mij);
                tile_mma<true>(mij, sm.S, sm.Bt, tm, 0, TS);
                acc_to_tile_t(wsM + tri_index(i, j) * TILE_ELEMS, mij, tm);
            }
        }

        // ------------------------------------------------------------------ K^-1 tiles and the contraction ---
        // P_ij = sum_{k >= i} M_ki' M_kj (i >= j): row operand M_ki' and column operand M_kj' are the stored tiles.
        // dlml/dtheta_s = -1/2 sum_ij (P - alpha alpha')_ij dK_ij/dtheta_s.
        for (int j = 0; j < nt; ++j) {
            for (int i = j; i < nt; ++i) {
                double acc[2][NCC];
                acc_zero(acc);
                ring_tile_mma<false>(
                    acc, sm.S, [&](int k) { return wsM + tri_index(k, i) * TILE_ELEMS; },
                    [&](int k) { return wsM + tri_index(k, j) * TILE_ELEMS; }, i, nt, i == j, tm, tid);
                int gi[2], gj[NCC];
                block_indices(tm, i, j, gi, gj);
                const double sym = (i == j) ? 1.0 : 2.0;
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int cc = 0; cc < NCC; ++cc) acc[mb][cc] = sym * (acc[mb][cc] - wsAl[gi[mb]] * wsAl[gj[cc]]);
                __syncthreads();  // the tiles in S / Bt are consumed: S parks the weights
                contract_grad_block(P, sm.sc, X, n, n, gi, j * TS, tm.t, acc, sm.S, tid, sm.gsum + warp * GPL_MAX_THETA);
            }
        }
        __syncthreads();
        if (tid < prm.p) {
            double g = 0.0;
#pragma unroll
            for (int w = 0; w < NWARPS; ++w) g += sm.gsum[w * GPL_MAX_THETA + tid];
            prm.dtheta[(size_t)b * prm.p + tid] = info ? NAN : -0.5 * g;
        }
        }  // GRAD
    }
}

__global__ void __launch_bounds__(NTHREADS, GPL_LML_CTAS_PER_SM) lml_batched_kernel(const __grid_constant__ LmlParams prm) {
    lml_batched_body<false>(prm);
}
__global__ void __launch_bounds__(NTHREADS, 3) lml_batched_grad_kernel(const __grid_constant__ LmlParams prm) {
    lml_batched_body<true>(prm);
}

size_t lml_smem_bytes(bool grad) { return grad ? sizeof(LmlSmem<true>) : sizeof(LmlSmem<false>); }

}  // namespace gpl
