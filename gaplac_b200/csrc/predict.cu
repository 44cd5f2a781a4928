// Posterior mean / variance at test points, and prior sampling, from a factor resident in HBM.
//
// Replaces mean_and_var(PosteriorGP, X*) [upstream AbstractGPs 0.5.12] (call site src/plotting.jl:12):
//   mean = K(X*, X) alpha ;  var = diag K(X*, X*) - colsumsq(U' \ K(X, X*))   (latent f; sigma2 not added back)
// and rand(FiniteGP) = U' z (call site CLI/src/sample.jl:25).
//
// predict_kernel: each CTA owns slabs of 64 test points.  K* is generated tile by tile in registers (never
// materialised), the triangular solve V = L^-1 K* is blocked on 64 x 64 tiles with the inverted diagonal
// tiles (all GEMM), and the column sums of squares / the mean are accumulated on the fly.
#include "kernels.h"
#include "kfun.cuh"
#include "tile.cuh"

namespace gpl {

namespace {
struct __align__(16) PredSmem {
    double A[TILE_ELEMS];  // at the end of a slab its first 16 KiB take the 32 partial column sums per column
    double Bt[TILE_ELEMS];
    ItemScalars sc;
};
}  // namespace

size_t predict_smem_bytes() { return sizeof(PredSmem); }

// NB: 8-column blocks of test points per slab (slab width 8 NB <= 64).  The host picks the width that balances the slabs
// over the SMs (20 000 points: 313 slabs of 64 put three slabs on 17 SMs and two on the rest; 417 slabs of 48 finish
// ~20 % sooner although each DMMA step feeds six n-blocks instead of eight).
template <int NB>
__device__ __forceinline__ void predict_body(const PredictParams &prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PredSmem &sm = *reinterpret_cast<PredSmem *>(smem_raw);
    const DevProgram &P = prm.prog;
    const int tid = threadIdx.x;
    const TMap tm = thread_map(tid);
    const int n = prm.n, nt = prm.nt, m = prm.m;
    double *const part = sm.A;  // 32 * TS doubles, used between the tile loops only
    double *wsV = prm.wsV + (size_t)blockIdx.x * nt * TILE_ELEMS;
    const int items = prm.items > 0 ? prm.items : 1;
    int cur = -1;

    constexpr int WS = 8 * NB, MASK = (1 << NB) - 1;  // slab width, active n-blocks
    static_assert(NB % 2 == 0 && NB >= 2 && NB <= 8, "whole 16-column quarters");
    const int nslab = (m + WS - 1) / WS;
    for (long long u = blockIdx.x; u < (long long)items * nslab; u += gridDim.x) {
        const int b = (int)(u / nslab), s = (int)(u - (long long)b * nslab);
        if (b != cur) {  // a new posterior: its hyperparameters
            __syncthreads();
            prepare_item_scalars(P, prm.theta + b * prm.theta_stride, &sm.sc, tid);
            __syncthreads();
            cur = b;
        }
        const double *tiles = prm.tiles + b * prm.tiles_stride, *winv = prm.winv + b * prm.winv_stride;
        const double *alpha = prm.alpha + b * prm.alpha_stride;
        double *mean = prm.mean + (size_t)b * m, *var = prm.var ? prm.var + (size_t)b * m : nullptr;
        const bool bad = prm.info && prm.info[b] != 0;  // not positive definite: NaN, like the lml's -Inf
        int gj[NCC];
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) gj[cc] = s * WS + col_of(tm, cc);
        double csq[NCC], cmean[NCC];
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) csq[cc] = cmean[cc] = 0.0;
        for (int i = 0; i < nt; ++i) {
            int gi[2];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) gi[mb] = i * TS + row_of(tm, mb);
            double acc[2][NCC];
            eval_block_acc<false>(P, sm.sc, prm.X, n, n, gi, prm.Xs, m, m, s * WS, tm.t, 0.0, acc, NB / 2);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const double al = gi[mb] < n ? alpha[gi[mb]] : 0.0;
#pragma unroll
                for (int cc = 0; cc < NCC; ++cc) cmean[cc] = fma(acc[mb][cc], al, cmean[cc]);
            }
            if (!prm.want_var) continue;
            // acc -= sum_{k<i} L_ik V_k : row operand L_ik, column operand V_k' (stored transposed)
            for (int k = 0; k < i; ++k) {
                __syncthreads();
                tile_load_async(sm.A, tiles + tri_index(i, k) * TILE_ELEMS, tid);
                tile_load_async(sm.Bt, wsV + (size_t)k * TILE_ELEMS, tid);
                cp_async_commit();
                cp_async_wait<0>();
                __syncthreads();
                tile_mma<true, MASK>(acc, sm.A, sm.Bt, tm, 0, TS);
            }
            // V_i = W_ii acc  (W lower triangular: rows of this warp need k < r0 + 16)
            __syncthreads();
            tile_load_async(sm.A, winv + (size_t)i * TILE_ELEMS, tid);
            cp_async_commit();
            acc_to_tile_t(sm.Bt, acc, tm);
            cp_async_wait<0>();
            __syncthreads();
            acc_zero(acc);
            tile_mma<false, MASK>(acc, sm.A, sm.Bt, tm, 0, tm.r0 + 16);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int cc = 0; cc < NCC; ++cc) csq[cc] = fma(acc[mb][cc], acc[mb][cc], csq[cc]);
            if (i + 1 < nt) acc_to_tile_t(wsV + (size_t)i * TILE_ELEMS, acc, tm);
        }
        // reduce the per-thread partials: 32 partials per column (4 warps x 8 lane groups g), fixed order
        const int slot = (tid >> 5) * 8 + tm.g;
        __syncthreads();
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) part[slot * TS + col_of(tm, cc)] = cmean[cc];
        __syncthreads();
        if (tid < WS && s * WS + tid < m) {
            double sum = 0.0;
#pragma unroll 8
            for (int gq = 0; gq < 32; ++gq) sum += part[gq * TS + tid];
            mean[s * WS + tid] = bad ? NAN : sum;
        }
        if (prm.want_var) {
            __syncthreads();
#pragma unroll
            for (int cc = 0; cc < NCC; ++cc) part[slot * TS + col_of(tm, cc)] = csq[cc];
            __syncthreads();
            if (tid < WS && s * WS + tid < m) {
                double sum = 0.0;
#pragma unroll 8
                for (int gq = 0; gq < 32; ++gq) sum += part[gq * TS + tid];
                // prior variance of the latent function at x*: Noise contributes 0 (SAME = false)
                int one_i[1] = {s * WS + tid}, one_j[1] = {s * WS + tid};
                double kss[1][1];
                eval_block<1, 1, false>(P, sm.sc, prm.Xs, m, m, one_i, prm.Xs, m, m, one_j, 0.0, kss);
                var[s * WS + tid] = bad ? NAN : kss[0][0] - sum;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NTHREADS, 3) predict_kernel(const __grid_constant__ PredictParams prm) { predict_body<8>(prm); }
__global__ void __launch_bounds__(NTHREADS, 3) predict_kernel_nb6(const __grid_constant__ PredictParams prm) { predict_body<6>(prm); }
__global__ void __launch_bounds__(NTHREADS, 3) predict_kernel_nb4(const __grid_constant__ PredictParams prm) { predict_body<4>(prm); }

// out (n x S) = L Z : grid = nt CTAs (tile row i), thread (row = tid % 64, sample lane = tid / 64: 2 samples at a time)
__global__ void __launch_bounds__(NTHREADS) sample_kernel(const double *tiles, int nt, int n, const double *Z, int S,
                                                          double *out) {
    __shared__ __align__(16) double T[TILE_ELEMS];
    const int tid = threadIdx.x, i = blockIdx.x;
    const int row = tid & (TS - 1), sl = tid >> 6;
    for (int s0 = 0; s0 < S; s0 += NTHREADS / TS) {
        const int s = s0 + sl;
        double acc = 0.0;
        for (int k = 0; k <= i; ++k) {
            __syncthreads();
            tile_load_async(T, tiles + tri_index(i, k) * TILE_ELEMS, tid);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            if (s < S) {
                for (int kk = 0; kk < TS; ++kk) {
                    const int g = k * TS + kk;
                    const double z = g < n ? Z[(size_t)s * n + g] : 0.0;
                    acc = fma(T[tidx(row, kk)], z, acc);
                }
            }
        }
        if (s < S && i * TS + row < n) out[(size_t)s * n + i * TS + row] = acc;
    }
}

}  // namespace gpl
