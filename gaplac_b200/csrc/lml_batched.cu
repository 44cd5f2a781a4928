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
// factored tiles streamed from the CTA's private workspace (L2-resident), then factored (diagonal tile) or
// multiplied by the inverse of the diagonal tile (below-diagonal tiles).  The forward solve z = L^-1 y and
// logdet ride along with the diagonal tiles.  The gradient phase forms M = L^-1 and K^-1 tile by tile and
// contracts (K^-1 - alpha alpha') with dK/dtheta generated on the fly.
#include "kernels.h"
#include "kfun.cuh"
#include "tile.cuh"

namespace gpl {

namespace {

constexpr double LOG2PI = 1.8378770664093454835606594728112;

struct __align__(16) LmlSmem {
    double A[TILE_ELEMS];
    double Bt[TILE_ELEMS];
    double W[TILE_ELEMS];
    ItemScalars sc;
    double colbuf[2 * TS];
    double rowbuf[2 * TS];
    double pivbuf[TS];
    double ybuf[TS];
    double L16s[256];
    double W16s[256];
    double gsum[GPL_MAX_THETA];
    double red[8];
    double logdet;
    int item;
    int info;
};

// conflict-free "diagonal" traversal of a column-major tile: thread c walks T[(c+s)%64][c]
__device__ __forceinline__ double col_dot_diag(const double *T, const double *v, int c) {
    double s = 0.0;
#pragma unroll 8
    for (int t = 0; t < TS; ++t) {
        const int m = (c + t) & (TS - 1);
        s = fma(T[c * TS + m], v[m], s);
    }
    return s;
}

}  // namespace

// V = 0: first version (one barrier per pivot, single-buffered tile loads) — kept as the in-tree A/B reference.
// V = 1: blocked warp-level diagonal factorisation, double-buffered half-tile loads, direct global stores.
template <int V>
__device__ __forceinline__ void lml_batched_body(const LmlParams &prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LmlSmem &sm = *reinterpret_cast<LmlSmem *>(smem_raw);
    const DevProgram &P = prm.prog;
    const int tid = threadIdx.x;
    const TMap tm = thread_map(tid);
    const int n = prm.n, nt = prm.nt;
    const long long ntri = tri_index(nt, 0);

    double *wsL = prm.ws + (size_t)blockIdx.x * prm.ws_stride;  // lower tiles of L
    double *wsW = wsL + ntri * TILE_ELEMS;                      // inverses of the diagonal tiles
    double *wsM = wsW + (size_t)nt * TILE_ELEMS;                // (L^-1)' tiles (gradient only)
    double *wsZ = prm.vec + (size_t)blockIdx.x * 2 * nt * TS;   // z = L^-1 y
    double *wsAl = wsZ + (size_t)nt * TS;                       // alpha = K^-1 y

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            sm.item = prm.counter ? (int)atomicAdd(prm.counter, 1u) : (int)blockIdx.x;
            sm.logdet = 0.0;
            sm.info = 0;
        }
        if (tid < GPL_MAX_THETA) sm.gsum[tid] = 0.0;
        __syncthreads();
        const int b = sm.item;
        if (b >= prm.B) break;
        const double *X = prm.X + (size_t)b * prm.x_stride;
        const double *Y = prm.Y + (size_t)b * prm.y_stride;
        const double *theta = prm.Theta + (size_t)b * prm.p;
        const double diag_add = prm.sigma2[(size_t)b * prm.sigma2_stride] + prm.jitter;
        prepare_item_scalars(P, theta, &sm.sc, tid);
        __syncthreads();

        // ------------------------------------------------------------------ factorisation ------------
        if (V == 0)
        for (int j = 0; j < nt; ++j) {
            for (int i = j; i < nt; ++i) {
                double acc[4][4];
                {
                    int gi[4], gj[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) gi[r] = i * TS + tm.m0 + r;
#pragma unroll
                    for (int c = 0; c < 4; ++c) gj[c] = j * TS + col_of(tm.cb, c);
                    eval_block<4, 4, true>(P, sm.sc, X, n, n, gi, X, n, n, gj, diag_add, acc);
                }
                const bool diag = (i == j);
                double ytmp = 0.0;
                if (diag && tid < TS) ytmp = (j * TS + tid < n) ? Y[j * TS + tid] : 0.0;
                for (int k = 0; k < j; ++k) {
                    tile_load_async(sm.A, wsL + tri_index(i, k) * TILE_ELEMS, tid);
                    if (!diag) tile_load_async(sm.Bt, wsL + tri_index(j, k) * TILE_ELEMS, tid);
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncthreads();
                    tile_gemm<true>(acc, sm.A, diag ? sm.A : sm.Bt, tm, 0, TS);
                    if (diag && tid < TS) {
                        const double *zk = wsZ + k * TS;
                        double s = 0.0;
#pragma unroll 8
                        for (int kk = 0; kk < TS; ++kk) s = fma(sm.A[kk * TS + tid], zk[kk], s);
                        ytmp -= s;
                    }
                    __syncthreads();
                }
                if (diag) {
                    double w[4][4];
                    const int fail = tile_potrf_inv(acc, w, tm, sm.colbuf, sm.rowbuf, sm.pivbuf, tid);
                    if (tid == 0 && fail >= 0 && sm.info == 0) sm.info = j * TS + fail + 1;
                    acc_to_smem(sm.A, acc, tm);
                    acc_to_smem(sm.W, w, tm);
                    if (tid < TS) sm.ybuf[tid] = ytmp;
                    __syncthreads();
                    if (tid < TS) {
                        // z_j = W y  (W lower triangular: columns c <= row)
                        double s = 0.0;
                        for (int c = 0; c <= tid; ++c) s = fma(sm.W[c * TS + tid], sm.ybuf[c], s);
                        wsZ[j * TS + tid] = s;
                    }
                    if (tid < 32) {
                        double lg = log(sm.pivbuf[tid]) + log(sm.pivbuf[tid + 32]);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
                        if (tid == 0) sm.logdet += lg;
                    }
                    tile_store(wsL + tri_index(j, j) * TILE_ELEMS, sm.A, tid);
                    if (prm.want_grad || prm.keep) tile_store(wsW + (size_t)j * TILE_ELEMS, sm.W, tid);
                    __syncthreads();
                } else {
                    // L_ij = T_ij * W_jj'  : X[m][n] = sum_k T[m][k] W[n][k], k <= n
                    acc_to_smem(sm.A, acc, tm);
                    __syncthreads();
                    double x[4][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) x[r][c] = 0.0;
                    const int kmax = ((tid >> 5) & 1) * 32 + 32;  // columns of this warp are < kmax
                    tile_gemm<false>(x, sm.A, sm.W, tm, 0, kmax);
                    __syncthreads();
                    acc_to_smem(sm.A, x, tm);
                    __syncthreads();
                    tile_store(wsL + tri_index(i, j) * TILE_ELEMS, sm.A, tid);
                    __syncthreads();
                }
            }
        }

        if (V == 1)
        for (int j = 0; j < nt; ++j) {
            for (int i = j; i < nt; ++i) {
                const bool diag = (i == j);
                const int Q = 2 * j;  // half-steps: (k, h) = (q >> 1, q & 1), 32 columns of L_ik / L_jk each
                const double *srcA = wsL + tri_index(i, 0) * TILE_ELEMS;  // tiles (i, 0..j-1) are contiguous
                const double *srcB = wsL + tri_index(j, 0) * TILE_ELEMS;
                // all readers of the staging buffers (previous tile) are done; start the first loads, then
                // generate the covariance tile while they are in flight
                __syncthreads();
                if (Q > 0) {
                    half_tile_load_async(sm.A, srcA, tid);
                    if (!diag) half_tile_load_async(sm.Bt, srcB, tid);
                    cp_async_commit();
                }
                double acc[4][4];
                {
                    int gi[4], gj[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) gi[r] = i * TS + tm.m0 + r;
#pragma unroll
                    for (int c = 0; c < 4; ++c) gj[c] = j * TS + col_of(tm.cb, c);
                    eval_block<4, 4, true>(P, sm.sc, X, n, n, gi, X, n, n, gj, diag_add, acc);
                }
                double ytmp = 0.0;
                if (diag && tid < TS) ytmp = (j * TS + tid < n) ? Y[j * TS + tid] : 0.0;
                for (int q = 0; q < Q; ++q) {
                    cp_async_wait<0>();
                    __syncthreads();  // half-step q landed for everyone; everyone finished half-step q-1
                    if (q + 1 < Q) {
                        const int nb = ((q + 1) & 1) * (TILE_ELEMS / 2);
                        half_tile_load_async(sm.A + nb, srcA + (size_t)(q + 1) * (TILE_ELEMS / 2), tid);
                        if (!diag) half_tile_load_async(sm.Bt + nb, srcB + (size_t)(q + 1) * (TILE_ELEMS / 2), tid);
                        cp_async_commit();
                    }
                    const double *a = sm.A + (q & 1) * (TILE_ELEMS / 2);
                    const double *bt = diag ? a : sm.Bt + (q & 1) * (TILE_ELEMS / 2);
                    tile_gemm<true>(acc, a, bt, tm, 0, TS / 2);
                    if (diag && tid < TS) {
                        const double *zk = wsZ + q * (TS / 2);
                        double s = 0.0;
#pragma unroll 8
                        for (int kk = 0; kk < TS / 2; ++kk) s = fma(a[kk * TS + tid], zk[kk], s);
                        ytmp -= s;
                    }
                }
                if (diag) {
                    __syncthreads();  // staging buffers become the factorisation scratch
                    double w[4][4];
                    const int fail = tile_potrf_inv_blocked(acc, w, tm, sm.A, sm.L16s, sm.W16s, sm.colbuf, sm.pivbuf, tid);
                    if (tid == 0 && fail >= 0 && sm.info == 0) sm.info = j * TS + fail + 1;
                    acc_to_global(wsL + tri_index(j, j) * TILE_ELEMS, acc, tm);
                    acc_to_smem(sm.W, w, tm);
                    if (tid < TS) sm.ybuf[tid] = ytmp;
                    __syncthreads();
                    if (tid < TS) {
                        double s = 0.0;
                        for (int c = 0; c <= tid; ++c) s = fma(sm.W[c * TS + tid], sm.ybuf[c], s);
                        wsZ[j * TS + tid] = s;
                    }
                    if (tid < 32) {
                        double lg = log(sm.pivbuf[tid]) + log(sm.pivbuf[tid + 32]);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
                        if (tid == 0) sm.logdet += lg;
                    }
                    if (prm.want_grad || prm.keep) tile_store(wsW + (size_t)j * TILE_ELEMS, sm.W, tid);
                } else {
                    // L_ij = T_ij * W_jj' through shared memory (T staged in the load buffer), result stored
                    // straight from registers
                    __syncthreads();
                    acc_to_smem(sm.A, acc, tm);
                    __syncthreads();
                    double x[4][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) x[r][c] = 0.0;
                    const int kmax = ((tid >> 5) & 1) * 32 + 32;
                    tile_gemm<false>(x, sm.A, sm.W, tm, 0, kmax);
                    acc_to_global(wsL + tri_index(i, j) * TILE_ELEMS, x, tm);
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
                tile_load_async(sm.A, wsL + tri_index(i, j) * TILE_ELEMS, tid);
                cp_async_commit();
                cp_async_wait<0>();
                if (tid < TS) sm.ybuf[tid] = wsAl[i * TS + tid];
                __syncthreads();
                if (tid < TS) rj -= col_dot_diag(sm.A, sm.ybuf, tid);
            }
            __syncthreads();
            tile_load_async(sm.A, wsW + (size_t)j * TILE_ELEMS, tid);
            cp_async_commit();
            cp_async_wait<0>();
            if (tid < TS) sm.ybuf[tid] = rj;
            __syncthreads();
            if (tid < TS) {
                const double a = col_dot_diag(sm.A, sm.ybuf, tid);  // W upper part is zero
                wsAl[j * TS + tid] = a;
                if (prm.dy && j * TS + tid < n) prm.dy[(size_t)b * n + j * TS + tid] = info ? NAN : -a;
            }
        }
        __syncthreads();
        if (!prm.want_grad) continue;

        // ------------------------------------------------------------------ M = L^-1, stored as M' tiles ----
        // M_jj = W_jj ; M_ij = -W_ii * sum_{k=j}^{i-1} L_ik M_kj  (i > j).  Tile (i,j) of M is kept
        // TRANSPOSED in wsM (element (r,c) at r*64 + c) so that it can be the B operand of the GEMM core.
        for (int j = 0; j < nt; ++j) {
            for (int i = j; i < nt; ++i) {
                double acc[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
                if (i == j) {
                    tile_load_async(sm.A, wsW + (size_t)j * TILE_ELEMS, tid);
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncthreads();
                    // transpose through registers: read column-major, write transposed
                    double t4[4][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) t4[r][c] = sm.A[col_of(tm.cb, c) * TS + tm.m0 + r];
                    __syncthreads();
                    acc_to_smem_t(sm.A, t4, tm);
                    __syncthreads();
                    tile_store(wsM + tri_index(j, j) * TILE_ELEMS, sm.A, tid);
                    __syncthreads();
                    continue;
                }
                for (int k = j; k < i; ++k) {
                    // S += L_ik * M_kj : A = L_ik (col-major), B[kk][n] = M_kj[kk][n] = transposed storage
                    tile_load_async(sm.A, wsL + tri_index(i, k) * TILE_ELEMS, tid);
                    tile_load_async(sm.Bt, wsM + tri_index(k, j) * TILE_ELEMS, tid);
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncthreads();
                    tile_gemm<false>(acc, sm.A, sm.Bt, tm, 0, TS);
                    __syncthreads();
                }
                // M_ij = -W_ii * S : A = W_ii (col-major), B[kk][n] = S[kk][n] (transposed store of acc)
                tile_load_async(sm.A, wsW + (size_t)i * TILE_ELEMS, tid);
                cp_async_commit();
                acc_to_smem_t(sm.Bt, acc, tm);
                cp_async_wait<0>();
                __syncthreads();
                double mij[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) mij[r][c] = 0.0;
                tile_gemm<true>(mij, sm.A, sm.Bt, tm, 0, TS);
                __syncthreads();
                acc_to_smem_t(sm.A, mij, tm);
                __syncthreads();
                tile_store(wsM + tri_index(i, j) * TILE_ELEMS, sm.A, tid);
                __syncthreads();
            }
        }

        // ------------------------------------------------------------------ K^-1 tiles and the contraction ---
        // P_ij = sum_{k >= i} M_ki' M_kj (i >= j).  With M' tiles (element (r, m) of M_ki at r*64 + m) both
        // operands are in GEMM-core layout.  dlml/dtheta_s = -1/2 sum_ij (P - alpha alpha')_ij dK_ij/dtheta_s.
        for (int j = 0; j < nt; ++j) {
            for (int i = j; i < nt; ++i) {
                double acc[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
                for (int k = i; k < nt; ++k) {
                    tile_load_async(sm.A, wsM + tri_index(k, i) * TILE_ELEMS, tid);
                    if (i != j) tile_load_async(sm.Bt, wsM + tri_index(k, j) * TILE_ELEMS, tid);
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncthreads();
                    tile_gemm<false>(acc, sm.A, (i == j) ? sm.A : sm.Bt, tm, 0, TS);
                    __syncthreads();
                }
                int gi[4], gj[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) gi[r] = i * TS + tm.m0 + r;
#pragma unroll
                for (int c = 0; c < 4; ++c) gj[c] = j * TS + col_of(tm.cb, c);
                const double sym = (i == j) ? 1.0 : 2.0;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        acc[r][c] = sym * (acc[r][c] - wsAl[gi[r]] * wsAl[gj[c]]);
                contract_grad_block<4, 4>(P, sm.sc, X, n, n, gi, gj, acc, sm.gsum);
            }
        }
        __syncthreads();
        if (tid < prm.p) prm.dtheta[(size_t)b * prm.p + tid] = info ? NAN : -0.5 * sm.gsum[tid];
    }
}

__global__ void __launch_bounds__(NTHREADS, 2) lml_batched_kernel(const __grid_constant__ LmlParams prm) {
    lml_batched_body<1>(prm);
}
__global__ void __launch_bounds__(NTHREADS, 2) lml_batched_kernel_v0(const __grid_constant__ LmlParams prm) {
    lml_batched_body<0>(prm);
}

size_t lml_smem_bytes() { return sizeof(LmlSmem); }

}  // namespace gpl
