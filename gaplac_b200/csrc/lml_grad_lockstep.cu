// Analytic gradient of the batched log marginal likelihood on the lockstep schedule.
//
// Replaces the ForwardDiff pass of the reference's mcmc model body (CLI/src/mcmc.jl:31-37: every leapfrog step of
// NUTS re-runs kernelmatrix + a generic Cholesky on Dual numbers, ~5 factorisations per gradient) with
//   dlml/dtheta_s = -1/2 sum_ij (K^-1 - alpha alpha')_ij dK_ij/dtheta_s ,   dlml/dy = -alpha        
// for every item of a batch at once.  The factorisation is the lockstep schedule of lml_lockstep.cu; the phases here
// run over the whole batch too, every CTA a pure stream of DMMAs (no pivot chain shares an SM with them):
//   lk_winv_kernel   grid B*nt      W_jj = L_jj^-1 (identity solved against L_jj and its 16 x 16 block inverses)
//   lk_minv_kernel   grid B*i       row i of M = L^-1:  M_ij = -W_ii sum_{k=j}^{i-1} L_ik M_kj   (one launch per tile row)
//   lk_alpha_kernel  grid B*nt      alpha_j = sum_{k>=j} M_kj' z_k     (= K^-1 y; dy = -alpha)
//   lk_gradc_kernel  grid B*ntri    P_ij = sum_{k>=i} M_ki' M_kj  (tile of K^-1), weights 2 (P - alpha alpha')_ij contracted
//                                   with dK/dtheta generated on the fly (kfun.cuh) -> per-tile partial sums
//   lk_gradsum_kernel               fixed-order sum of the partial sums -> dtheta
// M is kept as TRANSPOSED tiles: the tile product that forms (M_ij)' is then  -(sum_k (M_kj)' L_ik') W_ii' , i.e. the
// row operand is a stored (M_kj)' tile, the column operand a stored L_ik tile and the final triangular product takes its
// row operand from the accumulator registers (tile_trsm_w): no transposition through shared memory anywhere.  In
// column-major tile order the tiles (M_kj)', k = j.., are contiguous, as are L_ik, k = j.. in the row-major factor, so
// every operand stream is one contiguous run moved by a 4-slot cp.async ring (8 columns of both operands per slot).
// K^-1, M' M and dK/dtheta never exist in memory.  n^3 FLOPs per item with the factorisation (n^3/3 each phase).
#include "kernels.h"
#include "kfun.cuh"
#include "tile.cuh"

namespace gpl {

namespace {

constexpr int GNS = 4;            // ring slots
constexpr int GKC = 8;            // columns of each operand per slot
constexpr int GCH = GKC * TS;     // doubles per operand chunk (4 KiB)
constexpr int GSLOT = 2 * GCH;    // doubles per slot (8 KiB)
static_assert(GNS * GSLOT == TILE_ELEMS, "the ring is exactly one tile buffer");

struct __align__(16) WinvSmem {
    double S[TILE_ELEMS];
    double D[DSIZE];
};

struct __align__(16) GradSmem {
    double S[TILE_ELEMS];
    ItemScalars sc;
    double al[2 * TS];  // alpha_i | alpha_j
    double gsum[NWARPS * GPL_MAX_THETA];
    SepCtx sep;
    short kl[GPL_LK_KLMAX];  // tile rows k >= i whose M tiles are not exactly zero (zero-tile skipping)
    int nk;
};

// the first tile of a run is triangular: (M_jj)'(row, kk) = 0 for kk < row, so a warp whose rows start at r0 skips the
// 8-column stages that end at or before r0
__device__ __forceinline__ bool stage_is_zero(int q, const TMap &tm) { return q < TS / GKC && GKC * (q + 1) <= tm.r0; }

}  // namespace

size_t lk_winv_smem_bytes() { return sizeof(WinvSmem); }
size_t lk_grad_smem_bytes() { return sizeof(GradSmem); }

// ---- W_jj = L_jj^-1, stored as W (winv) and as W' (diagonal tile of the transposed-M array) ------------------------------
__global__ void __launch_bounds__(NTHREADS, 4) lk_winv_kernel(const __grid_constant__ LkGradParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WinvSmem &sm = *reinterpret_cast<WinvSmem *>(smem_raw);
    const int tid = threadIdx.x, nt = prm.nt;
    const int b = blockIdx.x / nt, j = blockIdx.x - b * nt;
    const TMap tm = thread_map(tid);
    const long long ntri = tri_index(nt, 0);
    tile_load_async(sm.S, prm.tiles + ((size_t)b * ntri + tri_index(j, j)) * TILE_ELEMS, tid);
    block_load_async<DSIZE * 8>(sm.D, prm.dblk + ((size_t)b * nt + j) * DSIZE, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    double e[2][NCC];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) e[mb][cc] = (row_of(tm, mb) == col_of(tm, cc)) ? 1.0 : 0.0;
    tile_trsm_ld(e, sm.S, sm.D, tm);  // e = I L^-T = W'
    acc_to_tile_t(prm.winv + ((size_t)b * nt + j) * TILE_ELEMS, e, tm);
    acc_to_tile(prm.minv + ((size_t)b * ntri + col_index(nt, j, j)) * TILE_ELEMS, e, tm);
}

// ---- row i of M = L^-1 (tiles j < i) -------------------------------------------------------------------------------------
template <bool SKIP>
__device__ __forceinline__ void lk_minv_body(const LkGradParams &prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *S = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x, nt = prm.nt, i = prm.i;
    const int b = blockIdx.x / i, j = blockIdx.x - b * i;  // the tiles of one item next to each other: they share row i of L in L2
    const TMap tm = thread_map(tid);
    const long long ntri = tri_index(nt, 0);
    const double *L = prm.tiles + (size_t)b * ntri * TILE_ELEMS;
    double *M = prm.minv + (size_t)b * ntri * TILE_ELEMS;
    const double *srcA = M + col_index(nt, j, j) * TILE_ELEMS;  // (M_kj)', k = j .. i-1
    const double *srcB = L + tri_index(i, j) * TILE_ELEMS;      // L_ik,    k = j .. i-1
    const double *srcW = prm.winv + ((size_t)b * nt + i) * TILE_ELEMS;
    // Zero-tile skipping (block-diagonal covariances): the run k = j .. i-1 shrinks to the k where neither (M_kj)' nor L_ik
    // is exactly zero; a tile of M that comes out exactly zero is flagged for the phases that read M.
    __shared__ short kl[SKIP ? GPL_LK_KLMAX : 1];
    __shared__ int nk;
    const int *zf = SKIP ? prm.zflag + (size_t)b * ntri : nullptr;
    int *mf = SKIP ? prm.mflag + (size_t)b * ntri : nullptr;
    int Q = (TS / GKC) * (i - j);
    bool tri_first = true;  // the first stages belong to the triangular (M_jj)'
    if (SKIP) {
        build_tile_list(kl, &nk, j, i, [&](int k) { return !(zf[tri_index(i, k)] || (k > j && mf[tri_index(k, j)])); }, tid);
        __syncthreads();
        Q = (TS / GKC) * nk;
        tri_first = nk > 0 && kl[0] == j;
    }
    auto issue = [&](int s) {  // always commits: the group count tracks the stage number
        double *dst = S + (s % GNS) * GSLOT;
        if (s < Q) {
            const size_t so = SKIP ? (size_t)((kl[s / (TS / GKC)] - j) * (TS / GKC) + s % (TS / GKC)) : (size_t)s;
            block_load_async<GCH * 8>(dst, srcA + so * GCH, tid);
            block_load_async<GCH * 8>(dst + GCH, srcB + so * GCH, tid);
        } else if (s < Q + GNS) {  // W_ii in 16-column chunks: Q is a multiple of GNS, so the ring becomes the whole tile
            block_load_async<GSLOT * 8>(dst, srcW + (size_t)(s - Q) * GSLOT, tid);
        }
        cp_async_commit();
    };
    double acc[2][NCC];
    acc_zero(acc);
#pragma unroll
    for (int s0 = 0; s0 < GNS - 1; ++s0) issue(s0);
    for (int q = 0; q < Q; ++q) {
        cp_async_wait<GNS - 2>();
        __syncthreads();
        issue(q + GNS - 1);
        const double *a = S + (q % GNS) * GSLOT;
        if (!(tri_first && stage_is_zero(q, tm))) tile_mma<true>(acc, a, a + GCH, tm, 0, GKC);
    }
    __syncthreads();  // the slot of the last update stage is free
    issue(Q + GNS - 1);
    cp_async_wait<0>();
    bool zero_tile = false;
    if (SKIP) {
        int nz = 0;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int cc = 0; cc < NCC; ++cc) nz |= (acc[mb][cc] != 0.0);
        zero_tile = !__syncthreads_or(nz);
        if (tid == 0) mf[tri_index(i, j)] = zero_tile ? 1 : 0;
    } else {
        __syncthreads();
    }
    if (!zero_tile) tile_trsm_w(acc, S, tm);  // (M_ij)' = acc W_ii'
    acc_to_tile(M + col_index(nt, i, j) * TILE_ELEMS, acc, tm);
}
__global__ void __launch_bounds__(NTHREADS, 4) lk_minv_kernel(const __grid_constant__ LkGradParams prm) { lk_minv_body<false>(prm); }
__global__ void __launch_bounds__(NTHREADS, 4) lk_minv_skip_kernel(const __grid_constant__ LkGradParams prm) { lk_minv_body<true>(prm); }

// ---- alpha = M' z --------------------------------------------------------------------------------------------------------
template <bool SKIP>
__device__ __forceinline__ void lk_alpha_body(const LkGradParams &prm) {
    __shared__ double part[TS];
    const int tid = threadIdx.x, nt = prm.nt;
    const int b = blockIdx.x / nt, j = blockIdx.x - b * nt;
    const long long ntri = tri_index(nt, 0);
    const double *run = prm.minv + ((size_t)b * ntri + col_index(nt, j, j)) * TILE_ELEMS;
    const double *z = prm.z + (size_t)b * nt * TS;
    const int row = tid & (TS - 1), half = tid >> 6;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const int *mf = SKIP ? prm.mflag + (size_t)b * ntri : nullptr;
    for (int k = j; k < nt; ++k) {
        if (SKIP && k > j && mf[tri_index(k, j)]) continue;  // an exactly-zero (M_kj)' adds nothing
        const double *T = run + (size_t)(k - j) * TILE_ELEMS;
        const double *zk = z + k * TS + 32 * half;
#pragma unroll 4
        for (int c = 0; c < 32; c += 4) {
            const int kk = 32 * half + c;
            s0 = fma(T[tidx(row, kk)], zk[c], s0);
            s1 = fma(T[tidx(row, kk + 1)], zk[c + 1], s1);
            s2 = fma(T[tidx(row, kk + 2)], zk[c + 2], s2);
            s3 = fma(T[tidx(row, kk + 3)], zk[c + 3], s3);
        }
    }
    const double s = (s0 + s1) + (s2 + s3);
    if (half) part[row] = s;
    __syncthreads();
    if (!half) {
        const double a = s + part[row];
        prm.alpha[(size_t)b * nt * TS + j * TS + row] = a;
        const int g = j * TS + row;
        if (prm.dy && g < prm.n) prm.dy[(size_t)b * prm.n + g] = prm.info[b] ? NAN : -a;
    }
}
__global__ void __launch_bounds__(NTHREADS) lk_alpha_kernel(const __grid_constant__ LkGradParams prm) { lk_alpha_body<false>(prm); }
__global__ void __launch_bounds__(NTHREADS) lk_alpha_skip_kernel(const __grid_constant__ LkGradParams prm) { lk_alpha_body<true>(prm); }

// ---- tiles of K^-1 contracted with dK/dtheta --------------------------------------------------------------------------------
#ifndef GPL_GRADC_CTAS
#define GPL_GRADC_CTAS 4  // 126 registers, no spills: four CTAs per SM like the other streaming kernels (150 registers at three)
#endif
template <bool SKIP>
__device__ __forceinline__ void lk_gradc_body(const LkGradParams &prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GradSmem &sm = *reinterpret_cast<GradSmem *>(smem_raw);
    const DevProgram &P = prm.prog;
    const int tid = threadIdx.x, warp = tid >> 5, nt = prm.nt, n = prm.n;
    const TMap tm = thread_map(tid);
    const long long ntri = tri_index(nt, 0);
    const int b = (int)(blockIdx.x / ntri);
    const long long t = blockIdx.x - (long long)b * ntri;
    int i, j;
    tri_unrank(t, i, j);
    const double *M = prm.minv + (size_t)b * ntri * TILE_ELEMS;
    const double *X = prm.X + (size_t)b * prm.x_stride;
    prepare_item_scalars(P, prm.Theta + (size_t)b * prm.p, &sm.sc, tid);
    if (tid < TS) sm.al[tid] = prm.alpha[(size_t)b * nt * TS + i * TS + tid];
    else sm.al[tid] = prm.alpha[(size_t)b * nt * TS + j * TS + tid - TS];
    if (tid < NWARPS * GPL_MAX_THETA) sm.gsum[tid] = 0.0;
    const bool same = (i == j);
    if (SKIP && !same) {  // programs with structural zeros only (see LkParams::zflag)
        __syncthreads();       // item scalars
        if (tile_cat_dead(P, sm.sc, X, n, prm.d, i, j, sm.S, tid)) {  // dK/dtheta vanishes on the tile: nothing to contract
            if (tid < prm.p) prm.gpart[((size_t)b * ntri + t) * prm.p + tid] = 0.0;
            return;
        }
    }
    const double *srcA = M + col_index(nt, i, i) * TILE_ELEMS;  // (M_ki)', k = i .. nt-1
    const double *srcB = M + col_index(nt, i, j) * TILE_ELEMS;  // (M_kj)', k = i .. nt-1
    // zero-tile skipping: the run k = i .. nt-1 shrinks to the k where neither (M_ki)' nor (M_kj)' is exactly zero
    const int *mf = SKIP ? prm.mflag + (size_t)b * ntri : nullptr;
    int Q = (TS / GKC) * (nt - i);
    bool tri_first = true, last_is_ragged = true;
    if (SKIP) {
        build_tile_list(sm.kl, &sm.nk, i, nt, [&](int k) { return !((k > i && mf[tri_index(k, i)]) || (k > j && mf[tri_index(k, j)])); }, tid);
        __syncthreads();
        Q = (TS / GKC) * sm.nk;
        tri_first = sm.nk > 0 && sm.kl[0] == i;
        last_is_ragged = sm.nk > 0 && sm.kl[sm.nk - 1] == nt - 1;
    }
    auto issue = [&](int s) {
        double *dst = sm.S + (s % GNS) * GSLOT;
        if (s < Q) {
            const size_t so = SKIP ? (size_t)((sm.kl[s / (TS / GKC)] - i) * (TS / GKC) + s % (TS / GKC)) : (size_t)s;
            block_load_async<GCH * 8>(dst, srcA + so * GCH, tid);
            if (!same) block_load_async<GCH * 8>(dst + GCH, srcB + so * GCH, tid);
        }
        cp_async_commit();
    };
    double acc[2][NCC];
    acc_zero(acc);
#pragma unroll
    for (int s0 = 0; s0 < GNS - 1; ++s0) issue(s0);
    // Ragged n: the last tile row of M ends in identity padding; in every (M_ki)' with k = nt - 1 the columns kk >= vlast
    // are zero (or the padding identity, which only reaches entries outside the matrix): those stages are skipped.
    const int vlast = n - (nt - 1) * TS;                                  // valid rows of the last tile (1 .. 64)
    const int q_dead = last_is_ragged ? Q - (TS / GKC) + (vlast + GKC - 1) / GKC : Q;  // first stage of the last tile that is all padding
    for (int q = 0; q < Q; ++q) {
        cp_async_wait<GNS - 2>();
        __syncthreads();
        issue(q + GNS - 1);
        const double *a = sm.S + (q % GNS) * GSLOT;
        if ((tri_first && stage_is_zero(q, tm)) || q >= q_dead) continue;
        // a diagonal tile of K^-1 is symmetric: warp w (rows 16w ..) only needs the columns up to its own diagonal block
        if (same) tile_mma<false, 0xFF, true>(acc, a, a, tm, 0, GKC, 2 * warp + 2);
        else tile_mma<false>(acc, a, a + GCH, tm, 0, GKC);
    }
    cp_async_wait<0>();
    __syncthreads();  // the ring is consumed: S parks the weights
    const SepCtx *sep = nullptr;
    if (prm.sep_col >= 0 && !same) {  // sorted inputs, block below the diagonal: OU leaves on that column in separable form
        int ns = 0, lf[2] = {0, 0};
        for (int f = 0; f < P.n_factors; ++f)
            if (P.f[f].kind == F_OU && P.f[f].col == prm.sep_col && ns < 2) lf[ns++] = f;
        const double *xc = X + (size_t)prm.sep_col * n;
        const int loc = tid & (TS - 1);
        const double c0 = xc[i * TS < n ? i * TS : n - 1];
        const int idx = (tid < TS ? i : j) * TS + loc;
        const double x = xc[idx < n ? idx : n - 1];
        for (int q = 0; q < ns; ++q) {
            const double a = sm.sc.a[lf[q]];
            if (tid < TS) sm.sep.u[q][loc] = fast_exp(a * (x - c0), sm.sc.etab);
            else sm.sep.v[q][loc] = fast_exp(a * (c0 - x), sm.sc.etab);
        }
        if (tid == 0) {
            sm.sep.n_sep = ns;
            sm.sep.leaf[0] = lf[0];
            sm.sep.leaf[1] = lf[1];
        }
        __syncthreads();
        sep = &sm.sep;
    }
    int gi[2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) gi[mb] = i * TS + row_of(tm, mb);
    // weights of sum_{all i, j} W_ij dK_ij from the lower triangle alone: 2 below the diagonal, 1 on it, 0 above it
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        const int r = row_of(tm, mb);
        const double ai = sm.al[r];
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) {
            const int cl = col_of(tm, cc);
            const double sym = !same ? 2.0 : (cl < r ? 2.0 : (cl == r ? 1.0 : 0.0));
            acc[mb][cc] = sym * fma(-ai, sm.al[TS + cl], acc[mb][cc]);
        }
    }
    // on a diagonal tile the 16-column quarters to the right of the warp's own diagonal block have zero weights
    contract_grad_block(P, sm.sc, X, n, n, gi, j * TS, tm.t, acc, sm.S, tid, sm.gsum + warp * GPL_MAX_THETA, same ? warp + 1 : 4, sep);
    __syncthreads();
    if (tid < prm.p) {
        double g = 0.0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) g += sm.gsum[w * GPL_MAX_THETA + tid];
        prm.gpart[((size_t)b * ntri + t) * prm.p + tid] = g;
    }
}
__global__ void __launch_bounds__(NTHREADS, GPL_GRADC_CTAS) lk_gradc_kernel(const __grid_constant__ LkGradParams prm) { lk_gradc_body<false>(prm); }
__global__ void __launch_bounds__(NTHREADS, GPL_GRADC_CTAS) lk_gradc_skip_kernel(const __grid_constant__ LkGradParams prm) { lk_gradc_body<true>(prm); }

__global__ void __launch_bounds__(128) lk_gradsum_kernel(const __grid_constant__ LkGradParams prm) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= prm.B) return;
    const long long ntri = tri_index(prm.nt, 0);
    const bool bad = prm.info[b] != 0;
    for (int s = 0; s < prm.p; ++s) {
        double g = 0.0;
        for (long long t = 0; t < ntri; ++t) g += prm.gpart[((size_t)b * ntri + t) * prm.p + s];
        prm.dtheta[(size_t)b * prm.p + s] = bad ? NAN : -0.5 * g;
    }
}

}  // namespace gpl
