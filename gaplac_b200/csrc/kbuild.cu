// Fused covariance construction: reads the input columns and the hyperparameters, writes K in FP64.
//
// Replaces kernelmatrix(k, RowVecs(X)) (+ sigma2 I) [upstream KernelFunctions / AbstractGPs cov(::FiniteGP)]
// reached from CLI/src/mcmc.jl:35, CLI/src/select.jl:43,47, CLI/src/sample.jl:25, src/plotting.jl:6, and the
// cross-covariance K(X, X*) of mean_and_var (src/plotting.jl:12).  The reference allocates one n x n
// temporary per formula node; here every entry is computed once in registers and stored once (8 B/entry).
#include "kernels.h"
#include "kfun.cuh"
#include "tile.cuh"

namespace gpl {

// Both kernels: one CTA of 128 threads per 64 x 64 block of K.
//   * the x-coordinates of the block's 64 rows and 64 columns are staged once in shared memory (every entry needs one of
//     each per leaf; the interpreter then reads them through broadcast / conflict-free shared loads instead of global ones);
//   * thread (warp w, lane l) owns rows 2l, 2l+1 and columns 16w .. 16w+15, evaluated as four 2 x 4 quarters through one
//     rolled copy of the interpreter (kfun.cuh);
//   * stores are 128-bit: the two rows of a thread are adjacent in memory (column-major K, and the swizzle of the tile
//     layout flips only bits 2-3 of the row), so a warp writes one whole 512-byte column of the block per instruction;
//   * blocks off the diagonal use the cross-covariance form (Noise terms and the diagonal bookkeeping drop out).
namespace {
struct __align__(16) CovSmem {
    ItemScalars sc;
    double xa[GPL_MAX_COLS][TS];
    double xb[GPL_MAX_COLS][TS];
};

__device__ __forceinline__ void stage_block_x(const DevProgram &P, CovSmem &sm, const double *__restrict__ Xa, int na, int row0,
                                              const double *__restrict__ Xb, int nb, int col0, int tid) {
    const int r = tid & (TS - 1);
    if (tid < TS) {
        const int gr = row0 + r < na ? row0 + r : na - 1;
        for (int c = 0; c < P.n_cols; ++c) sm.xa[c][r] = Xa[(size_t)c * na + gr];
    } else {
        const int gc = col0 + r < nb ? col0 + r : nb - 1;
        for (int c = 0; c < P.n_cols; ++c) sm.xb[c][r] = Xb[(size_t)c * nb + gc];
    }
}

// quarter h of this thread's 2 x 16 block: o[r][c] = K(row0 + 2 lane + r, col0 + 16 warp + 4 h + c)
template <bool SAME>
__device__ __forceinline__ void cov_quarter(const DevProgram &P, const CovSmem &sm, int na, int nb, int row0, int col0, int lane,
                                            int warp, int h, double diag_add, double (&o)[2][4]) {
    int gi[2], gj[4];
    gi[0] = row0 + 2 * lane;
    gi[1] = gi[0] + 1;
#pragma unroll
    for (int c = 0; c < 4; ++c) gj[c] = col0 + 16 * warp + 4 * h + c;
    // the staged columns stand in for X: element (row, col) of the block-local copy is sm.x?[col][row - row0]
    eval_block<2, 4, SAME>(P, sm.sc, &sm.xa[0][0] - row0, TS, na, gi, &sm.xb[0][0] - col0, TS, nb, gj, diag_add, o);
}
}  // namespace

// dense column-major output; grid = (ceil(na/64), ceil(nb/64))
__global__ void __launch_bounds__(NTHREADS) cov_dense_kernel(const __grid_constant__ CovParams prm) {
    __shared__ CovSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.x * TS, col0 = blockIdx.y * TS;
    prepare_item_scalars(prm.prog, prm.theta, &sm.sc, tid);
    stage_block_x(prm.prog, sm, prm.Xa, prm.na, row0, prm.Xb, prm.nb, col0, tid);
    __syncthreads();
    const bool on_diag = prm.same && blockIdx.x == blockIdx.y;
    const int g0 = row0 + 2 * lane;
    const bool vec_ok = (prm.na & 1) == 0 && g0 + 1 < prm.na;  // 16-byte aligned pair inside the matrix
#pragma unroll 1
    for (int h = 0; h < 4; ++h) {
        double o[2][4];
        if (on_diag) cov_quarter<true>(prm.prog, sm, prm.na, prm.nb, row0, col0, lane, warp, h, prm.diag_add, o);
        else cov_quarter<false>(prm.prog, sm, prm.na, prm.nb, row0, col0, lane, warp, h, 0.0, o);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int gc = col0 + 16 * warp + 4 * h + c;
            if (gc >= prm.nb) continue;
            double *dst = prm.K + (size_t)gc * prm.na + g0;
            if (vec_ok) {
                *reinterpret_cast<double2 *>(dst) = make_double2(o[0][c], o[1][c]);
            } else {
                if (g0 < prm.na) dst[0] = o[0][c];
                if (g0 + 1 < prm.na) dst[1] = o[1][c];
            }
        }
    }
}

// tile-major lower layout (identity padding beyond n); grid = nt(nt+1)/2 CTAs, one per tile
__global__ void __launch_bounds__(NTHREADS) cov_tiles_kernel(const __grid_constant__ CovTilesParams prm) {
    __shared__ CovSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long t = blockIdx.x;
    int i, j;
    tri_unrank(t, i, j);
    const int row0 = i * TS, col0 = j * TS;
    prepare_item_scalars(prm.prog, prm.theta, &sm.sc, tid);
    stage_block_x(prm.prog, sm, prm.X, prm.n, row0, prm.X, prm.n, col0, tid);
    __syncthreads();
    double *tile = prm.tiles + t * TILE_ELEMS;
#pragma unroll 1
    for (int h = 0; h < 4; ++h) {
        double o[2][4];
        if (i == j) cov_quarter<true>(prm.prog, sm, prm.n, prm.n, row0, col0, lane, warp, h, prm.diag_add, o);
        else cov_quarter<false>(prm.prog, sm, prm.n, prm.n, row0, col0, lane, warp, h, 0.0, o);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int lc = 16 * warp + 4 * h + c;  // (lc & 3) == c: rows 2l, 2l+1 stay an aligned pair under the swizzle
            *reinterpret_cast<double2 *>(tile + lc * TS + ((2 * lane) ^ (c << 2))) = make_double2(o[0][c], o[1][c]);
        }
    }
}

}  // namespace gpl
