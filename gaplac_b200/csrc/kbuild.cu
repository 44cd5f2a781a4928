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

// dense column-major output; grid = (ceil(na/64), ceil(nb/64)); each thread stores 4 consecutive rows
// (32 B) of 4 columns, a warp covers 128 B contiguous per column.
__global__ void __launch_bounds__(NTHREADS) cov_dense_kernel(const __grid_constant__ CovParams prm) {
    __shared__ ItemScalars sc;
    const int tid = threadIdx.x;
    prepare_item_scalars(prm.prog, prm.theta, &sc, tid);
    __syncthreads();
    const TMap tm = thread_map(tid);
    int gi[4], gj[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) gi[r] = blockIdx.x * TS + tm.m0 + r;
#pragma unroll
    for (int c = 0; c < 4; ++c) gj[c] = blockIdx.y * TS + col_of(tm.cb, c);
    double acc[4][4];
    if (prm.same)
        eval_block<4, 4, true>(prm.prog, sc, prm.Xa, prm.na, prm.na, gi, prm.Xb, prm.nb, prm.nb, gj, prm.diag_add, acc);
    else
        eval_block<4, 4, false>(prm.prog, sc, prm.Xa, prm.na, prm.na, gi, prm.Xb, prm.nb, prm.nb, gj, 0.0, acc);
    const bool vec_ok = (prm.na % 2 == 0) && (gi[3] < prm.na);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (gj[c] >= prm.nb) continue;
        double *col = prm.K + (size_t)gj[c] * prm.na;
        if (vec_ok) {
            *reinterpret_cast<double2 *>(col + gi[0]) = make_double2(acc[0][c], acc[1][c]);
            *reinterpret_cast<double2 *>(col + gi[2]) = make_double2(acc[2][c], acc[3][c]);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (gi[r] < prm.na) col[gi[r]] = acc[r][c];
        }
    }
}

// tile-major lower layout (identity padding beyond n); grid = nt(nt+1)/2 CTAs, one per tile
__global__ void __launch_bounds__(NTHREADS) cov_tiles_kernel(const __grid_constant__ CovTilesParams prm) {
    __shared__ ItemScalars sc;
    const int tid = threadIdx.x;
    prepare_item_scalars(prm.prog, prm.theta, &sc, tid);
    __syncthreads();
    // linear tile index -> (i, j), i >= j
    const long long t = blockIdx.x;
    int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (tri_index(i + 1, 0) <= t) ++i;
    while (tri_index(i, 0) > t) --i;
    const int j = (int)(t - tri_index(i, 0));
    const TMap tm = thread_map(tid);
    int gi[4], gj[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) gi[r] = i * TS + tm.m0 + r;
#pragma unroll
    for (int c = 0; c < 4; ++c) gj[c] = j * TS + col_of(tm.cb, c);
    double acc[4][4];
    eval_block<4, 4, true>(prm.prog, sc, prm.X, prm.n, prm.n, gi, prm.X, prm.n, prm.n, gj, prm.diag_add, acc);
    double *tile = prm.tiles + t * TILE_ELEMS;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double *p = tile + col_of(tm.cb, c) * TS + tm.m0;
        *reinterpret_cast<double2 *>(p) = make_double2(acc[0][c], acc[1][c]);
        *reinterpret_cast<double2 *>(p + 2) = make_double2(acc[2][c], acc[3][c]);
    }
}

}  // namespace gpl
