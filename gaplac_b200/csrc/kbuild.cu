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

// dense column-major output; grid = (ceil(na/64), ceil(nb/64)).  A warp stores, per (row block, column), 8
// consecutive rows (64 B) of 4 columns.
__global__ void __launch_bounds__(NTHREADS) cov_dense_kernel(const __grid_constant__ CovParams prm) {
    __shared__ ItemScalars sc;
    const int tid = threadIdx.x;
    prepare_item_scalars(prm.prog, prm.theta, &sc, tid);
    __syncthreads();
    const TMap tm = thread_map(tid);
    int gi[2], gj[NCC];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) gi[mb] = blockIdx.x * TS + row_of(tm, mb);
#pragma unroll
    for (int cc = 0; cc < NCC; ++cc) gj[cc] = blockIdx.y * TS + col_of(tm, cc);
    double acc[2][NCC];
    if (prm.same)
        eval_block_acc<true>(prm.prog, sc, prm.Xa, prm.na, prm.na, gi, prm.Xb, prm.nb, prm.nb, blockIdx.y * TS, tm.t,
                             prm.diag_add, acc);
    else
        eval_block_acc<false>(prm.prog, sc, prm.Xa, prm.na, prm.na, gi, prm.Xb, prm.nb, prm.nb, blockIdx.y * TS, tm.t, 0.0,
                              acc);
#pragma unroll
    for (int cc = 0; cc < NCC; ++cc) {
        if (gj[cc] >= prm.nb) continue;
        double *col = prm.K + (size_t)gj[cc] * prm.na;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
            if (gi[mb] < prm.na) col[gi[mb]] = acc[mb][cc];
    }
}

// tile-major lower layout (identity padding beyond n); grid = nt(nt+1)/2 CTAs, one per tile
__global__ void __launch_bounds__(NTHREADS) cov_tiles_kernel(const __grid_constant__ CovTilesParams prm) {
    __shared__ ItemScalars sc;
    const int tid = threadIdx.x;
    prepare_item_scalars(prm.prog, prm.theta, &sc, tid);
    __syncthreads();
    const long long t = blockIdx.x;
    int i, j;
    tri_unrank(t, i, j);
    const TMap tm = thread_map(tid);
    int gi[2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) gi[mb] = i * TS + row_of(tm, mb);
    double acc[2][NCC];
    eval_block_acc<true>(prm.prog, sc, prm.X, prm.n, prm.n, gi, prm.X, prm.n, prm.n, j * TS, tm.t, prm.diag_add, acc);
    acc_to_tile(prm.tiles + t * TILE_ELEMS, acc, tm);
}

}  // namespace gpl
