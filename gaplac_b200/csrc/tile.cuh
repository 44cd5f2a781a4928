// 64 x 64 FP64 tile primitives shared by the batched factorisation, the posterior/predict kernels and the
// large-n path.  One CTA of 256 threads owns one tile; each thread owns a 4 x 4 register block.
//
// Tile storage (global workspace and shared memory alike): column-major, dense, 64 x 64 doubles = 32 KiB,
// element (r, c) at c*64 + r.  A block-lower-triangular matrix of nt x nt tiles is stored tile-major:
// tile (i, j), i >= j, at offset (i(i+1)/2 + j) * 4096 doubles — so that one tile is one contiguous
// 32 KiB bulk copy.
//
// Thread -> entries map (tid = warp*32 + lane):  warp w: wr = w/2 (16 rows), wc = w%2 (32 cols);
// lane: lr = lane/8 (4 rows), lc = lane%8.  Rows m0 + {0..3}, m0 = wr*16 + lr*4; columns
// cb + {0, 1, 16, 17}, cb = wc*32 + lc*2.  With this interleave a warp's operand reads in the GEMM core
// (below) are one 128-byte wavefront per LDS.128: no shared-memory bank conflicts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpl {

constexpr int TS = 64;                 // tile edge
constexpr int TILE_ELEMS = TS * TS;    // 4096 doubles
constexpr int TILE_BYTES = TILE_ELEMS * 8;
constexpr int NTHREADS = 256;

__host__ __device__ __forceinline__ long long tri_index(int i, int j) { return (long long)i * (i + 1) / 2 + j; }

struct TMap {
    int m0;  // first row of the thread's block
    int cb;  // column base of the thread's block
};
__device__ __forceinline__ TMap thread_map(int tid) {
    const int w = tid >> 5, lane = tid & 31;
    TMap t;
    t.m0 = (w >> 1) * 16 + (lane >> 3) * 4;
    t.cb = (w & 1) * 32 + (lane & 7) * 2;
    return t;
}
__device__ __forceinline__ int col_of(int cb, int cc) { return cb + ((cc >> 1) << 4) + (cc & 1); }

// ---- global <-> shared tile movement ------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// asynchronous (LDGSTS) copy of one tile; caller commits / waits / syncs
__device__ __forceinline__ void tile_load_async(double *smem, const double *__restrict__ gmem, int tid) {
#pragma unroll
    for (int it = 0; it < TILE_BYTES / 16 / NTHREADS; ++it) {
        const int idx = it * NTHREADS + tid;  // 16-byte chunk index
        cp_async16(reinterpret_cast<char *>(smem) + idx * 16, reinterpret_cast<const char *>(gmem) + idx * 16);
    }
}

__device__ __forceinline__ void tile_store(double *__restrict__ gmem, const double *smem, int tid) {
#pragma unroll
    for (int it = 0; it < TILE_BYTES / 16 / NTHREADS; ++it) {
        const int idx = it * NTHREADS + tid;
        reinterpret_cast<double2 *>(gmem)[idx] = reinterpret_cast<const double2 *>(smem)[idx];
    }
}

// ---- register block <-> shared tile ---------------------------------------------------------------------
// column-major: element (row, col) at col*64 + row
__device__ __forceinline__ void acc_to_smem(double *T, const double (&acc)[4][4], TMap tm) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        double *p = T + col_of(tm.cb, cc) * TS + tm.m0;
        *reinterpret_cast<double2 *>(p) = make_double2(acc[0][cc], acc[1][cc]);
        *reinterpret_cast<double2 *>(p + 2) = make_double2(acc[2][cc], acc[3][cc]);
    }
}
// transposed: element (row, col) at row*64 + col
__device__ __forceinline__ void acc_to_smem_t(double *T, const double (&acc)[4][4], TMap tm) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double *p = T + (tm.m0 + r) * TS + tm.cb;
        *reinterpret_cast<double2 *>(p) = make_double2(acc[r][0], acc[r][1]);
        *reinterpret_cast<double2 *>(p + 16) = make_double2(acc[r][2], acc[r][3]);
    }
}

// ---- GEMM core ------------------------------------------------------------------------------------------
// acc[r][cc] (+|-)= sum_{k0 <= k < k1} A[k*64 + m0 + r] * B[k*64 + col_of(cc)]
// i.e. C (+|-)= A_tile * B_tile' for column-major tiles A (rows x k) and B (cols x k).
template <bool SUB>
__device__ __forceinline__ void tile_gemm(double (&acc)[4][4], const double *__restrict__ A,
                                          const double *__restrict__ B, TMap tm, int k0, int k1) {
    const double *pa = A + tm.m0;
    const double *pb = B + tm.cb;
#pragma unroll 8
    for (int k = k0; k < k1; ++k) {
        const double2 a01 = *reinterpret_cast<const double2 *>(pa + k * TS);
        const double2 a23 = *reinterpret_cast<const double2 *>(pa + k * TS + 2);
        const double2 b01 = *reinterpret_cast<const double2 *>(pb + k * TS);
        const double2 b23 = *reinterpret_cast<const double2 *>(pb + k * TS + 16);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y};
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = SUB ? fma(-a[r], b[c], acc[r][c]) : fma(a[r], b[c], acc[r][c]);
    }
}

// ---- Cholesky of the diagonal tile, with its inverse --------------------------------------------------------
// In: acc = lower part of the 64 x 64 SPD tile (upper part ignored).  Out: acc = L (zeros above the diagonal),
// w = L^-1 (lower triangular).  Right-looking, one __syncthreads per pivot: the owners of column c publish
// it, everybody scales by rsqrt(pivot) and applies the rank-1 update to its own registers; the same
// elementary transformation is applied to w (Gauss-Jordan on the identity), so L^-1 costs no extra barrier.
// colbuf/rowbuf: 2 x 64 doubles each (ping-pong), pivbuf: 64 doubles (pivots, for logdet).
// Returns -1, or the local index of the first non-positive pivot (the factor is then meaningless).
__device__ __forceinline__ int tile_potrf_inv(double (&acc)[4][4], double (&w)[4][4], TMap tm, double *colbuf,
                                              double *rowbuf, double *pivbuf, int tid) {
    int gr[4], gc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) gr[r] = tm.m0 + r;
#pragma unroll
    for (int c = 0; c < 4; ++c) gc[c] = col_of(tm.cb, c);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) w[r][c] = (gr[r] == gc[c]) ? 1.0 : 0.0;
    const int warp = tid >> 5;
    const int wrow_hi = (warp >> 1) * 16 + 15;  // last row covered by this warp
    const int wcol_lo = (warp & 1) * 32;        // first / last column covered by this warp
    const int wcol_hi = wcol_lo + 31;
    const bool warp_has_lower = wrow_hi >= wcol_lo;
    int fail = -1;
    for (int c = 0; c < TS; ++c) {
        double *cbuf = colbuf + (c & 1) * TS;
        double *rbuf = rowbuf + (c & 1) * TS;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
            if (gc[cc] == c) {
#pragma unroll
                for (int r = 0; r < 4; ++r) cbuf[gr[r]] = acc[r][cc];
            }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (gr[r] == c) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) rbuf[gc[cc]] = w[r][cc];
            }
        __syncthreads();
        double piv = cbuf[c];
        if (!(piv > 0.0)) {
            if (fail < 0) fail = c;
            piv = 1.0;
        }
        if (tid == 0) pivbuf[c] = piv;
        const double rs = rsqrt(piv);
        const bool do_s = warp_has_lower && wcol_hi >= c;  // column c itself or the trailing columns
        const bool do_w = wrow_hi >= c && wcol_lo <= c;    // rows >= c, columns <= c of the inverse
        if (do_s || do_w) {
            double lcol[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) lcol[r] = cbuf[gr[r]] * rs;
            if (do_s) {
                double lrow[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) lrow[cc] = cbuf[gc[cc]] * rs;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        if (gc[cc] == c) acc[r][cc] = (gr[r] >= c) ? lcol[r] : 0.0;
                        else if (gc[cc] > c) acc[r][cc] = fma(-lcol[r], lrow[cc], acc[r][cc]);
                    }
            }
            if (do_w) {
                double wrow[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) wrow[cc] = rbuf[gc[cc]] * rs;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        if (gr[r] == c) w[r][cc] = wrow[cc];
                        else if (gr[r] > c) w[r][cc] = fma(-lcol[r], wrow[cc], w[r][cc]);
                    }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
            if (gc[cc] > gr[r]) acc[r][cc] = 0.0;
    return fail;
}

// deterministic block-wide sum (256 threads), result valid in every thread; red = 8 doubles of shared memory
__device__ __forceinline__ double block_sum(double v, double *red, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NTHREADS / 32; ++i) s += red[i];
    return s;
}

}  // namespace gpl
