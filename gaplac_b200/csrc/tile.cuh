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

// ---- Cholesky of the diagonal tile, blocked (v1) -------------------------------------------------------------
// Same contract as tile_potrf_inv, but the 64 x 64 tile is processed in four 16-column panels:
//   1. the owners of the panel's columns (acc) and of the panel's rows of the running inverse (w) publish them;
//   2. warp 0 factors the 16 x 16 diagonal block in registers (one row per lane, pivots and columns exchanged
//      with warp shuffles: no block barrier inside) and inverts it;
//   3. all threads form the panel of L below the block (P * W16') and the new rows of the inverse (W16 * R);
//   4. all threads apply the rank-16 update to their register blocks of the trailing tile and of the inverse.
// Three block barriers per panel (12 per tile instead of 64) and the O(64^3) part runs as register-tiled FMAs.
// scratch: 8192 doubles (two ping-pong sets of P, R, Lp, Rp); L16s / W16s: 256 doubles each; rsbuf: 16; pivbuf: 64.
template <int H>
__device__ __forceinline__ void gemm16_half(double (&c)[4][4], const double *__restrict__ A,
                                            const double *__restrict__ B, TMap tm) {
    const double *pa = A + tm.m0;
    const double *pb = B + tm.cb + 16 * H;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const double2 a01 = *reinterpret_cast<const double2 *>(pa + k * TS);
        const double2 a23 = *reinterpret_cast<const double2 *>(pa + k * TS + 2);
        const double2 b = *reinterpret_cast<const double2 *>(pb + k * TS);
        c[0][2 * H] = fma(-a01.x, b.x, c[0][2 * H]);
        c[0][2 * H + 1] = fma(-a01.x, b.y, c[0][2 * H + 1]);
        c[1][2 * H] = fma(-a01.y, b.x, c[1][2 * H]);
        c[1][2 * H + 1] = fma(-a01.y, b.y, c[1][2 * H + 1]);
        c[2][2 * H] = fma(-a23.x, b.x, c[2][2 * H]);
        c[2][2 * H + 1] = fma(-a23.x, b.y, c[2][2 * H + 1]);
        c[3][2 * H] = fma(-a23.y, b.x, c[3][2 * H]);
        c[3][2 * H + 1] = fma(-a23.y, b.y, c[3][2 * H + 1]);
    }
}

template <int P_, int H>
__device__ __forceinline__ void potrf_panel_update(double (&acc)[4][4], double (&w)[4][4], TMap tm, int wr, int wc,
                                                   const double *Lp, const double *Rp) {
    const int pc = 2 * wc + H;  // 16-column panel this half of the thread's columns belongs to
    if (pc == P_) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double *src = Lp + (col_of(tm.cb, 2 * H + q) - 16 * P_) * TS + tm.m0;
            const double2 v01 = *reinterpret_cast<const double2 *>(src);
            const double2 v23 = *reinterpret_cast<const double2 *>(src + 2);
            acc[0][2 * H + q] = v01.x;
            acc[1][2 * H + q] = v01.y;
            acc[2][2 * H + q] = v23.x;
            acc[3][2 * H + q] = v23.y;
        }
    } else if (pc > P_ && wr >= pc) {
        gemm16_half<H>(acc, Lp, Lp, tm);
    }
    if (pc <= P_) {
        if (wr > P_) {
            gemm16_half<H>(w, Lp, Rp, tm);
        } else if (wr == P_) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double2 v = *reinterpret_cast<const double2 *>(Rp + (tm.m0 + r - 16 * P_) * TS + tm.cb + 16 * H);
                w[r][2 * H] = v.x;
                w[r][2 * H + 1] = v.y;
            }
        }
    }
}

template <int P_>
__device__ __forceinline__ void potrf_panel(double (&acc)[4][4], double (&w)[4][4], TMap tm, double *scratch,
                                            double *L16s, double *W16s, double *rsbuf, double *pivbuf, int tid, int &fail) {
    const int warp = tid >> 5, lane = tid & 31, wr = warp >> 1, wc = warp & 1;
    double *P = scratch + (P_ & 1) * 4096;  // 64 x 16 column-major: P[k*64 + row]
    double *R = P + 1024;                   // 16 x 64: R[k*64 + col]
    double *Lp = P + 2048;                  // 64 x 16 column-major
    double *Rp = P + 3072;                  // 16 x 64
    // 1. publish the panel's columns of the tile and the panel's rows of the running inverse
    if (wc == (P_ >> 1)) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int cc = (P_ & 1) * 2 + q;
            double *dst = P + (col_of(tm.cb, cc) - 16 * P_) * TS + tm.m0;
            *reinterpret_cast<double2 *>(dst) = make_double2(acc[0][cc], acc[1][cc]);
            *reinterpret_cast<double2 *>(dst + 2) = make_double2(acc[2][cc], acc[3][cc]);
        }
    }
    if (wr == P_) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            double *dst = R + (tm.m0 + r - 16 * P_) * TS + tm.cb;
            *reinterpret_cast<double2 *>(dst) = make_double2(w[r][0], w[r][1]);
            *reinterpret_cast<double2 *>(dst + 16) = make_double2(w[r][2], w[r][3]);
        }
    }
    __syncthreads();
    // 2. warp 0: Cholesky of the 16 x 16 diagonal block (row per lane, shuffles) and its inverse
    if (warp == 0) {
        const int r = lane & 15;
        double a[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) a[c] = P[c * TS + 16 * P_ + r];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            double piv = __shfl_sync(0xffffffffu, a[c], c);
            if (!(piv > 0.0)) {
                if (fail < 0) fail = 16 * P_ + c;
                piv = 1.0;
            }
            if (lane == 0) pivbuf[16 * P_ + c] = piv;
            const double rs = rsqrt(piv);
            if (lane == 0) rsbuf[c] = rs;
            const double l = a[c] * rs;
            a[c] = (r >= c) ? l : 0.0;
#pragma unroll
            for (int c2 = c + 1; c2 < 16; ++c2) {
                const double l2 = __shfl_sync(0xffffffffu, l, c2);
                a[c2] = fma(-l, l2, a[c2]);
            }
        }
        if (lane < 16) {
#pragma unroll
            for (int c = 0; c < 16; ++c) L16s[c * 16 + r] = a[c];
        }
        __syncwarp();
        // column r of the inverse by forward substitution (axpy form): x = L16^-1 e_r
        double x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = (i == r) ? 1.0 : 0.0;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            x[c] *= rsbuf[c];
#pragma unroll
            for (int i = c + 1; i < 16; ++i) x[i] = fma(-L16s[c * 16 + i], x[c], x[i]);
        }
        if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) W16s[i * 16 + r] = x[i];  // W16[i][r], row-major
        }
    }
    __syncthreads();
    // 3a. panel of L: rows below the block = P * W16', rows of the block = L16, rows above = 0
    {
        const int row = tid & 63, cg = tid >> 6;
        double out[4] = {0.0, 0.0, 0.0, 0.0};
        if (row >= 16 * (P_ + 1)) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const double pk = P[k * TS + row];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = 4 * cg + q;
                    if (k <= c) out[q] = fma(pk, W16s[c * 16 + k], out[q]);
                }
            }
        } else if (row >= 16 * P_) {
#pragma unroll
            for (int q = 0; q < 4; ++q) out[q] = L16s[(4 * cg + q) * 16 + (row - 16 * P_)];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) Lp[(4 * cg + q) * TS + row] = out[q];
    }
    // 3b. new rows of the inverse: Rp = W16 * R
    {
        const int col = tid & 63, kg = tid >> 6;
        double out[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
            const double rv = R[k2 * TS + col];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = 4 * kg + q;
                if (k2 <= k) out[q] = fma(W16s[k * 16 + k2], rv, out[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) Rp[(4 * kg + q) * TS + col] = out[q];
    }
    __syncthreads();
    // 4. rank-16 updates of the register blocks
    potrf_panel_update<P_, 0>(acc, w, tm, wr, wc, Lp, Rp);
    potrf_panel_update<P_, 1>(acc, w, tm, wr, wc, Lp, Rp);
}

__device__ __forceinline__ int tile_potrf_inv_blocked(double (&acc)[4][4], double (&w)[4][4], TMap tm,
                                                      double *scratch, double *L16s, double *W16s, double *rsbuf,
                                                      double *pivbuf, int tid) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) w[r][c] = (tm.m0 + r == col_of(tm.cb, c)) ? 1.0 : 0.0;
    int fail = -1;
    potrf_panel<0>(acc, w, tm, scratch, L16s, W16s, rsbuf, pivbuf, tid, fail);
    potrf_panel<1>(acc, w, tm, scratch, L16s, W16s, rsbuf, pivbuf, tid, fail);
    potrf_panel<2>(acc, w, tm, scratch, L16s, W16s, rsbuf, pivbuf, tid, fail);
    potrf_panel<3>(acc, w, tm, scratch, L16s, W16s, rsbuf, pivbuf, tid, fail);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (col_of(tm.cb, c) > tm.m0 + r) acc[r][c] = 0.0;
    return fail;  // meaningful in warp 0 (tid 0 reports it)
}

// register block -> global tile, column-major (each thread: 4 x 32 contiguous bytes)
__device__ __forceinline__ void acc_to_global(double *__restrict__ tile, const double (&acc)[4][4], TMap tm) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        double *p = tile + col_of(tm.cb, cc) * TS + tm.m0;
        *reinterpret_cast<double2 *>(p) = make_double2(acc[0][cc], acc[1][cc]);
        *reinterpret_cast<double2 *>(p + 2) = make_double2(acc[2][cc], acc[3][cc]);
    }
}

// asynchronous copy of one half (32 columns = 16 KiB) of a tile
__device__ __forceinline__ void half_tile_load_async(double *smem, const double *__restrict__ gmem, int tid) {
#pragma unroll
    for (int it = 0; it < TILE_BYTES / 2 / 16 / NTHREADS; ++it) {
        const int idx = it * NTHREADS + tid;
        cp_async16(reinterpret_cast<char *>(smem) + idx * 16, reinterpret_cast<const char *>(gmem) + idx * 16);
    }
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
