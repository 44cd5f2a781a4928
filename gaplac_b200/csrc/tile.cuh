// 64 x 64 FP64 tile primitives shared by the batched factorisation, the posterior/predict kernels and the
// large-n path.  One CTA of 256 threads (8 warps) owns one tile; the contraction runs on the FP64 tensor-core
// path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4): on sm_100a DMMA has the same peak as vector DFMA (37 TFLOP/s
// measured, tools/fp64_peak.cu) but its fragments are loaded from shared memory without the redundant
// broadcast reads a register-tiled DFMA loop needs, which is what bounded the first versions of these kernels
// (profiles/ncu_lml_r01_v1_*: shared-memory wavefronts at 62 % of peak, FP64 pipe at 41 %).
//
// Tile storage (global workspace and shared memory alike): 64 x 64 doubles = 32 KiB, dense, column-major with
// an XOR swizzle of the row index:  element (r, c) lives at  c*64 + (r ^ ((c & 3) << 2)).
// A block-lower-triangular matrix of nt x nt tiles is stored tile-major: tile (i, j), i >= j, at offset
// (i(i+1)/2 + j) * 4096 doubles, so one tile (or any run of whole columns of it) is one contiguous copy.
// The swizzle makes every fragment load (8 rows x 4 consecutive columns, one double per lane) hit 16 distinct
// 8-byte banks per half-warp: conflict-free LDS.64, 2 wavefronts per 256 bytes.
//
// GEMM form used everywhere:  C (+|-)= A * B'  with A = (rows x k) and B = (cols x k), both in tile format.
// Thread -> accumulator map (warp w: wr = w/2 -> 16 rows, wc = w%2 -> 32 columns; lane: g = lane/4, t = lane%4):
//   rows  row_of(mb)  = wr*16 + 8*mb + g,               mb = 0, 1
//   cols  col_of(cc)  = wc*32 + 8*(cc/2) + 2*t + cc%2,  cc = 0..7       -> acc[2][8]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpl {

constexpr int TS = 64;                 // tile edge
constexpr int TILE_ELEMS = TS * TS;    // 4096 doubles
constexpr int TILE_BYTES = TILE_ELEMS * 8;
constexpr int NTHREADS = 256;

__host__ __device__ __forceinline__ long long tri_index(int i, int j) { return (long long)i * (i + 1) / 2 + j; }

// position of element (r, c) inside a tile
__host__ __device__ __forceinline__ int tidx(int r, int c) { return c * TS + (r ^ ((c & 3) << 2)); }

struct TMap {
    int r0;  // first row of the warp's 16-row band
    int c0;  // first column of the warp's 32-column band
    int g;   // lane / 4
    int t;   // lane % 4
};
__device__ __forceinline__ TMap thread_map(int tid) {
    const int w = tid >> 5, lane = tid & 31;
    TMap m;
    m.r0 = (w >> 1) * 16;
    m.c0 = (w & 1) * 32;
    m.g = lane >> 2;
    m.t = lane & 3;
    return m;
}
__device__ __forceinline__ int row_of(const TMap &tm, int mb) { return tm.r0 + 8 * mb + tm.g; }
__device__ __forceinline__ int col_of(const TMap &tm, int cc) { return tm.c0 + 8 * (cc >> 1) + 2 * tm.t + (cc & 1); }

// ---- global <-> shared movement -------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// asynchronous (LDGSTS) copy of BYTES contiguous bytes by the whole CTA; caller commits / waits / syncs
template <int BYTES>
__device__ __forceinline__ void block_load_async(double *smem, const double *__restrict__ gmem, int tid) {
    static_assert(BYTES % (16 * NTHREADS) == 0, "whole 16-byte chunks per thread");
#pragma unroll
    for (int it = 0; it < BYTES / 16 / NTHREADS; ++it) {
        const int idx = it * NTHREADS + tid;
        cp_async16(reinterpret_cast<char *>(smem) + idx * 16, reinterpret_cast<const char *>(gmem) + idx * 16);
    }
}
__device__ __forceinline__ void tile_load_async(double *smem, const double *__restrict__ gmem, int tid) {
    block_load_async<TILE_BYTES>(smem, gmem, tid);
}
__device__ __forceinline__ void half_tile_load_async(double *smem, const double *__restrict__ gmem, int tid) {
    block_load_async<TILE_BYTES / 2>(smem, gmem, tid);
}
// 16 whole columns of a tile (8 KiB): the pipeline stage of the factorisation kernels
constexpr int KC = 16;
constexpr int CHUNK_ELEMS = KC * TS;
__device__ __forceinline__ void chunk_load_async(double *smem, const double *__restrict__ gmem, int tid) {
    block_load_async<CHUNK_ELEMS * 8>(smem, gmem, tid);
}

__device__ __forceinline__ void tile_store(double *__restrict__ gmem, const double *smem, int tid) {
#pragma unroll
    for (int it = 0; it < TILE_BYTES / 16 / NTHREADS; ++it) {
        const int idx = it * NTHREADS + tid;
        reinterpret_cast<double2 *>(gmem)[idx] = reinterpret_cast<const double2 *>(smem)[idx];
    }
}

// ---- register block <-> tile (shared or global) ------------------------------------------------------------
__device__ __forceinline__ void acc_zero(double (&acc)[2][8]) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) acc[mb][cc] = 0.0;
}
// tile element (row, col) <- acc
__device__ __forceinline__ void acc_to_tile(double *T, const double (&acc)[2][8], const TMap &tm) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) T[tidx(row_of(tm, mb), col_of(tm, cc))] = acc[mb][cc];
}
// transposed: tile element (col, row) <- acc
__device__ __forceinline__ void acc_to_tile_t(double *T, const double (&acc)[2][8], const TMap &tm) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) T[tidx(col_of(tm, cc), row_of(tm, mb))] = acc[mb][cc];
}
__device__ __forceinline__ void acc_from_tile(double (&acc)[2][8], const double *T, const TMap &tm) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) acc[mb][cc] = T[tidx(row_of(tm, mb), col_of(tm, cc))];
}
__device__ __forceinline__ void acc_from_tile_t(double (&acc)[2][8], const double *T, const TMap &tm) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) acc[mb][cc] = T[tidx(col_of(tm, cc), row_of(tm, mb))];
}

// ---- GEMM core on the FP64 tensor-core path ------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// acc (+|-)= A[rows, k0:k1] * B[cols, k0:k1]'   for the n-blocks nb with (NBMASK >> nb) & 1 (8 columns each).
// k0, k1 multiples of 4.  A and B are tiles (or leading column ranges of tiles) in tile format.
template <bool SUB, int NBMASK = 0xF>
__device__ __forceinline__ void tile_mma(double (&acc)[2][8], const double *__restrict__ A,
                                         const double *__restrict__ B, const TMap &tm, int k0, int k1) {
    // (k + t) & 3 == t for k % 4 == 0: the swizzle is a per-lane constant
    const int sw = tm.t << 2;
    const double *pa0 = A + tm.t * TS + ((tm.r0 + tm.g) ^ sw);
    const double *pa1 = A + tm.t * TS + ((tm.r0 + 8 + tm.g) ^ sw);
    const double *pb = B + tm.t * TS;
    const int bo0 = (tm.c0 + tm.g) ^ sw, bo1 = (tm.c0 + 8 + tm.g) ^ sw, bo2 = (tm.c0 + 16 + tm.g) ^ sw,
              bo3 = (tm.c0 + 24 + tm.g) ^ sw;
#pragma unroll 4
    for (int k = k0; k < k1; k += 4) {
        double a0 = pa0[k * TS], a1 = pa1[k * TS];
        if (SUB) {
            a0 = -a0;
            a1 = -a1;
        }
        if (NBMASK & 1) {
            const double b = pb[k * TS + bo0];
            dmma884(acc[0][0], acc[0][1], a0, b);
            dmma884(acc[1][0], acc[1][1], a1, b);
        }
        if (NBMASK & 2) {
            const double b = pb[k * TS + bo1];
            dmma884(acc[0][2], acc[0][3], a0, b);
            dmma884(acc[1][2], acc[1][3], a1, b);
        }
        if (NBMASK & 4) {
            const double b = pb[k * TS + bo2];
            dmma884(acc[0][4], acc[0][5], a0, b);
            dmma884(acc[1][4], acc[1][5], a1, b);
        }
        if (NBMASK & 8) {
            const double b = pb[k * TS + bo3];
            dmma884(acc[0][6], acc[0][7], a0, b);
            dmma884(acc[1][6], acc[1][7], a1, b);
        }
    }
}

// ---- Cholesky of the diagonal tile with its inverse, blocked ---------------------------------------------------
// In: acc = the 64 x 64 SPD tile (lower part used).  Out: acc = L (zeros above the diagonal), w = L^-1.
// The tile is processed in four 16-column panels (a rolled loop: the whole routine is ~20 KiB of SASS; the first,
// fully unrolled version was 270 KiB and made instruction fetch the second largest stall of the kernel):
//   1. the owners of the panel's columns (acc) and of the panel's rows of the running inverse (w) publish them;
//   2. warp 0 factors the 16 x 16 diagonal block in registers (one row per lane, pivots and columns exchanged
//      with warp shuffles: no block barrier inside) and inverts it by forward substitution;
//   3. all threads form the panel of L below the block (P * W16') and the new rows of the inverse (W16 * R);
//   4. all warps apply the rank-16 update to their register blocks of the trailing tile and of the inverse (DMMA).
// Three block barriers per panel.  The inverse rides along as a block Gauss-Jordan on the identity, so the
// triangular solves of the tiles below (L_ij = T_ij L_jj^-T) become plain GEMMs.
// scratch: 4096 doubles (P, R, Lp, Rp, each 64 x 16 in tile format; the three barriers order every reuse);
// L16s / W16s: 256 doubles each; rsbuf: 16 + 16 (reciprocal pivots, pivot-column exchange); pivbuf: 64 (pivots).
// Returns (in warp 0) -1 or the local index of the first non-positive pivot.
__device__ __forceinline__ int tile_potrf_inv(double (&acc)[2][8], double (&w)[2][8], const TMap &tm, double *scratch,
                                           double *L16s, double *W16s, double *rsbuf, double *pivbuf, int tid) {
    const int warp = tid >> 5, lane = tid & 31, wr = warp >> 1, wc = warp & 1;
    double *P = scratch;         // columns of the panel:           element (row, k) at tidx(row, k)
    double *R = P + 1024;        // rows of the inverse, as (col, k): tidx(col, k)
    double *Lp = P + 2048;       // panel of L, (row, k)
    double *Rp = P + 3072;       // new rows of the inverse, (col, k)
    double *colx = rsbuf + 16;   // pivot-column exchange of warp 0
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) w[mb][cc] = (row_of(tm, mb) == col_of(tm, cc)) ? 1.0 : 0.0;
    int fail = -1;
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        const int hs = p & 1;  // which column half of the owning warps holds the panel
        // 1. publish
        if (wc == (p >> 1)) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    P[tidx(row_of(tm, mb), col_of(tm, q) - tm.c0)] = hs ? acc[mb][4 + q] : acc[mb][q];  // panel-local column
        }
        if (wr == p) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) R[tidx(col_of(tm, cc), 8 * mb + tm.g)] = w[mb][cc];
        }
        __syncthreads();
        // 2. warp 0: 16 x 16 Cholesky, one row per lane, pivots and columns exchanged with warp shuffles
        if (warp == 0) {
            const int c_own = lane & 15;  // row owned by this lane (lanes 16..31 mirror 0..15)
            double a[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) a[c] = P[tidx(16 * p + c_own, c)];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                double piv = __shfl_sync(0xffffffffu, a[c], c);
                if (!(piv > 0.0)) {
                    if (fail < 0) fail = 16 * p + c;
                    piv = 1.0;
                }
                const double rs = rsqrt(piv);
                if (lane == 0) {
                    pivbuf[16 * p + c] = piv;
                    rsbuf[c] = rs;
                }
                const double l = a[c] * rs;
                a[c] = (c_own >= c) ? l : 0.0;
#pragma unroll
                for (int c2 = c + 1; c2 < 16; ++c2) {
                    const double l2 = __shfl_sync(0xffffffffu, l, c2);
                    a[c2] = fma(-l, l2, a[c2]);
                }
            }
            if (lane < 16) {
#pragma unroll
                for (int c = 0; c < 16; ++c) L16s[c * 16 + c_own] = a[c];  // L16[r][c], column-major
            }
            __syncwarp();
            double x[16];  // column c_own of L16^-1 (axpy-form forward substitution)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = (i == c_own) ? 1.0 : 0.0;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                x[c] *= rsbuf[c];
#pragma unroll
                for (int i = c + 1; i < 16; ++i) x[i] = fma(-L16s[c * 16 + i], x[c], x[i]);
            }
            if (lane < 16) {
#pragma unroll
                for (int i = 0; i < 16; ++i) W16s[i * 16 + c_own] = x[i];  // W16[i][c], row-major
            }
        }
        __syncthreads();
        // 3a. Lp: rows below the block = P * W16', rows of the block = L16, rows above = 0
        {
            const int row = tid & 63, cg = tid >> 6;
            double out[4] = {0.0, 0.0, 0.0, 0.0};
            if (row >= 16 * (p + 1)) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const double pk = P[tidx(row, k)];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int c = 4 * cg + q;
                        if (k <= c) out[q] = fma(pk, W16s[c * 16 + k], out[q]);
                    }
                }
            } else if (row >= 16 * p) {
#pragma unroll
                for (int q = 0; q < 4; ++q) out[q] = L16s[(4 * cg + q) * 16 + (row - 16 * p)];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) Lp[tidx(row, 4 * cg + q)] = out[q];
        }
        // 3b. Rp = W16 * R   (element (col, k) <- sum_{k2 <= k} W16[k][k2] R(col, k2))
        {
            const int col = tid & 63, kg = tid >> 6;
            double out[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                const double rv = R[tidx(col, k2)];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = 4 * kg + q;
                    if (k2 <= k) out[q] = fma(W16s[k * 16 + k2], rv, out[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) Rp[tidx(col, 4 * kg + q)] = out[q];
        }
        __syncthreads();
        // 4. register-block updates.  Column half h (n-blocks 2h, 2h+1) of this warp belongs to panel pc = 2 wc + h.
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pc = 2 * wc + h;
            if (pc == p) {  // finished columns of L
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        acc[mb][4 * h + q] = Lp[tidx(row_of(tm, mb), col_of(tm, 4 * h + q) - 16 * pc)];
            } else if (pc > p && wr >= pc) {  // trailing lower part: T -= Lp Lp'
                if (h == 0) tile_mma<true, 0x3>(acc, Lp, Lp, tm, 0, 16);
                else tile_mma<true, 0xC>(acc, Lp, Lp, tm, 0, 16);
            }
            if (pc <= p) {  // inverse: only columns <= the panel are non-zero in the new rows
                if (wr > p) {
                    if (h == 0) tile_mma<true, 0x3>(w, Lp, Rp, tm, 0, 16);
                    else tile_mma<true, 0xC>(w, Lp, Rp, tm, 0, 16);
                } else if (wr == p) {
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            w[mb][4 * h + q] = Rp[tidx(col_of(tm, 4 * h + q), 8 * mb + tm.g)];
                }
            }
        }
    }
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
            if (col_of(tm, cc) > row_of(tm, mb)) acc[mb][cc] = 0.0;
    return fail;
}

// ---- small helpers -----------------------------------------------------------------------------------------------
// (T' v)[c] for thread c < 64:  sum_m T(m, c) v[m]   (conflict-free: one column per thread, rows rotated)
__device__ __forceinline__ double tile_col_dot(const double *T, const double *v, int c) {
    double s = 0.0;
#pragma unroll 8
    for (int i = 0; i < TS; ++i) {
        const int m = (c + i) & (TS - 1);
        s = fma(T[tidx(m, c)], v[m], s);
    }
    return s;
}
// (T v)[r] for thread r < 64 over columns k0..k1-1:  sum_k T(r, k) v[k - k0]
__device__ __forceinline__ double tile_row_dot(const double *T, const double *v, int r, int k0, int k1) {
    double s = 0.0;
#pragma unroll 8
    for (int k = k0; k < k1; ++k) s = fma(T[tidx(r, k)], v[k - k0], s);
    return s;
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

// linear index t of a lower-triangular tile grid -> (i, j), i >= j
__device__ __forceinline__ void tri_unrank(long long t, int &i, int &j) {
    int ii = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (tri_index(ii + 1, 0) <= t) ++ii;
    while (tri_index(ii, 0) > t) --ii;
    i = ii;
    j = (int)(t - tri_index(ii, 0));
}

}  // namespace gpl
