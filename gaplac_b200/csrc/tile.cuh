// 64 x 64 FP64 tile primitives shared by the batched factorisation, the posterior/predict kernels and the
// large-n path.  One CTA of 128 threads (4 warps) owns one tile; the contraction runs on the FP64 tensor-core
// path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).  On sm_100a DMMA has the same peak as vector DFMA (37 TFLOP/s
// measured, tools/fp64_peak.cu) but its fragments are loaded from shared memory without the redundant
// broadcast reads a register-tiled DFMA loop needs (profiles/README.md, v1 -> v2).
//
// Why 4 warps with 16 x 64 warp tiles: (a) a CTA needs only 128 x <=128 registers and ~40 KiB of shared memory,
// so four independent GPs are resident per SM and the latency-bound stretches of one (pivot chains, barriers,
// first loads) are covered by the tensor work of the others; (b) every warp owns complete rows of the tile, so
// the triangular solve L_ij = T_ij L_jj^-T takes its row operand straight from the accumulator registers
// (quad shuffles), with no staging of T through shared memory.
//
// Tile storage (global workspace and shared memory alike): 64 x 64 doubles = 32 KiB, dense, column-major with
// an XOR swizzle of the row index:  element (r, c) lives at  c*64 + (r ^ ((c & 3) << 2)).
// A block-lower-triangular matrix of nt x nt tiles is stored tile-major: tile (i, j), i >= j, at offset
// (i(i+1)/2 + j) * 4096 doubles, so one tile (or any run of whole columns of it) is one contiguous copy.
// The swizzle makes every fragment load (8 rows x 4 consecutive columns, one double per lane) hit 16 distinct
// 8-byte banks per half-warp: conflict-free LDS.64, 2 wavefronts per 256 bytes.
//
// GEMM form used everywhere:  C (+|-)= A * B'  with A = (rows x k) and B = (cols x k), both in tile format.
// Thread -> accumulator map (warp w -> rows 16w..16w+15; lane: g = lane/4, t = lane%4):
//   rows  row_of(mb)  = 16*w + 8*mb + g,            mb = 0, 1
//   cols  col_of(cc)  = 8*(cc/2) + 2*t + cc%2,      cc = 0..15      -> acc[2][16]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpl {

constexpr int TS = 64;                 // tile edge
constexpr int TILE_ELEMS = TS * TS;    // 4096 doubles
constexpr int TILE_BYTES = TILE_ELEMS * 8;
constexpr int NTHREADS = 128;
constexpr int NWARPS = NTHREADS / 32;
constexpr int NCC = 16;                // accumulator columns per thread

__host__ __device__ __forceinline__ long long tri_index(int i, int j) { return (long long)i * (i + 1) / 2 + j; }

// position of element (r, c) inside a tile
__host__ __device__ __forceinline__ int tidx(int r, int c) { return c * TS + (r ^ ((c & 3) << 2)); }

struct TMap {
    int r0;  // first row of the warp's 16-row band
    int g;   // lane / 4
    int t;   // lane % 4
};
__device__ __forceinline__ TMap thread_map(int tid) {
    const int lane = tid & 31;
    TMap m;
    m.r0 = (tid >> 5) * 16;
    m.g = lane >> 2;
    m.t = lane & 3;
    return m;
}
__device__ __forceinline__ int row_of(const TMap &tm, int mb) { return tm.r0 + 8 * mb + tm.g; }
__device__ __forceinline__ int col_of(const TMap &tm, int cc) { return 8 * (cc >> 1) + 2 * tm.t + (cc & 1); }

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

// ---- bulk asynchronous copies (TMA engine, 1-D: cp.async.bulk) completing on shared-memory mbarriers --------------
// One thread moves a whole contiguous chunk with one instruction; consumers wait on the barrier's phase parity.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
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

__device__ __forceinline__ void tile_store(double *__restrict__ gmem, const double *smem, int tid) {
#pragma unroll
    for (int it = 0; it < TILE_BYTES / 16 / NTHREADS; ++it) {
        const int idx = it * NTHREADS + tid;
        reinterpret_cast<double2 *>(gmem)[idx] = reinterpret_cast<const double2 *>(smem)[idx];
    }
}

// ---- register block <-> tile (shared or global) ------------------------------------------------------------
__device__ __forceinline__ void acc_zero(double (&acc)[2][NCC]) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) acc[mb][cc] = 0.0;
}
// tile element (row, col) <- acc
__device__ __forceinline__ void acc_to_tile(double *T, const double (&acc)[2][NCC], const TMap &tm) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) T[tidx(row_of(tm, mb), col_of(tm, cc))] = acc[mb][cc];
}
// transposed: tile element (col, row) <- acc
__device__ __forceinline__ void acc_to_tile_t(double *T, const double (&acc)[2][NCC], const TMap &tm) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) T[tidx(col_of(tm, cc), row_of(tm, mb))] = acc[mb][cc];
}
__device__ __forceinline__ void acc_from_tile(double (&acc)[2][NCC], const double *T, const TMap &tm) {
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) acc[mb][cc] = T[tidx(row_of(tm, mb), col_of(tm, cc))];
}

#ifndef GPL_MMA_UNROLL
#define GPL_MMA_UNROLL 4
#endif
constexpr int MMA_UNROLL = GPL_MMA_UNROLL;  // k4-steps unrolled in tile_mma
// ---- GEMM core on the FP64 tensor-core path ------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// acc (+|-)= A[rows, k0:k1] * B[cols, k0:k1]'   for the n-blocks nb with (NBMASK >> nb) & 1 (8 columns each); with
// LIMIT additionally only nb < nbmax (warp-uniform run-time bound: diagonal tiles skip the blocks above the diagonal).
// k0, k1 multiples of 4.  A and B are tiles (or runs of whole columns starting at a multiple of 4) in tile format.
template <bool SUB, int NBMASK = 0xFF, bool LIMIT = false>
__device__ __forceinline__ void tile_mma(double (&acc)[2][NCC], const double *__restrict__ A,
                                         const double *__restrict__ B, const TMap &tm, int k0, int k1, int nbmax = 8) {
    // (k + t) & 3 == t for k % 4 == 0: the swizzle is a per-lane constant.  For the column operand the row index is
    // 8 nb + g: xor with sw = 4t flips bit 2 of g and (for t >= 2) bit 3, i.e. swaps odd and even n-blocks.
    const int sw = tm.t << 2;
    const double *pa0 = A + tm.t * TS + ((tm.r0 + tm.g) ^ sw);
    const double *pa1 = A + tm.t * TS + ((tm.r0 + 8 + tm.g) ^ sw);
    const double *pbe = B + tm.t * TS + (tm.g ^ (sw & 4)) + (sw & 8);        // even n-blocks
    const double *pbo = B + tm.t * TS + (tm.g ^ (sw & 4)) + (8 ^ (sw & 8));  // odd n-blocks
#pragma unroll MMA_UNROLL
    for (int k = k0; k < k1; k += 4) {
        double a0 = pa0[k * TS], a1 = pa1[k * TS];
        if (SUB) {
            a0 = -a0;
            a1 = -a1;
        }
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            if (((NBMASK >> nb) & 1) && (!LIMIT || nb < nbmax)) {
                const double b = ((nb & 1) ? pbo : pbe)[k * TS + 16 * (nb >> 1)];
                dmma884(acc[0][2 * nb], acc[0][2 * nb + 1], a0, b);
                dmma884(acc[1][2 * nb], acc[1][2 * nb + 1], a1, b);
            }
        }
    }
}

// ---- triangular solve from the accumulator registers -------------------------------------------------------------
// acc := acc * W'   for a lower-triangular W (64 x 64, tile format, in shared memory):
//   X[:, n] = sum_{k <= n} T[:, k] W[n, k].
// The row operand T is the accumulator itself: the A fragment of k-chunk kap (columns 4 kap .. 4 kap + 3) for lane
// (g, t) is T[row][4 kap + t], which lives in lane (g, 2 (kap & 1) + t / 2), register 2 (kap / 2) + (t & 1): two
// quad shuffles and a select.  Output n-blocks are produced in two groups, high half first, so the result can
// overwrite the accumulator in place (X[:, n-block] never needs T columns beyond its own block).
template <int NB_LO, int NB_HI>
__device__ __forceinline__ void trsm_group(double (&acc)[2][NCC], const double *__restrict__ W, const TMap &tm) {
    constexpr int NG = NB_HI - NB_LO;
    double x[2][2 * NG];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int q = 0; q < 2 * NG; ++q) x[mb][q] = 0.0;
    const int sw = tm.t << 2;
    const int lane_base = tm.g << 2;
    const double *pbe = W + tm.t * TS + (tm.g ^ (sw & 4)) + (sw & 8);
    const double *pbo = W + tm.t * TS + (tm.g ^ (sw & 4)) + (8 ^ (sw & 8));
#pragma unroll
    for (int kap = 0; kap < 2 * NB_HI; ++kap) {
        const int nbs = kap >> 1, h = kap & 1;
        const int src = lane_base | (2 * h + (tm.t >> 1));
        double a[2];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            const double v0 = __shfl_sync(0xffffffffu, acc[mb][2 * nbs], src);
            const double v1 = __shfl_sync(0xffffffffu, acc[mb][2 * nbs + 1], src);
            a[mb] = (tm.t & 1) ? v1 : v0;
        }
#pragma unroll
        for (int nb = (nbs > NB_LO ? nbs : NB_LO); nb < NB_HI; ++nb) {
            const double b = ((nb & 1) ? pbo : pbe)[4 * kap * TS + 16 * (nb >> 1)];
            dmma884(x[0][2 * (nb - NB_LO)], x[0][2 * (nb - NB_LO) + 1], a[0], b);
            dmma884(x[1][2 * (nb - NB_LO)], x[1][2 * (nb - NB_LO) + 1], a[1], b);
        }
    }
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int q = 0; q < 2 * NG; ++q) acc[mb][2 * NB_LO + q] = x[mb][q];
}
// acc := acc * W' for a full lower-triangular inverse tile W (used where the inverse exists anyway: large-n path)
__device__ __forceinline__ void tile_trsm_w(double (&acc)[2][NCC], const double *__restrict__ W, const TMap &tm) {
    trsm_group<4, 8>(acc, W, tm);
    trsm_group<0, 4>(acc, W, tm);
}

// ---- Cholesky of the diagonal tile, blocked -----------------------------------------------------------------------
// In: acc = the 64 x 64 SPD tile (lower part used).  Out: acc = L (zeros above the diagonal) and, in D, the inverses
// of the four 16 x 16 diagonal blocks of L (block p at D + p*DBLK, element (i, k) at i*DLD + k; DLD = 20 makes the
// DMMA fragment loads of these blocks bank-conflict free).  The tile is processed in four 16-column panels (a rolled
// loop keeps the routine small in the instruction cache):
//   1. every warp publishes its 16 rows of the panel's columns;
//   2. warp 0 factors the 16 x 16 diagonal block in registers (one row per lane, pivots and columns exchanged
//      with warp shuffles: no block barrier inside) and inverts it by forward substitution;
//   3. all threads form the panel of L below the block (P * W16');
//   4. all warps apply the rank-16 update to their register blocks of the trailing tile (DMMA).
// Three block barriers per panel.  The triangular solves that follow (tile_trsm_ld, tile_forward_solve) use L and the
// four block inverses, so the full 64 x 64 inverse is only formed where a caller needs it (tile_inverse_from_ld).
// scratch: 2048 doubles (P, Lp: 64 x 16 in tile format); L16s: 256; rsbuf: 16; pivbuf: 64 (pivots, for logdet).
// Returns (in the chain warp) -1 or the local index of the first non-positive pivot.
constexpr int DLD = 20;
constexpr int DBLK = 16 * DLD;    // 320 doubles per block inverse
constexpr int DSIZE = 4 * DBLK;   // 1280 doubles = 10 KiB per diagonal tile

template <int PC>
__device__ __forceinline__ void potrf_panel_update(double (&acc)[2][NCC], const TMap &tm, int p, int warp,
                                                   const double *Lp) {
    constexpr int MASK = 0x3 << (2 * PC);
    if (PC == p) {  // finished columns of L
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[mb][4 * PC + q] = Lp[tidx(row_of(tm, mb), col_of(tm, q))];
    } else if (PC > p && warp >= PC) {  // trailing lower part: T -= Lp Lp'
        tile_mma<true, MASK>(acc, Lp, Lp, tm, 0, 16);
    }
}

__device__ __forceinline__ int tile_potrf(double (&acc)[2][NCC], const TMap &tm, double *scratch, double *L16s,
                                          double *D, double *rsbuf, double *pivbuf, int tid, int chain_warp = 0) {
    // chain_warp: which warp runs the 16 x 16 pivot chains (callers rotate it over co-resident CTAs so that the
    // chains of different CTAs land on different SM sub-partitions); its return value carries `fail`
    const int warp = tid >> 5, lane = tid & 31;
    double *P = scratch;         // columns of the panel: element (row, k) at tidx(row, k)
    double *Lp = scratch + 1024; // panel of L, (row, k)
    int fail = -1;
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        double *W16s = D + p * DBLK;
        // 1. publish: the panel's columns are accumulator columns 4p..4p+3 of every thread
#pragma unroll
        for (int pp = 0; pp < 4; ++pp)
            if (pp == p) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int q = 0; q < 4; ++q) P[tidx(row_of(tm, mb), col_of(tm, q))] = acc[mb][4 * pp + q];
            }
        __syncthreads();
        // 2. one warp: 16 x 16 Cholesky, one row per lane, pivots and columns exchanged with warp shuffles
        if (warp == chain_warp) {
            const int r_own = lane & 15;  // row owned by this lane (lanes 16..31 mirror 0..15)
            double a[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) a[c] = P[tidx(16 * p + r_own, c)];
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
                a[c] = (r_own >= c) ? l : 0.0;
#pragma unroll
                for (int c2 = c + 1; c2 < 16; ++c2) {
                    const double l2 = __shfl_sync(0xffffffffu, l, c2);
                    a[c2] = fma(-l, l2, a[c2]);
                }
            }
            if (lane < 16) {
#pragma unroll
                for (int c = 0; c < 16; ++c) L16s[c * 16 + r_own] = a[c];  // L16[r][c], column-major
            }
            __syncwarp();
            double x[16];  // column r_own of L16^-1 (axpy-form forward substitution)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = (i == r_own) ? 1.0 : 0.0;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                x[c] *= rsbuf[c];
#pragma unroll
                for (int i = c + 1; i < 16; ++i) x[i] = fma(-L16s[c * 16 + i], x[c], x[i]);
            }
            if (lane < 16) {
#pragma unroll
                for (int i = 0; i < 16; ++i) W16s[i * DLD + r_own] = x[i];  // W16[i][c]
            }
        }
        __syncthreads();
        // 3. Lp (thread = one row, 8 of the 16 panel columns): rows below the block = P * W16', rows of the block =
        //    L16, rows above = 0
        {
            const int row = tid & 63, c8 = (tid >> 6) * 8;
            double out[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) out[q] = 0.0;
            if (row >= 16 * (p + 1)) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const double pk = P[tidx(row, k)];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (k <= c8 + q) out[q] = fma(pk, W16s[(c8 + q) * DLD + k], out[q]);
                    }
                }
            } else if (row >= 16 * p) {
#pragma unroll
                for (int q = 0; q < 8; ++q) out[q] = L16s[(c8 + q) * 16 + (row - 16 * p)];
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) Lp[tidx(row, c8 + q)] = out[q];
        }
        __syncthreads();
        // 4. register-block updates, per 16-column panel PC of the accumulator (n-blocks 2 PC, 2 PC + 1)
        potrf_panel_update<0>(acc, tm, p, warp, Lp);
        potrf_panel_update<1>(acc, tm, p, warp, Lp);
        potrf_panel_update<2>(acc, tm, p, warp, Lp);
        potrf_panel_update<3>(acc, tm, p, warp, Lp);
    }
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc)
            if (col_of(tm, cc) > row_of(tm, mb)) acc[mb][cc] = 0.0;
    return fail;
}

// ---- triangular solve with L and its block inverses, from the accumulator registers ------------------------------------
// acc := acc * L^-T  (X L' = T), L = 64 x 64 lower-triangular tile in shared memory (tile format), D = its four block
// inverses.  Blocked forward substitution over the four 16-column panels, everything in place:
//   X_p = (T_p - sum_{q<p} X_q L_pq') W16_p'
// The row operands (finished panels X_q, and T_p itself) come straight from the accumulator registers: the A fragment
// of k-chunk kap (columns 4 kap .. 4 kap + 3) for lane (g, t) is acc[row][4 kap + t], which lives in lane
// (g, 2 (kap & 1) + t / 2), register 2 (kap / 2) + (t & 1): two quad shuffles and a select.
__device__ __forceinline__ void acc_a_frag(const double (&acc)[2][NCC], int kap_static, const TMap &tm, double (&a)[2]) {
    const int src = (tm.g << 2) | (2 * (kap_static & 1) + (tm.t >> 1));
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        const double v0 = __shfl_sync(0xffffffffu, acc[mb][2 * (kap_static >> 1)], src);
        const double v1 = __shfl_sync(0xffffffffu, acc[mb][2 * (kap_static >> 1) + 1], src);
        a[mb] = (tm.t & 1) ? v1 : v0;
    }
}

template <int P_>
__device__ __forceinline__ void trsm_ld_panel(double (&acc)[2][NCC], const double *__restrict__ L,
                                              const double *__restrict__ D, const TMap &tm) {
    const int sw = tm.t << 2;
    // column operand rows 16 P_ + 8 nbl + g of L (n-blocks 2 P_, 2 P_ + 1)
    const double *pl0 = L + tm.t * TS + ((16 * P_ + tm.g) ^ sw);
    const double *pl1 = L + tm.t * TS + ((16 * P_ + 8 + tm.g) ^ sw);
    // T_p -= X_q L_pq'  for the finished panels q < P_ (k = columns 16 q .. 16 q + 15)
#pragma unroll
    for (int kap = 0; kap < 4 * P_; ++kap) {
        double a[2];
        acc_a_frag(acc, kap, tm, a);
        const double b0 = pl0[4 * kap * TS], b1 = pl1[4 * kap * TS];
        dmma884(acc[0][4 * P_], acc[0][4 * P_ + 1], -a[0], b0);
        dmma884(acc[1][4 * P_], acc[1][4 * P_ + 1], -a[1], b0);
        dmma884(acc[0][4 * P_ + 2], acc[0][4 * P_ + 3], -a[0], b1);
        dmma884(acc[1][4 * P_ + 2], acc[1][4 * P_ + 3], -a[1], b1);
    }
    // X_p = T_p W16_p'   (W16 lower triangular: k-chunk kl feeds n-block nbl >= kl / 2)
    const double *W16 = D + P_ * DBLK;
    double x[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
#pragma unroll
    for (int kl = 0; kl < 4; ++kl) {
        double a[2];
        acc_a_frag(acc, 4 * P_ + kl, tm, a);
#pragma unroll
        for (int nbl = kl >> 1; nbl < 2; ++nbl) {
            const double b = W16[(8 * nbl + tm.g) * DLD + 4 * kl + tm.t];
            dmma884(x[0][2 * nbl], x[0][2 * nbl + 1], a[0], b);
            dmma884(x[1][2 * nbl], x[1][2 * nbl + 1], a[1], b);
        }
    }
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[mb][4 * P_ + q] = x[mb][q];
}
__device__ __forceinline__ void tile_trsm_ld(double (&acc)[2][NCC], const double *__restrict__ L,
                                             const double *__restrict__ D, const TMap &tm) {
    trsm_ld_panel<0>(acc, L, D, tm);
    trsm_ld_panel<1>(acc, L, D, tm);
    trsm_ld_panel<2>(acc, L, D, tm);
    trsm_ld_panel<3>(acc, L, D, tm);
}

// Right-looking variant of the same solve, split so that the caller can stream L through a load pipeline 16 columns
// at a time:  for q = 0..3:  trsm_rl_solve<q> (X_q = T_q W16_q', registers and D only), then trsm_rl_update<q> with the
// chunk L[:, 16q .. 16q+15] (tile format, 64 x 16) to apply  T_p -= X_q L_pq'  to every later panel p > q.
template <int Q_>
__device__ __forceinline__ void trsm_rl_solve(double (&acc)[2][NCC], const double *__restrict__ D, const TMap &tm) {
    const double *W16 = D + Q_ * DBLK;
    double x[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
#pragma unroll
    for (int kl = 0; kl < 4; ++kl) {
        double a[2];
        acc_a_frag(acc, 4 * Q_ + kl, tm, a);
#pragma unroll
        for (int nbl = kl >> 1; nbl < 2; ++nbl) {
            const double b = W16[(8 * nbl + tm.g) * DLD + 4 * kl + tm.t];
            dmma884(x[0][2 * nbl], x[0][2 * nbl + 1], a[0], b);
            dmma884(x[1][2 * nbl], x[1][2 * nbl + 1], a[1], b);
        }
    }
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[mb][4 * Q_ + q] = x[mb][q];
}
template <int Q_>
__device__ __forceinline__ void trsm_rl_update(double (&acc)[2][NCC], const double *__restrict__ chunk, const TMap &tm) {
    const int sw = tm.t << 2;
    const double *pbe = chunk + tm.t * TS + (tm.g ^ (sw & 4)) + (sw & 8);        // even n-blocks
    const double *pbo = chunk + tm.t * TS + (tm.g ^ (sw & 4)) + (8 ^ (sw & 8));  // odd n-blocks
#pragma unroll
    for (int kl = 0; kl < 4; ++kl) {
        double a[2];
        acc_a_frag(acc, 4 * Q_ + kl, tm, a);
#pragma unroll
        for (int nb = 2 * (Q_ + 1); nb < 8; ++nb) {
            const double b = ((nb & 1) ? pbo : pbe)[4 * kl * TS + 16 * (nb >> 1)];
            dmma884(acc[0][2 * nb], acc[0][2 * nb + 1], -a[0], b);
            dmma884(acc[1][2 * nb], acc[1][2 * nb + 1], -a[1], b);
        }
    }
}

// z = L^-1 y for one 64-vector, all threads of the CTA (four block steps, two barriers each):
//   z_p = W16_p (y_p - sum_{q<p} L_pq z_q).  y is updated in place in `ybuf` (shared); the result replaces it.
__device__ __forceinline__ void tile_forward_solve(const double *__restrict__ L, const double *__restrict__ D,
                                                   double *ybuf, double *tmp, int tid) {
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        __syncthreads();
        if (tid >= 16 * p && tid < 16 * p + 16) {
            const int il = tid - 16 * p;
            const double *W16 = D + p * DBLK;
            double s = 0.0;
            for (int k = 0; k <= il; ++k) s = fma(W16[il * DLD + k], ybuf[16 * p + k], s);
            tmp[il] = s;
        }
        __syncthreads();
        if (tid >= 16 * p && tid < 16 * p + 16) ybuf[tid] = tmp[tid - 16 * p];
        else if (tid >= 16 * (p + 1) && tid < TS) {
            double s = ybuf[tid];
#pragma unroll
            for (int k = 0; k < 16; ++k) s = fma(-L[tidx(tid, 16 * p + k)], tmp[k], s);
            ybuf[tid] = s;
        }
    }
    __syncthreads();
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

// Compacted list of the tile indices k in [k0, k1) for which use_k(k) holds, built by warp 0 (one predicate evaluation per
// lane, one ballot per 32 indices); *nk receives the count.  A block barrier must follow before kl / nk are read.
template <class Pred>
__device__ __forceinline__ void build_tile_list(short *kl, int *nk, int k0, int k1, Pred use_k, int tid) {
    if (tid >= 32) return;
    int c = 0;
    for (int kb = k0; kb < k1; kb += 32) {
        const int k = kb + tid;
        const bool use = k < k1 && use_k(k);
        const unsigned m = __ballot_sync(0xffffffffu, use);
        if (use) kl[c + __popc(m & ((1u << tid) - 1u))] = (short)k;
        c += __popc(m);
    }
    if (tid == 0) *nk = c;
}

// deterministic block-wide sum, result valid in every thread; red = NWARPS doubles of shared memory
__device__ __forceinline__ double block_sum(double v, double *red, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NWARPS; ++i) s += red[i];
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
