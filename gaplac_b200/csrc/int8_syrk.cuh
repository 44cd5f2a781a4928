// INT8 split rank-K update on the tcgen05 tensor path:  C(i,j) -= sum_k A[i][k] A[j][k]  on 128 x 128 tiles of a lower
// triangle, from exact INT8 x INT8 -> INT32 products of 7-bit slices of A (Ozaki scheme).
//
//     A[i][k] = 2^e_i * sum_{s<S} q_s[i][k] * 2^(-7(s+1)),   q_s in [-127, 127]  (exact: truncation, not rounding)
//     A A^T   = 2^(e_i+e_j) * sum_g 2^(-7(g+2)) * sum_{s+t=g} q_s q_t^T          (groups g >= S dropped: < 2^(-7S) relative)
//
// sm_100a has no f64 kind on tcgen05 (the FP64 tensor path is DMMA.8x8x4, 37 TFLOP/s) but kind::i8 at a nominal 4.5 POP/s.
// One group g is summed in one TMEM accumulator ((g+1) * K * 127^2 < 2^31 for K <= 16384), read back with tcgen05.ld,
// converted exactly and accumulated in FP64 registers.  S = 8 keeps 56 bits below the row maximum: 36 INT8 products.
//
// Structure per CTA (persistent over output tiles, one CTA per SM): warp 0 = TMA producer (cp.async.bulk.tensor, 128B
// swizzle), warp 1 = MMA issuer (one elected lane) and TMEM owner (four INT32 accumulators = all 512 columns), warps 2..9 =
// epilogue (64 FP64 accumulators per thread).  The caller supplies the slices (S x n_rows x K int8, K-major, slice-major)
// and the row scales 2^e_i; C is either a column-major dense matrix or the library's tile-major lower storage.
//
// Header-only so that the microbenchmark (tools/ozaki) and the large-n factorisation (trail_int8.cu) share one kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>

namespace gpl_i8 {


constexpr int BM = 128;     // output tile rows = tcgen05 M
constexpr int BN = 128;     // output tile columns = tcgen05 N
constexpr int KB = 128;     // int8 elements (= bytes) of K per pipeline stage: one 128-byte swizzle row
constexpr int UK = 32;      // K of one tcgen05.mma.kind::i8
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 8;
constexpr int I8_THREADS = 32 * (2 + EPI_WARPS);
constexpr uint32_t TILE_BYTES = BM * KB;       // 16 KiB: one slice of 128 rows x 128 bytes of K
constexpr uint32_t STAGE_BYTES = 4 * TILE_BYTES;  // up to two row-side and two column-side slices per stage
constexpr int NSLOT = 4;                        // INT32 accumulators in TMEM (all 512 columns)
constexpr uint32_t TMEM_COLS = NSLOT * BN;
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + 256;
constexpr int MAX_STEPS = 16;

// The slice products are walked in 2 x 2 blocks {s0, s0+1} x {t0, t0+1} (s0, t0 even): one stage brings four slice
// tiles for up to four products, which halves the L2 -> shared-memory traffic per product of the plain pair order.
// A block feeds the groups g = s0+t0, +1, +2; blocks go by descending s0+t0, so the long groups of the next tile run
// while the epilogue of this one is still writing C.  The host writes the schedule; every role walks the same table.
struct Prod {
    uint8_t a, b, g, flags;  // a, b: which of the stage's row / column slices; flags: 1 = first product of g, 2 = last
};
// One tcgen05.mma chain of a stage.  Two products that share the row-side slice and whose groups sit in adjacent TMEM
// slots are issued as ONE N = 256 instruction (the two column-side slices are adjacent in the stage, 256 rows of B):
// the A tile is then read from shared memory once for both, 96 instead of 128 bytes per clock at full rate.
struct Issue {
    uint8_t a, b, slot, flags;  // flags: 1 = overwrite at the first K block, 2 = N = 256, 4 / 8 = commit group(s) after the last K block
};
struct Step {
    uint8_t na, nb, nprod, nissue;
    uint8_t sa[2], sb[2];  // slice indices to load
    Prod prod[4];
    Issue iss[4];
};
struct Schedule {
    int nsteps, S;
    uint8_t gorder[12];  // groups in the order they complete
    Step step[MAX_STEPS];
};

// instruction descriptor for kind::i8, dense, S32 accumulator, A and B signed 8-bit, both K-major, M = 128, N = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t IDESC_WIDE = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(2 * BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

static __device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
static __device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a lost arrival becomes a clean early exit with a breadcrumb (no trap, no hung GPU).  Returns false when
// this wait timed out or another role already gave up; every role then falls through to the common teardown.
static __device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int code, volatile int *abort_flag, volatile int *dbg) {
    // dbg[3] doubles as the "failed" flag the host turns into an error (the library sets info = -1 from it)
    unsigned long long t0 = 0;
    for (uint32_t spin = 1;; ++spin) {
        uint32_t ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (ok) return true;
        if ((spin & 63u) == 0) {
            if (*abort_flag) return false;
            const unsigned long long t = globaltimer();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 1000000000ull) {  // 1 s: a whole launch takes milliseconds; every wait here is on this CTA's own roles
                *abort_flag = 1;
                if (atomicCAS((int *)dbg, 0, code) == 0) {
                    dbg[1] = (int)blockIdx.x;
                    dbg[2] = (int)parity;
                    dbg[3] = -1;
                }
                return false;
            }
        }
    }
}
static __device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// One deterministic leader lane of a converged warp.  The issuing warps keep their control flow warp-uniform and
// predicate only the tcgen05 / TMA instructions with this, so that ptxas keeps descriptors in uniform registers
// (a divergent "if (lane == 0)" region costs a vote + five R2UR + branch loop around every single UTCIMMA).
static __device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
static __device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
static __device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
static __device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
static __device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// shared-memory matrix descriptor: K-major operand, rows of 128 bytes, 128B swizzle, 8-row groups 1024 bytes apart
// low word: start address >> 4 in bits [0,14), leading byte offset (unused with a swizzled K-major operand) = 1 in
// bits [16,30); high word: stride byte offset 1024 >> 4 between 8-row groups, descriptor version 1 (sm_100), SWIZZLE_128B
constexpr uint64_t DESC_HI = ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
static __device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Barriers {
    uint64_t full[STAGES], empty[STAGES], tfull[NSLOT], tempty[NSLOT];
    uint32_t tmem_base;
    int abort_flag;
};

// What one launch updates.  Output blocks are 128 x 128; block row / column indices are relative to the region the slices
// describe (slice row r = region row r).  Column blocks [cb0, cb1), every block row I >= J.
struct View {
    const double *rowscale;  // n_rows powers of two: 2^e_i
    double *C;               // layout 0: column-major dense, element (i, j) at C[j * ldc + i]
    long ldc;
    double *tiles;           // layout 1: tile-major lower storage of 64 x 64 tiles (tile.cuh: tri_index, tidx); the region
    int nt64, t0;            //           starts at tile row / column t0 of nt64
    int layout, mode;        // mode 0: C -= A A^T, 1: C = A A^T
    int n_rows, K;           // rows of the slice arrays (multiple of 128), K (multiple of 128)
    int cb0, cb1;
    int *dbg;                // 4 ints: breadcrumb of a timed-out barrier wait (code, CTA, parity); may be host-mapped
    int *info;               // optional: set to -1 on a timed-out wait (the library's "device-side failure" convention)
};

static __device__ __forceinline__ void block_of(const View &v, int t, int &bi, int &bj) {
    const int nb = v.n_rows / BM;
    int j = v.cb0;
    while (t >= nb - j) {
        t -= nb - j;
        ++j;
    }
    bi = j + t;
    bj = j;
}
static __host__ __device__ inline int blocks_in(int n_rows, int cb0, int cb1) {
    const int nb = n_rows / BM;
    int c = 0;
    for (int j = cb0; j < cb1; ++j) c += nb - j;
    return c;
}

// 10 warps are allocated as 12 (granularity 4), so the register file allows 65536 / 384 = 168 registers per thread
template <int LAYOUT>
static __global__ void __launch_bounds__(I8_THREADS, 1)
int8_syrk_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Schedule sch, const __grid_constant__ View vw) {
    const int n_rows = vw.n_rows, K = vw.K, mode = vw.mode;
    int *const dbg = vw.dbg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    Barriers *bars = reinterpret_cast<Barriers *>(smem + (size_t)STAGES * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = blocks_in(n_rows, vw.cb0, vw.cb1), nkb = K / KB;
    const int S = sch.S;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&bars->full[i], 1);
            mbar_init(&bars->empty[i], 1);
        }
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&bars->tfull[i], 1);
            mbar_init(&bars->tempty[i], EPI_WARPS);
        }
        bars->abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    volatile int *abortp = &bars->abort_flag;

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int ti, tj;
            block_of(vw, tile, ti, tj);
            for (int st = 0; st < sch.nsteps; ++st) {
                const Step &sp = sch.step[st];
                for (int kb = 0; kb < nkb; ++kb) {
                    if (!mbar_wait(&bars->empty[stage], phase ^ 1, 1, abortp, dbg)) goto done;
                    uint8_t *sa = smem + (size_t)stage * STAGE_BYTES;
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&bars->full[stage], (uint32_t)(sp.na + sp.nb) * TILE_BYTES);
                        for (int i = 0; i < sp.na; ++i)
                            tma_load_2d(sa + i * TILE_BYTES, &tmap, kb * KB, sp.sa[i] * n_rows + ti * BM, &bars->full[stage]);
                        for (int i = 0; i < sp.nb; ++i)
                            tma_load_2d(sa + (2 + i) * TILE_BYTES, &tmap, kb * KB, sp.sb[i] * n_rows + tj * BN, &bars->full[stage]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        int stage = 0;
        uint32_t phase = 0, spar = 0;  // spar: one use-parity bit per TMEM slot
        const uint32_t lo0 = ((smem_u32(smem) & 0x3FFFFu) >> 4) | 0x10000u;  // descriptor low word of stage 0, slice 0
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int st = 0; st < sch.nsteps; ++st) {
                // decode the step once; the K loop below is then straight-line issue with uniform-register arithmetic
                const Step &sp = sch.step[st];
                const int nissue = sp.nissue;
                uint32_t aoff[4], boff[4], dcol[4], idesc[4], first = 0, last = 0, overwrite = 0;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const Issue is = sp.iss[p < nissue ? p : 0];
                    aoff[p] = is.a * (TILE_BYTES >> 4);
                    boff[p] = (2 + is.b) * (TILE_BYTES >> 4);
                    dcol[p] = is.slot * BN;
                    idesc[p] = (is.flags & 2) ? IDESC_WIDE : IDESC;
                    if (p < nissue && (is.flags & 1)) {
                        first |= (is.flags & 2 ? 3u : 1u) << is.slot;  // slots that get a new tenant
                        overwrite |= 1u << p;                           // this chain starts them: no accumulate at K block 0
                    }
                    if (p < nissue && (is.flags & 4)) last |= 1u << is.slot;
                    if (p < nissue && (is.flags & 8)) last |= 2u << is.slot;
                }
#pragma unroll
                for (int slot = 0; slot < NSLOT; ++slot)
                    if (first >> slot & 1) {  // a new group takes the slot: its last tenant must be drained
                        if (!mbar_wait(&bars->tempty[slot], ((spar >> slot) & 1) ^ 1, 2, abortp, dbg)) goto done;
                    }
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb) {
                    if (!mbar_wait(&bars->full[stage], phase, 3, abortp, dbg)) goto done;
                    tc_fence_after();
                    const uint32_t lo = lo0 + (uint32_t)stage * (STAGE_BYTES >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int p = 0; p < 4; ++p)
                            if (p < nissue) {
                                const uint32_t fresh = (kb == 0 && (overwrite >> p & 1)) ? 0u : 1u;
#pragma unroll
                                for (int k4 = 0; k4 < KB / UK; ++k4)
                                    tc_mma_i8(tmem_base + dcol[p], DESC_HI | (uint64_t)(lo + aoff[p] + k4 * (UK / 16)),
                                              DESC_HI | (uint64_t)(lo + boff[p] + k4 * (UK / 16)), idesc[p], k4 == 0 ? fresh : 1u);
                            }
                        tc_commit(&bars->empty[stage]);  // frees the stage when these MMAs have read it
                        if (kb == nkb - 1) {
#pragma unroll
                            for (int slot = 0; slot < NSLOT; ++slot)
                                if (last >> slot & 1) tc_commit(&bars->tfull[slot]);  // group complete in TMEM
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                spar ^= last;
            }
        }
    } else {
        const int q = warp & 3;         // TMEM lane quarter this warp may read
        const int h = (warp - 2) >> 2;  // column half of the tile
        uint32_t spar = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int ti, tj;
            block_of(vw, tile, ti, tj);
            double acc[64];
#pragma unroll
            for (int c = 0; c < 64; ++c) acc[c] = 0.0;
            // The TMEM read leaves one row per lane, so a warp touches 256 contiguous bytes per column of a column-major C;
            // the library's 64 x 64 tiles are stored in the same direction (rows of a column contiguous, XOR-swizzled in
            // groups of four).
            const int row = ti * BM + q * 32 + lane;
            double *cp;
            bool valid = true;
            if (LAYOUT == 0) {
                cp = vw.C + (long)(tj * BN + h * 64) * vw.ldc + row;
            } else {
                const int ti64 = vw.t0 + 2 * ti + (q >> 1), tj64 = vw.t0 + 2 * tj + h;
                valid = ti64 < vw.nt64 && ti64 >= tj64;  // past the last tile row, or the upper tile of a diagonal block
                cp = vw.tiles + ((long long)ti64 * (ti64 + 1) / 2 + tj64) * 4096;
            }
            const int r64 = (q & 1) * 32 + lane;  // row inside the 64 x 64 tile (LAYOUT 1)
            auto at = [&](int c) -> double * { return LAYOUT == 0 ? cp + (long)c * vw.ldc : cp + c * 64 + (r64 ^ ((c & 3) << 2)); };
            if (mode == 0 && valid && (lane & 15) == 0) {  // pull this warp's part of the C tile into L2 while the products run
#pragma unroll 8
                for (int c = 0; c < 64; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(at(c)));
            }
            for (int gi = 0; gi < S; ++gi) {
                const int g = sch.gorder[gi];
                const uint32_t slot = (uint32_t)g & (NSLOT - 1);
                const double sc = __hiloint2double((1023 - 7 * (g + 2)) << 20, 0);
                if (!mbar_wait(&bars->tfull[slot], (spar >> slot) & 1, 4, abortp, dbg)) goto done;
                spar ^= 1u << slot;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * BN + h * 64;
#pragma unroll
                for (int part = 0; part < 4; ++part) {
                    uint32_t v[16];
                    tmem_ld16(taddr + part * 16, v);
                    if (part == 3) {  // everything this warp needs has left TMEM: hand the slot back before the arithmetic
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->tempty[slot]);
                    }
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        // exact int32 -> double on the FP64 pipe: 2^52 + 2^31 + x as a bit pattern, minus the bias
                        const double m = __hiloint2double(0x43300000, (int)(v[c] ^ 0x80000000u)) - 4503601774854144.0;
                        acc[part * 16 + c] = fma(m, sc, acc[part * 16 + c]);
                    }
                }
            }
            const double rs = vw.rowscale[row];
            const double cs_lo = vw.rowscale[tj * BN + h * 64 + lane], cs_hi = vw.rowscale[tj * BN + h * 64 + 32 + lane];
            if (mode == 0) {
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 8) {
                    double o[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (valid) o[c] = *at(c0 + c);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const double w = rs * __shfl_sync(0xffffffffu, c0 < 32 ? cs_lo : cs_hi, (c0 + c) & 31);  // 2^(e_i + e_j)
                        if (valid) *at(c0 + c) = o[c] - acc[c0 + c] * w;
                    }
                }
            } else {  // mode 1: C = A A^T (overwrite), for checking the product alone
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    const double w = rs * __shfl_sync(0xffffffffu, c < 32 ? cs_lo : cs_hi, c & 31);
                    if (valid) *at(c) = acc[c] * w;
                }
            }
        }
    }
done:
    if (vw.info && *abortp && threadIdx.x == 0) *vw.info = -1;
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// 2 x 2 slice blocks by descending s0 + t0; products with s + t >= S are dropped (below 2^(-7S) of the row maxima)
static inline Schedule make_schedule(int S) {
    Schedule sc = {};
    sc.S = S;
    int first_step[16], last_step[16], last_prod[16];
    for (int g = 0; g < 16; ++g) first_step[g] = last_step[g] = last_prod[g] = -1;
    for (int sum = ((S - 1) / 2) * 2; sum >= 0; sum -= 2)
        for (int s0 = 0; s0 <= sum; s0 += 2) {
            const int t0 = sum - s0;
            Step st = {};
            bool use_a[2] = {false, false}, use_b[2] = {false, false};
            for (int ds = 0; ds < 2; ++ds)
                for (int dt = 0; dt < 2; ++dt) {
                    const int a = s0 + ds, b = t0 + dt;
                    if (a < S && b < S && a + b <= S - 1) use_a[ds] = use_b[dt] = true;
                }
            if (!use_a[0] && !use_a[1]) continue;
            int ia[2] = {-1, -1}, ib[2] = {-1, -1};
            for (int d = 0; d < 2; ++d) {
                if (use_a[d]) { ia[d] = st.na; st.sa[st.na++] = (uint8_t)(s0 + d); }
                if (use_b[d]) { ib[d] = st.nb; st.sb[st.nb++] = (uint8_t)(t0 + d); }
            }
            for (int ds = 0; ds < 2; ++ds)
                for (int dt = 0; dt < 2; ++dt) {
                    const int a = s0 + ds, b = t0 + dt, g = a + b;
                    if (!(a < S && b < S && g <= S - 1)) continue;
                    Prod pr = {(uint8_t)ia[ds], (uint8_t)ib[dt], (uint8_t)g, 0};
                    if (first_step[g] < 0) { first_step[g] = sc.nsteps; pr.flags |= 1; }
                    last_step[g] = sc.nsteps;
                    last_prod[g] = st.nprod;
                    st.prod[st.nprod++] = pr;
                }
            sc.step[sc.nsteps++] = st;
        }
    for (int g = 0; g < S; ++g) sc.step[last_step[g]].prod[last_prod[g]].flags |= 2;
    int n = 0;
    for (int st = 0; st < sc.nsteps; ++st) {
        Step &sp = sc.step[st];
        // issues: per row-side slice, the products with column slice 0 and 1 -- fused when their TMEM slots are adjacent
        for (int a = 0; a < sp.na; ++a) {
            const Prod *p0 = nullptr, *p1 = nullptr;
            for (int p = 0; p < sp.nprod; ++p)
                if (sp.prod[p].a == a) (sp.prod[p].b == 0 ? p0 : p1) = &sp.prod[p];
            auto flags_of = [](const Prod *p, int commit_bit) { return (uint8_t)((p->flags & 1) | ((p->flags & 2) ? commit_bit : 0)); };
            if (p0 && p1 && (p0->g & (NSLOT - 1)) != NSLOT - 1 && (p0->flags & 1) == (p1->flags & 1)) {
                Issue is = {(uint8_t)a, 0, (uint8_t)(p0->g & (NSLOT - 1)), (uint8_t)(2 | flags_of(p0, 4) | flags_of(p1, 8))};
                sp.iss[sp.nissue++] = is;
            } else {
                if (p0) sp.iss[sp.nissue++] = Issue{(uint8_t)a, 0, (uint8_t)(p0->g & (NSLOT - 1)), flags_of(p0, 4)};
                if (p1) sp.iss[sp.nissue++] = Issue{(uint8_t)a, 1, (uint8_t)(p1->g & (NSLOT - 1)), flags_of(p1, 4)};
            }
        }
        // completion order = the order in which the issuing thread commits the groups of a step: by ascending slot
        for (int slot = 0; slot < NSLOT; ++slot)
            for (int p = 0; p < sp.nprod; ++p)
                if ((sp.prod[p].flags & 2) && (sp.prod[p].g & (NSLOT - 1)) == slot) sc.gorder[n++] = sp.prod[p].g;
    }
    return sc;
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Launch on `stream` with at most max_ctas persistent CTAs (0: one per SM; negative: -max_ctas output blocks per CTA, i.e.
// short-lived CTAs that hand their SM back to higher-priority streams often).  slices: S x n_rows x K int8.
// Returns 0, or a negative code (-1 driver entry point, -2 shape, -4 tensor map, -5 launch).
struct Runtime {
    EncodeTiledFn encode = nullptr;
    int sms = 0;
};
// Driver entry point and SM count once per process; the kernels' shared-memory attribute once per device.  This is also
// what forces the kernels' module to load: a caller that keeps a spinning persistent kernel resident must call it BEFORE
// that kernel starts (the lazy loading of a first launch waits for the device).
template <int LAYOUT>
static inline Runtime *prepare() {
    static std::mutex mu;
    static Runtime rt;
    static unsigned long long devices_done = 0;
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return nullptr;
    if (!rt.encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return nullptr;
        cudaDeviceGetAttribute(&rt.sms, cudaDevAttrMultiProcessorCount, dev);
        rt.encode = (EncodeTiledFn)fn;
    }
    if (!(devices_done >> dev & 1)) {
        if (cudaFuncSetAttribute(int8_syrk_kernel<LAYOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess)
            return nullptr;
        devices_done |= 1ull << dev;
    }
    return &rt;
}

template <int LAYOUT>
static inline int launch(const int8_t *slices, int S, View vw, int max_ctas, cudaStream_t stream) {
    Runtime *rt = prepare<LAYOUT>();
    if (!rt) return -1;
    const EncodeTiledFn encode = rt->encode;
    const int sms = rt->sms;
    if (vw.n_rows % BM || vw.K % KB || vw.K > 16384 || S < 1 || S > 9 || vw.cb0 < 0 || vw.cb1 > vw.n_rows / BM) return -2;
    const int ntiles = blocks_in(vw.n_rows, vw.cb0, vw.cb1);
    if (ntiles <= 0) return 0;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)vw.K, (cuuint64_t)S * vw.n_rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)vw.K};
    const cuuint32_t box[2] = {KB, BM};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t *>(slices), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -4;
    int grid = max_ctas > 0 && max_ctas < sms ? max_ctas : sms;
    if (max_ctas < 0) grid = (ntiles - max_ctas - 1) / -max_ctas;
    if (ntiles < grid) grid = ntiles;
    const Schedule sch = make_schedule(S);
    vw.layout = LAYOUT;
    int8_syrk_kernel<LAYOUT><<<grid, I8_THREADS, SMEM_BYTES, stream>>>(tmap, sch, vw);
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) fprintf(stderr, "int8_syrk launch: %s\n", cudaGetErrorString(err));
    return err == cudaSuccess ? 0 : -5;
}

}  // namespace gpl_i8
