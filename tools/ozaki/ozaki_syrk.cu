// Microbenchmark only: an Ozaki-split FP64 rank-K update on the INT8 tensor path.
//
//     C(i,j) -= sum_k A[i][k] * A[j][k]        (lower-triangle 128x128 tiles; C column-major, A row-major)
//
// which is the trailing update of the large-n blocked Cholesky (gaplac_b200/csrc/big.cu: big_trail_kernel does the
// same update with DMMA.8x8x4 at ~85 % of the 37 TFLOP/s FP64 pipe).  sm_100a has no f64 kind on tcgen05, but it does
// have kind::i8 at ~4.5 POP/s.  The split:
//
//     A[i][k] = 2^e_i * sum_{s<S} q_s[i][k] * 2^(-7(s+1)),   q_s in [-127, 127]  (exact: truncation, not rounding)
//     A A^T   = 2^(e_i+e_j) * sum_g 2^(-7(g+2)) * sum_{s+t=g} q_s q_t^T          (groups g >= S dropped: < 2^(-7S) relative)
//
// Every q_s q_t^T is an exact INT8 x INT8 -> INT32 product on tcgen05 (accumulator in TMEM); one group g is summed in
// the same TMEM accumulator ((g+1) * K * 127^2 < 2^31 for K <= 16384), read back with tcgen05.ld, converted and
// accumulated in FP64 registers.  S = 8 keeps 56 bits below the row maximum; S(S+1)/2 = 36 INT8 products.
//
// Structure per CTA (persistent over output tiles): warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle),
// warp 1 = MMA issuer (one thread) and TMEM owner (four INT32 accumulators), warps 2..9 = epilogue (64 FP64
// accumulators per thread).
// Not part of libgaplac_b200.so; nothing in the product calls it.
//
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC \
//        -o tools/ozaki/libozaki.so tools/ozaki/ozaki_syrk.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

namespace {

constexpr int BM = 128;     // output tile rows = tcgen05 M
constexpr int BN = 128;     // output tile columns = tcgen05 N
constexpr int KB = 128;     // int8 elements (= bytes) of K per pipeline stage: one 128-byte swizzle row
constexpr int UK = 32;      // K of one tcgen05.mma.kind::i8
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS = 32 * (2 + EPI_WARPS);
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

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a lost arrival becomes a clean early exit with a breadcrumb (no trap, no hung GPU).  Returns false when
// this wait timed out or another role already gave up; every role then falls through to the common teardown.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int code, volatile int *abort_flag, volatile int *dbg) {
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
            else if (t - t0 > 200000000ull) {  // 0.2 s: a whole launch takes a few milliseconds
                *abort_flag = 1;
                if (atomicCAS((int *)dbg, 0, code) == 0) {
                    dbg[1] = (int)blockIdx.x;
                    dbg[2] = (int)parity;
                }
                return false;
            }
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// One deterministic leader lane of a converged warp.  The issuing warps keep their control flow warp-uniform and
// predicate only the tcgen05 / TMA instructions with this, so that ptxas keeps descriptors in uniform registers
// (a divergent "if (lane == 0)" region costs a vote + five R2UR + branch loop around every single UTCIMMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// shared-memory matrix descriptor: K-major operand, rows of 128 bytes, 128B swizzle, 8-row groups 1024 bytes apart
// low word: start address >> 4 in bits [0,14), leading byte offset (unused with a swizzled K-major operand) = 1 in
// bits [16,30); high word: stride byte offset 1024 >> 4 between 8-row groups, descriptor version 1 (sm_100), SWIZZLE_128B
constexpr uint64_t DESC_HI = ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tile_of(int t, int &ti, int &tj) {
    int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(r + 1) * (r + 2) / 2 <= t) ++r;
    while ((long)r * (r + 1) / 2 > t) --r;
    ti = r;
    tj = t - r * (r + 1) / 2;
}

struct Barriers {
    uint64_t full[STAGES], empty[STAGES], tfull[NSLOT], tempty[NSLOT];
    uint32_t tmem_base;
    int abort_flag;
};

// 10 warps are allocated as 12 (granularity 4), so the register file allows 65536 / 384 = 168 registers per thread
__global__ void __launch_bounds__(NTHREADS, 1)
ozaki_syrk_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Schedule sch,
                  const double *__restrict__ rowscale, double *__restrict__ C, long ldc, int n_rows, int K, int mode, int *dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    Barriers *bars = reinterpret_cast<Barriers *>(smem + (size_t)STAGES * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = n_rows / BM, ntiles = nt * (nt + 1) / 2, nkb = K / KB;
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
            tile_of(tile, ti, tj);
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
            tile_of(tile, ti, tj);
            double acc[64];
#pragma unroll
            for (int c = 0; c < 64; ++c) acc[c] = 0.0;
            // C is addressed column-major (element (i, j) at C[j * ldc + i]): the TMEM read leaves one row per lane, so a
            // warp touches 256 contiguous bytes per column -- the same direction the product's 64 x 64 tiles are stored in
            const int row = ti * BM + q * 32 + lane;
            double *cp = C + (long)(tj * BN + h * 64) * ldc + row;
            if (mode == 0 && (lane & 15) == 0) {  // pull this warp's part of the C tile into L2 while the products run
#pragma unroll 8
                for (int c = 0; c < 64; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + (long)c * ldc));
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
            const double rs = rowscale[row];
            const double cs_lo = rowscale[tj * BN + h * 64 + lane], cs_hi = rowscale[tj * BN + h * 64 + 32 + lane];
            if (mode == 0) {
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 8) {
                    double o[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) o[c] = cp[(long)(c0 + c) * ldc];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const double w = rs * __shfl_sync(0xffffffffu, c0 < 32 ? cs_lo : cs_hi, (c0 + c) & 31);  // 2^(e_i + e_j)
                        cp[(long)(c0 + c) * ldc] = o[c] - acc[c0 + c] * w;
                    }
                }
            } else {  // mode 1: C = A A^T (overwrite), for checking the product alone
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    const double w = rs * __shfl_sync(0xffffffffu, c < 32 ? cs_lo : cs_hi, c & 31);
                    cp[(long)c * ldc] = acc[c] * w;
                }
            }
        }
    }
done:
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// One CTA per row: exponent of the row maximum, then S exact 7-bit truncation slices of every element.
__global__ void __launch_bounds__(128)
ozaki_split_kernel(const double *__restrict__ A, long lda, int n_rows, int K, int S, int8_t *__restrict__ out,
                   double *__restrict__ rowscale) {
    __shared__ double red[4];
    const int row = blockIdx.x, tid = threadIdx.x;
    const double *a = A + (long)row * lda;
    double mx = 0.0;
    for (int k = tid; k < K; k += 128) mx = fmax(mx, fabs(a[k]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
    const int e = mx > 0.0 ? ilogb(mx) + 1 : 0;  // |a| < 2^e
    if (tid == 0) rowscale[row] = ldexp(1.0, e);
    const double inv = ldexp(1.0, -e);
    for (int k4 = tid * 4; k4 < K; k4 += 512) {
        double r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = a[k4 + i] * inv;
        for (int s = 0; s < S; ++s) {
            char4 qv;
            int qi[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                r[i] *= 128.0;
                qi[i] = (int)r[i];  // truncation: the remainder keeps its sign and stays below 1 in magnitude
                r[i] -= (double)qi[i];
            }
            qv.x = (signed char)qi[0];
            qv.y = (signed char)qi[1];
            qv.z = (signed char)qi[2];
            qv.w = (signed char)qi[3];
            *reinterpret_cast<char4 *>(out + ((long)s * n_rows + row) * K + k4) = qv;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int *g_dbg_host = nullptr, *g_dbg_dev = nullptr;
int g_sms = 0;

int init_once() {
    if (g_encode) return 0;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
    if (cudaHostAlloc((void **)&g_dbg_host, 64, cudaHostAllocMapped) != cudaSuccess) return -2;
    for (int i = 0; i < 16; ++i) g_dbg_host[i] = 0;
    if (cudaHostGetDevicePointer((void **)&g_dbg_dev, g_dbg_host, 0) != cudaSuccess) return -3;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaFuncSetAttribute(ozaki_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess)
        return -4;
    g_encode = (EncodeTiledFn)fn;
    return 0;
}

// 2 x 2 slice blocks by descending s0 + t0; products with s + t >= S are dropped (below 2^(-7S) of the row maxima)
Schedule make_schedule(int S) {
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

}  // namespace

extern "C" {

// bytes of workspace for the slices (S * n_rows * K int8) followed by the row scales (n_rows doubles)
long ozaki_ws_bytes(int n_rows, int K, int S) { return (long)S * n_rows * K + (long)n_rows * 8 + 256; }

// 4 ints: code of the first barrier wait that timed out (1 producer/empty, 2 mma/tempty, 3 mma/full, 4 epilogue/tfull),
// CTA, parity; all zero when every launch so far ran to completion.  Call after synchronising.
void ozaki_last_debug(int *out) {
    for (int i = 0; i < 4; ++i) out[i] = g_dbg_host ? g_dbg_host[i] : 0;
}

// the product schedule for S slices as text (host only; lets the table be checked without a GPU)
int ozaki_schedule_text(int S, char *buf, int cap) {
    const Schedule sc = make_schedule(S);
    int n = snprintf(buf, cap, "S=%d steps=%d order=", S, sc.nsteps);
    for (int i = 0; i < S; ++i) n += snprintf(buf + n, cap - n, "%d ", sc.gorder[i]);
    for (int st = 0; st < sc.nsteps; ++st) {
        const Step &sp = sc.step[st];
        n += snprintf(buf + n, cap - n, "\n step %d: rows", st);
        for (int i = 0; i < sp.na; ++i) n += snprintf(buf + n, cap - n, " %d", sp.sa[i]);
        n += snprintf(buf + n, cap - n, " | cols");
        for (int i = 0; i < sp.nb; ++i) n += snprintf(buf + n, cap - n, " %d", sp.sb[i]);
        n += snprintf(buf + n, cap - n, " |");
        for (int p = 0; p < sp.nprod; ++p)
            n += snprintf(buf + n, cap - n, " (%d,%d)->g%d%s%s", sp.sa[sp.prod[p].a], sp.sb[sp.prod[p].b], sp.prod[p].g,
                          (sp.prod[p].flags & 1) ? "F" : "", (sp.prod[p].flags & 2) ? "L" : "");
        n += snprintf(buf + n, cap - n, " | issue");
        for (int p = 0; p < sp.nissue; ++p)
            n += snprintf(buf + n, cap - n, " %d:%d%s@%d", sp.sa[sp.iss[p].a], sp.sb[sp.iss[p].b], (sp.iss[p].flags & 2) ? "w" : "", sp.iss[p].slot);
    }
    return n;
}

// split only (what a panel kernel would emit as it stores L): A row-major n_rows x K, leading dimension lda
int ozaki_split(const double *A, long lda, int n_rows, int K, int S, void *ws, void *stream) {
    if (init_once()) return -1;
    if (n_rows % BM || K % KB || S < 1 || S > 9) return -2;
    int8_t *slices = (int8_t *)ws;
    double *rowscale = (double *)((char *)ws + (((long)S * n_rows * K + 255) & ~255L));
    ozaki_split_kernel<<<n_rows, 128, 0, (cudaStream_t)stream>>>(A, lda, n_rows, K, S, slices, rowscale);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

// the product on slices already in ws: mode 0  C -= A A^T,  mode 1  C = A A^T  (lower-triangle 128x128 tiles)
int ozaki_update(int n_rows, int K, int S, void *ws, double *C, long ldc, int mode, void *stream) {
    if (init_once()) return -1;
    if (n_rows % BM || K % KB || S < 1 || S > 9) return -2;
    int8_t *slices = (int8_t *)ws;
    double *rowscale = (double *)((char *)ws + (((long)S * n_rows * K + 255) & ~255L));
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)S * n_rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)K};
    const cuuint32_t box[2] = {KB, BM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, slices, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -4;
    const int nt = n_rows / BM, ntiles = nt * (nt + 1) / 2;
    const int grid = ntiles < g_sms ? ntiles : g_sms;
    const Schedule sch = make_schedule(S);
    ozaki_syrk_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tmap, sch, rowscale, C, ldc, n_rows, K, mode, g_dbg_dev);
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) fprintf(stderr, "ozaki_update launch: %s\n", cudaGetErrorString(err));
    return err == cudaSuccess ? 0 : -5;
}

}  // extern "C"
