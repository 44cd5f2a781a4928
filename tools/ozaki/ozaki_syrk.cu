// Microbenchmark only: an Ozaki-split FP64 rank-K update on the INT8 tensor path.
//
//     C[i][j] -= sum_k A[i][k] * A[j][k]        (lower-triangle 128x128 tiles, row-major C)
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
// warp 1 = MMA issuer (one thread) and TMEM owner, warps 2..9 = epilogue (64 FP64 accumulators per thread).
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
constexpr int STAGES = 6;
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS = 32 * (2 + EPI_WARPS);
constexpr uint32_t TILE_BYTES = BM * KB;  // 16 KiB per operand per stage
constexpr uint32_t STAGE_BYTES = 2 * TILE_BYTES;
constexpr uint32_t TMEM_COLS = 2 * BN;  // two INT32 accumulators (ping-pong between MMA and epilogue)
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + 256;

// instruction descriptor for kind::i8, dense, S32 accumulator, A and B signed 8-bit, both K-major, M = 128, N = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
                 : "memory");
}
// shared-memory matrix descriptor: K-major operand, rows of 128 bytes, 128B swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(const void *tile) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(tile) & 0x3FFFFu) >> 4);  // start address
    d |= (uint64_t)1 << 16;                              // leading byte offset (unused with a swizzled K-major operand)
    d |= (uint64_t)(1024u >> 4) << 32;                   // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
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
    uint64_t full[STAGES], empty[STAGES], tfull[2], tempty[2];
    uint32_t tmem_base;
    int abort_flag;
};

__global__ void __launch_bounds__(NTHREADS, 1)
ozaki_syrk_kernel(const __grid_constant__ CUtensorMap tmap, const double *__restrict__ rowscale, double *__restrict__ C,
                  long ldc, int n_rows, int K, int S, int mode, int *dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    Barriers *bars = reinterpret_cast<Barriers *>(smem + (size_t)STAGES * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = n_rows / BM, ntiles = nt * (nt + 1) / 2, nkb = K / KB;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&bars->full[i], 1);
            mbar_init(&bars->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
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
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                int ti, tj;
                tile_of(tile, ti, tj);
                for (int g = 0; g < S; ++g)
                    for (int s = 0; s <= g; ++s) {
                        const int t = g - s;
                        for (int kb = 0; kb < nkb; ++kb) {
                            if (!mbar_wait(&bars->empty[stage], phase ^ 1, 1, abortp, dbg)) goto done;
                            uint8_t *sa = smem + (size_t)stage * STAGE_BYTES;
                            mbar_arrive_expect_tx(&bars->full[stage], STAGE_BYTES);
                            tma_load_2d(sa, &tmap, kb * KB, s * n_rows + ti * BM, &bars->full[stage]);
                            tma_load_2d(sa + TILE_BYTES, &tmap, kb * KB, t * n_rows + tj * BN, &bars->full[stage]);
                            if (++stage == STAGES) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, gc = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int g = 0; g < S; ++g, ++gc) {
                    const uint32_t buf = gc & 1, bphase = (gc >> 1) & 1;
                    if (!mbar_wait(&bars->tempty[buf], bphase ^ 1, 2, abortp, dbg)) goto done;
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + buf * BN;
                    uint32_t accumulate = 0;
                    for (int s = 0; s <= g; ++s)
                        for (int kb = 0; kb < nkb; ++kb) {
                            if (!mbar_wait(&bars->full[stage], phase, 3, abortp, dbg)) goto done;
                            tc_fence_after();
                            const uint8_t *sa = smem + (size_t)stage * STAGE_BYTES;
                            const uint64_t ad = make_desc(sa), bd = make_desc(sa + TILE_BYTES);
#pragma unroll
                            for (int k4 = 0; k4 < KB / UK; ++k4) {
                                tc_mma_i8(tmem_d, ad + (uint64_t)(k4 * UK / 16), bd + (uint64_t)(k4 * UK / 16), accumulate);
                                accumulate = 1;
                            }
                            tc_commit(&bars->empty[stage]);  // frees the stage when these MMAs have read it
                            if (++stage == STAGES) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    tc_commit(&bars->tfull[buf]);  // group g complete in TMEM
                }
            }
        }
    } else {
        const int q = warp & 3;         // TMEM lane quarter this warp may read
        const int h = (warp - 2) >> 2;  // column half of the tile
        uint32_t gc = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int ti, tj;
            tile_of(tile, ti, tj);
            double acc[64];
#pragma unroll
            for (int c = 0; c < 64; ++c) acc[c] = 0.0;
            for (int g = 0; g < S; ++g, ++gc) {
                const uint32_t buf = gc & 1, bphase = (gc >> 1) & 1;
                const double sc = __hiloint2double((1023 - 7 * (g + 2)) << 20, 0);
                if (!mbar_wait(&bars->tfull[buf], bphase, 4, abortp, dbg)) goto done;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + h * 64;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld32(taddr + half * 32, v);
#pragma unroll
                    for (int c = 0; c < 32; ++c) acc[half * 32 + c] = fma(__int2double_rn((int)v[c]), sc, acc[half * 32 + c]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->tempty[buf]);
            }
            const int row = ti * BM + q * 32 + lane;
            const double rs = rowscale[row];
            const double *cs = rowscale + tj * BN + h * 64;
            double *crow = C + (long)row * ldc + tj * BN + h * 64;
            if (mode == 0) {
#pragma unroll
                for (int c = 0; c < 64; c += 2) {
                    double2 o = *reinterpret_cast<double2 *>(crow + c);
                    o.x -= acc[c] * rs * cs[c];
                    o.y -= acc[c + 1] * rs * cs[c + 1];
                    *reinterpret_cast<double2 *>(crow + c) = o;
                }
            } else {  // mode 1: C = A A^T (overwrite), for checking the product alone
#pragma unroll
                for (int c = 0; c < 64; c += 2)
                    *reinterpret_cast<double2 *>(crow + c) = make_double2(acc[c] * rs * cs[c], acc[c + 1] * rs * cs[c + 1]);
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

}  // namespace

extern "C" {

// bytes of workspace for the slices (S * n_rows * K int8) followed by the row scales (n_rows doubles)
long ozaki_ws_bytes(int n_rows, int K, int S) { return (long)S * n_rows * K + (long)n_rows * 8 + 256; }

// 4 ints: code of the first barrier wait that timed out (1 producer/empty, 2 mma/tempty, 3 mma/full, 4 epilogue/tfull),
// CTA, parity; all zero when every launch so far ran to completion.  Call after synchronising.
void ozaki_last_debug(int *out) {
    for (int i = 0; i < 4; ++i) out[i] = g_dbg_host ? g_dbg_host[i] : 0;
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
    ozaki_syrk_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tmap, rowscale, C, ldc, n_rows, K, S, mode, g_dbg_dev);
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // extern "C"
