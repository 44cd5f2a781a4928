// Microbenchmark of the tile update loop in isolation: how close does  acc -= A_chunk * B_chunk'  (DMMA, tile.cuh)
// get to the FP64 pipe peak as a function of CTAs per SM, stage width and prefetch depth?
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/mma_bench tools/mma_bench.cu
//   build/mma_bench
//
// Modes:  0 = operands resident in shared memory, no loads, no barriers      (DMMA issue ceiling at this occupancy)
//         1 = cp.async double buffering, one __syncthreads per stage          (the structure of lml_batched_kernel)
//         2 = cp.async 4-stage ring, wait_group<2>, one __syncthreads per stage
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../gaplac_b200/csrc/tile.cuh"

using namespace gpl;

#define CK(x)                                                                                       \
    do {                                                                                            \
        cudaError_t e = (x);                                                                        \
        if (e != cudaSuccess) {                                                                     \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                                \
        }                                                                                           \
    } while (0)

template <int MODE, int KCOLS, int STAGES>
__global__ void __launch_bounds__(NTHREADS) loop_kernel(const double *__restrict__ ws, long long ws_stride, int steps,
                                                        int tiles_per_cta, double *out) {
    extern __shared__ __align__(16) double sm[];
    constexpr int CH = KCOLS * TS;  // doubles per operand chunk
    const int tid = threadIdx.x;
    const TMap tm = thread_map(tid);
    const double *base = ws + (size_t)blockIdx.x * ws_stride;
    double acc[2][8];
    acc_zero(acc);
    if (MODE == 0) {
        for (int e = tid; e < 2 * CH; e += NTHREADS) sm[e] = 1e-3 * (e % 97);
        __syncthreads();
        for (int s = 0; s < steps; ++s) tile_mma<true>(acc, sm, sm + CH, tm, 0, KCOLS);
    } else {
        // chunk q of the row operand at base + q*CH (cyclic over tiles_per_cta tiles), column operand offset by half
        const long long span = (long long)tiles_per_cta * TILE_ELEMS;
        auto issue = [&](int q) {
            const int st = q % STAGES;
            const long long off = ((long long)q * CH) % (span / 2);
            block_load_async<CH * 8>(sm + (size_t)st * 2 * CH, base + off, tid);
            block_load_async<CH * 8>(sm + (size_t)st * 2 * CH + CH, base + span / 2 + off, tid);
            cp_async_commit();
        };
        for (int q = 0; q < STAGES - 1 && q < steps; ++q) issue(q);
        for (int q = 0; q < steps; ++q) {
            if (STAGES == 2) cp_async_wait<0>();
            else cp_async_wait<STAGES - 2>();
            __syncthreads();
            if (q + STAGES - 1 < steps) issue(q + STAGES - 1);
            else cp_async_commit();  // keep the group count uniform
            const double *a = sm + (size_t)(q % STAGES) * 2 * CH;
            tile_mma<true>(acc, a, a + CH, tm, 0, KCOLS);
        }
    }
    double s = 0;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) s += acc[mb][cc];
    if (s == 12345.678) out[0] = s;
}

template <int MODE, int KCOLS, int STAGES>
void run(const char *name, int ctas_per_sm, int sms, const double *ws, long long ws_stride, int tiles_per_cta, double *out) {
    const size_t smem = (size_t)(MODE == 0 ? 1 : STAGES) * 2 * KCOLS * TS * 8;
    auto k = loop_kernel<MODE, KCOLS, STAGES>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NTHREADS, smem));
    if (occ < ctas_per_sm) {
        printf("%-34s ctas/SM %d: does not fit (occ %d)\n", name, ctas_per_sm, occ);
        return;
    }
    const int grid = sms * ctas_per_sm;
    const int steps = 4096 * 16 / KCOLS;  // same number of MMAs for every stage width
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) k<<<grid, NTHREADS, smem>>>(ws, ws_stride, steps, tiles_per_cta, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k<<<grid, NTHREADS, smem>>>(ws, ws_stride, steps, tiles_per_cta, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64 * 64 * (double)KCOLS * steps * grid;
    printf("%-34s ctas/SM %d: %7.2f TFLOP/s  (%.2f ms)\n", name, ctas_per_sm, flops / ms * 1e-9, ms);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int tiles_per_cta = 36;  // one n = 512 factor: 1.15 MB per CTA, like the real workspace
    const long long ws_stride = (long long)tiles_per_cta * TILE_ELEMS;
    const int max_ctas = sms * 6;
    double *ws, *out;
    CK(cudaMalloc(&ws, (size_t)max_ctas * ws_stride * 8));
    CK(cudaMemset(ws, 0, (size_t)max_ctas * ws_stride * 8));
    CK(cudaMalloc(&out, 64));
    for (int c = 1; c <= 6; ++c) run<0, 16, 1>("resident operands, k=16 per call", c, sms, ws, ws_stride, tiles_per_cta, out);
    for (int c = 1; c <= 4; ++c) run<0, 32, 1>("resident operands, k=32 per call", c, sms, ws, ws_stride, tiles_per_cta, out);
    for (int c = 1; c <= 3; ++c) run<1, 32, 2>("cp.async 2 stages of 32 cols", c, sms, ws, ws_stride, tiles_per_cta, out);
    for (int c = 1; c <= 6; ++c) run<1, 16, 2>("cp.async 2 stages of 16 cols", c, sms, ws, ws_stride, tiles_per_cta, out);
    for (int c = 1; c <= 3; ++c) run<2, 16, 4>("cp.async 4 stages of 16 cols", c, sms, ws, ws_stride, tiles_per_cta, out);
    for (int c = 1; c <= 6; ++c) run<2, 8, 4>("cp.async 4 stages of 8 cols", c, sms, ws, ws_stride, tiles_per_cta, out);
    return 0;
}
