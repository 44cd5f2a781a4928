// FP64 roofline probe for B200 (sm_100a): the denominators DESIGN.md / bench.py quote.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/fp64_peak tools/fp64_peak.cu -lcublas -lcusolver
//   build/fp64_peak [--lib]      (prints one JSON object; --lib adds the cuBLAS / cuSOLVER yardsticks)
//
// Measures (CUDA events, after warm-up, best of several repetitions):
//   dfma        vector DFMA issue rate, 8 independent chains per thread
//   dmma_*      FP64 tensor-core path: mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 .f64
//   exp         FP64 exp() evaluations per second (the covariance-build inner op)
//   dgemm       cuBLAS DGEMM 8192^3            (yardstick)
//   potrf       cuSOLVER Dpotrf n=8192          (yardstick for config 5)
//   potrf_b     cuSOLVER DpotrfBatched 512 x 512, batch 1024 (yardstick for the headline config)
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) dfma_kernel(double *out, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) exp_kernel(double *out, double a) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = -(threadIdx.x * 1e-3 + i) * a;
    double s = 0;
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s += exp(x[i]);
            x[i] -= 1e-4;
        }
    }
    if (s == 12345.678) out[0] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs
__global__ void __launch_bounds__(256) dmma884_kernel(double *out, double a, double b) {
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

// m16n8k4: A 2 regs, B 1 reg, C 4 regs
__global__ void __launch_bounds__(256) dmma1684_kernel(double *out, double a, double b) {
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile(
                "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                : "d"(a), "d"(b), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 12345.678) out[0] = s;
}

// m16n8k8: A 4 regs, B 2 regs, C 4 regs
__global__ void __launch_bounds__(256) dmma1688_kernel(double *out, double a, double b) {
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile(
                "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                "{%0,%1,%2,%3};\n"
                : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 12345.678) out[0] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C 4 regs
__global__ void __launch_bounds__(256) dmma16816_kernel(double *out, double a, double b) {
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
double best_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    bool lib = argc > 1 && !strcmp(argv[1], "--lib");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *out;
    CK(cudaMalloc(&out, 64));
    const int grid = sms * 8;
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);

    // sustained: run each probe long enough (>= ~0.2 s total) that clocks settle; report best and a long-run figure
    {
        double ms = best_ms([&] { dfma_kernel<<<grid, 256>>>(out, 1.0000001, 1e-9); });
        double flops = 2.0 * 8 * ITERS * 256.0 * grid;
        printf(", \"dfma_tflops\": %.3f", flops / ms * 1e-9);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        const int reps = 400;
        for (int i = 0; i < reps; ++i) dfma_kernel<<<grid, 256>>>(out, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float t;
        CK(cudaEventElapsedTime(&t, e0, e1));
        printf(", \"dfma_tflops_sustained\": %.3f, \"dfma_sustained_window_ms\": %.1f", flops * reps / t * 1e-9, t);
    }
    {
        double ms = best_ms([&] { dmma884_kernel<<<grid, 256>>>(out, 1.0000001, 1e-9); });
        double flops = 2.0 * 8 * 8 * 4 * 4 * ITERS * 8.0 * grid;  // per warp: 4 mma of 256 FMA
        printf(", \"dmma_m8n8k4_tflops\": %.3f", flops / ms * 1e-9);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        const int reps = 400;
        for (int i = 0; i < reps; ++i) dmma884_kernel<<<grid, 256>>>(out, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float t;
        CK(cudaEventElapsedTime(&t, e0, e1));
        printf(", \"dmma_m8n8k4_tflops_sustained\": %.3f", flops * reps / t * 1e-9);
    }
    {
        double ms = best_ms([&] { dmma1684_kernel<<<grid, 256>>>(out, 1.0000001, 1e-9); });
        double flops = 2.0 * 16 * 8 * 4 * 4 * ITERS * 8.0 * grid;
        printf(", \"dmma_m16n8k4_tflops\": %.3f", flops / ms * 1e-9);
    }
    {
        double ms = best_ms([&] { dmma1688_kernel<<<grid, 256>>>(out, 1.0000001, 1e-9); });
        double flops = 2.0 * 16 * 8 * 8 * 4 * ITERS * 8.0 * grid;
        printf(", \"dmma_m16n8k8_tflops\": %.3f", flops / ms * 1e-9);
    }
    {
        double ms = best_ms([&] { dmma16816_kernel<<<grid, 256>>>(out, 1.0000001, 1e-9); });
        double flops = 2.0 * 16 * 8 * 16 * 4 * ITERS * 8.0 * grid;
        printf(", \"dmma_m16n8k16_tflops\": %.3f", flops / ms * 1e-9);
    }
    {
        double ms = best_ms([&] { exp_kernel<<<grid, 256>>>(out, 1.0); });
        double evals = 4.0 * (ITERS / 8) * 256.0 * grid;
        printf(", \"exp_gevals_per_s\": %.2f", evals / ms * 1e-6);
    }
    if (lib) {
        const int n = 8192;
        double *A, *B, *C;
        CK(cudaMalloc(&A, (size_t)n * n * 8));
        CK(cudaMalloc(&B, (size_t)n * n * 8));
        CK(cudaMalloc(&C, (size_t)n * n * 8));
        std::vector<double> h((size_t)n * n);
        for (size_t i = 0; i < h.size(); ++i) h[i] = (double)((i * 2654435761u) % 1000) * 1e-3;
        CK(cudaMemcpy(A, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
        cublasHandle_t hb;
        cublasCreate(&hb);
        const double one = 1.0, zero = 0.0;
        double ms = best_ms([&] { cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n); }, 5);
        printf(", \"cublas_dgemm_8192_tflops\": %.3f", 2.0 * n * n * (double)n / ms * 1e-9);
        // SPD matrix: C = A A' / n + n I  -> potrf
        cusolverDnHandle_t hs;
        cusolverDnCreate(&hs);
        int lwork = 0;
        cusolverDnDpotrf_bufferSize(hs, CUBLAS_FILL_MODE_LOWER, n, C, n, &lwork);
        double *work;
        int *info;
        CK(cudaMalloc(&work, (size_t)lwork * 8));
        CK(cudaMalloc(&info, 4096 * sizeof(int)));
        auto make_spd = [&] {
            cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, A, n, &zero, C, n);
            // add a large diagonal through a rank-0 trick: scale + axpy on the diagonal with stride n+1
            const double big = 1e6;
            std::vector<double> ones(1, big);
            double *dbig;
            cudaMalloc(&dbig, 8);
            cudaMemcpy(dbig, ones.data(), 8, cudaMemcpyHostToDevice);
            cublasDaxpy(hb, n, &one, dbig, 0, C, n + 1);
            cudaFree(dbig);
        };
        double best = 1e30;
        for (int r = 0; r < 3; ++r) {
            make_spd();
            CK(cudaDeviceSynchronize());
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0);
            cusolverDnDpotrf(hs, CUBLAS_FILL_MODE_LOWER, n, C, n, work, lwork, info);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float t;
            cudaEventElapsedTime(&t, e0, e1);
            if (t < best) best = t;
        }
        int hinfo = -1;
        cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
        printf(", \"cusolver_potrf_8192_ms\": %.3f, \"cusolver_potrf_8192_tflops\": %.3f, \"cusolver_potrf_info\": %d", best,
               (double)n * n * n / 3.0 / best * 1e-9, hinfo);
        // batched 512 x 512, batch 1024: all matrices = leading block of C (SPD), freshly copied each repetition
        const int nb = 512, batch = 1024;
        double *M;
        CK(cudaMalloc(&M, (size_t)batch * nb * nb * 8));
        std::vector<double *> hp(batch);
        for (int b = 0; b < batch; ++b) hp[b] = M + (size_t)b * nb * nb;
        double **dp;
        CK(cudaMalloc(&dp, batch * sizeof(double *)));
        CK(cudaMemcpy(dp, hp.data(), batch * sizeof(double *), cudaMemcpyHostToDevice));
        make_spd();
        double bestb = 1e30;
        for (int r = 0; r < 3; ++r) {
            for (int b = 0; b < batch; ++b)
                cudaMemcpy2DAsync(hp[b], nb * 8, C, (size_t)n * 8, nb * 8, nb, cudaMemcpyDeviceToDevice);
            CK(cudaDeviceSynchronize());
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0);
            cusolverDnDpotrfBatched(hs, CUBLAS_FILL_MODE_LOWER, nb, dp, nb, info, batch);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float t;
            cudaEventElapsedTime(&t, e0, e1);
            if (t < bestb) bestb = t;
        }
        printf(", \"cusolver_potrf_batched_512x1024_ms\": %.3f, \"cusolver_potrf_batched_512_per_s\": %.1f", bestb,
               batch / bestb * 1e3);
    }
    printf("}\n");
    return 0;
}
