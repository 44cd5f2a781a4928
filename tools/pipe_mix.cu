// Does a dependent FP64 chain slow down when other warps on the SM stream DMMAs, and do DMMA and DFMA add up?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/pipe_mix tools/pipe_mix.cu && build/pipe_mix
// Kernel: CTA of 32*(1+NMMA) threads on every SM; warp 0 runs a dependent DFMA chain (or an 8-way independent one) and
// records clocks per op; the other warps stream independent DMMAs until warp 0 is done.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s\n", cudaGetErrorString(e)); exit(1);} } while (0)

template <int ILP>
__global__ void mix_kernel(int n_ops, double a, double b, long long *clk_out, double *sink, unsigned long long *mma_count) {
    __shared__ volatile int done;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        double x[ILP];
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
        const long long t0 = clock64();
        for (int it = 0; it < n_ops; ++it) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
        }
        const long long t1 = clock64();
        double s = 0;
#pragma unroll
        for (int i = 0; i < ILP; ++i) s += x[i];
        if (threadIdx.x == 0) {
            clk_out[blockIdx.x] = t1 - t0;
            done = 1;
        }
        if (s == 1234.5) sink[0] = s;
    } else {
        double c[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = threadIdx.x;
        unsigned long long cnt = 0;
        while (!done) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            cnt += 32;
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
        if (s == 1234.5) sink[0] = s;
        if ((threadIdx.x & 31) == 0) atomicAdd(mma_count, cnt);
    }
}

template <int ILP>
void run(int nmma_warps, int sms) {
    long long *clk;
    double *sink;
    unsigned long long *cnt;
    CK(cudaMalloc(&clk, sms * 8));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMalloc(&cnt, 8));
    CK(cudaMemset(cnt, 0, 8));
    const int n_ops = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mix_kernel<ILP><<<sms, 32 * (1 + nmma_warps)>>>(n_ops, 1.0000001, 1e-9, clk, sink, cnt);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(cnt, 0, 8));
    cudaEventRecord(e0);
    mix_kernel<ILP><<<sms, 32 * (1 + nmma_warps)>>>(n_ops, 1.0000001, 1e-9, clk, sink, cnt);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[256];
    CK(cudaMemcpy(h, clk, sms * 8, cudaMemcpyDeviceToHost));
    unsigned long long hc;
    CK(cudaMemcpy(&hc, cnt, 8, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    const double mma_tf = 2.0 * 256 * (double)hc / (ms * 1e-3) * 1e-12;
    const double dfma_tf = 2.0 * 32 * ILP * (double)n_ops * sms / (ms * 1e-3) * 1e-12;
    printf("ILP %d, %2d DMMA warps/SM: %7.1f clk per DFMA step (%.2f clk per op); DMMA %.2f TF + DFMA %.3f TF\n", ILP, nmma_warps,
           avg / n_ops, avg / n_ops / ILP, mma_tf, dfma_tf);
    cudaFree(clk); cudaFree(sink); cudaFree(cnt);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    for (int nm : {0, 3, 4, 7, 8, 15}) run<1>(nm, sms);
    for (int nm : {0, 3, 7, 15}) run<8>(nm, sms);
    return 0;
}
