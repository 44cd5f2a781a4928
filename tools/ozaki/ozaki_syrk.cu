// Microbenchmark: the Ozaki-split FP64 rank-K update on the INT8 tensor path, on dense matrices.
//
//     C(i,j) -= sum_k A[i][k] * A[j][k]        (lower-triangle 128x128 tiles; C column-major, A row-major)
//
// The kernel is gaplac_b200/csrc/int8_syrk.cuh (tcgen05.mma.kind::i8 + TMEM + TMA); this file adds
// the split of a dense row-major A into slices and the C entry points tools/ozaki/ozaki_bench.py drives.
//
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC \
//        -I gaplac_b200/csrc -o tools/ozaki/libozaki.so tools/ozaki/ozaki_syrk.cu
#include "int8_syrk.cuh"

namespace {

using namespace gpl_i8;

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

int *g_dbg_host = nullptr, *g_dbg_dev = nullptr;

int init_once() {
    if (g_dbg_host) return 0;
    if (cudaHostAlloc((void **)&g_dbg_host, 64, cudaHostAllocMapped) != cudaSuccess) return -2;
    for (int i = 0; i < 16; ++i) g_dbg_host[i] = 0;
    if (cudaHostGetDevicePointer((void **)&g_dbg_dev, g_dbg_host, 0) != cudaSuccess) return -3;
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
    View vw = {};
    vw.rowscale = (const double *)((char *)ws + (((long)S * n_rows * K + 255) & ~255L));
    vw.C = C;
    vw.ldc = ldc;
    vw.layout = 0;
    vw.mode = mode;
    vw.n_rows = n_rows;
    vw.K = K;
    vw.cb0 = 0;
    vw.cb1 = n_rows / BM;
    vw.dbg = g_dbg_dev;
    return launch<0>((const int8_t *)ws, S, vw, 0, (cudaStream_t)stream);
}

}  // extern "C"
