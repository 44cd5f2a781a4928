// Trailing update of the large-n Cholesky on the INT8 tensor path (option "trail_int8").
//
// The right-looking factorisation of big.cu applies every finished panel to the whole trailing matrix with DMMA
// (big_trail_kernel, K = 256 per launch).  With this option the columns beyond the current block of I8_BLOCK tile columns
// are left alone until the block is complete; then its part of L (rows below the block, K = 64 * I8_BLOCK columns) is cut
// into 7-bit slices (i8_split_tiles_kernel) and applied in one pass by the tcgen05 kernel of int8_syrk.cuh, which reads and
// writes the library's 64 x 64 tiles directly.  Slices = 8 keeps 56 bits below each row's maximum (FP64-equivalent).
#include "int8_syrk.cuh"
#include "kernels.h"
#include "tile.cuh"

namespace gpl {

// Slices of the rows of L below tile row t0, tile columns [c0, c0 + kt): S slices, K-major (out[(s * rows_pad + row) * K + k],
// K = 64 * kt), and the row scales 2^e (e: |L[row][k]| < 2^e over the whole block).  Two launches of one CTA per TILE (so a
// block of 16 tile columns under 112 tile rows is 1792 CTAs, not 112): the row maxima first (atomic max on the bit patterns
// of |x|: non-negative doubles order like their bits), then the split.  Tile rows past nt (the region is padded to a
// multiple of 128 rows) get zero slices and scale 1.
__global__ void __launch_bounds__(NTHREADS)
i8_rowmax_kernel(const double *__restrict__ tiles, int nt, int t0, int c0, int kt, unsigned long long *__restrict__ maxbits) {
    __shared__ double red[TS];
    const int tid = threadIdx.x, tr = blockIdx.x / kt, tk = blockIdx.x - tr * kt, ti = t0 + tr;
    if (ti >= nt) return;
    const double *src = tiles + tri_index(ti, c0 + tk) * TILE_ELEMS;
    const int r = tid & (TS - 1), half = tid >> 6;
    double mx = 0.0;
#pragma unroll 8
    for (int c = half * 32; c < half * 32 + 32; ++c) mx = fmax(mx, fabs(src[tidx(r, c)]));
    if (half) red[r] = mx;
    __syncthreads();
    if (!half) atomicMax(maxbits + (size_t)tr * TS + r, (unsigned long long)__double_as_longlong(fmax(mx, red[r])));
}

__global__ void __launch_bounds__(NTHREADS)
i8_split_tiles_kernel(const double *__restrict__ tiles, int nt, int t0, int c0, int kt, int S, int rows_pad,
                      const unsigned long long *__restrict__ maxbits, signed char *__restrict__ out, double *__restrict__ rowscale) {
    __shared__ __align__(16) double T[TILE_ELEMS];
    const int tid = threadIdx.x, tr = blockIdx.x / kt, tk = blockIdx.x - tr * kt, ti = t0 + tr, row0 = tr * TS, K = kt * TS;
    const int row = tid >> 1, kh = tid & 1;  // two threads per row, 32 columns of the tile each
    if (ti >= nt) {
        if (tk == 0 && tid < TS) rowscale[row0 + tid] = 1.0;
        for (int s = 0; s < S; ++s) {
            uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t)s * rows_pad + row0 + row) * K + tk * TS + kh * 32);
            dst[0] = make_uint4(0, 0, 0, 0);
            dst[1] = make_uint4(0, 0, 0, 0);
        }
        return;
    }
    const double *src = tiles + tri_index(ti, c0 + tk) * TILE_ELEMS;
    for (int idx = tid; idx < TILE_ELEMS / 2; idx += NTHREADS)
        reinterpret_cast<double2 *>(T)[idx] = reinterpret_cast<const double2 *>(src)[idx];
    const double mx = __longlong_as_double((long long)maxbits[row0 + row]);
    const int e = mx > 0.0 ? ilogb(mx) + 1 : 0;
    if (tk == 0 && kh == 0) rowscale[row0 + row] = ldexp(1.0, e);
    const double inv = ldexp(1.0, -e);
    __syncthreads();
#pragma unroll
    for (int chunk = 0; chunk < 2; ++chunk) {
        const int cbase = kh * 32 + chunk * 16;
        double r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = T[tidx(row, cbase + i)] * inv;
        for (int s = 0; s < S; ++s) {
            unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                r[i] *= 128.0;
                const int q = (int)r[i];  // truncation: the remainder keeps its sign and stays below 1 in magnitude
                r[i] -= (double)q;
                w[i >> 2] |= (unsigned)(q & 0xff) << ((i & 3) * 8);
            }
            *reinterpret_cast<uint4 *>(out + ((size_t)s * rows_pad + row0 + row) * K + tk * TS + cbase) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// Loads the kernels (see gpl_i8::prepare): call before the factorisation's persistent worker kernel is started.
int i8_prepare() {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, i8_split_tiles_kernel) != cudaSuccess || cudaFuncGetAttributes(&fa, i8_rowmax_kernel) != cudaSuccess) return -1;
    return gpl_i8::prepare<1>() ? 0 : -1;
}

size_t i8_slices_bytes(int nt, int t0, int kt, int S) {
    const size_t rows_pad = (((size_t)(nt - t0) * TS + 127) / 128) * 128;
    return (size_t)S * rows_pad * kt * TS;
}
// row scales (doubles) followed by the bit patterns of the row maxima (scratch of the split)
size_t i8_scale_bytes(int nt, int t0) { return 2 * ((((size_t)(nt - t0) * TS + 127) / 128) * 128) * sizeof(double); }

int i8_split_tiles(const double *tiles, int nt, int t0, int c0, int kt, int S, signed char *slices, double *rowscale, cudaStream_t st) {
    const int rows_pad = (((nt - t0) * TS + 127) / 128) * 128;
    unsigned long long *maxbits = reinterpret_cast<unsigned long long *>(rowscale + rows_pad);
    if (cudaMemsetAsync(maxbits, 0, (size_t)rows_pad * sizeof(unsigned long long), st) != cudaSuccess) return -1;
    i8_rowmax_kernel<<<rows_pad / TS * kt, NTHREADS, 0, st>>>(tiles, nt, t0, c0, kt, maxbits);
    i8_split_tiles_kernel<<<rows_pad / TS * kt, NTHREADS, 0, st>>>(tiles, nt, t0, c0, kt, S, rows_pad, maxbits, slices, rowscale);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// tiles(i, l) -= sum over the block's columns, for 128-wide column blocks [cb0, cb1) of the region and every block row below
int i8_trail(double *tiles, int nt, int t0, int kt, int S, const signed char *slices, const double *rowscale, int cb0, int cb1,
             int max_ctas, int *info, int *dbg, cudaStream_t st) {
    gpl_i8::View vw = {};
    vw.rowscale = rowscale;
    vw.tiles = tiles;
    vw.nt64 = nt;
    vw.t0 = t0;
    vw.layout = 1;
    vw.mode = 0;
    vw.n_rows = (((nt - t0) * TS + 127) / 128) * 128;
    vw.K = kt * TS;
    vw.cb0 = cb0;
    vw.cb1 = cb1 < vw.n_rows / 128 ? cb1 : vw.n_rows / 128;
    vw.dbg = dbg;
    vw.info = info;
    return gpl_i8::launch<1>(reinterpret_cast<const int8_t *>(slices), S, vw, max_ctas, st);
}

}  // namespace gpl
