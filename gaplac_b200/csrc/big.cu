// Single large-n model: blocked Cholesky over the whole GPU on tile-major storage, with the forward solve
// fused into the panel kernels (BASELINE config 5: n = 8192; also posterior fits with n > 1024).
//
// Replaces cholesky(Symmetric(K_y)) -> dpotrf('U') + U' \ y -> dtrtrs + logdet [upstream LinearAlgebra /
// OpenBLAS via AbstractGPs logpdf / posterior]; call sites CLI/src/select.jl:49-52, src/plotting.jl:8.
//
// Two-level blocking (one CTA of 128 threads per tile, DMMA tile primitives of tile.cuh).  A panel is PANEL tile columns (256 matrix columns).  Inside a panel the tile columns are
// factored left-looking (big_diag_kernel: 1 CTA; big_col_kernel: one CTA per tile below the diagonal); after
// a panel, big_trail_kernel applies its rank-256 update to every remaining tile (one CTA per tile, K = 256 so
// the trailing matrix moves through HBM once per panel, not once per tile column).  The host overlaps the (serial,
// latency-bound) factorisation of panel P+1 with the trailing update of panel P on two streams (look-ahead, api.cu).
#include <cstdio>

#include "kernels.h"
#include "tile.cuh"

#ifndef GPL_TRAIL_KC
#define GPL_TRAIL_KC 16  // columns per stage of the trailing-update ring
#define GPL_TRAIL_NS 3   // ring slots (48 KiB: still four CTAs per SM; measured n = 8192: 16 x 3 -> 7.90 ms, 16 x 2 -> 8.05,
                         // 8 x 4 -> 8.25, 16 x 4 (three CTAs per SM) -> 8.27; the LDGSTS ring it replaced: 8.07)
#endif

namespace gpl {

namespace {
struct __align__(16) BigSmem {
    double A[TILE_ELEMS];
    double Bt[TILE_ELEMS];
    double W[TILE_ELEMS];
    double D[DSIZE];
    double rsbuf[16];
    double pivbuf[TS];
    double ybuf[TS];
    double L16s[256];
};

// the column kernel of the worker protocol: two tile buffers are enough (74 KiB and <= 128 registers, so that three CTAs
// of the trailing update still fit next to a waiting column CTA; with the 106 KiB layout only two did)
struct __align__(16) ColSmem {
    double A[TILE_ELEMS];
    double Bt[TILE_ELEMS];
    double D[DSIZE];
    double ybuf[TS];
    double tmp[16];
};
}  // namespace

size_t big_smem_bytes() { return sizeof(BigSmem); }
size_t big_col_smem_bytes() { return sizeof(ColSmem); }

__global__ void __launch_bounds__(NTHREADS) dense_to_tiles_kernel(const double *__restrict__ A, int n, int nt,
                                                                  double *__restrict__ tiles) {
    const long long t = blockIdx.x;
    int i, j;
    tri_unrank(t, i, j);
    double *tile = tiles + t * TILE_ELEMS;
    for (int e = threadIdx.x; e < TILE_ELEMS; e += NTHREADS) {
        const int r = e & (TS - 1), c = e >> 6;
        int gr = i * TS + r, gc = j * TS + c;
        double v;
        if (gr >= n || gc >= n) {
            v = (gr == gc) ? 1.0 : 0.0;
        } else {
            if (gr < gc) {  // upper part of a diagonal tile: mirror the lower triangle
                const int tmp = gr;
                gr = gc;
                gc = tmp;
            }
            v = A[(size_t)gc * n + gr];
        }
        tile[tidx(r, c)] = v;
    }
}

__global__ void __launch_bounds__(NTHREADS) tiles_to_upper_kernel(const double *__restrict__ tiles, int n, int nt,
                                                                  double *__restrict__ U) {
    __shared__ double T[TS][TS + 1];
    const int ti = blockIdx.x, tj = blockIdx.y;  // block (ti, tj) of U: rows ti*64.., cols tj*64..
    if (ti <= tj) {
        const double *tile = tiles + tri_index(tj, ti) * TILE_ELEMS;  // L block (tj, ti)
        for (int e = threadIdx.x; e < TILE_ELEMS; e += NTHREADS) T[e >> 6][e & (TS - 1)] = tile[tidx(e & (TS - 1), e >> 6)];  // T[c][r]
    }
    __syncthreads();
    for (int e = threadIdx.x; e < TILE_ELEMS; e += NTHREADS) {
        const int a = e & (TS - 1), b = e >> 6;  // U local (a, b): row a, column b
        const int ga = ti * TS + a, gb = tj * TS + b;
        if (ga >= n || gb >= n) continue;
        // U(a, b) = L(b, a): L tile element (row b, col a) = T[a][b]
        U[(size_t)gb * n + ga] = (ti <= tj && ga <= gb) ? T[a][b] : 0.0;
    }
}

// tile (j, j): left-looking update with the panel's earlier columns, Cholesky + inverse, forward-solve block
__global__ void __launch_bounds__(NTHREADS) big_diag_kernel(BigParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BigSmem &sm = *reinterpret_cast<BigSmem *>(smem_raw);
    const int tid = threadIdx.x, j = prm.j;
    const TMap tm = thread_map(tid);
    double *Tjj = prm.tiles + tri_index(j, j) * TILE_ELEMS;
    tile_load_async(sm.A, Tjj, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    double acc[2][NCC];
    acc_from_tile(acc, sm.A, tm);
    for (int k = prm.k0; k < j; ++k) {
        __syncthreads();
        tile_load_async(sm.A, prm.tiles + tri_index(j, k) * TILE_ELEMS, tid);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        tile_mma<true>(acc, sm.A, sm.A, tm, 0, TS);
    }
    __syncthreads();
    const int fail = tile_potrf(acc, tm, sm.A, sm.L16s, sm.D, sm.rsbuf, sm.pivbuf, tid);
    if (tid == 0 && fail >= 0) atomicCAS(prm.info, 0, j * TS + fail + 1);
    acc_to_tile(Tjj, acc, tm);
    __syncthreads();
    acc_to_tile(sm.A, acc, tm);  // L_jj in shared memory for the inverse
    if (prm.y && tid < TS) sm.ybuf[tid] = prm.y[j * TS + tid];
    __syncthreads();
    // full inverse W_jj = L_jj^-1 (X L' = I gives W'): the column kernel and the backward solve use it
    double e[2][NCC];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) e[mb][cc] = (row_of(tm, mb) == col_of(tm, cc)) ? 1.0 : 0.0;
    tile_trsm_ld(e, sm.A, sm.D, tm);
    acc_to_tile_t(sm.W, e, tm);
    __syncthreads();
    tile_store(prm.winv + (size_t)j * TILE_ELEMS, sm.W, tid);
    if (tid < TS) prm.pivlog[j * TS + tid] = log(sm.pivbuf[tid]);
    // z_j = W_jj y_j (y_j already carries the updates of all earlier tile columns)
    if (prm.y && tid < TS) prm.y[j * TS + tid] = tile_row_dot(sm.W, sm.ybuf, tid, 0, tid + 1);
}

// tiles (i, j), i > j: left-looking update inside the panel, then L_ij = T_ij W_jj', then y_i -= L_ij z_j
__global__ void __launch_bounds__(NTHREADS) big_col_kernel(BigParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BigSmem &sm = *reinterpret_cast<BigSmem *>(smem_raw);
    const int tid = threadIdx.x, j = prm.j, i = prm.j + 1 + blockIdx.x;
    const TMap tm = thread_map(tid);
    double *Tij = prm.tiles + tri_index(i, j) * TILE_ELEMS;
    tile_load_async(sm.A, Tij, tid);
    tile_load_async(sm.W, prm.winv + (size_t)j * TILE_ELEMS, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    double acc[2][NCC];
    acc_from_tile(acc, sm.A, tm);
    for (int k = prm.k0; k < j; ++k) {
        __syncthreads();
        tile_load_async(sm.A, prm.tiles + tri_index(i, k) * TILE_ELEMS, tid);
        tile_load_async(sm.Bt, prm.tiles + tri_index(j, k) * TILE_ELEMS, tid);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        tile_mma<true>(acc, sm.A, sm.Bt, tm, 0, TS);
    }
    tile_trsm_w(acc, sm.W, tm);  // L_ij = T_ij W_jj', row operand straight from the accumulator registers
    acc_to_tile(Tij, acc, tm);
    if (prm.y) {
        __syncthreads();
        acc_to_tile(sm.A, acc, tm);
        __syncthreads();
        if (tid < TS) prm.y[i * TS + tid] -= tile_row_dot(sm.A, prm.y + j * TS, tid, 0, TS);
    }
}

// ---- look-ahead protocol -----------------------------------------------------------------------------------------
// The diagonal tiles are the serial chain of the factorisation.  With look-ahead they are factored by ONE persistent CTA
// (big_worker_kernel) that owns a whole SM for the duration (200 KiB of dynamic shared memory: no DMMA-streaming CTA can
// share its sub-partitions and starve its pivot chains, tools/pipe_mix.cu), while the host streams the column and
// trailing kernels around it.  Synchronisation is through three flag arrays in global memory:
//   panel_ready[P]  set by the host (stream-ordered memset) once every earlier panel's update of panel P's columns is done
//   tiledone[i]     = j + 1 once tile (i, j) is final                            (written by big_col_flag_kernel)
//   rowdone[i]      = j + 1 once, in addition, y_i carries column j              (written by big_col_flag_kernel)
//   diagdone[j]     1 once L_jj and its block inverses are stored, 2 once z_j is   (written by the worker)
//   abort           set by whoever waits longer than BIG_WAIT_NS for a hand-off: every later wait returns at once, the
//                   factorisation finishes with garbage and info = -1, and the host reports GPL_ERR_CUDA.  Forward progress
//                   of the protocol needs the worker CTA and at least one column CTA resident at the same time; a bounded wait
//                   turns a lost hand-off (or a device shared with something that keeps them apart) into an error, not a hang.
constexpr unsigned long long BIG_WAIT_NS = 2000000000ull;  // 2 s; hand-offs normally take microseconds
__device__ __forceinline__ int ld_flag(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void wait_flag_ge(const int *p, int v, int tid, int *abort_flag, int *info) {
    if (tid == 0 && ld_flag(p) < v) {
        const unsigned long long t0 = global_ns();
        unsigned spins = 0;
        while (ld_flag(p) < v && ld_flag(abort_flag) == 0) {
            __nanosleep(40);
            if ((++spins & 1023u) == 0 && global_ns() - t0 > BIG_WAIT_NS) {
                atomicExch(abort_flag, 1);
                atomicExch(info, -1);
            }
        }
    }
    if (tid == 0) __threadfence();
    __syncthreads();
}

__global__ void __launch_bounds__(NTHREADS) big_worker_kernel(BigParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BigSmem &sm = *reinterpret_cast<BigSmem *>(smem_raw);
    const int tid = threadIdx.x, nt = prm.nt;
    const TMap tm = thread_map(tid);
    int *panel_ready = prm.flags, *rowdone = prm.flags + BIG_MAXP, *diagdone = rowdone + nt, *tiledone = diagdone + nt;
    int *abort_flag = tiledone + nt;
    constexpr int PANEL_ = BIG_PANEL;
#ifdef GPL_BIG_PROFILE
    long long t_wait_panel = 0, t_wait_row = 0, t_begin = clock64(), t_pre = 0, t_last = 0, t_potrf = 0, t_pub = 0, t_fwd = 0;
#endif
    for (int j = 0; j < nt; ++j) {
        const int k0 = (j / PANEL_) * PANEL_;
#ifdef GPL_BIG_PROFILE
        long long ta = clock64();
#endif
        if (j == k0) wait_flag_ge(panel_ready + j / PANEL_, 1, tid, abort_flag, prm.info);
#ifdef GPL_BIG_PROFILE
        long long tb = clock64();
        t_wait_panel += tb - ta;
#endif
        // In-panel left-looking update of the diagonal tile.  Everything but the last term is applied BEFORE waiting
        // for tile (j, j-1) - the only input that is still being produced (by the column kernel of column j-1, which
        // started when L_{j-1,j-1} was published) - so that only one 64-column update and the factorisation itself sit
        // on the serial path.
        double *Tjj = prm.tiles + tri_index(j, j) * TILE_ELEMS;
        double acc[2][NCC];
        if (j - k0 >= 2) wait_flag_ge(tiledone + j, j - 1, tid, abort_flag, prm.info);  // tiles (j, k0..j-2) are final (normally long since)
        __syncthreads();
        tile_load_async(sm.A, Tjj, tid);
        cp_async_commit();
        if (j - k0 >= 2) {
            tile_load_async(sm.Bt, prm.tiles + tri_index(j, k0) * TILE_ELEMS, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        acc_from_tile(acc, sm.A, tm);
        for (int k = k0; k < j - 1; ++k) {  // tile k sits in Bt / A alternately
            double *cur = ((k - k0) & 1) ? sm.A : sm.Bt;
            double *nxt = ((k - k0) & 1) ? sm.Bt : sm.A;
            cp_async_wait<0>();
            __syncthreads();  // tile k landed; everyone done reading `nxt` (previous step / accumulator load)
            if (k + 1 < j - 1) {
                tile_load_async(nxt, prm.tiles + tri_index(j, k + 1) * TILE_ELEMS, tid);
                cp_async_commit();
            }
            tile_mma<true>(acc, cur, cur, tm, 0, TS);
        }
#ifdef GPL_BIG_PROFILE
        long long tc = clock64();
        t_pre += tc - tb;
#endif
        if (j > 0) wait_flag_ge(tiledone + j, j, tid, abort_flag, prm.info);
#ifdef GPL_BIG_PROFILE
        long long td = clock64();
        t_wait_row += td - tc;
#endif
        if (j > k0) {  // the last term: tile (j, j-1)
            double *cur = ((j - 1 - k0) & 1) ? sm.A : sm.Bt;
            __syncthreads();
            tile_load_async(cur, prm.tiles + tri_index(j, j - 1) * TILE_ELEMS, tid);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            tile_mma<true>(acc, cur, cur, tm, 0, TS);
        }
        __syncthreads();
#ifdef GPL_BIG_PROFILE
        long long te = clock64();
        t_last += te - td;
#endif
        const int fail = tile_potrf(acc, tm, sm.A, sm.L16s, sm.D, sm.rsbuf, sm.pivbuf, tid);
#ifdef GPL_BIG_PROFILE
        long long tf = clock64();
        t_potrf += tf - te;
#endif
        if (tid == 0 && fail >= 0) atomicCAS(prm.info, 0, j * TS + fail + 1);
        // Publish L_jj and the inverses of its 16 x 16 diagonal blocks; that is all the column kernel needs for its solve.
        // The full inverse W_jj (backward substitution, posterior) is formed after the factorisation by big_winv_kernel,
        // off this serial path.
        acc_to_tile(Tjj, acc, tm);
        for (int t = tid; t < DSIZE; t += NTHREADS) prm.dblk[(size_t)j * DSIZE + t] = sm.D[t];
        __threadfence();
        __syncthreads();
        if (tid == 0) *reinterpret_cast<volatile int *>(diagdone + j) = 1;  // the solves below the tile can start
#ifdef GPL_BIG_PROFILE
        long long tg = clock64();
        t_pub += tg - tf;
#endif
        if (tid < TS) prm.pivlog[j * TS + tid] = log(sm.pivbuf[tid]);
        if (prm.y) {  // z_j while the column CTAs load L_jj and solve: they need it only for their right-hand sides
            acc_to_tile(sm.A, acc, tm);
            if (j > 0) wait_flag_ge(rowdone + j, j, tid, abort_flag, prm.info);  // y_j carries every earlier column (set after tiledone)
            if (tid < TS) sm.ybuf[tid] = __ldcg(prm.y + j * TS + tid);
            tile_forward_solve(sm.A, sm.D, sm.ybuf, sm.rsbuf, tid);  // z_j = L_jj^-1 y_j
            if (tid < TS) prm.y[j * TS + tid] = sm.ybuf[tid];
            __threadfence();
            __syncthreads();
            if (tid == 0) *reinterpret_cast<volatile int *>(diagdone + j) = 2;  // ... and z_j is stored
        }
#ifdef GPL_BIG_PROFILE
        t_fwd += clock64() - tg;
#endif
    }
#ifdef GPL_BIG_PROFILE
    if (tid == 0)
        printf("worker: total %lld clk, waiting for panel_ready %lld, for rowdone %lld, working %lld (load + early updates %lld, "
               "last update %lld, potrf %lld, publish %lld, pivlog + forward solve %lld)\n",
               clock64() - t_begin, t_wait_panel, t_wait_row, clock64() - t_begin - t_wait_panel - t_wait_row, t_pre, t_last,
               t_potrf, t_pub, t_fwd);
#endif
}

__global__ void __launch_bounds__(NTHREADS, 4) big_col_flag_kernel(BigParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ColSmem &sm = *reinterpret_cast<ColSmem *>(smem_raw);
    const int tid = threadIdx.x, j = prm.j, i = prm.j + 1 + blockIdx.x, nt = prm.nt;
    const TMap tm = thread_map(tid);
    int *rowdone = prm.flags + BIG_MAXP, *diagdone = rowdone + nt, *tiledone = diagdone + nt, *abort_flag = tiledone + nt;
    double *Tij = prm.tiles + tri_index(i, j) * TILE_ELEMS;
    // the tile itself receives no further outside update: fetch it while the worker is still on the diagonal tile
    tile_load_async(sm.A, Tij, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    double acc[2][NCC];
    acc_from_tile(acc, sm.A, tm);
    // in-panel update first: its inputs - tiles (i, k) and (j, k), k0 <= k < j - were final before this kernel was
    // launched (stream order behind the previous column's kernel), so it overlaps the worker's factorisation of L_jj
    for (int k = prm.k0; k < j; ++k) {
        __syncthreads();
        tile_load_async(sm.A, prm.tiles + tri_index(i, k) * TILE_ELEMS, tid);
        tile_load_async(sm.Bt, prm.tiles + tri_index(j, k) * TILE_ELEMS, tid);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        tile_mma<true>(acc, sm.A, sm.Bt, tm, 0, TS);
    }
    wait_flag_ge(diagdone + j, 1, tid, abort_flag, prm.info);  // L_jj and its block inverses are stored
    __syncthreads();  // the last update's operands are consumed: Bt takes L_jj
    tile_load_async(sm.Bt, prm.tiles + tri_index(j, j) * TILE_ELEMS, tid);
    block_load_async<DSIZE * 8>(sm.D, prm.dblk + (size_t)j * DSIZE, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    tile_trsm_ld(acc, sm.Bt, sm.D, tm);
    acc_to_tile(Tij, acc, tm);
    __threadfence();
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile int *>(tiledone + i) = j + 1;  // the worker's next update needs only the tile
    if (prm.y) {
        acc_to_tile(sm.A, acc, tm);
        wait_flag_ge(diagdone + j, 2, tid, abort_flag, prm.info);  // z_j (published after L_jj; normally long since)
        if (tid < TS) sm.ybuf[tid] = __ldcg(prm.y + j * TS + tid);
        __syncthreads();
        if (tid < TS) prm.y[i * TS + tid] = __ldcg(prm.y + i * TS + tid) - tile_row_dot(sm.A, sm.ybuf, tid, 0, TS);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile int *>(rowdone + i) = j + 1;
}


// ---- look-ahead protocol, second version (round 2): the worker also solves tile (j+1, j) --------------------------------
// In the first version every in-panel column cost a round trip on the serial path: worker factors L_jj -> the column CTA
// of row j+1 loads it, solves its tile, stores, raises a flag -> the worker loads that tile back, applies the last update
// and factors again (measured: 23.8 us of worker work + 8.5 us of waiting per column).  Here the CTA of row j+1 prepares,
// while the worker is still busy with L_jj, everything that does not need L_jj - the tile (j+1, j) and the diagonal tile
// (j+1, j+1), both updated with the panel's earlier columns - and raises prep[j+1]; the worker then solves the one tile
// itself (it has L_jj and its block inverses in shared memory), stores it, and applies the rank-64 update to the diagonal
// tile it keeps in registers for the next factorisation: no hand-off remains on the path inside a panel.  The forward
// solves z_j = L_jj^-1 y_j move off the worker to that same CTA (the last column's stays with the worker).  At a panel
// boundary (column j+1 starts a new panel: its tiles still await the trailing update) the column CTA of row j+1 solves its
// tile as before.
//   prep[i] = j + 1: T'_{i,j} and T'_{i,i} (i = j + 1) are stored          (written by big_col2_kernel, read by the worker)
//   tiledone[i] = j + 1: L_{i,j} is stored  (by the worker for i = j + 1 inside a panel, by big_col2_kernel otherwise)
__global__ void __launch_bounds__(NTHREADS) big_worker2_kernel(BigParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BigSmem &sm = *reinterpret_cast<BigSmem *>(smem_raw);
    const int tid = threadIdx.x, nt = prm.nt;
    const TMap tm = thread_map(tid);
    int *panel_ready = prm.flags, *rowdone = prm.flags + BIG_MAXP, *diagdone = rowdone + nt, *tiledone = diagdone + nt;
    int *abort_flag = tiledone + nt, *prep = abort_flag + 1, *prepd = prep + nt;
    constexpr int PANEL_ = BIG_PANEL;
    double acc[2][NCC];
    bool have_acc = false;  // acc already holds the fully updated diagonal tile of this column
#ifdef GPL_BIG_PROFILE
    long long w_panel = 0, w_load = 0, w_potrf = 0, w_pub = 0, w_prep = 0, w_solve = 0, w_begin = clock64(), w_t;
#define W_MARK(acc_) do { const long long now_ = clock64(); acc_ += now_ - w_t; w_t = now_; } while (0)
#else
#define W_MARK(acc_)
#endif
    for (int j = 0; j < nt; ++j) {
#ifdef GPL_BIG_PROFILE
        w_t = clock64();
#endif
        const int k0 = (j / PANEL_) * PANEL_, j1 = (k0 + PANEL_ < nt) ? k0 + PANEL_ : nt;
        double *Tjj = prm.tiles + tri_index(j, j) * TILE_ELEMS;
        if (!have_acc) {  // first column of a panel: the trailing kernels have applied every earlier panel
            wait_flag_ge(panel_ready + j / PANEL_, 1, tid, abort_flag, prm.info);
            W_MARK(w_panel);
            tile_load_async(sm.A, Tjj, tid);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            acc_from_tile(acc, sm.A, tm);
            __syncthreads();  // sm.A becomes the factorisation scratch
            W_MARK(w_load);
        }
        const int fail = tile_potrf(acc, tm, sm.A, sm.L16s, sm.D, sm.rsbuf, sm.pivbuf, tid);
        W_MARK(w_potrf);
        if (tid == 0 && fail >= 0) atomicCAS(prm.info, 0, j * TS + fail + 1);
        acc_to_tile(Tjj, acc, tm);
        for (int t = tid; t < DSIZE; t += NTHREADS) prm.dblk[(size_t)j * DSIZE + t] = sm.D[t];
        __threadfence();
        __syncthreads();
        if (tid == 0) *reinterpret_cast<volatile int *>(diagdone + j) = 1;  // the solves below the tile can start
        if (tid < TS) prm.pivlog[j * TS + tid] = log(sm.pivbuf[tid]);
        W_MARK(w_pub);
        have_acc = false;
        if (j + 1 < j1) {  // same panel: tile (j+1, j) and the next diagonal tile are this CTA's business
            const int i = j + 1;
            double *Tij = prm.tiles + tri_index(i, j) * TILE_ELEMS;
            wait_flag_ge(prep + i, j + 1, tid, abort_flag, prm.info);   // T'_{j+1,j} is stored (row CTA of the column kernel)
            wait_flag_ge(prepd + i, j + 1, tid, abort_flag, prm.info);  // T'_{j+1,j+1} is stored (its extra CTA)
            W_MARK(w_prep);
            tile_load_async(sm.Bt, Tij, tid);
            tile_load_async(sm.W, prm.tiles + tri_index(i, i) * TILE_ELEMS, tid);
            cp_async_commit();
            acc_to_tile(sm.A, acc, tm);  // L_jj for the solve (the scratch is dead: barrier above)
            cp_async_wait<0>();
            __syncthreads();
            double t2[2][NCC];
            acc_from_tile(t2, sm.Bt, tm);
            tile_trsm_ld(t2, sm.A, sm.D, tm);  // L_{j+1,j} = T'_{j+1,j} L_jj^-T
            acc_to_tile(Tij, t2, tm);
            __syncthreads();  // every thread has read its part of sm.Bt: it now takes L_{j+1,j} for the update
            acc_to_tile(sm.Bt, t2, tm);
            __threadfence();
            __syncthreads();
            if (tid == 0) *reinterpret_cast<volatile int *>(tiledone + i) = j + 1;
            acc_from_tile(acc, sm.W, tm);                    // T'_{j+1,j+1}: updated with the panel's columns before j
            tile_mma<true>(acc, sm.Bt, sm.Bt, tm, 0, TS);    // ... and now with column j
            __syncthreads();  // sm.A (next scratch), sm.Bt, sm.W are free again
            W_MARK(w_solve);
            have_acc = true;
        }
    }
#ifdef GPL_BIG_PROFILE
    if (tid == 0)
        printf("worker2: total %lld clk: waiting for panel_ready %lld, loading first tiles %lld, potrf %lld, publish %lld, waiting for prep %lld, "
               "solve + update %lld\n", clock64() - w_begin, w_panel, w_load, w_potrf, w_pub, w_prep, w_solve);
#endif
    if (prm.y) {  // z of the last column: no column kernel exists for it
        const int j = nt - 1;
        acc_to_tile(sm.A, acc, tm);
        if (j > 0) wait_flag_ge(rowdone + j, j, tid, abort_flag, prm.info);
        else __syncthreads();
        if (tid < TS) sm.ybuf[tid] = __ldcg(prm.y + j * TS + tid);
        tile_forward_solve(sm.A, sm.D, sm.ybuf, sm.rsbuf, tid);
        if (tid < TS) prm.y[j * TS + tid] = sm.ybuf[tid];
        __threadfence();
        __syncthreads();
        if (tid == 0) *reinterpret_cast<volatile int *>(diagdone + j) = 2;
    }
}

__global__ void __launch_bounds__(NTHREADS, 3) big_col2_kernel(BigParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ColSmem &sm = *reinterpret_cast<ColSmem *>(smem_raw);
    const int tid = threadIdx.x, j = prm.j, i = prm.j + 1 + blockIdx.x, nt = prm.nt;
    const TMap tm = thread_map(tid);
    int *rowdone = prm.flags + BIG_MAXP, *diagdone = rowdone + nt, *tiledone = diagdone + nt, *abort_flag = tiledone + nt;
    int *prep = abort_flag + 1, *prepd = prep + nt;
    if (i == nt) {  // the extra CTA of an in-panel column: the diagonal tile of column j + 1 with the updates of the panel's
                    // columns before j (the row CTA prepares the tile (j+1, j) meanwhile: the two together take as long
                    // as the worker's factorisation of L_jj, either alone fits inside it)
        const int r = j + 1;
        double *Trr = prm.tiles + tri_index(r, r) * TILE_ELEMS;
        tile_load_async(sm.A, Trr, tid);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        double d[2][NCC];
        acc_from_tile(d, sm.A, tm);
        for (int k = prm.k0; k < j; ++k) {
            __syncthreads();
            tile_load_async(sm.A, prm.tiles + tri_index(r, k) * TILE_ELEMS, tid);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            tile_mma<true>(d, sm.A, sm.A, tm, 0, TS);
        }
        acc_to_tile(Trr, d, tm);
        __threadfence();
        __syncthreads();
        if (tid == 0) *reinterpret_cast<volatile int *>(prepd + r) = j + 1;
        return;
    }
    const bool next_diag = (i == j + 1);
    const bool same_panel = next_diag && (j + 1 < prm.j1);  // the worker solves this tile
    double *Tij = prm.tiles + tri_index(i, j) * TILE_ELEMS;
    tile_load_async(sm.A, Tij, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    double acc[2][NCC];
    acc_from_tile(acc, sm.A, tm);
    // in-panel update: tiles (i, k) and (j, k), k0 <= k < j, were final before this kernel was launched (stream order)
    for (int k = prm.k0; k < j; ++k) {
        __syncthreads();
        tile_load_async(sm.A, prm.tiles + tri_index(i, k) * TILE_ELEMS, tid);
        tile_load_async(sm.Bt, prm.tiles + tri_index(j, k) * TILE_ELEMS, tid);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        tile_mma<true>(acc, sm.A, sm.Bt, tm, 0, TS);
    }
    if (same_panel) {
        acc_to_tile(Tij, acc, tm);  // T'_{j+1,j}: everything but the solve
        __threadfence();
        __syncthreads();
        if (tid == 0) *reinterpret_cast<volatile int *>(prep + i) = j + 1;
    }
    wait_flag_ge(diagdone + j, 1, tid, abort_flag, prm.info);  // L_jj and its block inverses are stored
    __syncthreads();
    tile_load_async(sm.Bt, prm.tiles + tri_index(j, j) * TILE_ELEMS, tid);
    block_load_async<DSIZE * 8>(sm.D, prm.dblk + (size_t)j * DSIZE, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (!same_panel) {
        tile_trsm_ld(acc, sm.Bt, sm.D, tm);
        acc_to_tile(Tij, acc, tm);
        __threadfence();
        __syncthreads();
        if (tid == 0) *reinterpret_cast<volatile int *>(tiledone + i) = j + 1;
    }
    if (prm.y) {
        if (next_diag) {  // this CTA owns the forward solve of column j: z_j = L_jj^-1 y_j
            if (j > 0) wait_flag_ge(rowdone + j, j, tid, abort_flag, prm.info);  // y_j carries every earlier column
            if (tid < TS) sm.ybuf[tid] = __ldcg(prm.y + j * TS + tid);
            tile_forward_solve(sm.Bt, sm.D, sm.ybuf, sm.tmp, tid);
            if (tid < TS) prm.y[j * TS + tid] = sm.ybuf[tid];
            __threadfence();
            __syncthreads();
            if (tid == 0) *reinterpret_cast<volatile int *>(diagdone + j) = 2;
        } else {
            wait_flag_ge(diagdone + j, 2, tid, abort_flag, prm.info);
            if (tid < TS) sm.ybuf[tid] = __ldcg(prm.y + j * TS + tid);
        }
    }
    if (same_panel) {  // the worker stores L_{j+1,j}: wait for it (also orders this kernel's end behind that store)
        wait_flag_ge(tiledone + i, j + 1, tid, abort_flag, prm.info);
        if (prm.y) {
            tile_load_async(sm.A, Tij, tid);
            cp_async_commit();
            cp_async_wait<0>();
        }
    } else if (prm.y) {
        __syncthreads();
        acc_to_tile(sm.A, acc, tm);
    }
    if (prm.y) {
        __syncthreads();
        if (tid < TS) prm.y[i * TS + tid] = __ldcg(prm.y + i * TS + tid) - tile_row_dot(sm.A, sm.ybuf, tid, 0, TS);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile int *>(rowdone + i) = j + 1;
}

// W_jj = L_jj^-1 for every diagonal tile at once (worker protocol: the serial path only produced the block inverses)
__global__ void __launch_bounds__(NTHREADS) big_winv_kernel(BigParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BigSmem &sm = *reinterpret_cast<BigSmem *>(smem_raw);
    const int tid = threadIdx.x, j = blockIdx.x;
    const TMap tm = thread_map(tid);
    tile_load_async(sm.A, prm.tiles + tri_index(j, j) * TILE_ELEMS, tid);
    block_load_async<DSIZE * 8>(sm.D, prm.dblk + (size_t)j * DSIZE, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    double e[2][NCC];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) e[mb][cc] = (row_of(tm, mb) == col_of(tm, cc)) ? 1.0 : 0.0;
    tile_trsm_ld(e, sm.A, sm.D, tm);
    acc_to_tile_t(prm.winv + (size_t)j * TILE_ELEMS, e, tm);
}

// trailing tiles (i, l), i >= l >= j1: T_il -= sum_{k0 <= k < j1} L_ik L_lk'
// Operands move with bulk asynchronous copies (cp.async.bulk, one instruction per 4 KiB chunk, issued by thread 0) through
// a ring of slots; "full" mbarriers (transaction bytes) hand a slot to the consumers, "empty" mbarriers (one arrival per
// warp) hand it back, so the warps run decoupled: there is no block barrier in the loop.
__global__ void __launch_bounds__(NTHREADS, 4) big_trail_kernel(BigParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *S = reinterpret_cast<double *>(smem_raw);  // NS slots of KC columns of both operands
    constexpr int KC = GPL_TRAIL_KC, CH = KC * TS, NS = GPL_TRAIL_NS;
    __shared__ __align__(8) unsigned long long full_bar[NS], empty_bar[NS];
    const int tid = threadIdx.x;
    const TMap tm = thread_map(tid);
    // linear block index -> tile (i, l), l0 <= l < l1, l <= i < nt (column by column)
    int idx = blockIdx.x, l = prm.l0;
    while (idx >= prm.nt - l) {
        idx -= prm.nt - l;
        ++l;
    }
    const int i = l + idx;
    const bool diag = (i == l);
    double *Til = prm.tiles + tri_index(i, l) * TILE_ELEMS;
    const int Q = (TS / KC) * (prm.j1 - prm.k0);  // the panel's tiles of a tile row are contiguous
    const double *srcA = prm.tiles + tri_index(i, prm.k0) * TILE_ELEMS;
    const double *srcB = prm.tiles + tri_index(l, prm.k0) * TILE_ELEMS;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NWARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int s) {  // thread 0 only
        const int slot = s % NS;
        double *dst = S + slot * 2 * CH;
        mbar_arrive_expect_tx(&full_bar[slot], (diag ? 1u : 2u) * CH * 8u);
        bulk_load(dst, srcA + (size_t)s * CH, CH * 8, &full_bar[slot]);
        if (!diag) bulk_load(dst + CH, srcB + (size_t)s * CH, CH * 8, &full_bar[slot]);
    };
    if (tid == 0)
        for (int s = 0; s < NS - 1 && s < Q; ++s) issue(s);
    double acc[2][NCC];
    acc_from_tile(acc, Til, tm);  // straight from global / L2 into the accumulator registers
    for (int q = 0; q < Q; ++q) {
        if (tid == 0 && q + NS - 1 < Q) {
            // slot of stage q - 1: every warp has finished reading it (its use number (q - 1) / NS)
            if (q >= 1) mbar_wait(&empty_bar[(q - 1) % NS], ((q - 1) / NS) & 1);
            issue(q + NS - 1);
        }
        mbar_wait(&full_bar[q % NS], (q / NS) & 1);
        const double *a = S + (q % NS) * 2 * CH;
        tile_mma<true>(acc, a, diag ? a : a + CH, tm, 0, KC);
        // generic-proxy reads of the slot must be ordered before the async-proxy refill: without this fence the bulk copy
        // overwrote chunks that a warp was still reading (wrong factors at n = 8192, run-to-run different)
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty_bar[q % NS]);
    }
    acc_to_tile(Til, acc, tm);
}
size_t big_trail_smem_bytes() { return (size_t)GPL_TRAIL_NS * 2 * GPL_TRAIL_KC * TS * 8; }

// backward substitution, one launch per tile row i (descending), grid = i + 1: every CTA recomputes
// alpha_i = W_ii' r_i (r_i is final and read-only in this launch), CTA j < i applies r_j -= L_ij' alpha_i,
// CTA j == i writes alpha_i to the output vector.
__global__ void __launch_bounds__(NTHREADS) big_backward_kernel(const double *tiles, const double *winv, int i,
                                                                double *r, double *alpha) {
    __shared__ __align__(16) double T[TILE_ELEMS];
    __shared__ double v[TS], a[TS];
    const int tid = threadIdx.x;
    const int j = blockIdx.x;  // 0..i
    tile_load_async(T, winv + (size_t)i * TILE_ELEMS, tid);
    cp_async_commit();
    cp_async_wait<0>();
    if (tid < TS) v[tid] = r[i * TS + tid];
    __syncthreads();
    if (tid < TS) a[tid] = tile_col_dot(T, v, tid);
    __syncthreads();
    if (j == i) {
        if (tid < TS) alpha[i * TS + tid] = a[tid];
        return;
    }
    tile_load_async(T, tiles + tri_index(i, j) * TILE_ELEMS, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (tid < TS) r[j * TS + tid] -= tile_col_dot(T, a, tid);
}

__global__ void __launch_bounds__(NTHREADS) big_reduce_kernel(const double *pivlog, const double *z, int len,
                                                              double *out) {
    __shared__ double red[NWARPS];
    const int tid = threadIdx.x;
    double ld = 0.0, q = 0.0;
    for (int t = tid; t < len; t += NTHREADS) {
        ld += pivlog[t];
        if (z) q = fma(z[t], z[t], q);
    }
    ld = block_sum(ld, red, tid);
    q = block_sum(q, red, tid);
    if (tid == 0) {
        out[0] = ld;
        out[1] = q;
    }
}

}  // namespace gpl
