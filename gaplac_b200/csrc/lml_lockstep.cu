// Batched log-marginal-likelihood, "lockstep" schedule: all GPs of the batch advance tile column by tile column
// through two kernels per column, so that the latency-bound pivot chains of the diagonal tiles never share an SM
// sub-partition with streams of tensor instructions.
//
// Why not one fused kernel per GP (lml_batched.cu)?  Measured on B200 (tools/pipe_mix.cu, profiles/README.md): a
// warp issuing dependent FP64 operations slows from 8 to 73 clocks per operation when ONE other warp on its SM
// sub-partition streams DMMAs, and is starved outright (2*10^4 clocks per operation) by two of them.  In the fused
// kernel the diagonal-tile Cholesky of one GP (64 sequential pivots per tile) ran exactly in that situation next to
// the update loops of its three co-resident GPs: 19 % of every CTA's life went into 8 % of its arithmetic.
//
// Replaces, per batch item, logpdf(FiniteGP, y) [upstream AbstractGPs 0.5.12]: kernelmatrix -> + sigma2 I ->
// cholesky (dpotrf) -> U' \ y (dtrtrs) -> logdet; call sites CLI/src/select.jl:49-50, CLI/src/mcmc.jl:35.
//
// Per tile column j (left-looking blocked Cholesky on 64 x 64 tiles, workspace = every item's lower tiles in HBM):
//   lk_diag_kernel   grid B          T_jj = K_jj - sum_k L_jk L_jk'   (covariance tile generated in registers, DMMA
//                                    update from the streamed tiles), y_j - sum_k L_jk z_k
//   lk_potrf_warp_kernel  one warp per item: L_jj = chol(T_jj), its four 16 x 16 block inverses, z_j, logdet and z'z sums
//   lk_below_kernel  grid B*(nt-1-j) L_ij = (K_ij - sum_k L_ik L_jk') L_jj^-T   (K-gen, DMMA update, streamed solve)
// The first and third kernel are pure throughput kernels (every phase saturates the FP64 pipe); the second is
// latency-bound but runs thousands of independent chains with nothing else on the machine.
#include "kernels.h"
#include "kfun.cuh"
#include "tile.cuh"

namespace gpl {

namespace {

constexpr double LOG2PI = 1.8378770664093454835606594728112;
constexpr int LKC = 16;        // columns of one ring slot when it holds a single operand
constexpr int LK_NS = 4;       // ring slots of the operand pipelines (cp.async groups, LK_NS - 1 stages in flight)
constexpr int LCH = LKC * TS;  // doubles per ring slot (8 KiB); S = LK_NS slots = 32 KiB
static_assert(LK_NS * LCH == TILE_ELEMS, "the ring fills the staging buffer exactly");

struct __align__(16) StepSmem {
    double S[TILE_ELEMS];
    double D[DSIZE];
    ItemScalars sc;
    double zs[GPL_LK_ZMAX];  // z_k of the earlier tile columns (diag kernel)
    SepCtx sep;              // separable OU factors of the current block (lk_below_kernel, sorted inputs)
    short kl[GPL_LK_KLMAX];  // tile columns k < j whose tiles are not structurally zero (zero-tile skipping)
    int nk;
};

// The earlier tile columns k < j this tile's update has to visit: all of them, or (with zero flags) those where neither
// operand tile is exactly zero.  Warp 0 fills sm.kl / sm.nk (build_tile_list); a block barrier must follow.
__device__ __forceinline__ void build_klist(StepSmem &sm, const int *zf, int i, int j, bool both, int tid) {
    build_tile_list(sm.kl, &sm.nk, 0, j, [&](int k) { return !(zf && (zf[tri_index(i, k)] || (both && zf[tri_index(j, k)]))); }, tid);
}

__device__ __forceinline__ const double *item_ptr(const double *base, long long stride, int b) {
    return base + (size_t)b * stride;
}

}  // namespace

static_assert(sizeof(StepSmem) <= 56 * 1024, "four CTAs of the lockstep kernels per SM");
size_t lk_step_smem_bytes() { return sizeof(StepSmem); }

// ---- diagonal tile of column j: covariance + update, right-hand side update ----------------------------------------
// Only the lower triangle of a diagonal tile is needed (36 of its 64 8 x 8 blocks).  To balance them over the four
// warps this kernel uses its own row map: warp w owns row block w (accumulator row 0, needs n-blocks 0..w) and row
// block 7 - w (accumulator row 1, needs n-blocks 0..7-w): 9 blocks per warp.  The potrf kernel re-loads the tile in
// the standard map, so nothing else sees this layout.
template <int W, int KCOLS>
__device__ __forceinline__ void diag_mma(double (&acc)[2][NCC], const double *__restrict__ A, const TMap &tm) {
    const int sw = tm.t << 2;
    const double *pa0 = A + tm.t * TS + ((8 * W + tm.g) ^ sw);
    const double *pa1 = A + tm.t * TS + ((8 * (7 - W) + tm.g) ^ sw);
    const double *pbe = A + tm.t * TS + (tm.g ^ (sw & 4)) + (sw & 8);
    const double *pbo = A + tm.t * TS + (tm.g ^ (sw & 4)) + (8 ^ (sw & 8));
#pragma unroll 4
    for (int k = 0; k < KCOLS; k += 4) {
        const double a0 = -pa0[k * TS], a1 = -pa1[k * TS];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            if (nb <= W || nb <= 7 - W) {
                const double b = ((nb & 1) ? pbo : pbe)[k * TS + 16 * (nb >> 1)];
                if (nb <= W) dmma884(acc[0][2 * nb], acc[0][2 * nb + 1], a0, b);
                if (nb <= 7 - W) dmma884(acc[1][2 * nb], acc[1][2 * nb + 1], a1, b);
            }
        }
    }
}

// Diagonal tile of column j of item b: covariance, update with the tiles (j, 0..j-1), right-hand side update.  Item
// scalars must be ready in sm.sc (and a __syncthreads() must lie between that and this call).  Shared by lk_diag_kernel
// (column 0) and lk_below_kernel, whose CTA for the tile (j, j-1) runs it as soon as that tile - the last input - is
// stored: the fixed cost of this phase (covariance code, tile store: ~100 us per launch when it ran alone) then
// overlaps the DMMA streams of the other CTAs instead of leaving the FP64 pipe idle.
__device__ __forceinline__ void diag_tile_phase(const LkParams &prm, StepSmem &sm, int j, int b, int tid) {
    const DevProgram &P = prm.prog;
    const int warp = tid >> 5;
    const TMap tm = thread_map(tid);
    const int n = prm.n, nt = prm.nt;
    const long long ntri = tri_index(nt, 0);
    double *wsL = prm.tiles + (size_t)b * ntri * TILE_ELEMS;
    double *zb = prm.z + (size_t)b * nt * TS;
    const double *X = item_ptr(prm.X, prm.x_stride, b);
    const double *Y = item_ptr(prm.Y, prm.y_stride, b);
    const double diag_add = prm.sigma2[(size_t)b * prm.sigma2_stride] + prm.jitter;

    // only one operand is staged: a ring of LK_NS slots of 16 columns, LK_NS - 1 stages in flight (one barrier per stage)
    constexpr int DKC = 16, DCH = DKC * TS;
    static_assert(DCH == LCH, "one ring slot per stage");
    // stage s = chunk s % 4 of the tile (j, kl[s / 4]): only the tiles of row j that are not exactly zero are visited
    // (without flags - programs that cannot produce exact zeros - the walk is the plain one: no list, no extra barrier)
    const int *zf = prm.zflag ? prm.zflag + (size_t)b * ntri : nullptr;
    int Q = (TS / DKC) * j;
    if (zf) {
        build_klist(sm, zf, j, j, false, tid);
        __syncthreads();
        Q = (TS / DKC) * sm.nk;
    }
    const double *srcA = wsL + tri_index(j, 0) * TILE_ELEMS;  // tiles (j, 0..j-1) are contiguous
    auto stage_off = [&](int s) { return zf ? (size_t)(sm.kl[s / (TS / DKC)] * (TS / DKC) + s % (TS / DKC)) : (size_t)s; };
    auto issue = [&](int s) {  // always commits, so that the group count tracks the stage number
        if (s < Q) block_load_async<DCH * 8>(sm.S + (s % LK_NS) * DCH, srcA + stage_off(s) * DCH, tid);
        cp_async_commit();
    };
    const bool z_in_smem = j * TS <= GPL_LK_ZMAX;
    if (z_in_smem)
        for (int t = tid; t < j * TS; t += NTHREADS) sm.zs[t] = zb[t];
    const double *zsrc = z_in_smem ? sm.zs : zb;
    const int rows[2] = {8 * warp + tm.g, 8 * (7 - warp) + tm.g};  // this phase's row map
    double acc[2][NCC];
    {
        int gi[2] = {j * TS + rows[0], j * TS + rows[1]};
        // quarters of 16 columns: row block 7 - w needs columns up to 63 - 8w
        double *const slot[4] = {sm.S, sm.S + LCH, sm.S + 2 * LCH, sm.S + 3 * LCH};
        eval_block_acc_scr<true, 4, GPL_LK_CW>(P, sm.sc, X, n, n, gi, X, n, n, j * TS, tm.t, diag_add, slot, tid, acc,
                                    (63 - 8 * warp) / 16 + 1);
        __syncthreads();  // every thread has read its quarters back (and zs is complete)
#pragma unroll
        for (int s0 = 0; s0 < LK_NS - 1; ++s0) issue(s0);
    }
    // right-hand side update y_j - sum_k L_jk z_k: every thread takes one row and half of a stage's columns, as two
    // interleaved chains (a 32-long dependent DFMA chain in two warps only made the other two wait at the barrier:
    // FP64 chains crawl while DMMA streams share the pipe, tools/pipe_mix.cu)
    double yp0 = 0.0, yp1 = 0.0;
    const int yrow = tid & (TS - 1), yhalf = (tid >> 6) * (DKC / 2);
    for (int q = 0; q < Q; ++q) {
        cp_async_wait<LK_NS - 2>();  // stage q landed (for this thread's copies)
        __syncthreads();             // ... for everyone's; and everyone is done with stage q - 1
        issue(q + LK_NS - 1);        // into the slot of stage q - 1
        const double *a = sm.S + (q % LK_NS) * DCH;
        switch (warp) {
        case 0: diag_mma<0, DKC>(acc, a, tm); break;
        case 1: diag_mma<1, DKC>(acc, a, tm); break;
        case 2: diag_mma<2, DKC>(acc, a, tm); break;
        default: diag_mma<3, DKC>(acc, a, tm); break;
        }
        {
            const double *zq = zsrc + stage_off(q) * DKC + yhalf;
#pragma unroll
            for (int k = 0; k < DKC / 2; k += 2) {
                yp0 = fma(a[tidx(yrow, yhalf + k)], zq[k], yp0);
                yp1 = fma(a[tidx(yrow, yhalf + k + 1)], zq[k + 1], yp1);
            }
        }
    }
    double *Tjj = wsL + tri_index(j, j) * TILE_ELEMS;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int cc = 0; cc < NCC; ++cc) Tjj[tidx(rows[mb], col_of(tm, cc))] = acc[mb][cc];
    __syncthreads();  // last stage consumed: S is free for the two partial sums per row
    sm.S[tid] = yp0 + yp1;
    __syncthreads();
    if (tid < TS) zb[j * TS + tid] = ((j * TS + tid < n) ? Y[j * TS + tid] : 0.0) - (sm.S[tid] + sm.S[tid + TS]);
}

__global__ void __launch_bounds__(NTHREADS, 4) lk_diag_kernel(const __grid_constant__ LkParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StepSmem &sm = *reinterpret_cast<StepSmem *>(smem_raw);
    const int tid = threadIdx.x, b = blockIdx.x;
    prepare_item_scalars(prm.prog, prm.Theta + (size_t)b * prm.p, &sm.sc, tid);
    __syncthreads();
    diag_tile_phase(prm, sm, prm.j, b, tid);
}

// ---- Cholesky of the diagonal tiles of column j, one WARP per GP ----------------------------------------------------------
// The factorisation of a 64 x 64 tile is a chain of 64 dependent pivots; with one CTA per tile three of four warps sat
// at barriers while the chain warp ran, and a tile took ~59 k clocks (ncu: 45 % of warp time in the barrier after the
// chain).  Here every warp owns one GP's tile in shared memory and runs the whole blocked algorithm by itself
// (__syncwarp only): five independent chains per SM, no idle warps.  Per 16-column panel: the 16 x 16 diagonal block
// is factored in registers (one row per lane, warp shuffles), inverted by forward substitution, the rows below are
// solved against it in place, the right-hand side is advanced (z = L^-1 y), and the trailing lower blocks take their
// rank-16 update through DMMA with accumulators loaded from / stored to the shared tile.
constexpr int PW_WARPS = 7;
// The warp's scratch lives in blocks of the tile that lie above the diagonal (rows 0..15 of columns >= 16 are never
// part of L): the inverse of the current 16 x 16 block in columns 16..31 (row i of it in column 16 + i, rotated by 2i
// so that "one row per lane" reads spread over the banks), the right-hand side in columns 32..35, the pivots in
// 36..39, their reciprocal square roots in column 40.  That makes a GP cost exactly one 32 KiB tile: 7 per SM.
struct __align__(16) PotrfWarpSmem {
    double T[TILE_ELEMS];
};
__device__ __forceinline__ int pw_w(int i, int k) { return (16 + i) * TS + ((k + 2 * i) & 15); }
__device__ __forceinline__ int pw_y(int m) { return (32 + (m >> 4)) * TS + (m & 15); }
__device__ __forceinline__ int pw_piv(int m) { return (36 + (m >> 4)) * TS + (m & 15); }
__device__ __forceinline__ int pw_rs(int c) { return 40 * TS + c; }
size_t lk_potrf_warp_smem_bytes() { return sizeof(PotrfWarpSmem) * PW_WARPS; }  // per GP: / PW_WARPS
// rows 16 P .. 63 of the 16 columns of panel P, shared tile -> workspace, 16 bytes per lane and step (P is a compile-time
// constant: the chunk -> (column, offset) split is a multiply-shift, not the integer division it was)
template <int P_>
__device__ __forceinline__ void pw_store_panel(double *__restrict__ Tjj, const double *T, int lane) {
    constexpr int cpc = 32 - 8 * P_;  // 16-byte chunks per column
#pragma unroll
    for (int it = 0; it < cpc / 2; ++it) {
        const int q = it * 32 + lane, col = q / cpc, off = q - col * cpc;
        const int e = (16 * P_ + col) * TS + 16 * P_ + 2 * off;
        *reinterpret_cast<double2 *>(Tjj + e) = *reinterpret_cast<const double2 *>(T + e);
    }
}
int lk_potrf_warp_items_per_cta() { return PW_WARPS; }

__global__ void __launch_bounds__(32 * PW_WARPS, 1) lk_potrf_warp_kernel(const __grid_constant__ LkPotrfParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;  // the launch decides how many GPs share a CTA (<= PW_WARPS)
    if (b >= prm.B) return;  // whole warp
    PotrfWarpSmem &sm = reinterpret_cast<PotrfWarpSmem *>(smem_raw)[warp];
    const int nt = prm.nt, j = prm.j, g = lane >> 2, t = lane & 3;
    const long long ntri = tri_index(nt, 0);
    double *Tjj = prm.tiles + ((size_t)b * ntri + tri_index(j, j)) * TILE_ELEMS;
    double *zb = prm.z + (size_t)b * nt * TS;
    double *Dg = prm.dblk + ((size_t)b * nt + j) * DSIZE;
    // Only the lower part of the tile moves: rows 16p..63 of the 16 columns of panel p (62.5 % of the tile; the rows
    // above stay scratch).  One cp.async group per panel: the first 16 x 16 block is factored while panels 1..3 land.
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) {
        constexpr int dummy = 0;
        (void)dummy;
        const int cpc = 32 - 8 * pp;  // 16-byte chunks per column
#pragma unroll
        for (int it = 0; it < (16 * (32 - 8 * pp)) / 32; ++it) {
            const int q = it * 32 + lane, col = q / cpc, off = q - col * cpc;
            const int e = (16 * pp + col) * TS + 16 * pp + 2 * off;
            cp_async16(sm.T + e, Tjj + e);
        }
        cp_async_commit();
    }
    sm.T[pw_y(lane)] = zb[j * TS + lane];
    sm.T[pw_y(lane + 32)] = zb[j * TS + lane + 32];
    cp_async_wait<3>();
    __syncwarp();
    int fail = -1;
    const int r_own = lane & 15;
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        // 16 x 16 diagonal block: one row per lane (lanes 16..31 mirror), pivots and columns through shuffles
        double a[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) a[c] = sm.T[tidx(16 * p + r_own, 16 * p + c)];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            double piv = __shfl_sync(0xffffffffu, a[c], c);
            if (!(piv > 0.0)) {
                if (fail < 0) fail = 16 * p + c;
                piv = 1.0;
            }
            const double rs = rsqrt(piv);
            if (lane == 0) {
                sm.T[pw_piv(16 * p + c)] = piv;
                sm.T[pw_rs(c)] = rs;
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
            for (int c = 0; c < 16; ++c) {
                sm.T[tidx(16 * p + r_own, 16 * p + c)] = a[c];
            }
        }
        __syncwarp();
        {
            double x[16];  // column r_own of L16^-1 (axpy-form forward substitution)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = (i == r_own) ? 1.0 : 0.0;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                x[c] *= sm.T[pw_rs(c)];
#pragma unroll
                for (int i = c + 1; i < 16; ++i) x[i] = fma(-sm.T[tidx(16 * p + i, 16 * p + c)], x[c], x[i]);
            }
            if (lane < 16) {
#pragma unroll
                for (int i = 0; i < 16; ++i) sm.T[pw_w(i, r_own)] = x[i];
            }
        }
        __syncwarp();
        for (int e = lane; e < 256; e += 32) Dg[p * DBLK + (e >> 4) * DLD + (e & 15)] = sm.T[pw_w(e >> 4, e & 15)];
        // rows below the block: L[rows, panel] = T[rows, panel] * W16' on the tensor path, 8 rows at a time, in place
        // (ncu, round 2: the scalar version - one row per lane, 136 FMAs each - was 15 % of this kernel's instructions)
        {
            double bw[2][4];  // W16 as the column operand: bw[nb][kk] = W16[8 nb + g][4 kk + t]; rows 0..7 are zero beyond column 7
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) bw[nb][kk] = (nb == 0 && kk >= 2) ? 0.0 : sm.T[pw_w(8 * nb + g, 4 * kk + t)];
            for (int rb = 2 * (p + 1); rb < 8; ++rb) {
                double av[4], x0[2] = {0.0, 0.0}, x1[2] = {0.0, 0.0};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) av[kk] = sm.T[tidx(8 * rb + g, 16 * p + 4 * kk + t)];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) dmma884(x0[0], x0[1], av[kk], bw[0][kk]);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) dmma884(x1[0], x1[1], av[kk], bw[1][kk]);
                // every lane's operands are in registers before the (warp-synchronous) products: the stores cannot overtake a load
                sm.T[tidx(8 * rb + g, 16 * p + 2 * t)] = x0[0];
                sm.T[tidx(8 * rb + g, 16 * p + 2 * t + 1)] = x0[1];
                sm.T[tidx(8 * rb + g, 16 * p + 8 + 2 * t)] = x1[0];
                sm.T[tidx(8 * rb + g, 16 * p + 8 + 2 * t + 1)] = x1[1];
            }
        }
        if (p == 0) cp_async_wait<0>();  // panels 1..3 are needed from here on
        __syncwarp();
        // panel p is final (L): rows 16p..63 of its columns go home now, spreading the stores over the kernel
        switch (p) {
        case 0: pw_store_panel<0>(Tjj, sm.T, lane); break;
        case 1: pw_store_panel<1>(Tjj, sm.T, lane); break;
        case 2: pw_store_panel<2>(Tjj, sm.T, lane); break;
        default: pw_store_panel<3>(Tjj, sm.T, lane); break;
        }
        // right-hand side: z_p = W16 y_p, then y_below -= L[below, panel] z_p
        double zp = 0.0;
        {
            double z1 = 0.0;  // W16 row is zero beyond the diagonal: no triangular predicate needed
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const double2 w = *reinterpret_cast<const double2 *>(&sm.T[pw_w(r_own, 2 * k2)]);
                const double2 yv = *reinterpret_cast<const double2 *>(&sm.T[pw_y(16 * p + 2 * k2)]);
                zp = fma(w.x, yv.x, zp);
                z1 = fma(w.y, yv.y, z1);
            }
            zp += z1;
        }
        __syncwarp();
        if (lane < 16) sm.T[pw_y(16 * p + lane)] = zp;
        __syncwarp();
        for (int r = 16 * (p + 1) + lane; r < TS; r += 32) {
            double s = sm.T[pw_y(r)];
#pragma unroll
            for (int k = 0; k < 16; ++k) s = fma(-sm.T[tidx(r, 16 * p + k)], sm.T[pw_y(16 * p + k)], s);
            sm.T[pw_y(r)] = s;
        }
        // trailing lower blocks (8 x 8): T[rb, cb] -= L[rb, panel] L[cb, panel]'
        for (int rb = 2 * (p + 1); rb < 8; ++rb) {
            const int cb0 = 2 * (p + 1);
            double av[4], cv[6][2], bv[6][4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) av[kk] = -sm.T[tidx(8 * rb + g, 16 * p + 4 * kk + t)];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                if (cb0 + q <= rb) {
                    const int cb = cb0 + q;
                    cv[q][0] = sm.T[tidx(8 * rb + g, 8 * cb + 2 * t)];
                    cv[q][1] = sm.T[tidx(8 * rb + g, 8 * cb + 2 * t + 1)];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) bv[q][kk] = sm.T[tidx(8 * cb + g, 16 * p + 4 * kk + t)];
                }
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int q = 0; q < 6; ++q)
                    if (cb0 + q <= rb) dmma884(cv[q][0], cv[q][1], av[kk], bv[q][kk]);
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                if (cb0 + q <= rb) {
                    const int cb = cb0 + q;
                    sm.T[tidx(8 * rb + g, 8 * cb + 2 * t)] = cv[q][0];
                    sm.T[tidx(8 * rb + g, 8 * cb + 2 * t + 1)] = cv[q][1];
                }
            }
        }
        __syncwarp();
    }
    // results: z_j, running z'z and logdet, failure report, lml after the last column (L_jj went home panel by panel)
    const double z0 = sm.T[pw_y(lane)], z1 = sm.T[pw_y(lane + 32)];
    zb[j * TS + lane] = z0;
    zb[j * TS + lane + 32] = z1;
    double zq = fma(z0, z0, z1 * z1), lg = log(sm.T[pw_piv(lane)]) + log(sm.T[pw_piv(lane + 32)]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        zq += __shfl_xor_sync(0xffffffffu, zq, o);
        lg += __shfl_xor_sync(0xffffffffu, lg, o);
    }
    if (lane == 0) {
        int inf = prm.info[b];
        if (fail >= 0 && inf == 0) {
            inf = j * TS + fail + 1;
            prm.info[b] = inf;
        }
        const double q = (j ? prm.acc2[2 * b] : 0.0) + zq, l = (j ? prm.acc2[2 * b + 1] : 0.0) + lg;
        prm.acc2[2 * b] = q;
        prm.acc2[2 * b + 1] = l;
        if (j == nt - 1) prm.lml[b] = inf ? -INFINITY : -0.5 * ((double)prm.n * LOG2PI + l + q);
    }
}

// ---- tiles below the diagonal of column j -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 4) lk_below_kernel(const __grid_constant__ LkParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StepSmem &sm = *reinterpret_cast<StepSmem *>(smem_raw);
    const DevProgram &P = prm.prog;
    const int tid = threadIdx.x;
    const TMap tm = thread_map(tid);
    const int n = prm.n, nt = prm.nt, j = prm.j;
    const int nbelow = nt - 1 - j;
    // Block order: first the B CTAs of row j + 1 (they also form the next diagonal tile and run about twice as long:
    // started first, they are not what the last wave waits for), then the other rows, the tiles of one GP next to each
    // other (L2 reuse of its row-j operand).
    const int Bn = gridDim.x / nbelow;
    int b, i;
    if ((int)blockIdx.x < Bn) {
        b = blockIdx.x;
        i = j + 1;
    } else {
        const int r = blockIdx.x - Bn;
        b = r / (nbelow - 1);
        i = j + 2 + r % (nbelow - 1);
    }
    const long long ntri = tri_index(nt, 0);
    double *wsL = prm.tiles + (size_t)b * ntri * TILE_ELEMS;
    const double *X = item_ptr(prm.X, prm.x_stride, b);
    {  // pull this tile's inputs into L1 before the covariance code asks for them (its first loads stalled ~9 % of the
       // j = 0 launch on L2 round trips: profiles/ README)
        const int row = (tid < TS ? i : j) * TS + (tid & (TS - 1));
        const double *px = X + (row < n ? row : n - 1);
        for (int c = 0; c < prm.d; ++c) asm volatile("prefetch.global.L1 [%0];" ::"l"(px + (size_t)c * n));
    }
    prepare_item_scalars(P, prm.Theta + (size_t)b * prm.p, &sm.sc, tid);

    // Operand ring: LK_NS slots of 8 KiB.  Update stage s < Q holds 8 columns of both operands (A = tiles (i, 0..j-1),
    // B = tiles (j, 0..j-1)); the three stages after the last update hold columns 0..47 of L_jj as 16-column chunks for
    // the triangular solve.  LK_NS - 1 stages are in flight; one barrier per stage.  (Measured and dropped at the end of
    // round 2: the same ring with cp.async.bulk issued by thread 0 and full / empty mbarriers instead of LDGSTS + block
    // barriers, as in big_trail_kernel - identical bits, headline step 8.73 -> 8.87 ms: with four CTAs per SM the block
    // barriers are already hidden, and thread 0's issue work lands on the critical warp.)
    constexpr int RKC = 8, RCH = RKC * TS;
    static_assert(2 * RCH == LCH && LK_NS * LCH == TILE_ELEMS, "ring slots fill S");
    int *zf = prm.zflag ? prm.zflag + (size_t)b * ntri : nullptr;
    if (zf) build_klist(sm, zf, i, j, true, tid);  // read after the barrier below ("item scalars")
    const double *srcA = wsL + tri_index(i, 0) * TILE_ELEMS;  // tiles (i, 0..j-1)
    const double *srcB = wsL + tri_index(j, 0) * TILE_ELEMS;  // tiles (j, 0..j-1)
    const double *srcL = wsL + tri_index(j, j) * TILE_ELEMS;
    int Q = (TS / RKC) * j;  // with flags: reset after the barrier to 8 stages per visited tile column
    auto issue = [&](int s) {  // always commits, so that the group count tracks the stage number
        double *dst = sm.S + (s % LK_NS) * LCH;
        if (s < Q) {
            const size_t so = zf ? (size_t)(sm.kl[s / (TS / RKC)] * (TS / RKC) + s % (TS / RKC)) : (size_t)s;
            block_load_async<RCH * 8>(dst, srcA + so * RCH, tid);
            block_load_async<RCH * 8>(dst + RCH, srcB + so * RCH, tid);
        } else if (s < Q + 3) {
            block_load_async<LCH * 8>(dst, srcL + (size_t)(s - Q) * LCH, tid);
        }
        cp_async_commit();
    };
    // block inverses of L_jj ride in the first commit group
    {
        const double *Dg = prm.dblk + ((size_t)b * nt + j) * DSIZE;
        static_assert(DSIZE * 8 % (16 * NTHREADS) == 0, "D in whole 16-byte chunks per thread");
        block_load_async<DSIZE * 8>(sm.D, Dg, tid);
    }
    __syncthreads();  // item scalars, k-list
    if (zf) Q = (TS / RKC) * sm.nk;
    // Inputs sorted by column sep_col (the host sorted them: the likelihood does not depend on the order of the
    // observations): every row of this block lies at or above every column, so the OU leaves on that column factor
    // into row and column parts - one exponential per thread here instead of 32 per leaf in the evaluation below.
    const SepCtx *sep = nullptr;
    if (prm.sep_col >= 0) {
        int ns = 0, lf[2] = {0, 0};
        for (int f = 0; f < P.n_factors; ++f)
            if (P.f[f].kind == F_OU && P.f[f].col == prm.sep_col && ns < 2) lf[ns++] = f;
        const double *xc = X + (size_t)prm.sep_col * n;
        const int loc = tid & (TS - 1);
        const int r_first = i * TS < n ? i * TS : n - 1;
        const double c0 = xc[r_first];  // smallest row coordinate of the block
        const int idx = (tid < TS ? i : j) * TS + loc;
        const double x = xc[idx < n ? idx : n - 1];
        for (int q = 0; q < ns; ++q) {
            const double a = sm.sc.a[lf[q]];  // -1 / l
            if (tid < TS) sm.sep.u[q][loc] = fast_exp(a * (x - c0), sm.sc.etab);
            else sm.sep.v[q][loc] = fast_exp(a * (c0 - x), sm.sc.etab);
        }
        if (tid == 0) {
            sm.sep.n_sep = ns;
            sm.sep.leaf[0] = lf[0];
            sm.sep.leaf[1] = lf[1];
        }
        __syncthreads();
        sep = &sm.sep;
    }
    // (Measured and dropped: letting the warps whose 16 rows lie entirely in the identity padding of the last tile row skip
    // their arithmetic.  A CTA takes as long as its busiest warp, so n = 300 gained nothing, and the extra predicates in
    // the update loop cost the headline shape 2 %.)
    double acc[2][NCC];
    {
        int gi[2];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) gi[mb] = i * TS + row_of(tm, mb);
        // i > j: no entry of this tile is on the diagonal, so the cross-covariance form applies (Noise terms and the
        // diagonal bookkeeping drop out; rows / columns >= n read as 0 either way)
        double *const slot[4] = {sm.S, sm.S + LCH, sm.S + 2 * LCH, sm.S + 3 * LCH};
        // rows and columns that share no category under the program's Cat factors: the tile of K is zero whatever theta is
        if (zf && tile_cat_dead(P, sm.sc, X, n, prm.d, i, j, sm.S, tid)) acc_zero(acc);
        else eval_block_acc_scr<false, 4, GPL_LK_CW>(P, sm.sc, X, n, n, gi, X, n, n, j * TS, tm.t, 0.0, slot, tid, acc, 4, sep);
        __syncthreads();  // quarters read back: S is free for the ring
    }
#pragma unroll
    for (int s0 = 0; s0 < LK_NS - 1; ++s0) issue(s0);  // D rides in the first group
    for (int q = 0; q < Q; ++q) {
        cp_async_wait<LK_NS - 2>();
        __syncthreads();
        issue(q + LK_NS - 1);
        const double *a = sm.S + (q % LK_NS) * LCH;
        tile_mma<true>(acc, a, a + RCH, tm, 0, RKC);
    }
    // An exactly zero T_ij (rows and columns of different groups under a Cat(...) product, and nothing to subtract) stays
    // zero through the solve: store it, flag it, and leave the three L_jj stages unused.
    bool zero_tile = false;
    if (zf) {
        int nz = 0;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int cc = 0; cc < NCC; ++cc) nz |= (acc[mb][cc] != 0.0);
        zero_tile = !__syncthreads_or(nz);
        if (tid == 0) zf[tri_index(i, j)] = zero_tile ? 1 : 0;
    }
    if (zero_tile) {
        cp_async_wait<0>();  // the prefetched L_jj chunks and D
    } else {
    // L_ij = T_ij L_jj^-T, right-looking over the four 16-column panels; chunk c of L_jj is ring stage Q + c
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            cp_async_wait<LK_NS - 2>();
            __syncthreads();
            issue(Q + c + LK_NS - 1);  // nothing left to load: an empty group keeps the count in step
            const double *l = sm.S + ((Q + c) % LK_NS) * LCH;
            if (c == 0) {
                trsm_rl_solve<0>(acc, sm.D, tm);
                trsm_rl_update<0>(acc, l, tm);
            } else if (c == 1) {
                trsm_rl_solve<1>(acc, sm.D, tm);
                trsm_rl_update<1>(acc, l, tm);
            } else {
                trsm_rl_solve<2>(acc, sm.D, tm);
                trsm_rl_update<2>(acc, l, tm);
            }
        }
        trsm_rl_solve<3>(acc, sm.D, tm);
    }
    acc_to_tile(wsL + tri_index(i, j) * TILE_ELEMS, acc, tm);
    if (i == j + 1) {      // the tile just stored completes row j + 1: its diagonal tile can be formed now
        __syncthreads();   // the stores above are visible to the whole CTA; S is free
        diag_tile_phase(prm, sm, j + 1, b, tid);
    }
}

// ---- sorting the observations by one input column -----------------------------------------------------------------------
// One CTA, bitonic sort of (key, index) pairs in shared memory; ties keep no particular order (any order of equal
// coordinates is a valid sort).  npow2 = n rounded up to a power of two (<= 8192); padding keys are +inf.
__global__ void __launch_bounds__(1024) lk_sort_perm_kernel(const double *xcol, int n, int npow2, int *perm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *key = reinterpret_cast<double *>(smem_raw);
    int *idx = reinterpret_cast<int *>(key + npow2);
    for (int e = threadIdx.x; e < npow2; e += blockDim.x) {
        const double v = e < n ? xcol[e] : INFINITY;
        key[e] = v == v ? v : INFINITY;  // NaN coordinates sort last (the item's covariance is NaN whatever their place)
        idx[e] = e;
    }
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1)
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int e = threadIdx.x; e < npow2; e += blockDim.x) {
                const int o = e ^ jj;
                if (o > e) {
                    const bool up = (e & k) == 0;
                    const double a = key[e], b = key[o];
                    const int ia = idx[e], ib = idx[o];
                    const bool gt = a > b || (a == b && ia > ib);  // index as tie-break: a strict total order
                    if (gt == up) {
                        key[e] = b;
                        key[o] = a;
                        idx[e] = ib;
                        idx[o] = ia;
                    }
                }
            }
            __syncthreads();
        }
    for (int e = threadIdx.x; e < n; e += blockDim.x) perm[e] = idx[e];
}

__global__ void lk_permute_kernel(const double *in, double *out, const int *perm, int n, long long ncols, int inverse) {
    const long long total = (long long)n * ncols;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long c = e / n;
        const int i = (int)(e - c * n);
        if (inverse) out[c * n + perm[i]] = in[e];
        else out[e] = in[c * n + perm[i]];
    }
}

// ---- batched posteriors from the lockstep workspace ----------------------------------------------------------------
// After the factorisation of B items (L tiles, block inverses and z = L^-1 y in the workspace) one CTA per item forms
// what mean_and_var needs: the inverses of the diagonal tiles W_jj (identity solved against L_jj) and
// alpha = L^-T z by backward substitution over tiles, alpha_j = W_jj' (z_j - sum_{i>j} L_ij' alpha_i).
// Replaces posterior(fx, y) [upstream AbstractGPs] for every row of an MCMC chain at once (gaplac_b200/chain.py).
size_t lk_post_smem_bytes() { return (size_t)(TILE_ELEMS + DSIZE + 2 * TS) * sizeof(double); }

__global__ void __launch_bounds__(NTHREADS) lk_post_kernel(const __grid_constant__ LkPostParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *S = reinterpret_cast<double *>(smem_raw), *D = S + TILE_ELEMS, *rbuf = D + DSIZE, *abuf = rbuf + TS;
    const int tid = threadIdx.x, b = blockIdx.x, nt = prm.nt;
    const TMap tm = thread_map(tid);
    const long long ntri = tri_index(nt, 0);
    const double *L = prm.tiles + (size_t)b * ntri * TILE_ELEMS;
    const double *Dg = prm.dblk + (size_t)b * nt * DSIZE;
    const double *z = prm.z + (size_t)b * nt * TS;
    double *W = prm.winv + (size_t)b * nt * TILE_ELEMS;
    double *alpha = prm.alpha + (size_t)b * nt * TS;
    for (int j = 0; j < nt; ++j) {
        __syncthreads();
        tile_load_async(S, L + tri_index(j, j) * TILE_ELEMS, tid);
        block_load_async<DSIZE * 8>(D, Dg + (size_t)j * DSIZE, tid);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        double e[2][NCC];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int cc = 0; cc < NCC; ++cc) e[mb][cc] = (row_of(tm, mb) == col_of(tm, cc)) ? 1.0 : 0.0;
        tile_trsm_ld(e, S, D, tm);
        acc_to_tile_t(W + (size_t)j * TILE_ELEMS, e, tm);
    }
    for (int j = nt - 1; j >= 0; --j) {
        double rj = 0.0;
        if (tid < TS) rj = z[j * TS + tid];
        for (int i = nt - 1; i > j; --i) {
            __syncthreads();
            tile_load_async(S, L + tri_index(i, j) * TILE_ELEMS, tid);
            cp_async_commit();
            if (tid < TS) abuf[tid] = alpha[i * TS + tid];  // written by this thread in an earlier step
            cp_async_wait<0>();
            __syncthreads();
            if (tid < TS) rj -= tile_col_dot(S, abuf, tid);
        }
        __syncthreads();  // also orders the W stores above before this CTA's loads of them
        tile_load_async(S, W + (size_t)j * TILE_ELEMS, tid);
        cp_async_commit();
        if (tid < TS) rbuf[tid] = rj;
        cp_async_wait<0>();
        __syncthreads();
        if (tid < TS) alpha[j * TS + tid] = tile_col_dot(S, rbuf, tid);  // W upper part is zero
    }
}

}  // namespace gpl
