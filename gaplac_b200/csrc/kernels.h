// Kernel parameter blocks and host-side launchers (internal; the public surface is include/gaplac_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "mcmc_core.h"
#include "program.h"

namespace gpl {

// ---- batched fused lml (lml_batched.cu) ---------------------------------------------------------------------
struct LmlParams {
    DevProgram prog;
    int n, d, nt, B, p;
    int want_grad, keep;
    int sigma2_stride;
    long long x_stride, y_stride;  // doubles between items (0 = shared)
    const double *X, *Y, *Theta, *sigma2;
    double jitter;
    double *ws;            // per-CTA tile workspace
    long long ws_stride;   // doubles
    double *vec;           // per-CTA vectors: z and alpha, 2 * nt * 64 doubles
    double *lml, *dtheta, *dy;
    int *info;
    unsigned int *counter;  // dynamic work distribution (NULL: item = blockIdx.x)
};
__global__ void lml_batched_kernel(const __grid_constant__ LmlParams prm);       // lml (+ keep): 3 CTAs / SM
__global__ void lml_batched_grad_kernel(const __grid_constant__ LmlParams prm);  // lml + gradient: 2 CTAs / SM
size_t lml_smem_bytes(bool grad);

// ---- batched lml, lockstep schedule (lml_lockstep.cu): three kernels per tile column over the whole batch ------------
#ifndef GPL_LK_CW
#define GPL_LK_CW 4  // covariance entries per row and interpretation step in the lockstep kernels (4 or 8)
#endif
#define GPL_LK_KLMAX 512  // tile rows the lockstep kernels' column list has room for (n <= 32768)
#ifndef GPL_LK_ZMAX
#define GPL_LK_ZMAX 1024  // rows of z staged in shared memory by the diagonal-tile kernel
#endif
struct LkParams {
    DevProgram prog;
    int n, d, nt, p, j;
    int sigma2_stride;
    long long x_stride, y_stride;
    const double *X, *Y, *Theta, *sigma2;
    double jitter;
    double *tiles;  // B x ntri tiles (tile-major lower factor of every item)
    double *dblk;   // B x nt x DSIZE: block inverses of the diagonal tiles
    double *z;      // B x nt*64: right-hand side / z = L^-1 y
    int sep_col;    // >= 0: X is sorted by this column; its OU leaves (at most two) use the separable form below the diagonal
    int *zflag;     // B x ntri ints or NULL: 1 = the stored tile (i, j), i > j, of L is exactly zero (block-diagonal covariances:
                    // Cat(...) * k products on grouped rows); updates with such tiles and their triangular solves are skipped
};
// sorting the observations by one input column (the log marginal likelihood does not depend on their order)
__global__ void lk_sort_perm_kernel(const double *xcol, int n, int npow2, int *perm);           // 1 CTA, bitonic
// out[c*n + i] = in[c*n + perm[i]] (inverse: out[c*n + perm[i]] = in[c*n + i]) for c < ncols
__global__ void lk_permute_kernel(const double *in, double *out, const int *perm, int n, long long ncols, int inverse);
struct LkPotrfParams {
    int n, nt, j, B;
    double *tiles, *dblk, *z;
    double *acc2;  // B x 2: running z'z and logdet
    double *lml;
    int *info;
};
__global__ void lk_diag_kernel(const __grid_constant__ LkParams prm);        // grid B
__global__ void lk_below_kernel(const __grid_constant__ LkParams prm);       // grid B * (nt - 1 - j)
size_t lk_step_smem_bytes();
__global__ void lk_potrf_warp_kernel(const __grid_constant__ LkPotrfParams prm);  // one warp per item
size_t lk_potrf_warp_smem_bytes();
int lk_potrf_warp_items_per_cta();

// ---- analytic gradient on the lockstep schedule (lml_grad_lockstep.cu) -----------------------------------------------
// After the lockstep factorisation of a batch (L tiles, block inverses, z in the workspace) the gradient phases run as
// batch-wide launches as well: W_jj = L_jj^-1 for every diagonal tile, M = L^-1 tile row by tile row, alpha = M' z,
// then K^-1 = M' M tile by tile contracted on the fly with dK/dtheta.  M is stored as TRANSPOSED tiles in column-major
// tile order (tile (k, j) at col_index(nt, k, j)) so that every operand stream of these phases is one contiguous run.
__host__ __device__ inline long long col_index(int nt, int k, int j) { return (long long)j * nt - (long long)j * (j - 1) / 2 + (k - j); }
struct LkGradParams {
    DevProgram prog;
    int n, d, nt, p, i;  // i: tile row of the current lk_minv launch
    long long x_stride;
    const double *X, *Theta;
    const double *tiles, *dblk, *z;  // factor workspace of the lockstep schedule
    double *winv;                    // B x nt tiles: W_jj
    double *minv;                    // B x ntri tiles: (M_kj)' at col_index(nt, k, j)
    double *alpha;                   // B x nt*64: alpha = K^-1 y
    double *gpart;                   // B x ntri x p: per-tile partial sums of sum_ij W_ij dK_ij/dtheta
    double *dtheta, *dy;             // outputs (dy may be NULL)
    const int *info;
    int B;
    int sep_col;  // >= 0: X is sorted by this column (see LkParams): separable OU factors in the contraction below the diagonal
    const int *zflag;  // B x ntri or NULL: exactly-zero tiles of L (LkParams::zflag)
    int *mflag;        // B x ntri or NULL: exactly-zero tiles of M = L^-1 below the diagonal, written by lk_minv_kernel
};
__global__ void lk_winv_kernel(const __grid_constant__ LkGradParams prm);     // grid B * nt
__global__ void lk_minv_kernel(const __grid_constant__ LkGradParams prm);     // grid B * i  (row i, tiles j < i)
__global__ void lk_alpha_kernel(const __grid_constant__ LkGradParams prm);    // grid B * nt
__global__ void lk_gradc_kernel(const __grid_constant__ LkGradParams prm);    // grid B * ntri
__global__ void lk_gradsum_kernel(const __grid_constant__ LkGradParams prm);  // grid ceil(B / 128)
// the same three phases with zero-tile skipping (zflag / mflag set): separate instantiations, the dense ones stay as they were
__global__ void lk_minv_skip_kernel(const __grid_constant__ LkGradParams prm);
__global__ void lk_alpha_skip_kernel(const __grid_constant__ LkGradParams prm);
__global__ void lk_gradc_skip_kernel(const __grid_constant__ LkGradParams prm);
size_t lk_winv_smem_bytes();
size_t lk_grad_smem_bytes();

// ---- batched on-device sampler (mcmc.cu; per-chain logic in mcmc_core.h) -----------------------------------------------
struct McmcDevParams {
    McmcConfig cfg;
    int B, chain_offset, n_rec;
    ChainState *state;      // B
    double *vec;            // B x vec_stride: the vectors of every chain
    long long vec_stride;
    const double *q0;       // B x dim initial positions
    const double *Y;        // observations, chain b at b * y_stride (0: shared)
    long long y_stride;
    double *theta_eval;     // B x p: hyperparameters of the next evaluation
    double *y_eval;         // B x n: latent vectors of the next evaluation (latent model)
    const double *lml, *dtheta, *dy;  // results of the last evaluation
    const int *info;
    double *theta_out, *lp_out, *accept_out, *eps_out, *q_out;
    int *depth_out, *nleap_out, *div_out;
    unsigned int *done;     // chains that reached MC_DONE / MC_FAILED
    // Evaluation slots: the batched log-density call runs over n_slots items; slot s belongs to chain slot_chain[s].
    // theta_eval, y_eval, lml, dtheta, dy, info are indexed by SLOT, everything else by chain.  At start slot = chain;
    // once enough chains have finished the driver compacts the active chains into the first slots (mcmc_compact_kernel)
    // so that finished chains stop costing evaluations.
    int *slot_chain;
    int n_slots;
    int *n_active;          // device scalar written by mcmc_compact_kernel
};
__global__ void mcmc_init_kernel(const __grid_constant__ McmcDevParams prm);     // one warp per chain
__global__ void mcmc_advance_kernel(const __grid_constant__ McmcDevParams prm);  // one warp per slot
__global__ void mcmc_compact_kernel(const __grid_constant__ McmcDevParams prm);  // 1 CTA: slot_chain <- active chains, in order
__global__ void mcmc_reemit_kernel(const __grid_constant__ McmcDevParams prm);   // one warp per slot: pending point -> its new slot
// rows of a per-chain input (width doubles each) into slot order: dst[s] = src[slot_chain[s]]
__global__ void mcmc_gather_kernel(const double *src, double *dst, const int *slot_chain, int n_slots, long long width);
__global__ void mcmc_status_kernel(const ChainState *state, int B, int *status);
size_t mcmc_state_bytes();

// ---- covariance construction (kbuild.cu) ----------------------------------------------------------------------
struct CovParams {
    DevProgram prog;
    int na, nb, d, p;
    int same;  // 1: K(X,X) + diag_add I ; 0: cross-covariance K(Xa, Xb)
    const double *Xa, *Xb, *theta;
    double diag_add;
    double *K;  // na x nb column-major
};
__global__ void cov_dense_kernel(const __grid_constant__ CovParams prm);

// K_y of a program straight into the tile-major lower layout used by the large-n factorisation
struct CovTilesParams {
    DevProgram prog;
    int n, nt, d, p;
    const double *X, *theta;
    double diag_add;
    double *tiles;
};
__global__ void cov_tiles_kernel(const __grid_constant__ CovTilesParams prm);

// ---- layout conversion (big.cu) -----------------------------------------------------------------------------------
// dense column-major lower triangle (ld = n) -> tile-major lower, identity padding
__global__ void dense_to_tiles_kernel(const double *__restrict__ A, int n, int nt, double *__restrict__ tiles);
// tile-major lower factor L -> dense upper factor U = L' (zeros below the diagonal), column-major ld = n
__global__ void tiles_to_upper_kernel(const double *__restrict__ tiles, int n, int nt, double *__restrict__ U);

// ---- large-n blocked Cholesky on tile-major storage (big.cu) --------------------------------------------------------
struct BigParams {
    double *tiles;  // lower tiles, in place: A on entry, L on exit
    double *winv;   // nt tiles: inverses of the diagonal tiles
    double *dblk;   // nt x DSIZE: inverses of the 16 x 16 diagonal blocks (worker protocol; NULL otherwise)
    double *pivlog; // nt * 64 doubles: log of the pivots
    int *info;      // single int, first failing pivot (1-based) or 0
    double *y;      // optional padded right-hand side (nt*64): y on entry, z = L^-1 y on exit; NULL to skip
    int nt;
    int j;          // current tile column
    int k0;         // first tile column of the current panel (left-looking inside the panel)
    int j1;         // trailing update: one past the last tile column of the finished panel
    int l0, l1;     // trailing update: tile columns [l0, l1) are updated by this launch (all rows i >= l)
    int *flags;     // look-ahead worker protocol: panel_ready[BIG_MAXP] | rowdone[nt] | diagdone[nt] | tiledone[nt] | abort | prep[nt] | prepd[nt]
};
constexpr int BIG_MAXP = 64;  // panels the flag block has room for
#ifndef GPL_BIG_PANEL
#define GPL_BIG_PANEL 4
#endif
constexpr int BIG_PANEL = GPL_BIG_PANEL;  // tile columns per panel of the large-n factorisation
__global__ void big_diag_kernel(BigParams prm);   // 1 CTA
__global__ void big_col_kernel(BigParams prm);    // nt - j - 1 CTAs
__global__ void big_trail_kernel(BigParams prm);  // one CTA per trailing tile (i >= l, l0 <= l < l1)
// look-ahead protocol: one persistent CTA factors every diagonal tile in turn; the column kernel waits for it per column
__global__ void big_worker_kernel(BigParams prm);    // 1 CTA for the whole factorisation (owns an SM)
__global__ void big_col_flag_kernel(BigParams prm);
__global__ void big_worker2_kernel(BigParams prm);  // second version: the worker also solves tile (j+1, j) inside a panel
__global__ void big_col2_kernel(BigParams prm);
__global__ void big_winv_kernel(BigParams prm);  // grid nt: W_jj = L_jj^-1 from L_jj and its block inverses  // nt - j - 1 CTAs; spins on diagdone[j], publishes rowdone[i]
size_t big_smem_bytes();
size_t big_col_smem_bytes();
size_t big_trail_smem_bytes();

// ---- trailing update on the INT8 tensor path (trail_int8.cu; option "trail_int8") ------------------------------------------
constexpr int I8_BLOCK = 16;  // tile columns per deferred update: K = 1024
int i8_prepare();  // loads the kernels; must run before a spinning persistent kernel is resident
size_t i8_slices_bytes(int nt, int t0, int kt, int S);
size_t i8_scale_bytes(int nt, int t0);
// rows of L below tile row t0, tile columns [c0, c0 + kt) -> S slices (K-major) + row scales
int i8_split_tiles(const double *tiles, int nt, int t0, int c0, int kt, int S, signed char *slices, double *rowscale, cudaStream_t st);
// tiles(i, l) -= L_i L_l^T over those columns, for the 128-wide column blocks [cb0, cb1) of the region starting at t0
int i8_trail(double *tiles, int nt, int t0, int kt, int S, const signed char *slices, const double *rowscale, int cb0, int cb1,
             int max_ctas, int *info, int *dbg, cudaStream_t st);

// backward substitution alpha = L^-T z, one launch per tile row i (descending), grid = i + 1 CTAs; r = z on entry
// of the first launch and is updated in place, alpha receives the finished blocks
__global__ void big_backward_kernel(const double *tiles, const double *winv, int i, double *r, double *alpha);
// logdet = sum pivlog, quad = z'z -> out[0] = logdet, out[1] = quad
__global__ void big_reduce_kernel(const double *pivlog, const double *z, int len, double *out);

// ---- posterior prediction and sampling (predict.cu) ------------------------------------------------------------------
struct PredictParams {
    DevProgram prog;
    int n, nt, d, p, m;
    int want_var;
    const double *X, *theta, *Xs;  // Xs: m x d column-major
    const double *tiles, *winv, *alpha;
    double *wsV;          // per-CTA workspace, nt tiles
    double *mean, *var;
    // batched form (several posteriors over the same X, e.g. the rows of an MCMC chain): item b reads theta, tiles, winv,
    // alpha at b * stride and writes mean / var at b * m; `info` (optional) marks items whose factorisation failed
    int items;            // 0 or 1: single posterior
    long long theta_stride, tiles_stride, winv_stride, alpha_stride;
    const int *info;
};
// Batched posteriors on top of the lockstep workspace: per item b the inverses of the diagonal tiles (from L_jj and its
// block inverses) and alpha = L^-T z.  grid B.
struct LkPostParams {
    int nt;
    const double *tiles, *dblk, *z;  // lockstep workspace: B x ntri tiles, B x nt x DSIZE, B x nt*64
    double *winv;                    // B x nt tiles
    double *alpha;                   // B x nt*64
};
__global__ void lk_post_kernel(const __grid_constant__ LkPostParams prm);
size_t lk_post_smem_bytes();
__global__ void predict_kernel(const __grid_constant__ PredictParams prm);      // slabs of 64 test points
__global__ void predict_kernel_nb6(const __grid_constant__ PredictParams prm);  // 48
__global__ void predict_kernel_nb4(const __grid_constant__ PredictParams prm);  // 32
size_t predict_smem_bytes();

// out (n x S) = L Z
__global__ void sample_kernel(const double *tiles, int nt, int n, const double *Z, int S, double *out);

}  // namespace gpl
