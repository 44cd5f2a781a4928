// C ABI of libgaplac_b200.so (include/gaplac_b200.h): contexts, workspace pool, launch orchestration.
// No torch types, no C++ types across the boundary; every entry point returns a gpl_status.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/gaplac_b200.h"
#include "kernels.h"
#include "tile.cuh"

using namespace gpl;

namespace {

constexpr double LOG2PI = 1.8378770664093454835606594728112;
constexpr int PANEL = BIG_PANEL;  // tile columns per panel of the large-n factorisation (kernels.h)
constexpr int SMALL_MAX_N = 128;  // posterior fits up to this n run on the one-CTA fused kernel (n-sweep, profiles/sweep_large_r02.txt: the
                                  // panel path is faster from n = 256 on: 0.21 vs 0.29 ms there, 0.38 vs 0.91 ms at n = 512)

thread_local std::string g_last_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct gpl_ctx {
    int device = 0;
    int sm_count = 0;
    int clock_khz = 0;
    char name[128] = {0};
    cudaStream_t stream = nullptr;      // the stream host entry points run on (own_stream, or the caller's: gpl_set_stream)
    cudaStream_t own_stream = nullptr;
    cudaStream_t s_panel = nullptr, s_trail = nullptr, s_worker = nullptr, s_i8 = nullptr;  // look-ahead streams of the large-n factorisation
    DevBuf bigFlags, bigD, lkW, lkAlpha, dStage, i8Slices, i8Scale;
    void *hStage = nullptr;  // pinned host staging block of the small-call path (gpl_lml_batched)
    std::vector<gpl_post *> livePosts;  // posteriors created on this context and not yet freed (gpl_destroy detaches them)
    cudaEvent_t wsEvent = nullptr;      // completion of the last call that used the shared workspace (any stream)
    bool wsEventSet = false;
    std::vector<std::pair<void *, size_t>> postFree;  // device blocks of freed posteriors (cudaMalloc / cudaFree cost
                                                      // milliseconds next to multi-GB workspaces: a refit reuses them)
    uint64_t launches = 0;
    std::string err;
    std::mutex mu;
    int lml_variant = 0;
    int chol_variant = 0;
    bool attr_sort = false;
    bool attr_lml = false, attr_big = false, attr_pred = false, attr_lk = false, attr_post = false;
    size_t lk_ws_limit = (size_t)24 << 30;  // lockstep workspace cap in bytes; larger batches run in chunks
    int ou_separable = 1;                   // 1: sort the observations by the OU column and use the separable form (lockstep lml)
    int trail_int8 = -1;                    // large-n trailing updates: -1 auto (INT8 split path, 9 slices, from n = 6144 on), 0 FP64 DMMA only,
                                            // 5..9: that many slices from n = 4096 on
    int zero_tile_skip = 1;                 // lockstep factorisation: skip updates with / solves of exactly-zero tiles: 1 when the program
                                            // can produce them (a Cat factor in every term but the noise), 2 always, 0 never,
                                            // 3: as 1, and group the observations by the shared category column first
    int poison_ws = 0;                      // 1: fill the whole workspace with NaN payloads before every call (hygiene tests)
    int profile_events = 0;                 // 1: time every lockstep launch with CUDA events (bench.py roofline pass)
    double lk_ms[7] = {0, 0, 0, 0, 0, 0, 0};  // last instrumented call: total ms in diag / potrf / below / winv / minv /
    int lk_launches[7] = {0, 0, 0, 0, 0, 0, 0};  // alpha / contraction kernels
    // grow-only device buffers
    DevBuf lkTiles, lkD, lkZ, lkAcc, lkM, lkGpart, lkPerm, lkXs, lkYs, lkDyS, lkZero, lkMzero;
    DevBuf ws, vec, counter, bX, bY, bTheta, bSigma, bLml, bDtheta, bDy, bInfo, bMisc, bK, bXs, bMean, bVar, bWsV;
};

struct gpl_prog {
    DevProgram dev;
};

struct gpl_post {
    gpl_ctx *ctx = nullptr;
    DevProgram prog;
    int n = 0, d = 0, nt = 0, p = 0;
    double *dX = nullptr, *dtheta = nullptr;
    double *tiles = nullptr, *winv = nullptr, *alpha = nullptr;
    void *block = nullptr;  // one device allocation behind dX | dtheta | tiles | alpha (recycled through the context)
    size_t block_bytes = 0;
    double lml = 0.0;
};

namespace {

int fail(gpl_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}

#define CU(ctx, call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(ctx, GPL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                            \
    } while (0)

// The context owns ONE grow-only workspace.  Calls may arrive on different streams (the *_dev entry points take the
// caller's stream): every call that touches the workspace first makes its stream wait for the previous such call and
// records its own completion when it has enqueued everything, so that workspace reuse is ordered across streams.
inline void all_buffers(gpl_ctx *ctx, std::vector<DevBuf *> &out) {
    out = {&ctx->lkZero, &ctx->lkMzero, &ctx->i8Slices, &ctx->i8Scale, &ctx->bigFlags, &ctx->bigD, &ctx->lkW, &ctx->lkAlpha, &ctx->dStage, &ctx->lkTiles, &ctx->lkD, &ctx->lkZ, &ctx->lkAcc,
           &ctx->lkM, &ctx->lkGpart, &ctx->lkPerm, &ctx->lkXs, &ctx->lkYs, &ctx->lkDyS, &ctx->ws, &ctx->vec, &ctx->counter, &ctx->bX, &ctx->bY, &ctx->bTheta, &ctx->bSigma,
           &ctx->bLml, &ctx->bDtheta, &ctx->bDy, &ctx->bInfo, &ctx->bMisc, &ctx->bK, &ctx->bXs, &ctx->bMean, &ctx->bVar, &ctx->bWsV};
}
struct WsOrder {
    gpl_ctx *c;
    cudaStream_t st;
    WsOrder(gpl_ctx *ctx, cudaStream_t s) : c(ctx), st(s) {
        if (c->wsEventSet) cudaStreamWaitEvent(st, c->wsEvent, 0);
        if (c->poison_ws) {  // every byte 0xFF = NaN payloads: a kernel that reads workspace it did not write shows up in the results
            std::vector<DevBuf *> bufs;
            all_buffers(c, bufs);
            for (DevBuf *b : bufs)
                if (b->p) cudaMemsetAsync(b->p, 0xFF, b->cap, st);
        }
    }
    ~WsOrder() {
        if (c->wsEvent && cudaEventRecord(c->wsEvent, st) == cudaSuccess) c->wsEventSet = true;
    }
};

int ensure(gpl_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return GPL_OK;
    if (b.p) CU(ctx, cudaFree(b.p));  // (cudaFree synchronises the device: nothing in flight can still read the old block)
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    CU(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return GPL_OK;
}

template <typename T>
T *ptr(DevBuf &b) {
    return static_cast<T *>(b.p);
}

int check_prog_args(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, int p) {
    if (!ctx || !prog) return fail(ctx, GPL_ERR_ARG, "null context or program");
    if (n <= 0 || d <= 0) return fail(ctx, GPL_ERR_ARG, "n=%d d=%d must be positive", n, d);
    if (d < prog->dev.n_cols) return fail(ctx, GPL_ERR_ARG, "program reads column %d but d=%d", prog->dev.n_cols - 1, d);
    if (p < prog->dev.n_theta) return fail(ctx, GPL_ERR_ARG, "program uses %d hyperparameter slots but p=%d", prog->dev.n_theta, p);
    if (p > GPL_MAX_THETA) return fail(ctx, GPL_ERR_LIMIT, "p=%d exceeds %d", p, GPL_MAX_THETA);
    return GPL_OK;
}

// ---- launch helpers ------------------------------------------------------------------------------------------
int launch_lml_lockstep(gpl_ctx *ctx, const DevProgram &prog, int n, int d, const double *dX, int x_batched,
                        const double *dY, int y_batched, const double *dTheta, int p, const double *dsigma2,
                        int sigma2_batched, double jitter, int B, double *dlml, int *dinfo, cudaStream_t st,
                        double *ddtheta = nullptr, double *ddy = nullptr, int want_grad = 0, int allow_sort = 1);
int launch_lml(gpl_ctx *ctx, const DevProgram &prog, int n, int d, const double *dX, int x_batched, const double *dY,
               int y_batched, const double *dTheta, int p, const double *dsigma2, int sigma2_batched, double jitter,
               int B, double *dlml, double *ddtheta, double *ddy, int *dinfo, int want_grad, int keep,
               double *keep_ws, double *keep_vec, cudaStream_t st) {
    const int nt = (n + TS - 1) / TS;
    // (Measured and dropped, tools/small_n_ab.py: sending value + gradient of one-tile models (n <= 64) to the fused per-item
    // kernel - one launch instead of six.  45 vs 58 us for B <= 64, but 0.146 vs 0.110 ms at B = 1024 and 2.0 vs 1.3 ms at
    // B = 16384: the README chain gains 11 %, every larger batch loses; an item's bits would also depend on the kernel choice.)
    if (!keep && ctx->lml_variant != 1)
        return launch_lml_lockstep(ctx, prog, n, d, dX, x_batched, dY, y_batched, dTheta, p, dsigma2, sigma2_batched, jitter, B,
                                   dlml, dinfo, st, ddtheta, ddy, want_grad);
    const long long ntri = tri_index(nt, 0);
    const long long tiles_per_cta = ntri + nt + (want_grad ? ntri : 0);
    const size_t smem = lml_smem_bytes(want_grad != 0);
    bool &attr_set = ctx->attr_lml;
    if (!attr_set) {
        CU(ctx, cudaFuncSetAttribute(lml_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)lml_smem_bytes(false)));
        CU(ctx, cudaFuncSetAttribute(lml_batched_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)lml_smem_bytes(true)));
        attr_set = true;
    }
    int occ = 0;
    if (want_grad) CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lml_batched_grad_kernel, NTHREADS, smem));
    else CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lml_batched_kernel, NTHREADS, smem));
    if (occ < 1) return fail(ctx, GPL_ERR_CUDA, "lml kernel does not fit on an SM (smem %zu)", smem);
    int grid = ctx->sm_count * occ;
    if (grid > B) grid = B;
    LmlParams prm;
    prm.prog = prog;
    prm.n = n;
    prm.d = d;
    prm.nt = nt;
    prm.B = B;
    prm.p = p;
    prm.want_grad = want_grad;
    prm.keep = keep;
    prm.sigma2_stride = sigma2_batched ? 1 : 0;
    prm.x_stride = x_batched ? (long long)n * d : 0;
    prm.y_stride = y_batched ? n : 0;
    prm.X = dX;
    prm.Y = dY;
    prm.Theta = dTheta;
    prm.sigma2 = dsigma2;
    prm.jitter = jitter;
    prm.ws_stride = tiles_per_cta * TILE_ELEMS;
    if (keep) {
        prm.ws = keep_ws;
        prm.vec = keep_vec;
        grid = 1;
    } else {
        int rc = ensure(ctx, ctx->ws, (size_t)grid * prm.ws_stride * sizeof(double));
        if (rc) return rc;
        rc = ensure(ctx, ctx->vec, (size_t)grid * 2 * nt * TS * sizeof(double));
        if (rc) return rc;
        prm.ws = ptr<double>(ctx->ws);
        prm.vec = ptr<double>(ctx->vec);
    }
    int rc = ensure(ctx, ctx->counter, sizeof(unsigned int));
    if (rc) return rc;
    prm.counter = ptr<unsigned int>(ctx->counter);
    CU(ctx, cudaMemsetAsync(prm.counter, 0, sizeof(unsigned int), st));
    prm.lml = dlml;
    prm.dtheta = ddtheta;
    prm.dy = ddy;
    prm.info = dinfo;
    if (want_grad) lml_batched_grad_kernel<<<grid, NTHREADS, smem, st>>>(prm);
    else lml_batched_kernel<<<grid, NTHREADS, smem, st>>>(prm);
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    return GPL_OK;
}

// lockstep schedule: 2 kernels per tile column over the batch; with want_grad the gradient phases of
// lml_grad_lockstep.cu follow (W_jj, M = L^-1 row by row, alpha, K^-1 tiles contracted with dK/dtheta)
int launch_lml_lockstep(gpl_ctx *ctx, const DevProgram &prog, int n, int d, const double *dX, int x_batched,
                        const double *dY, int y_batched, const double *dTheta, int p, const double *dsigma2,
                        int sigma2_batched, double jitter, int B, double *dlml, int *dinfo, cudaStream_t st,
                        double *ddtheta, double *ddy, int want_grad, int allow_sort) {
    const int nt = (n + TS - 1) / TS;
    const long long ntri = tri_index(nt, 0);
    if (!ctx->attr_lk) {
        CU(ctx, cudaFuncSetAttribute(lk_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lk_step_smem_bytes()));
        CU(ctx, cudaFuncSetAttribute(lk_below_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lk_step_smem_bytes()));
        CU(ctx, cudaFuncSetAttribute(lk_potrf_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)lk_potrf_warp_smem_bytes()));
        CU(ctx, cudaFuncSetAttribute(lk_winv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lk_winv_smem_bytes()));
        CU(ctx, cudaFuncSetAttribute(lk_minv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES));
        CU(ctx, cudaFuncSetAttribute(lk_gradc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lk_grad_smem_bytes()));
        CU(ctx, cudaFuncSetAttribute(lk_minv_skip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES));
        CU(ctx, cudaFuncSetAttribute(lk_gradc_skip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lk_grad_smem_bytes()));
        ctx->attr_lk = true;
    }
    const size_t grad_item = want_grad ? (size_t)(ntri + nt) * TILE_BYTES + (size_t)nt * TS * 8 + (size_t)ntri * (p > 0 ? p : 1) * 8 : 0;
    const size_t per_item = (size_t)ntri * TILE_BYTES + (size_t)nt * DSIZE * 8 + (size_t)nt * TS * 8 + 16 + grad_item;
    int Bc = (int)(ctx->lk_ws_limit / per_item);
    if (Bc < 1) Bc = 1;
    if (Bc > B) Bc = B;
    int rc;
    if ((rc = ensure(ctx, ctx->lkTiles, (size_t)Bc * ntri * TILE_BYTES)) || (rc = ensure(ctx, ctx->lkD, (size_t)Bc * nt * DSIZE * 8)) ||
        (rc = ensure(ctx, ctx->lkZ, (size_t)Bc * nt * TS * 8)) || (rc = ensure(ctx, ctx->lkAcc, (size_t)Bc * 16)) ||
        (rc = ensure(ctx, ctx->lkZero, (size_t)Bc * ntri * sizeof(int))))
        return rc;
    if (nt > GPL_LK_KLMAX) return fail(ctx, GPL_ERR_LIMIT, "batched path: n = %d exceeds %d", n, GPL_LK_KLMAX * TS);
    if (want_grad && ((rc = ensure(ctx, ctx->lkM, (size_t)Bc * ntri * TILE_BYTES)) || (rc = ensure(ctx, ctx->lkW, (size_t)Bc * nt * TILE_BYTES)) ||
                      (rc = ensure(ctx, ctx->lkAlpha, (size_t)Bc * nt * TS * 8)) ||
                      (rc = ensure(ctx, ctx->lkGpart, (size_t)Bc * ntri * (p > 0 ? p : 1) * 8))))
        return rc;
    int *info_dev = dinfo;
    if (!info_dev) {
        if ((rc = ensure(ctx, ctx->bInfo, (size_t)B * 4))) return rc;
        info_dev = ptr<int>(ctx->bInfo);
    }
    CU(ctx, cudaMemsetAsync(info_dev, 0, (size_t)B * sizeof(int), st));
    // Separable OU leaves: when the program has one or two OU leaves on one input column and X is shared by the batch, the
    // observations are sorted by that column first (lml, dtheta are invariant; dy is scattered back at the end): every
    // block below the diagonal then has all its rows at or above all its columns and those leaves cost one multiply per
    // entry instead of an exponential (kfun.cuh SepCtx).  The decision depends on the model alone (n, the program), never on
    // the batch: an item's bits must not depend on how many items travel with it (parts of a multi-device call, the
    // shrinking batch of the sampler).  From four tile columns on the three small launches of the sort cost < 5 % of a call.
    // Exact zero tiles arise when every term but the noise carries a Cat(...) factor: K is block-diagonal once the rows are
    // grouped by a category column that all those terms share.  Such programs get the flags of LkParams::zflag; with
    // zero_tile_skip = 3 their observations are also grouped here (a stable sort by that column; rows that arrive grouped in
    // ascending category order keep their order and their bits) and the separable OU form is not used (rows of different
    // groups are not ordered along the OU column).
    bool cat_everywhere = prog.n_terms > 0;
    int group_col = -1;
    {
        unsigned long long common = ~0ull;  // category columns (< 64) present in every term but the noise
        int real_terms = 0;
        for (int t = 0; t < prog.n_terms; ++t) {
            if (prog.has_noise >> t & 1) continue;
            ++real_terms;
            unsigned long long cols = 0;
            bool has_cat = false;
            for (int f = prog.term_begin[t]; f < prog.term_begin[t + 1]; ++f)
                if (prog.f[f].kind == F_CAT) {
                    has_cat = true;
                    if (prog.f[f].col < 64) cols |= 1ull << prog.f[f].col;
                }
            cat_everywhere &= has_cat;
            common &= cols;
        }
        cat_everywhere &= real_terms > 0;
        if (cat_everywhere && common)
            for (int c = 0; c < 64 && group_col < 0; ++c)
                if (common >> c & 1) group_col = c;
    }
    const bool use_zflags = ctx->zero_tile_skip == 2 || (ctx->zero_tile_skip && cat_everywhere);
    int sep_col = -1, sort_col = -1;
    // grouping is opt-in (zero_tile_skip = 3): three extra launches and two permutations per call cost the sampler's small
    // batches 12 % (C3, rows already grouped), so by default the caller's row order is taken as it comes
    if (ctx->zero_tile_skip == 3 && cat_everywhere && group_col >= 0 && allow_sort && !x_batched && n <= 8192 && nt >= 2) sort_col = group_col;
    // (a program that gets the zero flags keeps the caller's row order - presumably grouped - instead of the OU sort, which
    // would interleave the groups: skipping most tiles is worth far more than one exponential per entry)
    else if (!(use_zflags && cat_everywhere) && ctx->ou_separable && allow_sort && !x_batched && n <= 8192 &&
             (nt >= 4 || (ctx->ou_separable == 2 && nt >= 2))) {
        int cnt = 0;
        for (int f = 0; f < prog.n_factors; ++f)
            if (prog.f[f].kind == F_OU) {
                if (sep_col < 0) sep_col = prog.f[f].col;
                if (prog.f[f].col == sep_col) ++cnt;
            }
        if (cnt < 1 || cnt > 2) sep_col = -1;
        sort_col = sep_col;
    }
    double *ddy_user = ddy;
    if (sort_col >= 0) {
        const size_t ny = (size_t)n * (y_batched ? B : 1);
        if ((rc = ensure(ctx, ctx->lkPerm, (size_t)n * 4)) || (rc = ensure(ctx, ctx->lkXs, (size_t)n * d * 8)) ||
            (rc = ensure(ctx, ctx->lkYs, ny * 8)) || (ddy && (rc = ensure(ctx, ctx->lkDyS, (size_t)n * B * 8))))
            return rc;
        int npow2 = 1;
        while (npow2 < n) npow2 <<= 1;
        const size_t sort_smem = (size_t)npow2 * 12;
        if (!ctx->attr_sort) {
            CU(ctx, cudaFuncSetAttribute(lk_sort_perm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12));
            ctx->attr_sort = true;
        }
        int *perm = ptr<int>(ctx->lkPerm);
        lk_sort_perm_kernel<<<1, npow2 < 1024 ? (npow2 < 32 ? 32 : npow2) : 1024, sort_smem, st>>>(dX + (size_t)sort_col * n, n, npow2, perm);
        lk_permute_kernel<<<64, 256, 0, st>>>(dX, ptr<double>(ctx->lkXs), perm, n, d, 0);
        lk_permute_kernel<<<ny > 65536 ? 592 : 64, 256, 0, st>>>(dY, ptr<double>(ctx->lkYs), perm, n, (long long)(ny / n), 0);
        ctx->launches += 3;
        dX = ptr<double>(ctx->lkXs);
        dY = ptr<double>(ctx->lkYs);
        if (ddy) ddy = ptr<double>(ctx->lkDyS);
    }
    LkParams prm;
    prm.sep_col = sep_col;
    prm.zflag = use_zflags ? ptr<int>(ctx->lkZero) : nullptr;
    prm.prog = prog;
    prm.n = n;
    prm.d = d;
    prm.nt = nt;
    prm.p = p;
    prm.sigma2_stride = sigma2_batched ? 1 : 0;
    prm.x_stride = x_batched ? (long long)n * d : 0;
    prm.y_stride = y_batched ? n : 0;
    prm.jitter = jitter;
    prm.tiles = ptr<double>(ctx->lkTiles);
    prm.dblk = ptr<double>(ctx->lkD);
    prm.z = ptr<double>(ctx->lkZ);
    LkPotrfParams pp;
    pp.n = n;
    pp.nt = nt;
    pp.tiles = prm.tiles;
    pp.dblk = prm.dblk;
    pp.z = prm.z;
    pp.acc2 = ptr<double>(ctx->lkAcc);
    std::vector<cudaEvent_t> evs;
    std::vector<int> ev_kind;
    LkGradParams gp;
    gp.prog = prog;
    gp.n = n;
    gp.d = d;
    gp.nt = nt;
    gp.p = p;
    gp.i = 0;
    gp.x_stride = prm.x_stride;
    gp.tiles = prm.tiles;
    gp.dblk = prm.dblk;
    gp.z = prm.z;
    gp.winv = ptr<double>(ctx->lkW);
    gp.minv = ptr<double>(ctx->lkM);
    gp.alpha = ptr<double>(ctx->lkAlpha);
    gp.gpart = ptr<double>(ctx->lkGpart);
    gp.sep_col = sep_col;
    gp.zflag = prm.zflag;
    gp.mflag = nullptr;
    if (want_grad && prm.zflag) {
        if ((rc = ensure(ctx, ctx->lkMzero, (size_t)Bc * ntri * sizeof(int)))) return rc;
        gp.mflag = ptr<int>(ctx->lkMzero);
        gp.zflag = ptr<int>(ctx->lkZero);  // (ensure may have moved nothing: same buffer as prm.zflag)
    }
    auto mark = [&](int kind) {  // kind: 0 diag, 1 potrf, 2 below, 3..6 gradient phases (winv, minv, alpha, contraction), -1 start
        if (!ctx->profile_events) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        evs.push_back(e);
        ev_kind.push_back(kind);
    };
    for (int off = 0; off < B; off += Bc) {
        const int nb = (B - off < Bc) ? B - off : Bc;
        prm.X = dX + (size_t)off * prm.x_stride;
        prm.Y = dY + (size_t)off * prm.y_stride;
        prm.Theta = dTheta + (size_t)off * p;
        prm.sigma2 = dsigma2 + (size_t)off * prm.sigma2_stride;
        pp.lml = dlml + off;
        pp.info = info_dev + off;
        mark(-1);
        for (int j = 0; j < nt; ++j) {
            prm.j = j;
            pp.j = j;
            if (j == 0) {  // later diagonal tiles are formed by lk_below_kernel of the previous column
                lk_diag_kernel<<<nb, NTHREADS, lk_step_smem_bytes(), st>>>(prm);
                ctx->launches++;
                mark(0);
            }
            pp.B = nb;
            {
                // up to 7 GPs per CTA (one warp each); small batches are spread over the SMs instead
                const int ipc_max = lk_potrf_warp_items_per_cta();
                int ipc = (nb + ctx->sm_count - 1) / ctx->sm_count;
                ipc = ipc < 1 ? 1 : (ipc > ipc_max ? ipc_max : ipc);
                lk_potrf_warp_kernel<<<(nb + ipc - 1) / ipc, 32 * ipc, lk_potrf_warp_smem_bytes() / ipc_max * ipc, st>>>(pp);
            }
            mark(1);
            ctx->launches++;
            if (j + 1 < nt) {
                lk_below_kernel<<<(unsigned)((size_t)nb * (nt - 1 - j)), NTHREADS, lk_step_smem_bytes(), st>>>(prm);
                mark(2);
                ctx->launches++;
            }
        }
        if (want_grad) {
            gp.X = prm.X;
            gp.Theta = prm.Theta;
            gp.info = pp.info;
            gp.B = nb;
            gp.dtheta = ddtheta ? ddtheta + (size_t)off * p : nullptr;
            gp.dy = ddy ? ddy + (size_t)off * n : nullptr;
            lk_winv_kernel<<<(unsigned)((size_t)nb * nt), NTHREADS, lk_winv_smem_bytes(), st>>>(gp);
            ctx->launches++;
            mark(3);
            for (int i = 1; i < nt; ++i) {
                gp.i = i;
                if (gp.mflag) lk_minv_skip_kernel<<<(unsigned)((size_t)nb * i), NTHREADS, TILE_BYTES, st>>>(gp);
                else lk_minv_kernel<<<(unsigned)((size_t)nb * i), NTHREADS, TILE_BYTES, st>>>(gp);
                ctx->launches++;
                mark(4);
            }
            if (gp.mflag) lk_alpha_skip_kernel<<<(unsigned)((size_t)nb * nt), NTHREADS, 0, st>>>(gp);
            else lk_alpha_kernel<<<(unsigned)((size_t)nb * nt), NTHREADS, 0, st>>>(gp);
            ctx->launches++;
            mark(5);
            if (p > 0 && ddtheta) {
                if (gp.mflag) lk_gradc_skip_kernel<<<(unsigned)((size_t)nb * ntri), NTHREADS, lk_grad_smem_bytes(), st>>>(gp);
                else lk_gradc_kernel<<<(unsigned)((size_t)nb * ntri), NTHREADS, lk_grad_smem_bytes(), st>>>(gp);
                lk_gradsum_kernel<<<(nb + 127) / 128, 128, 0, st>>>(gp);
                ctx->launches += 2;
                mark(6);
            }
        }
    }
    if (sort_col >= 0 && ddy_user) {  // dlml/dy back to the caller's order of the observations
        lk_permute_kernel<<<592, 256, 0, st>>>(ddy, ddy_user, ptr<int>(ctx->lkPerm), n, B, 1);
        ctx->launches++;
    }
    CU(ctx, cudaGetLastError());
    if (ctx->profile_events) {
        CU(ctx, cudaStreamSynchronize(st));
        for (int k = 0; k < 7; ++k) ctx->lk_ms[k] = 0.0, ctx->lk_launches[k] = 0;
        for (size_t e = 1; e < evs.size(); ++e) {
            if (ev_kind[e] >= 0) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, evs[e - 1], evs[e]);
                ctx->lk_ms[ev_kind[e]] += ms;
                ctx->lk_launches[ev_kind[e]]++;
            }
        }
        for (cudaEvent_t e : evs) cudaEventDestroy(e);
    }
    return GPL_OK;
}

// large-n factorisation of tile-major `tiles` (in place).  y (padded, nt*64) optional: becomes z = L^-1 y.
int big_factor(gpl_ctx *ctx, double *tiles, double *winv, double *pivlog, int *dinfo, double *y, int nt,
               cudaStream_t st) {
    const size_t smem = big_smem_bytes();
    bool &attr_set = ctx->attr_big;
    if (!attr_set) {
        CU(ctx, cudaFuncSetAttribute(big_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CU(ctx, cudaFuncSetAttribute(big_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(ctx, cudaFuncSetAttribute(big_trail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)big_trail_smem_bytes()));
        attr_set = true;
    }
    CU(ctx, cudaMemsetAsync(dinfo, 0, sizeof(int), st));
    BigParams prm;
    prm.tiles = tiles;
    prm.winv = winv;
    prm.dblk = nullptr;
    prm.pivlog = pivlog;
    prm.info = dinfo;
    prm.y = y;
    prm.nt = nt;
    prm.l0 = prm.l1 = 0;
    prm.flags = nullptr;
    const int NP = (nt + PANEL - 1) / PANEL;
    auto cols_tiles = [&](int l0, int l1) {  // number of tiles (i, l), l0 <= l < l1, l <= i < nt
        long long c = 0;
        for (int l = l0; l < l1; ++l) c += nt - l;
        return c;
    };
    auto factor_panel = [&](int P, cudaStream_t s, size_t diag_smem) {
        const int k0 = P * PANEL, j1 = (k0 + PANEL < nt) ? k0 + PANEL : nt;
        prm.k0 = k0;
        prm.j1 = j1;
        for (int j = k0; j < j1; ++j) {
            prm.j = j;
            big_diag_kernel<<<1, NTHREADS, diag_smem, s>>>(prm);
            ctx->launches++;
            if (j + 1 < nt) {
                big_col_kernel<<<nt - j - 1, NTHREADS, smem, s>>>(prm);
                ctx->launches++;
            }
        }
    };
    auto trail = [&](int P, int l0, int l1, cudaStream_t s) {  // update tile columns [l0, l1) with panel P
        if (l0 >= l1) return;
        prm.k0 = P * PANEL;
        prm.j1 = prm.k0 + PANEL;
        prm.l0 = l0;
        prm.l1 = l1;
        big_trail_kernel<<<(unsigned)cols_tiles(l0, l1), NTHREADS, big_trail_smem_bytes(), s>>>(prm);
        ctx->launches++;
    };
    if (NP <= 2 || ctx->chol_variant == 2) {
        // small problems: one stream, no look-ahead
        for (int P = 0; P < NP; ++P) {
            factor_panel(P, st, smem);
            trail(P, (P + 1) * PANEL, nt, st);
        }
        CU(ctx, cudaGetLastError());
        return GPL_OK;
    }
    // Look-ahead (depth 1).  A persistent worker CTA (big_worker_kernel, its own high-priority stream, a whole SM of
    // shared memory) factors the diagonal tiles in turn; s_panel carries the column kernels (they spin on the worker's
    // flag per column) and the update of the next panel's columns; s_trail carries the rest of the trailing update.
    if (!ctx->s_panel) {
        int lo = 0, hi = 0;
        CU(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(ctx, cudaStreamCreateWithPriority(&ctx->s_worker, cudaStreamNonBlocking, hi));
        CU(ctx, cudaStreamCreateWithPriority(&ctx->s_panel, cudaStreamNonBlocking, hi));
        CU(ctx, cudaStreamCreateWithPriority(&ctx->s_trail, cudaStreamNonBlocking, hi < lo - 1 ? lo - 1 : lo));
        CU(ctx, cudaStreamCreateWithPriority(&ctx->s_i8, cudaStreamNonBlocking, lo));
        CU(ctx, cudaFuncSetAttribute(big_worker_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CU(ctx, cudaFuncSetAttribute(big_col_flag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)big_col_smem_bytes()));
        CU(ctx, cudaFuncSetAttribute(big_worker2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CU(ctx, cudaFuncSetAttribute(big_col2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_col_smem_bytes()));
        CU(ctx, cudaFuncSetAttribute(big_winv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const bool use_worker = NP <= BIG_MAXP && nt <= 140 && ctx->chol_variant != 3;
    const size_t nflags = (size_t)BIG_MAXP + 5 * (size_t)nt + 1;  // ... + the abort flag of the bounded waits + prep[nt] + prepd[nt]
    int rc = ensure(ctx, ctx->bigFlags, nflags * sizeof(int));
    if (rc) return rc;
    prm.flags = ptr<int>(ctx->bigFlags);
    if ((rc = ensure(ctx, ctx->bigD, (size_t)nt * DSIZE * sizeof(double)))) return rc;
    prm.dblk = ptr<double>(ctx->bigD);
    CU(ctx, cudaMemsetAsync(prm.flags, 0, nflags * sizeof(int), st));
    const size_t diag_smem = 200 * 1024;
    std::vector<cudaEvent_t> eF(NP), eB(NP);
    for (int P = 0; P < NP; ++P) {
        CU(ctx, cudaEventCreateWithFlags(&eF[P], ctx->profile_events == 2 ? cudaEventDefault : cudaEventDisableTiming));
        CU(ctx, cudaEventCreateWithFlags(&eB[P], ctx->profile_events == 2 ? cudaEventDefault : cudaEventDisableTiming));
    }
    cudaEvent_t e0, eW;
    CU(ctx, cudaEventCreateWithFlags(&e0, ctx->profile_events == 2 ? cudaEventDefault : cudaEventDisableTiming));
    CU(ctx, cudaEventCreateWithFlags(&eW, cudaEventDisableTiming));
    CU(ctx, cudaEventRecord(e0, st));
    CU(ctx, cudaStreamWaitEvent(ctx->s_panel, e0, 0));
    CU(ctx, cudaStreamWaitEvent(ctx->s_trail, e0, 0));
    auto panel_cols = [&](int P) -> int {  // column kernels of panel P under the worker protocol
        const int k0 = P * PANEL, j1 = (k0 + PANEL < nt) ? k0 + PANEL : nt;
        prm.k0 = k0;
        prm.j1 = j1;
        CU(ctx, cudaMemsetAsync(prm.flags + P, 1, sizeof(int), ctx->s_panel));  // panel_ready[P]
        for (int j = k0; j < j1 && j + 1 < nt; ++j) {
            prm.j = j;
            if (ctx->chol_variant == 4) big_col_flag_kernel<<<nt - j - 1, NTHREADS, big_col_smem_bytes(), ctx->s_panel>>>(prm);
            else  // + one CTA that prepares the next diagonal tile when column j + 1 lies in the same panel
                big_col2_kernel<<<nt - j - 1 + (j + 1 < j1 ? 1 : 0), NTHREADS, big_col_smem_bytes(), ctx->s_panel>>>(prm);
            ctx->launches++;
        }
        return (int)GPL_OK;
    };
    // Option "trail_int8" (S slices): the update of everything beyond the current block of I8_BLOCK tile columns is deferred
    // until the block is complete and then applied in one pass on the INT8 tensor path (K = 1024); inside the block the
    // panels are applied with DMMA as before.  One SM stays with the worker CTA.
    constexpr int i8_block = I8_BLOCK;  // tile columns per deferred INT8 pass (B200 sweep, n = 8192: 8 -> 6.5 ms, 16 -> 6.35, 32 -> 6.6)
    static_assert(I8_BLOCK % BIG_PANEL == 0 && I8_BLOCK % 2 == 0, "blocks are whole panels and whole 128-column blocks");
    int i8S = 0;
    if (ctx->trail_int8 > 0 && nt >= 4 * i8_block) i8S = ctx->trail_int8;
    else if (ctx->trail_int8 < 0 && nt >= 6 * i8_block) i8S = 9;
    // 9 slices (63 bits below each row's maximum) match the FP64 path on ill-conditioned covariances as well: SqExp l = 3 +
    // 1e-6 noise at n = 6144 deviates from LAPACK by 1.5e-10 (DMMA path 2.2e-10); 8 slices give 4.3e-9 there (they are within
    // 1e-14 of the FP64 path on well-conditioned problems such as C5, and 10-20 % faster: option trail_int8 = 8).
    // Measured against DMMA with 9 slices: n = 6144 1.02x, 7168 1.11x, 8192 1.19x, 12288 1.36x, 16384 1.42x.
    int *i8dbg = nullptr;
    if (i8S && i8_prepare()) {
        if (ctx->trail_int8 > 0) return fail(ctx, GPL_ERR_CUDA, "INT8 trailing update: kernels or cuTensorMapEncodeTiled unavailable");
        i8S = 0;  // automatic mode: stay on the FP64 tensor path
    }
    if (i8S) {
        if ((rc = ensure(ctx, ctx->i8Slices, i8_slices_bytes(nt, i8_block, i8_block, i8S)))) return rc;
        if ((rc = ensure(ctx, ctx->i8Scale, i8_scale_bytes(nt, i8_block) + 4 * sizeof(int)))) return rc;
        i8dbg = reinterpret_cast<int *>(ptr<char>(ctx->i8Scale) + i8_scale_bytes(nt, i8_block));
        CU(ctx, cudaMemsetAsync(i8dbg, 0, 4 * sizeof(int), st));
    }
    if (use_worker) {
        CU(ctx, cudaStreamWaitEvent(ctx->s_worker, e0, 0));
        if (ctx->chol_variant == 4) big_worker_kernel<<<1, NTHREADS, diag_smem, ctx->s_worker>>>(prm);  // round-1 protocol
        else big_worker2_kernel<<<1, NTHREADS, diag_smem, ctx->s_worker>>>(prm);
        ctx->launches++;
        CU(ctx, cudaEventRecord(eW, ctx->s_worker));
        if ((rc = panel_cols(0))) return rc;
    } else {
        factor_panel(0, ctx->s_panel, diag_smem);
    }
    CU(ctx, cudaEventRecord(eF[0], ctx->s_panel));
    int sms = 0;
    CU(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    // persistent CTAs of the deferred passes; the next block's chain and DMMA updates get the rest (B200 sweep, n = 8192:
    // 64 -> 7.0 ms, 98 -> 6.35, 128 -> 6.55; one output block per CTA instead of persistent CTAs: 6.6; profiles/i8_trail_r02.txt)
    const int i8_ctas = (sms * 2) / 3;
    bool have_i8 = false;
    cudaEvent_t eSplit = nullptr, eI8 = nullptr;
    if (i8S) {
        CU(ctx, cudaEventCreateWithFlags(&eSplit, cudaEventDisableTiming));
        CU(ctx, cudaEventCreateWithFlags(&eI8, cudaEventDisableTiming));
        CU(ctx, cudaStreamWaitEvent(ctx->s_i8, e0, 0));
    }
    for (int P = 0; P + 1 < NP; ++P) {
        const int n0 = (P + 1) * PANEL, n1 = (n0 + PANEL < nt) ? n0 + PANEL : nt;  // columns of panel P + 1
        const int blk_end = i8S ? ((P * PANEL) / i8_block + 1) * i8_block : nt;  // DMMA updates stop at the end of the block
        if (P >= 1) CU(ctx, cudaStreamWaitEvent(ctx->s_panel, eB[P - 1], 0));
        if (i8S && n0 >= blk_end && n0 < nt) {
            // panel P closes a block: slices of its rows below, then three INT8 passes in the order the chain needs them --
            // the columns of the next panel at once (all SMs but the worker's, on the chain's stream), the other columns of
            // the next block on the DMMA stream (ordered with the in-block updates of the same tiles that follow), the rest
            // of the trailing matrix on its own stream while the next block is being factored on the remaining SMs
            const int c0 = blk_end - i8_block, ncb = (nt - blk_end + 1) / 2;
            const int cb_a = PANEL / 2 < ncb ? PANEL / 2 : ncb, cb_b = i8_block / 2 < ncb ? i8_block / 2 : ncb;
            signed char *sl = ptr<signed char>(ctx->i8Slices);
            double *sc = ptr<double>(ctx->i8Scale);
            if (have_i8) CU(ctx, cudaStreamWaitEvent(ctx->s_panel, eI8, 0));  // the slices and the tiles of the previous pass
            if (i8_split_tiles(tiles, nt, blk_end, c0, i8_block, i8S, sl, sc, ctx->s_panel)) return fail(ctx, GPL_ERR_CUDA, "i8_split_tiles launch failed");
            CU(ctx, cudaEventRecord(eSplit, ctx->s_panel));
            int irc = i8_trail(tiles, nt, blk_end, i8_block, i8S, sl, sc, 0, cb_a, sms - 1, dinfo, i8dbg, ctx->s_panel);
            CU(ctx, cudaStreamWaitEvent(ctx->s_trail, eSplit, 0));
            if (!irc && cb_a < cb_b) {
                irc = i8_trail(tiles, nt, blk_end, i8_block, i8S, sl, sc, cb_a, cb_b, i8_ctas, dinfo, i8dbg, ctx->s_trail);
                ctx->launches++;
            }
            CU(ctx, cudaEventRecord(eB[P], ctx->s_trail));
            if (!irc && cb_b < ncb) {
                CU(ctx, cudaStreamWaitEvent(ctx->s_i8, eB[P], 0));  // one persistent INT8 kernel at a time next to the chain
                irc = i8_trail(tiles, nt, blk_end, i8_block, i8S, sl, sc, cb_b, ncb, i8_ctas, dinfo, i8dbg, ctx->s_i8);
                ctx->launches++;
            }
            CU(ctx, cudaEventRecord(eI8, ctx->s_i8));
            have_i8 = true;
            if (irc) return fail(ctx, GPL_ERR_CUDA, "INT8 trailing update: launch failed (%d)", irc);
            ctx->launches += 2;
            if (use_worker) {
                if ((rc = panel_cols(P + 1))) return rc;
            } else {
                factor_panel(P + 1, ctx->s_panel, diag_smem);
            }
            CU(ctx, cudaEventRecord(eF[P + 1], ctx->s_panel));
            continue;
        }
        trail(P, n0, n1, ctx->s_panel);
        if (use_worker) {
            if ((rc = panel_cols(P + 1))) return rc;
        } else {
            factor_panel(P + 1, ctx->s_panel, diag_smem);
        }
        CU(ctx, cudaEventRecord(eF[P + 1], ctx->s_panel));
        CU(ctx, cudaStreamWaitEvent(ctx->s_trail, eF[P], 0));
        trail(P, n1, n1 > blk_end ? n1 : (blk_end < nt ? blk_end : nt), ctx->s_trail);
        CU(ctx, cudaEventRecord(eB[P], ctx->s_trail));
    }
    if (have_i8) CU(ctx, cudaStreamWaitEvent(st, eI8, 0));
    CU(ctx, cudaStreamWaitEvent(st, eF[NP - 1], 0));
    CU(ctx, cudaStreamWaitEvent(st, eB[NP - 2], 0));
    if (use_worker) {
        CU(ctx, cudaStreamWaitEvent(st, eW, 0));
        big_winv_kernel<<<nt, NTHREADS, smem, st>>>(prm);
        ctx->launches++;
    }
    CU(ctx, cudaGetLastError());
    if (ctx->profile_events == 2) {  // debug: when each panel was factored / its trailing update finished (ms from start)
        CU(ctx, cudaStreamSynchronize(st));
        if (i8dbg) {
            int h[4] = {0, 0, 0, 0};
            cudaMemcpy(h, i8dbg, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "int8 trailing update: barrier breadcrumb (code, cta, parity, failed) = %d %d %d %d\n", h[0], h[1], h[2], h[3]);
        }
        for (int P = 0; P < NP; ++P) {
            float tf = 0.f, tb = 0.f;
            cudaEventElapsedTime(&tf, e0, eF[P]);
            if (P + 1 < NP) cudaEventElapsedTime(&tb, e0, eB[P]);
            fprintf(stderr, "panel %2d factored %.3f  trail done %.3f\n", P, tf, tb);
        }
    }
    for (int P = 0; P < NP; ++P) {
        cudaEventDestroy(eF[P]);
        cudaEventDestroy(eB[P]);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(eW);
    if (eSplit) cudaEventDestroy(eSplit);
    if (eI8) cudaEventDestroy(eI8);
    return GPL_OK;
}

int launch_cov(gpl_ctx *ctx, const DevProgram &prog, int na, int nb, int d, int p, int same, const double *dXa,
               const double *dXb, const double *dtheta, double diag_add, double *dK, cudaStream_t st) {
    CovParams prm;
    prm.prog = prog;
    prm.na = na;
    prm.nb = nb;
    prm.d = d;
    prm.p = p;
    prm.same = same;
    prm.Xa = dXa;
    prm.Xb = dXb;
    prm.theta = dtheta;
    prm.diag_add = diag_add;
    prm.K = dK;
    dim3 grid((na + TS - 1) / TS, (nb + TS - 1) / TS);
    cov_dense_kernel<<<grid, NTHREADS, 0, st>>>(prm);
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    return GPL_OK;
}

}  // namespace

// ================================================================================================================
extern "C" {

int gpl_abi_version(void) { return GPL_ABI_VERSION; }

const char *gpl_last_error(gpl_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int gpl_init(int device, gpl_ctx **out) {
    if (!out) return fail(nullptr, GPL_ERR_ARG, "gpl_init: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, GPL_ERR_CUDA, "gpl_init: no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) return fail(nullptr, GPL_ERR_ARG, "gpl_init: device %d out of range (%d devices)", device, count);
    gpl_ctx *ctx = new (std::nothrow) gpl_ctx();
    if (!ctx) return fail(nullptr, GPL_ERR_ARG, "gpl_init: out of host memory");
    ctx->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->wsEvent, cudaEventDisableTiming)) != cudaSuccess) {
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return fail(nullptr, GPL_ERR_CUDA, "gpl_init: %s", cudaGetErrorString(e));
    }
    if (prop.major < 10) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return fail(nullptr, GPL_ERR_CUDA, "gpl_init: device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    }
    ctx->own_stream = ctx->stream;
    ctx->sm_count = prop.multiProcessorCount;
    cudaDeviceGetAttribute(&ctx->clock_khz, cudaDevAttrClockRate, device);
    snprintf(ctx->name, sizeof(ctx->name), "%s", prop.name);
    *out = ctx;
    return GPL_OK;
}

int gpl_destroy(gpl_ctx *ctx) {
    if (!ctx) return GPL_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();  // *_dev calls may still be in flight on caller streams
    // Posteriors that outlive their context are detached: their device block is released here, the host handle stays
    // valid for gpl_posterior_free (which then only deletes it); every other call on such a handle returns GPL_ERR_ARG.
    for (gpl_post *post : ctx->livePosts) {
        if (post->block) cudaFree(post->block);
        post->block = nullptr;
        post->tiles = post->winv = post->alpha = post->dX = post->dtheta = nullptr;
        post->ctx = nullptr;
    }
    ctx->livePosts.clear();
    if (ctx->wsEvent) cudaEventDestroy(ctx->wsEvent);
    std::vector<DevBuf *> bufs;
    all_buffers(ctx, bufs);
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (auto &blk : ctx->postFree) cudaFree(blk.first);
    if (ctx->hStage) cudaFreeHost(ctx->hStage);
    if (ctx->s_worker) cudaStreamDestroy(ctx->s_worker);
    if (ctx->s_panel) cudaStreamDestroy(ctx->s_panel);
    if (ctx->s_trail) cudaStreamDestroy(ctx->s_trail);
    if (ctx->s_i8) cudaStreamDestroy(ctx->s_i8);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return GPL_OK;
}

// Give the grow-only workspace back (the buffers re-grow on the next call that needs them).
int gpl_release_workspace(gpl_ctx *ctx) {
    if (!ctx) return fail(ctx, GPL_ERR_ARG, "gpl_release_workspace: null context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaDeviceSynchronize());
    std::vector<DevBuf *> bufs;
    all_buffers(ctx, bufs);
    for (DevBuf *b : bufs) {
        if (b->p) cudaFree(b->p);
        b->p = nullptr;
        b->cap = 0;
    }
    for (auto &blk : ctx->postFree) cudaFree(blk.first);
    ctx->postFree.clear();
    return GPL_OK;
}

int gpl_set_stream(gpl_ctx *ctx, void *stream) {
    if (!ctx) return fail(ctx, GPL_ERR_ARG, "gpl_set_stream: null context");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = stream ? (cudaStream_t)stream : ctx->own_stream;
    return GPL_OK;
}

uint64_t gpl_launch_count(gpl_ctx *ctx) { return ctx ? ctx->launches : 0; }

int gpl_set_option(gpl_ctx *ctx, const char *key, int value) {
    if (!ctx || !key) return fail(ctx, GPL_ERR_ARG, "gpl_set_option: null argument");
    if (!strcmp(key, "lml_variant")) ctx->lml_variant = value;
    else if (!strcmp(key, "chol_variant")) ctx->chol_variant = value;
    else if (!strcmp(key, "lk_ws_limit_mb")) ctx->lk_ws_limit = (size_t)value << 20;
    else if (!strcmp(key, "profile_events")) ctx->profile_events = value;
    else if (!strcmp(key, "poison_ws")) ctx->poison_ws = value;
    else if (!strcmp(key, "zero_tile_skip")) ctx->zero_tile_skip = value;
    else if (!strcmp(key, "trail_int8")) {
        if (value != 0 && value != -1 && (value < 5 || value > 9)) return fail(ctx, GPL_ERR_ARG, "trail_int8: -1 (auto), 0 (off) or 5..9 slices");
        ctx->trail_int8 = value;
    }
    else if (!strcmp(key, "ou_separable")) ctx->ou_separable = value;
    else return fail(ctx, GPL_ERR_ARG, "gpl_set_option: unknown key '%s'", key);
    return GPL_OK;
}

int gpl_device_info(gpl_ctx *ctx, char *name, int len, int *sm_count, int *clock_khz) {
    if (!ctx) return fail(ctx, GPL_ERR_ARG, "null context");
    if (name && len > 0) snprintf(name, len, "%s", ctx->name);
    if (sm_count) *sm_count = ctx->sm_count;
    if (clock_khz) *clock_khz = ctx->clock_khz;
    return GPL_OK;
}

// ---- program ---------------------------------------------------------------------------------------------------
int gpl_program_create(gpl_ctx *ctx, const gpl_op *ops, int n_ops, gpl_prog **out) {
    if (!out) return fail(ctx, GPL_ERR_ARG, "gpl_program_create: out is NULL");
    *out = nullptr;
    gpl_prog *p = new (std::nothrow) gpl_prog();
    if (!p) return fail(ctx, GPL_ERR_ARG, "out of host memory");
    char msg[192];
    int rc = compile_program(ops, n_ops, &p->dev, msg);
    if (rc) {
        delete p;
        return fail(ctx, rc, "%s", msg);
    }
    *out = p;
    return GPL_OK;
}
int gpl_program_destroy(gpl_prog *prog) {
    delete prog;
    return GPL_OK;
}
int gpl_program_n_theta(const gpl_prog *prog) { return prog ? prog->dev.n_theta : GPL_ERR_ARG; }
int gpl_program_n_cols(const gpl_prog *prog) { return prog ? prog->dev.n_cols : GPL_ERR_ARG; }

// ---- covariance ------------------------------------------------------------------------------------------------
int gpl_cov_dev(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *dX, const double *dtheta, int p,
                double sigma2, double jitter, double *dK, void *stream) {
    int rc = check_prog_args(ctx, prog, n, d, p);
    if (rc) return rc;
    if (!dX || !dK || (p > 0 && !dtheta)) return fail(ctx, GPL_ERR_ARG, "gpl_cov_dev: null pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    return launch_cov(ctx, prog->dev, n, n, d, p, 1, dX, dX, dtheta, sigma2 + jitter, dK, (cudaStream_t)stream);
}

int gpl_cov(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, const double *theta, int p,
            double sigma2, double jitter, double *K) {
    int rc = check_prog_args(ctx, prog, n, d, p);
    if (rc) return rc;
    if (!X || !K || (p > 0 && !theta)) return fail(ctx, GPL_ERR_ARG, "gpl_cov: null pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    WsOrder order(ctx, st);
    if ((rc = ensure(ctx, ctx->bX, (size_t)n * d * 8)) || (rc = ensure(ctx, ctx->bTheta, (size_t)(p + 1) * 8)) ||
        (rc = ensure(ctx, ctx->bK, (size_t)n * n * 8)))
        return rc;
    CU(ctx, cudaMemcpyAsync(ctx->bX.p, X, (size_t)n * d * 8, cudaMemcpyHostToDevice, st));
    if (p > 0) CU(ctx, cudaMemcpyAsync(ctx->bTheta.p, theta, (size_t)p * 8, cudaMemcpyHostToDevice, st));
    rc = launch_cov(ctx, prog->dev, n, n, d, p, 1, ptr<double>(ctx->bX), ptr<double>(ctx->bX), ptr<double>(ctx->bTheta),
                    sigma2 + jitter, ptr<double>(ctx->bK), st);
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(K, ctx->bK.p, (size_t)n * n * 8, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return GPL_OK;
}

int gpl_cross_cov(gpl_ctx *ctx, const gpl_prog *prog, int n, int m, int d, const double *X, const double *Xs,
                  const double *theta, int p, double *Ks) {
    int rc = check_prog_args(ctx, prog, n, d, p);
    if (rc) return rc;
    if (m <= 0 || !X || !Xs || !Ks || (p > 0 && !theta)) return fail(ctx, GPL_ERR_ARG, "gpl_cross_cov: bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    WsOrder order(ctx, st);
    if ((rc = ensure(ctx, ctx->bX, (size_t)n * d * 8)) || (rc = ensure(ctx, ctx->bXs, (size_t)m * d * 8)) ||
        (rc = ensure(ctx, ctx->bTheta, (size_t)(p + 1) * 8)) || (rc = ensure(ctx, ctx->bK, (size_t)n * m * 8)))
        return rc;
    CU(ctx, cudaMemcpyAsync(ctx->bX.p, X, (size_t)n * d * 8, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->bXs.p, Xs, (size_t)m * d * 8, cudaMemcpyHostToDevice, st));
    if (p > 0) CU(ctx, cudaMemcpyAsync(ctx->bTheta.p, theta, (size_t)p * 8, cudaMemcpyHostToDevice, st));
    rc = launch_cov(ctx, prog->dev, n, m, d, p, 0, ptr<double>(ctx->bX), ptr<double>(ctx->bXs), ptr<double>(ctx->bTheta),
                    0.0, ptr<double>(ctx->bK), st);
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(Ks, ctx->bK.p, (size_t)n * m * 8, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return GPL_OK;
}

// ---- batched lml -----------------------------------------------------------------------------------------------
int gpl_lml_batched_dev(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *dX, int x_batched,
                        const double *dY, int y_batched, const double *dTheta, int p, const double *dsigma2,
                        int sigma2_batched, double jitter, int B, double *dlml, double *ddtheta, double *ddy,
                        int *dinfo, void *stream) {
    int rc = check_prog_args(ctx, prog, n, d, p);
    if (rc) return rc;
    if (B <= 0) return fail(ctx, GPL_ERR_ARG, "gpl_lml_batched: B=%d", B);
    if (!dX || !dY || !dsigma2 || !dlml || (p > 0 && !dTheta)) return fail(ctx, GPL_ERR_ARG, "gpl_lml_batched: null pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    WsOrder order(ctx, (cudaStream_t)stream);
    const int want_grad = ddtheta != nullptr;
    return launch_lml(ctx, prog->dev, n, d, dX, x_batched, dY, y_batched, dTheta, p, dsigma2, sigma2_batched, jitter, B,
                      dlml, ddtheta, ddy, dinfo, want_grad, 0, nullptr, nullptr, (cudaStream_t)stream);
}

int gpl_lml_batched(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, int x_batched, const double *Y,
                    int y_batched, const double *Theta, int p, const double *sigma2, int sigma2_batched, double jitter,
                    int B, double *lml, double *dtheta, double *dy, int *info) {
    int rc = check_prog_args(ctx, prog, n, d, p);
    if (rc) return rc;
    if (B <= 0) return fail(ctx, GPL_ERR_ARG, "gpl_lml_batched: B=%d", B);
    if (!X || !Y || !sigma2 || !lml || (p > 0 && !Theta)) return fail(ctx, GPL_ERR_ARG, "gpl_lml_batched: null pointer");
    if (dy && !dtheta && p > 0) return fail(ctx, GPL_ERR_ARG, "gpl_lml_batched: dy requires dtheta");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    WsOrder order(ctx, st);
    const size_t xb = (size_t)n * d * (x_batched ? B : 1) * 8, yb = (size_t)n * (y_batched ? B : 1) * 8;
    const size_t tb = (size_t)(p > 0 ? p : 1) * B * 8, sb = (size_t)(sigma2_batched ? B : 1) * 8;
    if ((rc = ensure(ctx, ctx->bX, xb)) || (rc = ensure(ctx, ctx->bY, yb)) || (rc = ensure(ctx, ctx->bTheta, tb)) ||
        (rc = ensure(ctx, ctx->bSigma, sb)) || (rc = ensure(ctx, ctx->bLml, (size_t)B * 8)) ||
        (rc = ensure(ctx, ctx->bInfo, (size_t)B * 4)))
        return rc;
    const bool grad = dtheta != nullptr || dy != nullptr;
    {
        // Small calls (one log-density + gradient evaluation of an MCMC step: a few KB) are bound by driver calls, not by
        // bytes: pack every input into one pinned block -> one H2D copy, and every output into one D2H copy.
        constexpr size_t STAGE = 1 << 20;
        auto up = [](size_t v) { return (v + 255) / 256 * 256; };
        const size_t o_x = 0, o_y = o_x + up(xb), o_t = o_y + up(yb), o_s = o_t + up((size_t)p * B * 8), in_bytes = o_s + up(sb);
        const size_t r_lml = up(in_bytes), r_info = r_lml + up((size_t)B * 8), r_dth = r_info + up((size_t)B * 4),
                     r_dy = r_dth + (grad ? up((size_t)(p > 0 ? p : 1) * B * 8) : 0), end = r_dy + (grad ? up((size_t)n * B * 8) : 0);
        if (end <= STAGE) {
            if (!ctx->hStage) CU(ctx, cudaHostAlloc(&ctx->hStage, STAGE, cudaHostAllocDefault));
            if ((rc = ensure(ctx, ctx->dStage, STAGE))) return rc;
            char *h = static_cast<char *>(ctx->hStage), *dv = static_cast<char *>(ctx->dStage.p);
            memcpy(h + o_x, X, xb);
            memcpy(h + o_y, Y, yb);
            if (p > 0) memcpy(h + o_t, Theta, (size_t)p * B * 8);
            memcpy(h + o_s, sigma2, sb);
            CU(ctx, cudaMemcpyAsync(dv, h, in_bytes, cudaMemcpyHostToDevice, st));
            rc = launch_lml(ctx, prog->dev, n, d, reinterpret_cast<double *>(dv + o_x), x_batched,
                            reinterpret_cast<double *>(dv + o_y), y_batched, reinterpret_cast<double *>(dv + o_t), p,
                            reinterpret_cast<double *>(dv + o_s), sigma2_batched, jitter, B,
                            reinterpret_cast<double *>(dv + r_lml), grad ? reinterpret_cast<double *>(dv + r_dth) : nullptr,
                            grad ? reinterpret_cast<double *>(dv + r_dy) : nullptr, reinterpret_cast<int *>(dv + r_info),
                            grad ? 1 : 0, 0, nullptr, nullptr, st);
            if (rc) return rc;
            CU(ctx, cudaMemcpyAsync(h + r_lml, dv + r_lml, end - r_lml, cudaMemcpyDeviceToHost, st));
            CU(ctx, cudaStreamSynchronize(st));
            memcpy(lml, h + r_lml, (size_t)B * 8);
            if (info) memcpy(info, h + r_info, (size_t)B * 4);
            if (dtheta && p > 0) memcpy(dtheta, h + r_dth, (size_t)p * B * 8);
            if (dy) memcpy(dy, h + r_dy, (size_t)n * B * 8);
            return GPL_OK;
        }
    }
    if (grad && ((rc = ensure(ctx, ctx->bDtheta, tb)) || (rc = ensure(ctx, ctx->bDy, (size_t)n * B * 8)))) return rc;
    CU(ctx, cudaMemcpyAsync(ctx->bX.p, X, xb, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->bY.p, Y, yb, cudaMemcpyHostToDevice, st));
    if (p > 0) CU(ctx, cudaMemcpyAsync(ctx->bTheta.p, Theta, (size_t)p * B * 8, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->bSigma.p, sigma2, sb, cudaMemcpyHostToDevice, st));
    rc = launch_lml(ctx, prog->dev, n, d, ptr<double>(ctx->bX), x_batched, ptr<double>(ctx->bY), y_batched,
                    ptr<double>(ctx->bTheta), p, ptr<double>(ctx->bSigma), sigma2_batched, jitter, B, ptr<double>(ctx->bLml),
                    grad ? ptr<double>(ctx->bDtheta) : nullptr, grad ? ptr<double>(ctx->bDy) : nullptr,
                    ptr<int>(ctx->bInfo), grad ? 1 : 0, 0, nullptr, nullptr, st);
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(lml, ctx->bLml.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
    if (info) CU(ctx, cudaMemcpyAsync(info, ctx->bInfo.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    if (dtheta && p > 0) CU(ctx, cudaMemcpyAsync(dtheta, ctx->bDtheta.p, (size_t)p * B * 8, cudaMemcpyDeviceToHost, st));
    if (dy) CU(ctx, cudaMemcpyAsync(dy, ctx->bDy.p, (size_t)n * B * 8, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return GPL_OK;
}

// Per-phase device time of the last instrumented call (option "profile_events" = 1): the optional stats struct of
// SURVEY.md section 5.  Phases of gpl_lml_batched(_dev): 0 lk_diag, 1 lk_potrf_warp, 2 lk_below, 3 lk_winv, 4 lk_minv,
// 5 lk_alpha, 6 lk_gradc + lk_gradsum.  gpl_lml_large: 0 covariance build, 1 factorisation + forward solve.
int gpl_last_timing(gpl_ctx *ctx, gpl_timing *out) {
    if (!ctx || !out) return fail(ctx, GPL_ERR_ARG, "gpl_last_timing: null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    out->n_phases = 7;
    for (int k = 0; k < 8; ++k) {
        out->ms[k] = k < 7 ? ctx->lk_ms[k] : 0.0;
        out->launches[k] = k < 7 ? ctx->lk_launches[k] : 0;
    }
    return GPL_OK;
}

// Debug (profile builds only, -DGPL_LML_PROFILE): run the plain lml kernel with `raw` (device, grid*8 doubles)
// receiving per-CTA phase clock totals.  Not part of the public header.
int gpl_debug_phase_profile(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *dX, const double *dY,
                            const double *dTheta, int p, const double *dsigma2, int B, double *dlml, int *dinfo,
                            double *raw, int *grid_out) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    // want_grad = 0 but dtheta carries the raw buffer
    int rc = launch_lml(ctx, prog->dev, n, d, dX, 0, dY, 0, dTheta, p, dsigma2, 0, 0.0, B, dlml, raw, nullptr, dinfo, 0, 0,
                        nullptr, nullptr, ctx->stream);
    if (grid_out) {
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lml_batched_kernel, NTHREADS, lml_smem_bytes(false));
        *grid_out = ctx->sm_count * occ < B ? ctx->sm_count * occ : B;
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

// ---- posterior ---------------------------------------------------------------------------------------------------
static void posterior_release(gpl_post *post) {  // caller holds the context lock (or there is no context)
    gpl_ctx *ctx = post->ctx;
    if (ctx) {
        cudaSetDevice(ctx->device);
        for (size_t k = 0; k < ctx->livePosts.size(); ++k)
            if (ctx->livePosts[k] == post) {
                ctx->livePosts.erase(ctx->livePosts.begin() + k);
                break;
            }
    }
    if (post->block) {
        if (ctx && ctx->postFree.size() < 8) {
            cudaStreamSynchronize(ctx->stream);  // nothing in flight reads it any more
            ctx->postFree.emplace_back(post->block, post->block_bytes);
        } else {
            cudaFree(post->block);
        }
    }
    delete post;
}
int gpl_posterior_free(gpl_post *post) {
    if (!post) return GPL_OK;
    if (post->ctx) {
        std::lock_guard<std::mutex> lk(post->ctx->mu);
        posterior_release(post);
    } else {
        posterior_release(post);
    }
    return GPL_OK;
}

static int posterior_fit_impl(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, const double *y,
                              const double *theta, int p, double sigma2, double jitter, gpl_post *post) {
    cudaStream_t st = ctx->stream;
    const int nt = (n + TS - 1) / TS;
    const long long ntri = tri_index(nt, 0);
    post->ctx = ctx;
    post->prog = prog->dev;
    post->n = n;
    post->d = d;
    post->nt = nt;
    post->p = p;
    int rc;
    {
        // one block: [tiles: ntri L tiles + nt W tiles (the fused kernel's workspace layout)] [z | alpha] [X] [theta]
        auto up = [](size_t v) { return (v + 255) / 256 * 256; };
        const size_t b_tiles = up((size_t)(ntri + nt) * TILE_BYTES), b_alpha = up((size_t)2 * nt * TS * 8),
                     b_x = up((size_t)n * d * 8), b_th = up((size_t)(p + 1) * 8);
        const size_t need = b_tiles + b_alpha + b_x + b_th;
        int best = -1;
        for (int k = 0; k < (int)ctx->postFree.size(); ++k)
            if (ctx->postFree[k].second >= need && ctx->postFree[k].second <= 2 * need &&
                (best < 0 || ctx->postFree[k].second < ctx->postFree[best].second))
                best = k;
        if (best >= 0) {
            post->block = ctx->postFree[best].first;
            post->block_bytes = ctx->postFree[best].second;
            ctx->postFree.erase(ctx->postFree.begin() + best);
        } else {
            CU(ctx, cudaMalloc(&post->block, need));
            post->block_bytes = need;
        }
        char *base = static_cast<char *>(post->block);
        post->tiles = reinterpret_cast<double *>(base);
        post->alpha = reinterpret_cast<double *>(base + b_tiles);
        post->dX = reinterpret_cast<double *>(base + b_tiles + b_alpha);
        post->dtheta = reinterpret_cast<double *>(base + b_tiles + b_alpha + b_x);
        post->winv = nullptr;
    }
    CU(ctx, cudaMemcpyAsync(post->dX, X, (size_t)n * d * 8, cudaMemcpyHostToDevice, st));
    if (p > 0) CU(ctx, cudaMemcpyAsync(post->dtheta, theta, (size_t)p * 8, cudaMemcpyHostToDevice, st));
    double *winv = post->tiles + ntri * TILE_ELEMS;
    double *zvec = post->alpha, *avec = post->alpha + (size_t)nt * TS;
    if ((rc = ensure(ctx, ctx->bMisc, 64 + (size_t)nt * TS * 8))) return rc;
    double *dres = ptr<double>(ctx->bMisc);               // [0]=lml or logdet, [1]=quad
    int *dinfo = reinterpret_cast<int *>(dres + 4);
    double *pivlog = dres + 8;
    if ((rc = ensure(ctx, ctx->bY, (size_t)nt * TS * 8)) || (rc = ensure(ctx, ctx->bSigma, 8))) return rc;
    CU(ctx, cudaMemsetAsync(ctx->bY.p, 0, (size_t)nt * TS * 8, st));
    CU(ctx, cudaMemcpyAsync(ctx->bY.p, y, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    double h_lml = 0.0;
    int h_info = 0;
    if (n <= SMALL_MAX_N && ctx->chol_variant == 0) {
        CU(ctx, cudaMemcpyAsync(ctx->bSigma.p, &sigma2, 8, cudaMemcpyHostToDevice, st));
        rc = launch_lml(ctx, prog->dev, n, d, post->dX, 0, ptr<double>(ctx->bY), 0, post->dtheta, p, ptr<double>(ctx->bSigma),
                        0, jitter, 1, dres, nullptr, nullptr, dinfo, 0, 1, post->tiles, post->alpha, st);
        if (rc) return rc;
        CU(ctx, cudaMemcpyAsync(&h_lml, dres, 8, cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaMemcpyAsync(&h_info, dinfo, 4, cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaStreamSynchronize(st));
    } else {
        CovTilesParams cp;
        cp.prog = prog->dev;
        cp.n = n;
        cp.nt = nt;
        cp.d = d;
        cp.p = p;
        cp.X = post->dX;
        cp.theta = post->dtheta;
        cp.diag_add = sigma2 + jitter;
        cp.tiles = post->tiles;
        cov_tiles_kernel<<<(unsigned)ntri, NTHREADS, 0, st>>>(cp);
        ctx->launches++;
        CU(ctx, cudaMemcpyAsync(zvec, ctx->bY.p, (size_t)nt * TS * 8, cudaMemcpyDeviceToDevice, st));
        if ((rc = big_factor(ctx, post->tiles, winv, pivlog, dinfo, zvec, nt, st))) return rc;
        big_reduce_kernel<<<1, NTHREADS, 0, st>>>(pivlog, zvec, nt * TS, dres);
        ctx->launches++;
        // backward substitution: r starts as a copy of z (kept), alpha written block by block
        if ((rc = ensure(ctx, ctx->bDy, (size_t)nt * TS * 8))) return rc;
        CU(ctx, cudaMemcpyAsync(ctx->bDy.p, zvec, (size_t)nt * TS * 8, cudaMemcpyDeviceToDevice, st));
        for (int i = nt - 1; i >= 0; --i) {
            big_backward_kernel<<<i + 1, NTHREADS, 0, st>>>(post->tiles, winv, i, ptr<double>(ctx->bDy), avec);
            ctx->launches++;
        }
        CU(ctx, cudaGetLastError());
        double h_res[2];
        CU(ctx, cudaMemcpyAsync(h_res, dres, 16, cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaMemcpyAsync(&h_info, dinfo, 4, cudaMemcpyDeviceToHost, st));
        CU(ctx, cudaStreamSynchronize(st));
        h_lml = -0.5 * ((double)n * LOG2PI + h_res[0] + h_res[1]);
    }
    if (h_info < 0) return fail(ctx, GPL_ERR_CUDA, "large-n factorisation: a hand-off between its kernels timed out (device busy?)");
    if (h_info != 0)
        return fail(ctx, GPL_ERR_NOTPD, "covariance not positive definite: pivot %d (PosDefException(%d))", h_info, h_info);
    post->lml = h_lml;
    return GPL_OK;
}

int gpl_posterior_fit(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, const double *y,
                      const double *theta, int p, double sigma2, double jitter, gpl_post **out) {
    if (!out) return fail(ctx, GPL_ERR_ARG, "gpl_posterior_fit: out is NULL");
    *out = nullptr;
    int rc = check_prog_args(ctx, prog, n, d, p);
    if (rc) return rc;
    if (!X || !y || (p > 0 && !theta)) return fail(ctx, GPL_ERR_ARG, "gpl_posterior_fit: null pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    WsOrder order(ctx, ctx->stream);
    gpl_post *post = new (std::nothrow) gpl_post();
    if (!post) return fail(ctx, GPL_ERR_ARG, "out of host memory");
    rc = posterior_fit_impl(ctx, prog, n, d, X, y, theta, p, sigma2, jitter, post);
    if (rc) {
        cudaStreamSynchronize(ctx->stream);
        posterior_release(post);
        return rc;
    }
    ctx->livePosts.push_back(post);
    *out = post;
    return GPL_OK;
}

int gpl_posterior_logpdf(gpl_post *post, double *lml) {
    if (!post || !lml) return fail(nullptr, GPL_ERR_ARG, "gpl_posterior_logpdf: null argument");
    *lml = post->lml;
    return GPL_OK;
}

int gpl_posterior_alpha(gpl_post *post, double *alpha) {
    if (!post || !alpha) return fail(nullptr, GPL_ERR_ARG, "gpl_posterior_alpha: null argument");
    gpl_ctx *ctx = post->ctx;
    if (!ctx) return fail(nullptr, GPL_ERR_ARG, "posterior handle outlived its context (gpl_destroy was called)");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    WsOrder order(ctx, ctx->stream);
    CU(ctx, cudaMemcpyAsync(alpha, post->alpha + (size_t)post->nt * TS, (size_t)post->n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return GPL_OK;
}

int gpl_posterior_factor(gpl_post *post, double *U) {
    if (!post || !U) return fail(nullptr, GPL_ERR_ARG, "gpl_posterior_factor: null argument");
    gpl_ctx *ctx = post->ctx;
    if (!ctx) return fail(nullptr, GPL_ERR_ARG, "posterior handle outlived its context (gpl_destroy was called)");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    WsOrder order(ctx, ctx->stream);
    const int n = post->n, nt = post->nt;
    int rc = ensure(ctx, ctx->bK, (size_t)n * n * 8);
    if (rc) return rc;
    dim3 grid(nt, nt);
    tiles_to_upper_kernel<<<grid, NTHREADS, 0, ctx->stream>>>(post->tiles, n, nt, ptr<double>(ctx->bK));
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(U, ctx->bK.p, (size_t)n * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return GPL_OK;
}

int gpl_posterior_mean_var(gpl_post *post, int m, const double *Xs, double *mean, double *var) {
    if (!post || !Xs || !mean || m <= 0) return fail(nullptr, GPL_ERR_ARG, "gpl_posterior_mean_var: bad argument");
    gpl_ctx *ctx = post->ctx;
    if (!ctx) return fail(nullptr, GPL_ERR_ARG, "posterior handle outlived its context (gpl_destroy was called)");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    WsOrder order(ctx, st);
    const int nt = post->nt, d = post->d;
    const size_t smem = predict_smem_bytes();
    bool &attr_set = ctx->attr_pred;
    if (!attr_set) {
        CU(ctx, cudaFuncSetAttribute(predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(ctx, cudaFuncSetAttribute(predict_kernel_nb6, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(ctx, cudaFuncSetAttribute(predict_kernel_nb4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int occ = 0;
    CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, predict_kernel, NTHREADS, smem));
    if (occ < 1) return fail(ctx, GPL_ERR_CUDA, "predict kernel does not fit on an SM");
    // slab width (64, 48 or 32 test points): the one whose busiest SM has the fewest points (ties: the wider)
    int nb_blocks = 8;
    {
        long long best = -1;
        for (int nb : {8, 6, 4}) {
            const long long ns = (m + 8 * nb - 1) / (8 * nb);
            const long long cost = ((ns + ctx->sm_count - 1) / ctx->sm_count) * nb;
            if (best < 0 || cost < best) {
                best = cost;
                nb_blocks = nb;
            }
        }
    }
    const int nslab = (m + 8 * nb_blocks - 1) / (8 * nb_blocks);
    int grid = ctx->sm_count * occ;
    if (grid > nslab) grid = nslab;
    int rc;
    if ((rc = ensure(ctx, ctx->bXs, (size_t)m * d * 8)) || (rc = ensure(ctx, ctx->bMean, (size_t)m * 8)) ||
        (rc = ensure(ctx, ctx->bVar, (size_t)m * 8)) || (rc = ensure(ctx, ctx->bWsV, (size_t)grid * nt * TILE_BYTES)))
        return rc;
    CU(ctx, cudaMemcpyAsync(ctx->bXs.p, Xs, (size_t)m * d * 8, cudaMemcpyHostToDevice, st));
    PredictParams prm;
    prm.prog = post->prog;
    prm.n = post->n;
    prm.nt = nt;
    prm.d = d;
    prm.p = post->p;
    prm.m = m;
    prm.want_var = var != nullptr;
    prm.items = 0;
    prm.theta_stride = prm.tiles_stride = prm.winv_stride = prm.alpha_stride = 0;
    prm.info = nullptr;
    prm.X = post->dX;
    prm.theta = post->dtheta;
    prm.Xs = ptr<double>(ctx->bXs);
    prm.tiles = post->tiles;
    prm.winv = post->tiles + tri_index(nt, 0) * TILE_ELEMS;
    prm.alpha = post->alpha + (size_t)nt * TS;
    prm.wsV = ptr<double>(ctx->bWsV);
    prm.mean = ptr<double>(ctx->bMean);
    prm.var = ptr<double>(ctx->bVar);
    if (nb_blocks == 8) predict_kernel<<<grid, NTHREADS, smem, st>>>(prm);
    else if (nb_blocks == 6) predict_kernel_nb6<<<grid, NTHREADS, smem, st>>>(prm);
    else predict_kernel_nb4<<<grid, NTHREADS, smem, st>>>(prm);
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(mean, ctx->bMean.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    if (var) CU(ctx, cudaMemcpyAsync(var, ctx->bVar.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return GPL_OK;
}

// ---- batched posteriors + predictions over the rows of an MCMC chain -------------------------------------------------
// Replaces, for every chain row b at once,  post_b = posterior(FiniteGP(GP(kernel(theta_b)), X, sigma2_b), y);
// mean_and_var(post_b, Xs)  - the loop behind the `predict` / `fitplot` commands (CLI/src/main.jl:8-16, output columns
// test/pred.jl:11-14) and src/plotting.jl:6-12.  All rows are factored by the lockstep schedule in one go; lk_post_kernel
// forms the diagonal-tile inverses and alpha per row; the prediction kernel runs over (row, slab of test points) pairs.
int gpl_predict_batched(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, const double *y,
                        const double *Theta, int p, const double *sigma2, int sigma2_batched, double jitter, int B, int m,
                        const double *Xs, double *mean, double *var, double *lml, int *info) {
    int rc = check_prog_args(ctx, prog, n, d, p);
    if (rc) return rc;
    if (B <= 0 || m <= 0) return fail(ctx, GPL_ERR_ARG, "gpl_predict_batched: B=%d m=%d", B, m);
    if (!X || !y || !sigma2 || !Xs || !mean || (p > 0 && !Theta)) return fail(ctx, GPL_ERR_ARG, "gpl_predict_batched: null pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    WsOrder order(ctx, st);
    const int nt = (n + TS - 1) / TS;
    const long long ntri = tri_index(nt, 0);
    // rows per pass: the factor workspace cap ("lk_ws_limit_mb") bounds how many posteriors are resident at once; longer
    // chains run in passes of Bc rows (same bits as in one piece: every row is independent)
    const size_t per_item = (size_t)ntri * TILE_BYTES + (size_t)nt * DSIZE * 8 + (size_t)nt * TS * 8 + 16 +
                            (size_t)nt * TILE_BYTES + (size_t)nt * TS * 8;
    int Bc = (int)(ctx->lk_ws_limit / per_item);
    Bc = Bc < 1 ? 1 : (Bc > B ? B : Bc);
    const size_t tb = (size_t)(p > 0 ? p : 1) * B * 8, sb = (size_t)(sigma2_batched ? B : 1) * 8;
    if ((rc = ensure(ctx, ctx->bX, (size_t)n * d * 8)) || (rc = ensure(ctx, ctx->bY, (size_t)n * 8)) ||
        (rc = ensure(ctx, ctx->bTheta, tb)) || (rc = ensure(ctx, ctx->bSigma, sb)) || (rc = ensure(ctx, ctx->bLml, (size_t)B * 8)) ||
        (rc = ensure(ctx, ctx->bInfo, (size_t)B * 4)) || (rc = ensure(ctx, ctx->bXs, (size_t)m * d * 8)) ||
        (rc = ensure(ctx, ctx->bMean, (size_t)m * B * 8)) || (rc = ensure(ctx, ctx->bVar, (size_t)m * B * 8)) ||
        (rc = ensure(ctx, ctx->lkW, (size_t)Bc * nt * TILE_BYTES)) || (rc = ensure(ctx, ctx->lkAlpha, (size_t)Bc * nt * TS * 8)))
        return rc;
    CU(ctx, cudaMemcpyAsync(ctx->bX.p, X, (size_t)n * d * 8, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->bY.p, y, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    if (p > 0) CU(ctx, cudaMemcpyAsync(ctx->bTheta.p, Theta, (size_t)p * B * 8, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->bSigma.p, sigma2, sb, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->bXs.p, Xs, (size_t)m * d * 8, cudaMemcpyHostToDevice, st));
    if (!ctx->attr_post) {
        CU(ctx, cudaFuncSetAttribute(lk_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lk_post_smem_bytes()));
        ctx->attr_post = true;
    }
    const size_t smem = predict_smem_bytes();
    if (!ctx->attr_pred) {
        CU(ctx, cudaFuncSetAttribute(predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(ctx, cudaFuncSetAttribute(predict_kernel_nb6, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(ctx, cudaFuncSetAttribute(predict_kernel_nb4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->attr_pred = true;
    }
    int occ = 0;
    CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, predict_kernel, NTHREADS, smem));
    if (occ < 1) return fail(ctx, GPL_ERR_CUDA, "predict kernel does not fit on an SM");
    for (int off = 0; off < B; off += Bc) {
        const int nb = (B - off < Bc) ? B - off : Bc;
        const size_t s_off = sigma2_batched ? (size_t)off : 0;
        rc = launch_lml_lockstep(ctx, prog->dev, n, d, ptr<double>(ctx->bX), 0, ptr<double>(ctx->bY), 0,
                                 ptr<double>(ctx->bTheta) + (size_t)off * p, p, ptr<double>(ctx->bSigma) + s_off, sigma2_batched,
                                 jitter, nb, ptr<double>(ctx->bLml) + off, ptr<int>(ctx->bInfo) + off, st, nullptr, nullptr, 0,
                                 /*allow_sort=*/0);  // the prediction kernels read X in the caller's order
        if (rc) return rc;
        LkPostParams pq;
        pq.nt = nt;
        pq.tiles = ptr<double>(ctx->lkTiles);
        pq.dblk = ptr<double>(ctx->lkD);
        pq.z = ptr<double>(ctx->lkZ);
        pq.winv = ptr<double>(ctx->lkW);
        pq.alpha = ptr<double>(ctx->lkAlpha);
        lk_post_kernel<<<nb, NTHREADS, lk_post_smem_bytes(), st>>>(pq);
        ctx->launches++;
        int nb_blocks = 8;
        {
            long long best = -1;
            for (int w : {8, 6, 4}) {
                const long long units = (long long)nb * ((m + 8 * w - 1) / (8 * w));
                const long long cost = ((units + ctx->sm_count - 1) / ctx->sm_count) * w;
                if (best < 0 || cost < best) {
                    best = cost;
                    nb_blocks = w;
                }
            }
        }
        const long long units = (long long)nb * ((m + 8 * nb_blocks - 1) / (8 * nb_blocks));
        int grid = ctx->sm_count * occ;
        if (grid > units) grid = (int)units;
        if ((rc = ensure(ctx, ctx->bWsV, (size_t)grid * nt * TILE_BYTES))) return rc;
        PredictParams prm;
        prm.prog = prog->dev;
        prm.n = n;
        prm.nt = nt;
        prm.d = d;
        prm.p = p;
        prm.m = m;
        prm.want_var = var != nullptr;
        prm.items = nb;
        prm.theta_stride = p;
        prm.tiles_stride = ntri * TILE_ELEMS;
        prm.winv_stride = (long long)nt * TILE_ELEMS;
        prm.alpha_stride = (long long)nt * TS;
        prm.info = ptr<int>(ctx->bInfo) + off;
        prm.X = ptr<double>(ctx->bX);
        prm.theta = ptr<double>(ctx->bTheta) + (size_t)off * p;
        prm.Xs = ptr<double>(ctx->bXs);
        prm.tiles = ptr<double>(ctx->lkTiles);
        prm.winv = ptr<double>(ctx->lkW);
        prm.alpha = ptr<double>(ctx->lkAlpha);
        prm.wsV = ptr<double>(ctx->bWsV);
        prm.mean = ptr<double>(ctx->bMean) + (size_t)off * m;
        prm.var = ptr<double>(ctx->bVar) + (size_t)off * m;
        if (nb_blocks == 8) predict_kernel<<<grid, NTHREADS, smem, st>>>(prm);
        else if (nb_blocks == 6) predict_kernel_nb6<<<grid, NTHREADS, smem, st>>>(prm);
        else predict_kernel_nb4<<<grid, NTHREADS, smem, st>>>(prm);
        ctx->launches++;
        CU(ctx, cudaGetLastError());
    }
    CU(ctx, cudaMemcpyAsync(mean, ctx->bMean.p, (size_t)m * B * 8, cudaMemcpyDeviceToHost, st));
    if (var) CU(ctx, cudaMemcpyAsync(var, ctx->bVar.p, (size_t)m * B * 8, cudaMemcpyDeviceToHost, st));
    if (lml) CU(ctx, cudaMemcpyAsync(lml, ctx->bLml.p, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
    if (info) CU(ctx, cudaMemcpyAsync(info, ctx->bInfo.p, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return GPL_OK;
}

// ---- batched on-device sampler ---------------------------------------------------------------------------------------
int gpl_mcmc_nuts(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, int x_batched, const double *Y,
                  int y_batched, int p, const double *lo, const double *hi, const double *sigma2, int sigma2_batched,
                  double jitter, int B, const double *q0, const gpl_mcmc_opts *opts, double *theta, double *lp, double *q,
                  double *accept, double *eps, int *depth, int *n_leapfrog, int *divergent, int *status,
                  long long *n_grad_evals) {
    int rc = check_prog_args(ctx, prog, n, d, p);
    if (rc) return rc;
    if (!opts || !X || !Y || !sigma2 || !q0 || !theta || !lp || B <= 0 || (p > 0 && (!lo || !hi)))
        return fail(ctx, GPL_ERR_ARG, "gpl_mcmc_nuts: null pointer or B <= 0");
    if (p > MC_MAX_P) return fail(ctx, GPL_ERR_LIMIT, "gpl_mcmc_nuts: p=%d exceeds %d", p, MC_MAX_P);
    if (opts->n_samples <= 0) return fail(ctx, GPL_ERR_ARG, "gpl_mcmc_nuts: n_samples=%d", opts->n_samples);
    for (int k = 0; k < p; ++k)
        if (!(hi[k] > lo[k])) return fail(ctx, GPL_ERR_ARG, "gpl_mcmc_nuts: prior bounds of slot %d are not lo < hi", k);
    McmcDevParams mp;
    memset(&mp, 0, sizeof(mp));
    McmcConfig &c = mp.cfg;
    c.n = n;
    c.p = p;
    c.latent = opts->latent ? 1 : 0;
    c.dim = p + (c.latent ? n : 0);
    if (c.dim < 1) return fail(ctx, GPL_ERR_ARG, "gpl_mcmc_nuts: nothing to sample (p = 0 and latent = 0)");
    c.max_depth = opts->max_depth > 0 ? opts->max_depth : MC_MAX_DEPTH;
    if (c.max_depth > MC_MAX_DEPTH) return fail(ctx, GPL_ERR_LIMIT, "gpl_mcmc_nuts: max_depth %d exceeds %d", c.max_depth, MC_MAX_DEPTH);
    c.n_samples = opts->n_samples;
    c.n_adapt = opts->n_adapt >= 0 ? opts->n_adapt : (opts->n_samples / 2 < 1000 ? opts->n_samples / 2 : 1000);
    c.search_eps = opts->search_eps ? 1 : 0;
    c.adapt_mass = opts->adapt_mass ? 1 : 0;
    c.record_warmup = opts->record_warmup ? 1 : 0;
    c.record_q = q ? 1 : 0;
    c.delta = opts->delta > 0 ? opts->delta : 0.65;
    c.max_dh = opts->max_dh > 0 ? opts->max_dh : 1000.0;
    c.obs_sd = opts->obs_sd > 0 ? opts->obs_sd : 1.0;
    c.eps0 = opts->eps0 > 0 ? opts->eps0 : 0.1;
    c.seed = opts->seed;
    for (int k = 0; k < p; ++k) c.lo[k] = lo[k], c.hi[k] = hi[k];
    mc_setup_windows(c);
    const int T = c.n_adapt + c.n_samples, n_rec = c.n_samples + (c.record_warmup ? c.n_adapt : 0), dim = c.dim;
    const int pp = p > 0 ? p : 1;

    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    WsOrder order(ctx, st);
    // one device block for everything the sampler owns (released before returning)
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t vec_stride = (size_t)mc_vectors_per_chain(c.max_depth) * dim;
    const size_t xb = (size_t)n * d * (x_batched ? B : 1) * 8, yb = (size_t)n * (y_batched ? B : 1) * 8,
                 sb = (size_t)(sigma2_batched ? B : 1) * 8;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off += up(bytes);
        return o;
    };
    const size_t o_state = take((size_t)B * mcmc_state_bytes()), o_vec = take((size_t)B * vec_stride * 8), o_q0 = take((size_t)B * dim * 8),
                 o_x = take(xb), o_y = take(yb), o_s2 = take(sb), o_the = take((size_t)B * pp * 8), o_ye = take((size_t)B * n * 8),
                 o_lml = take((size_t)B * 8), o_dth = take((size_t)B * pp * 8), o_dy = take((size_t)B * n * 8), o_info = take((size_t)B * 4),
                 o_th = take((size_t)B * n_rec * pp * 8), o_lp = take((size_t)B * n_rec * 8), o_acc = take((size_t)B * n_rec * 8),
                 o_eps = take((size_t)B * n_rec * 8), o_qo = take(q ? (size_t)B * n_rec * dim * 8 : 8), o_dep = take((size_t)B * n_rec * 4),
                 o_nl = take((size_t)B * n_rec * 4), o_div = take((size_t)B * n_rec * 4), o_done = take(256), o_stat = take((size_t)B * 4),
                 o_slot = take((size_t)B * 4),
                 // slot-ordered copies of the per-chain inputs of the evaluation (used after a compaction)
                 o_xs = take(x_batched ? xb : 8), o_ys = take((y_batched && !opts->latent) ? yb : 8), o_s2s = take(sigma2_batched ? sb : 8);
    char *blk = nullptr;
    CU(ctx, cudaMalloc((void **)&blk, off));
    struct Guard {  // frees the block (and the graph objects) on every exit path
        char *p;
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ge = nullptr;
        ~Guard() {
            if (ge) cudaGraphExecDestroy(ge);
            if (g) cudaGraphDestroy(g);
            cudaFree(p);
        }
    } guard{blk};
    auto dptr = [&](size_t o) { return reinterpret_cast<double *>(blk + o); };
    auto iptr = [&](size_t o) { return reinterpret_cast<int *>(blk + o); };
    CU(ctx, cudaMemsetAsync(blk + o_done, 0, 256, st));
    CU(ctx, cudaMemcpyAsync(blk + o_q0, q0, (size_t)B * dim * 8, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(blk + o_x, X, xb, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(blk + o_y, Y, yb, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(blk + o_s2, sigma2, sb, cudaMemcpyHostToDevice, st));
    mp.B = B;
    mp.chain_offset = opts->chain_offset;
    mp.n_rec = n_rec;
    mp.state = reinterpret_cast<ChainState *>(blk + o_state);
    mp.vec = dptr(o_vec);
    mp.vec_stride = (long long)vec_stride;
    mp.q0 = dptr(o_q0);
    mp.Y = dptr(o_y);
    mp.y_stride = y_batched ? n : 0;
    mp.theta_eval = dptr(o_the);
    mp.y_eval = dptr(o_ye);
    mp.lml = dptr(o_lml);
    mp.dtheta = dptr(o_dth);
    mp.dy = dptr(o_dy);
    mp.info = iptr(o_info);
    mp.theta_out = dptr(o_th);
    mp.lp_out = dptr(o_lp);
    mp.accept_out = dptr(o_acc);
    mp.eps_out = dptr(o_eps);
    mp.q_out = q ? dptr(o_qo) : nullptr;
    mp.depth_out = iptr(o_dep);
    mp.nleap_out = iptr(o_nl);
    mp.div_out = iptr(o_div);
    mp.done = reinterpret_cast<unsigned int *>(blk + o_done);
    mp.n_active = reinterpret_cast<int *>(blk + o_done + 64);
    mp.slot_chain = iptr(o_slot);
    mp.n_slots = B;
    const int warps_per_block = 4;
    auto warp_grid = [&](int items) { return (items + warps_per_block - 1) / warps_per_block; };
    mcmc_init_kernel<<<warp_grid(B), 32 * warps_per_block, 0, st>>>(mp);
    ctx->launches++;
    int n_slots = B;
    const double *xs = dptr(o_x), *ys = dptr(o_y), *s2s = dptr(o_s2);  // evaluation inputs in slot order
    // one step: evaluate log-density + gradient at every chain's emitted point, then advance every chain
    const int saved_profile = ctx->profile_events;
    ctx->profile_events = 0;
    auto step = [&]() -> int {
        int r = launch_lml(ctx, prog->dev, n, d, xs, x_batched, c.latent ? dptr(o_ye) : ys, c.latent ? 1 : y_batched, dptr(o_the), p,
                           s2s, sigma2_batched, jitter, n_slots, dptr(o_lml), dptr(o_dth), dptr(o_dy), iptr(o_info), 1, 0, nullptr,
                           nullptr, st);
        if (r) return r;
        mcmc_advance_kernel<<<warp_grid(n_slots), 32 * warps_per_block, 0, st>>>(mp);
        ctx->launches++;
        return GPL_OK;
    };
    const uint64_t launches_before = ctx->launches;
    rc = step();  // the first step also sizes the workspaces (no allocation may happen inside a capture)
    uint64_t launches_per_step = ctx->launches - launches_before;  // re-counted at every capture (fewer slots, fewer chunks)
    long long steps = 1, evals = B;  // evaluations actually launched: slots per step (finished chains leave at compactions)
    unsigned int done = 0;
    if (rc == GPL_OK) {
        // capture one step and replay it; the host only polls the done counter every `poll` steps
        bool graph_ok = false;
        auto capture = [&]() {
            if (guard.ge) cudaGraphExecDestroy(guard.ge);
            if (guard.g) cudaGraphDestroy(guard.g);
            guard.ge = nullptr;
            guard.g = nullptr;
            graph_ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (graph_ok) {
                const uint64_t before = ctx->launches;
                const int r = step();
                launches_per_step = ctx->launches - before;
                ctx->launches = before;  // captured, not launched
                cudaError_t e = cudaStreamEndCapture(st, &guard.g);
                graph_ok = r == GPL_OK && e == cudaSuccess && guard.g && cudaGraphInstantiate(&guard.ge, guard.g, 0) == cudaSuccess;
                if (!graph_ok) cudaGetLastError();
            }
        };
        capture();
        const long long max_steps = (long long)T * ((1LL << c.max_depth) + 1) + 64;
        const int poll = 16;
        while (rc == GPL_OK && done < (unsigned)B && steps < max_steps) {
            for (int k = 0; k < poll && rc == GPL_OK; ++k) {
                if (graph_ok) {
                    if (cudaGraphLaunch(guard.ge, st) != cudaSuccess) rc = fail(ctx, GPL_ERR_CUDA, "gpl_mcmc_nuts: graph launch failed");
                    ctx->launches += launches_per_step;
                } else {
                    rc = step();
                }
                ++steps;
                evals += n_slots;
            }
            if (rc) break;
            CU(ctx, cudaMemcpyAsync(&done, mp.done, sizeof(done), cudaMemcpyDeviceToHost, st));
            CU(ctx, cudaStreamSynchronize(st));
            // Compaction: chains finish at different times (their trees differ); once a quarter of the slots idle, the
            // active chains move to the first slots and the evaluation batch shrinks to them.
            const int active = B - (int)done;
            if (active > 0 && n_slots >= 32 && active <= n_slots - n_slots / 4) {
                mcmc_compact_kernel<<<1, 256, 0, st>>>(mp);
                int na = 0;
                CU(ctx, cudaMemcpyAsync(&na, mp.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
                CU(ctx, cudaStreamSynchronize(st));
                if (na < 1 || na > n_slots) return fail(ctx, GPL_ERR_CUDA, "gpl_mcmc_nuts: compaction returned %d active chains", na);
                n_slots = mp.n_slots = na;
                mcmc_reemit_kernel<<<warp_grid(n_slots), 32 * warps_per_block, 0, st>>>(mp);
                ctx->launches += 2;
                auto gather = [&](size_t src_off, size_t dst_off, long long width, const double *&cur) {
                    mcmc_gather_kernel<<<256, 256, 0, st>>>(dptr(src_off), dptr(dst_off), mp.slot_chain, n_slots, width);
                    ctx->launches++;
                    cur = dptr(dst_off);
                };
                if (x_batched) gather(o_x, o_xs, (long long)n * d, xs);
                if (y_batched && !c.latent) gather(o_y, o_ys, n, ys);
                if (sigma2_batched) gather(o_s2, o_s2s, 1, s2s);
                CU(ctx, cudaGetLastError());
                capture();
            }
        }
    }
    ctx->profile_events = saved_profile;
    if (rc) return rc;
    mcmc_status_kernel<<<(B + 127) / 128, 128, 0, st>>>(mp.state, B, iptr(o_stat));
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    if (p > 0) CU(ctx, cudaMemcpyAsync(theta, blk + o_th, (size_t)B * n_rec * p * 8, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaMemcpyAsync(lp, blk + o_lp, (size_t)B * n_rec * 8, cudaMemcpyDeviceToHost, st));
    if (q) CU(ctx, cudaMemcpyAsync(q, blk + o_qo, (size_t)B * n_rec * dim * 8, cudaMemcpyDeviceToHost, st));
    if (accept) CU(ctx, cudaMemcpyAsync(accept, blk + o_acc, (size_t)B * n_rec * 8, cudaMemcpyDeviceToHost, st));
    if (eps) CU(ctx, cudaMemcpyAsync(eps, blk + o_eps, (size_t)B * n_rec * 8, cudaMemcpyDeviceToHost, st));
    if (depth) CU(ctx, cudaMemcpyAsync(depth, blk + o_dep, (size_t)B * n_rec * 4, cudaMemcpyDeviceToHost, st));
    if (n_leapfrog) CU(ctx, cudaMemcpyAsync(n_leapfrog, blk + o_nl, (size_t)B * n_rec * 4, cudaMemcpyDeviceToHost, st));
    if (divergent) CU(ctx, cudaMemcpyAsync(divergent, blk + o_div, (size_t)B * n_rec * 4, cudaMemcpyDeviceToHost, st));
    if (status) CU(ctx, cudaMemcpyAsync(status, blk + o_stat, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    if (n_grad_evals) *n_grad_evals = evals;
    return GPL_OK;
}

// ---- sample --------------------------------------------------------------------------------------------------------
int gpl_sample(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, const double *theta, int p,
               double sigma2, double jitter, const double *Z, int S, double *out) {
    if (!Z || !out || S <= 0) return fail(ctx, GPL_ERR_ARG, "gpl_sample: bad argument");
    int rc = check_prog_args(ctx, prog, n, d, p);
    if (rc) return rc;
    if (!X || (p > 0 && !theta)) return fail(ctx, GPL_ERR_ARG, "gpl_sample: null pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    WsOrder order(ctx, ctx->stream);
    gpl_post *post = new (std::nothrow) gpl_post();
    if (!post) return fail(ctx, GPL_ERR_ARG, "out of host memory");
    double *zeros = new (std::nothrow) double[n]();
    rc = zeros ? posterior_fit_impl(ctx, prog, n, d, X, zeros, theta, p, sigma2, jitter, post) : GPL_ERR_ARG;
    delete[] zeros;
    if (rc == GPL_OK) {
        cudaStream_t st = ctx->stream;
        rc = ensure(ctx, ctx->bK, (size_t)2 * n * S * 8);
        if (rc == GPL_OK) {
            double *dZ = ptr<double>(ctx->bK), *dOut = dZ + (size_t)n * S;
            cudaError_t e = cudaMemcpyAsync(dZ, Z, (size_t)n * S * 8, cudaMemcpyHostToDevice, st);
            sample_kernel<<<post->nt, NTHREADS, 0, st>>>(post->tiles, post->nt, n, dZ, S, dOut);
            ctx->launches++;
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(out, dOut, (size_t)n * S * 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) rc = fail(ctx, GPL_ERR_CUDA, "gpl_sample: %s", cudaGetErrorString(e));
        }
    }
    cudaStreamSynchronize(ctx->stream);
    posterior_release(post);
    return rc;
}

// ---- large-n ---------------------------------------------------------------------------------------------------------
static int chol_logdet_dev_locked(gpl_ctx *ctx, int n, double *dA, int want_factor, double *dlogdet, int *dinfo,
                                  cudaStream_t st) {
    const int nt = (n + TS - 1) / TS;
    const long long ntri = tri_index(nt, 0);
    int rc;
    if ((rc = ensure(ctx, ctx->ws, (size_t)(ntri + nt) * TILE_BYTES)) || (rc = ensure(ctx, ctx->bMisc, 64 + (size_t)nt * TS * 8)))
        return rc;
    double *tiles = ptr<double>(ctx->ws), *winv = tiles + ntri * TILE_ELEMS;
    double *dres = ptr<double>(ctx->bMisc), *pivlog = dres + 8;
    dense_to_tiles_kernel<<<(unsigned)ntri, NTHREADS, 0, st>>>(dA, n, nt, tiles);
    ctx->launches++;
    if ((rc = big_factor(ctx, tiles, winv, pivlog, dinfo, nullptr, nt, st))) return rc;
    big_reduce_kernel<<<1, NTHREADS, 0, st>>>(pivlog, nullptr, nt * TS, dres);
    ctx->launches++;
    CU(ctx, cudaMemcpyAsync(dlogdet, dres, 8, cudaMemcpyDeviceToDevice, st));
    if (want_factor) {
        dim3 grid(nt, nt);
        tiles_to_upper_kernel<<<grid, NTHREADS, 0, st>>>(tiles, n, nt, dA);
        ctx->launches++;
    }
    CU(ctx, cudaGetLastError());
    return GPL_OK;
}

int gpl_chol_logdet_dev(gpl_ctx *ctx, int n, double *dA, int want_factor, double *dlogdet, int *dinfo, void *stream) {
    if (!ctx || !dA || !dlogdet || !dinfo || n <= 0) return fail(ctx, GPL_ERR_ARG, "gpl_chol_logdet_dev: bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    WsOrder order(ctx, (cudaStream_t)stream);
    return chol_logdet_dev_locked(ctx, n, dA, want_factor, dlogdet, dinfo, (cudaStream_t)stream);
}

int gpl_chol_logdet(gpl_ctx *ctx, int n, double *A, int want_factor, double *logdet, int *info) {
    if (!ctx || !A || !logdet || n <= 0) return fail(ctx, GPL_ERR_ARG, "gpl_chol_logdet: bad argument");
    // one critical section from the upload to the download: the staging buffer bK belongs to this call throughout
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    WsOrder order(ctx, st);
    int rc;
    if ((rc = ensure(ctx, ctx->bK, (size_t)n * n * 8)) || (rc = ensure(ctx, ctx->bLml, 16)) || (rc = ensure(ctx, ctx->bInfo, 4)))
        return rc;
    CU(ctx, cudaMemcpyAsync(ctx->bK.p, A, (size_t)n * n * 8, cudaMemcpyHostToDevice, st));
    if ((rc = chol_logdet_dev_locked(ctx, n, ptr<double>(ctx->bK), want_factor, ptr<double>(ctx->bLml), ptr<int>(ctx->bInfo), st)))
        return rc;
    int h_info = 0;
    CU(ctx, cudaMemcpyAsync(logdet, ctx->bLml.p, 8, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaMemcpyAsync(&h_info, ctx->bInfo.p, 4, cudaMemcpyDeviceToHost, st));
    if (want_factor) CU(ctx, cudaMemcpyAsync(A, ctx->bK.p, (size_t)n * n * 8, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    if (info) *info = h_info;
    if (h_info) *logdet = NAN;
    if (h_info < 0) return fail(ctx, GPL_ERR_CUDA, "large-n factorisation: a hand-off between its kernels timed out (device busy?)");
    return GPL_OK;
}

int gpl_lml_large(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, const double *y,
                  const double *theta, int p, double sigma2, double jitter, double *lml, double *logdet, int *info) {
    int rc = check_prog_args(ctx, prog, n, d, p);
    if (rc) return rc;
    if (!X || !y || !lml || (p > 0 && !theta)) return fail(ctx, GPL_ERR_ARG, "gpl_lml_large: null pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    WsOrder order(ctx, st);
    const int nt = (n + TS - 1) / TS;
    const long long ntri = tri_index(nt, 0);
    if ((rc = ensure(ctx, ctx->ws, (size_t)(ntri + nt) * TILE_BYTES)) || (rc = ensure(ctx, ctx->bMisc, 64 + (size_t)nt * TS * 8)) ||
        (rc = ensure(ctx, ctx->bX, (size_t)n * d * 8)) || (rc = ensure(ctx, ctx->bTheta, (size_t)(p + 1) * 8)) ||
        (rc = ensure(ctx, ctx->bY, (size_t)nt * TS * 8)))
        return rc;
    double *tiles = ptr<double>(ctx->ws), *winv = tiles + ntri * TILE_ELEMS;
    double *dres = ptr<double>(ctx->bMisc), *pivlog = dres + 8;
    int *dinfo = reinterpret_cast<int *>(dres + 4);
    CU(ctx, cudaMemcpyAsync(ctx->bX.p, X, (size_t)n * d * 8, cudaMemcpyHostToDevice, st));
    if (p > 0) CU(ctx, cudaMemcpyAsync(ctx->bTheta.p, theta, (size_t)p * 8, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemsetAsync(ctx->bY.p, 0, (size_t)nt * TS * 8, st));
    CU(ctx, cudaMemcpyAsync(ctx->bY.p, y, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CovTilesParams cp;
    cp.prog = prog->dev;
    cp.n = n;
    cp.nt = nt;
    cp.d = d;
    cp.p = p;
    cp.X = ptr<double>(ctx->bX);
    cp.theta = ptr<double>(ctx->bTheta);
    cp.diag_add = sigma2 + jitter;
    cp.tiles = tiles;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (ctx->profile_events) {
        for (auto &e : ev) cudaEventCreate(&e);
        cudaEventRecord(ev[0], st);
    }
    cov_tiles_kernel<<<(unsigned)ntri, NTHREADS, 0, st>>>(cp);
    ctx->launches++;
    if (ctx->profile_events) cudaEventRecord(ev[1], st);
    if ((rc = big_factor(ctx, tiles, winv, pivlog, dinfo, ptr<double>(ctx->bY), nt, st))) return rc;
    if (ctx->profile_events) cudaEventRecord(ev[2], st);
    big_reduce_kernel<<<1, NTHREADS, 0, st>>>(pivlog, ptr<double>(ctx->bY), nt * TS, dres);
    ctx->launches++;
    CU(ctx, cudaGetLastError());
    double h_res[2];
    int h_info = 0;
    CU(ctx, cudaMemcpyAsync(h_res, dres, 16, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaMemcpyAsync(&h_info, dinfo, 4, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    if (ctx->profile_events) {  // lk_ms[0] = covariance build, lk_ms[1] = factorisation + forward solve (device time)
        float a_ms = 0.f, b_ms = 0.f;
        cudaEventElapsedTime(&a_ms, ev[0], ev[1]);
        cudaEventElapsedTime(&b_ms, ev[1], ev[2]);
        ctx->lk_ms[0] = a_ms;
        ctx->lk_ms[1] = b_ms;
        ctx->lk_ms[2] = 0.0;
        for (auto &e : ev) cudaEventDestroy(e);
    }
    if (h_info < 0) return fail(ctx, GPL_ERR_CUDA, "large-n factorisation: a hand-off between its kernels timed out (device busy?)");
    if (info) *info = h_info;
    if (logdet) *logdet = h_info ? NAN : h_res[0];
    *lml = h_info ? -INFINITY : -0.5 * ((double)n * LOG2PI + h_res[0] + h_res[1]);
    return GPL_OK;
}

}  // extern "C"
