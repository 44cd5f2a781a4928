/*
 * gaplac_b200.h — C ABI of libgaplac_b200.so, the B200 (sm_100a) backend for GaPLAC's Gaussian-process
 * marginal-likelihood / posterior hot path.
 *
 * The reference (biobakery/GaPLAC, pure Julia) has no FFI for this path; the boundary this library sits
 * behind is the AbstractGPs/KernelFunctions call surface GaPLAC uses.  Each entry point
 * below names the reference call it replaces (paths relative to the reference tree).  A Julia host binds
 * them with `ccall` (julia/GaPLACB200.jl); in this repository the same symbols are bound
 * with Python ctypes (gaplac_b200/_lib.py).
 *
 * Conventions
 *   - all matrices column-major (Julia / LAPACK): X is n x d with leading dimension n, Theta is p x B,
 *     Y is n x B, K is n x n;
 *   - all host pointers are caller-owned; host entry points copy H2D/D2H on the context's stream and
 *     return after the results are in the caller's buffers (blocking, like LAPACK);
 *   - *_dev entry points take DEVICE pointers and a CUDA stream (cudaStream_t passed as void*) and return
 *     after enqueueing (asynchronous); the caller synchronises the stream;
 *   - every function returns GPL_OK (0) or a negative gpl_status; nothing throws or aborts;
 *     gpl_last_error() gives the message.  Per-item numerical failures (matrix not positive definite) are
 *     NOT API errors: they are reported LAPACK-style in info[b] (1-based failing pivot) with lml[b] = -Inf
 *     (the Julia shim turns info != 0 into PosDefException(info), as `cholesky` does [upstream]).
 *   - there is no CPU fallback: without a CUDA device gpl_init fails with GPL_ERR_CUDA.
 */
#ifndef GAPLAC_B200_H
#define GAPLAC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPL_ABI_VERSION 2 /* 2: + gpl_mcmc_nuts, gpl_multi_*, gpl_set_stream, gpl_last_timing, gpl_release_workspace */

typedef enum gpl_status {
    GPL_OK = 0,
    GPL_ERR_ARG = -1,     /* bad argument (null pointer, size, malformed program) */
    GPL_ERR_CUDA = -2,    /* CUDA runtime error, or no usable device */
    GPL_ERR_LIMIT = -3,   /* program too large (terms / factors / slots) or n beyond a kernel limit */
    GPL_ERR_NOTPD = -4    /* single-model calls only: covariance not positive definite (info in message) */
} gpl_status;

/* Node kinds of the kernel-program: the leaves of src/gp_parts.jl:21-47 (SqExp, Linear, OU, Cat) with
 * the k(x,x') of src/abstractgp_translations.jl:8-15 and src/gp_parts.jl:11-13, the two components the
 * README/legacy fixtures name but the current src/ lacks (Constant, Noise), and the two
 * operations of src/gp_parts.jl:55,59 (`+` -> :add, `*` -> :multiply). */
typedef enum gpl_kind {
    GPL_SQEXP = 0,    /* exp(-(x-x')^2 / (2 l^2))      hyperparameter l (> 0) */
    GPL_OU = 1,       /* exp(-|x-x'| / l)               hyperparameter l (> 0).  A FIXED l <= 0 is rejected by
                         gpl_program_create; an l <= 0 (or NaN) arriving through a theta slot makes that item's covariance
                         NaN: info[b] = 1, lml[b] = -Inf, NaN gradient / predictions (the reference's ScaleTransform
                         throws for l <= 0 [upstream]; a sampler sees a rejected proposal) */
    GPL_LINEAR = 2,   /* x x' + c                       hyperparameter c */
    GPL_CAT = 3,      /* x == x' ? 1 : 0                none */
    GPL_CONSTANT = 4, /* c                              hyperparameter c */
    GPL_NOISE = 5,    /* delta_ij by row index on K(X,X); 0 on cross-covariances and on diag K(X*,X*) */
    GPL_ADD = 6,      /* pops two, pushes lhs + rhs */
    GPL_MUL = 7       /* pops two, pushes lhs * rhs */
} gpl_kind;

/* One postfix instruction.  The formula AST (src/gp_parts.jl:3-9) is flattened left to right, so the
 * i-th leaf reads column `col` exactly as `kernel()` binds the i-th leaf to the i-th column with
 * SelectTransform (src/abstractgp_translations.jl:45-71).  After a node's value is computed it is
 * multiplied by a variance: theta[var_slot] if var_slot >= 0, else `var` (1.0 = reference semantics). */
typedef struct gpl_op {
    int32_t kind;       /* gpl_kind */
    int32_t col;        /* input column of a leaf (ignored by CONSTANT, NOISE, ADD, MUL) */
    int32_t theta_slot; /* slot of l / c in the per-item hyperparameter vector; -1: use `value` */
    int32_t var_slot;   /* slot of the variance multiplier; -1: use `var` */
    double value;       /* fixed hyperparameter when theta_slot < 0 */
    double var;         /* fixed variance multiplier when var_slot < 0 */
} gpl_op;

#define GPL_MAX_OPS 64
#define GPL_MAX_TERMS 16   /* additive terms after expansion to a sum of products */
#define GPL_MAX_FACTORS 48 /* leaf / parameter factors over all terms */
#define GPL_MAX_THETA 16   /* hyperparameter slots per item */
#define GPL_MAX_COLS 16    /* input columns d */

typedef struct gpl_ctx gpl_ctx;   /* device, stream, workspace pool */
typedef struct gpl_prog gpl_prog; /* compiled kernel-program */
typedef struct gpl_post gpl_post; /* posterior: Cholesky factor and alpha resident in HBM */

/* ---- context ------------------------------------------------------------------------------------- */
/* device = CUDA ordinal, or -1 for the current device.  Host entry points on one context serialise (mutex + the
 * context's stream).  The *_dev entry points enqueue on the CALLER's stream and return; the context owns one workspace,
 * so every call first makes its stream wait (cudaStreamWaitEvent) for the previous call that used the workspace, on
 * whatever stream that was: calls on one context never overlap on the device, they only pipeline.  Use one context per
 * concurrent chain for real concurrency.  gpl_destroy synchronises the device, releases the device memory of any
 * posterior still alive and detaches it: gpl_posterior_free on such a handle is still valid (and required), every other
 * call on it returns GPL_ERR_ARG. */
int gpl_init(int device, gpl_ctx **out);
int gpl_destroy(gpl_ctx *ctx);
/* Run the host entry points of this context on the caller's CUDA stream (cudaStream_t as void*) instead of the
 * context's own one; NULL restores it.  The calls stay blocking; a host that records CUDA events on its stream (or
 * orders other work on it, e.g. CUDA.jl's task stream) then sees the library's copies and kernels in that order. */
int gpl_set_stream(gpl_ctx *ctx, void *stream);
/* The context's workspace is grow-only (a 4096 x n=512 gradient batch holds ~11 GB): this releases all of it (and the
 * recycled posterior blocks); the next call re-allocates what it needs.  Live posteriors are not touched. */
int gpl_release_workspace(gpl_ctx *ctx);
const char *gpl_last_error(gpl_ctx *ctx); /* ctx may be NULL: last error of the calling thread */
int gpl_abi_version(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t gpl_launch_count(gpl_ctx *ctx);
/* tuning knobs for experiments; unknown keys return GPL_ERR_ARG.  Keys: "lml_variant" (0 lockstep schedule, 1 fused
 * per-item kernel of round 1), "chol_variant" (1: force the multi-CTA large-n path; 2: its one-stream form; 3: look-ahead streams without the worker CTA;
 * 4: the round-1 worker protocol), "trail_int8" (trailing updates of the large-n factorisation on the INT8 tensor path
 * - tcgen05.mma.kind::i8 on 7-bit slices of L, exact INT32 products, FP64 accumulation: -1 (default) automatic = 9 slices,
 * 63 bits below each row's maximum (as accurate as the FP64 path on ill-conditioned covariances too), from n = 6144 on;
 * 0: FP64 DMMA updates at every n; 5..9: that many slices from n = 4096 on - on well-conditioned problems 8 slices stay within
 * 1e-14 of the FP64 path and are 10-20 % faster, 7 slices keep the lml to ~1e-12, 6 to ~1e-10), "lk_ws_limit_mb" (workspace cap, default 24576),
 * "ou_separable" (default 1: batched log-densities with n > 192 whose program has one or two OU leaves on one column of a
 * shared X sort the observations by that column - the likelihood does not depend on their order; dy is returned in the
 * caller's order - and evaluate those leaves in separable form below the diagonal; 0: off; 2: from n > 64 on),
 * "zero_tile_skip" (default 1: batched factorisations of programs whose every term but the noise carries a Cat factor -
 * block-diagonal covariances once the rows are grouped, e.g. Cat(:subject) * SqExp(:time) - skip the updates with, and the
 * triangular solves of, tiles of L that are exactly zero; results keep every bit; 2: flags for every program (short length
 * scales underflow to exact zeros too); 3: as 1, and the observations are first grouped by the category column those terms
 * share - for rows that do not arrive grouped; lml and dtheta do not depend on the order, dy is returned in the caller's;
 * 0: off),
 * "profile_events" (1: per-phase CUDA-event timing, see gpl_last_timing), "poison_ws" (1: the context fills its whole
 * workspace with NaN payloads before every call - a debugging aid: results must not change) */
int gpl_set_option(gpl_ctx *ctx, const char *key, int value);
/* Optional per-call statistics.  With option "profile_events" = 1 the library brackets every
 * kernel launch of gpl_lml_batched(_dev) / gpl_lml_large with CUDA events (and synchronises the stream at the end of the
 * call); gpl_last_timing returns the device time per phase of the last such call.  Phases of the batched calls:
 * 0 lk_diag, 1 lk_potrf_warp, 2 lk_below (factorisation), 3 lk_winv, 4 lk_minv, 5 lk_alpha, 6 lk_gradc (gradient);
 * gpl_lml_large: 0 covariance build, 1 factorisation + forward solve. */
typedef struct gpl_timing {
    int32_t n_phases;
    int32_t launches[8];
    double ms[8];
} gpl_timing;
int gpl_last_timing(gpl_ctx *ctx, gpl_timing *out);
/* device facts for reports: name (<= len bytes), SM count, SM clock kHz */
int gpl_device_info(gpl_ctx *ctx, char *name, int len, int *sm_count, int *clock_khz);

/* ---- kernel program: replaces makekernel/_convert2eq/kernel (src/abstractgp_translations.jl:8-71) -- */
int gpl_program_create(gpl_ctx *ctx, const gpl_op *ops, int n_ops, gpl_prog **out);
int gpl_program_destroy(gpl_prog *prog);
int gpl_program_n_theta(const gpl_prog *prog); /* 1 + highest slot referenced */
int gpl_program_n_cols(const gpl_prog *prog);  /* 1 + highest column referenced */

/* ---- covariance construction: replaces kernelmatrix(k, RowVecs(X)) + sigma2*I ----------------------
 * [upstream KernelFunctions] reached from FiniteGP at CLI/src/mcmc.jl:35, CLI/src/select.jl:43,47,
 * CLI/src/sample.jl:25, src/plotting.jl:6.  K (n x n) = K(X,X) + (sigma2 + jitter) I. */
int gpl_cov(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, const double *theta, int p,
            double sigma2, double jitter, double *K);
int gpl_cov_dev(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *dX, const double *dtheta, int p,
                double sigma2, double jitter, double *dK, void *stream);
/* cross-covariance K(X, Xs) (n x m), Noise contributes 0: the K* of mean_and_var (src/plotting.jl:12) */
int gpl_cross_cov(gpl_ctx *ctx, const gpl_prog *prog, int n, int m, int d, const double *X, const double *Xs,
                  const double *theta, int p, double *Ks);

/* ---- batched log marginal likelihood: replaces logpdf(FiniteGP, y) ------------------------------------
 * [upstream AbstractGPs]; call sites CLI/src/select.jl:49-50 and, per leapfrog step, CLI/src/mcmc.jl:35.
 *   lml[b] = -1/2 ( n log 2pi + logdet K_b + y_b' K_b^-1 y_b ),  K_b = K(X_b,X_b; theta_b) + (sigma2_b + jitter) I
 * Batch modes: X shared (x_batched = 0, n x d) or per item (n x d x B); Y shared (n) or per item (n x B);
 * sigma2 shared (1 value) or per item (B values); Theta always p x B.
 * Optional analytic gradient (the reference gets it by ForwardDiff through the model body,
 * CLI/src/mcmc.jl:31-37): dtheta (p x B) = dlml/dtheta, dy (n x B) = dlml/dy = -K^-1 y.  NULL to skip. */
int gpl_lml_batched(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, int x_batched,
                    const double *Y, int y_batched, const double *Theta, int p, const double *sigma2,
                    int sigma2_batched, double jitter, int B, double *lml, double *dtheta, double *dy, int *info);
int gpl_lml_batched_dev(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *dX, int x_batched,
                        const double *dY, int y_batched, const double *dTheta, int p, const double *dsigma2,
                        int sigma2_batched, double jitter, int B, double *dlml, double *ddtheta, double *ddy,
                        int *dinfo, void *stream);

/* ---- posterior: replaces posterior(FiniteGP, y) and mean_and_var(PosteriorGP, X*) ----------------------
 * [upstream AbstractGPs]; call sites CLI/src/select.jl:51-52, src/plotting.jl:8 and src/plotting.jl:12.
 * The factor and alpha = K^-1 y stay in HBM inside the handle. */
int gpl_posterior_fit(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, const double *y,
                      const double *theta, int p, double sigma2, double jitter, gpl_post **out);
int gpl_posterior_free(gpl_post *post);
int gpl_posterior_logpdf(gpl_post *post, double *lml);  /* logpdf from the same factorisation */
int gpl_posterior_alpha(gpl_post *post, double *alpha); /* n values */
int gpl_posterior_factor(gpl_post *post, double *U);    /* n x n upper factor (K = U'U), zeros below */
/* mean[m] = K(X*,X) alpha ; var[m] = diag K(X*,X*) - colsumsq(U' \ K(X,X*)) (latent f; sigma2 not added).
 * Xs is m x d column-major. var may be NULL. Tiled over m: K* is never materialised. */
int gpl_posterior_mean_var(gpl_post *post, int m, const double *Xs, double *mean, double *var);

/* ---- batched posteriors + predictions: the loop behind the `predict` / `fitplot` commands ----------------
 * For every row b of a chain of hyperparameter draws (Theta p x B; sigma2 shared or per row):
 *   post_b = posterior(FiniteGP(GP(kernel(theta_b)), X, sigma2_b), y);  mean_and_var(post_b, Xs)
 * (src/plotting.jl:6-12 per row; commands stubbed at CLI/src/main.jl:8-16, output columns test/pred.jl:11-14).
 * X (n x d) and y (n) are shared by the rows.  mean, var: m x B column-major (row b at b*m; var may be NULL);
 * lml (B, optional) and info (B, optional) as in gpl_lml_batched: a row whose covariance is not positive
 * definite has info[b] != 0, lml[b] = -Inf and NaN predictions.  The rows are factored in lockstep batches of as many
 * rows as the factor workspace cap ("lk_ws_limit_mb") holds; longer chains run in several passes inside the call. */
int gpl_predict_batched(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, const double *y,
                        const double *Theta, int p, const double *sigma2, int sigma2_batched, double jitter, int B,
                        int m, const double *Xs, double *mean, double *var, double *lml, int *info);

/* ---- batched on-device sampler: replaces sample(m, NUTS(0.65), N) on the mcmc model body -------------------------------
 * CLI/src/mcmc.jl:31-41 [upstream Turing 0.21.1 / AdvancedHMC 0.3.5].  The model of every chain b (position in
 * unconstrained space q = (u, fx)):
 *     theta_k = lo_k + (hi_k - lo_k) sigmoid(u_k)      hyperparameter slots, Uniform(lo_k, hi_k) priors   (mcmc.jl:32)
 *     fx ~ N(0, K(X; theta) + (sigma2 + jitter) I)     latent function values                              (mcmc.jl:35)
 *     Y_b ~ N(fx, obs_sd^2 I)                          observations                                        (mcmc.jl:36)
 * (latent = 0: Y_b ~ N(0, K + (sigma2 + jitter) I) directly, hyperparameters only.)  B independent chains - one per
 * column of Y (y_batched: the features of a table) or B chains of the same response - advance in lockstep: one batched
 * log-density + analytic-gradient evaluation per leapfrog step for all chains, then one state-machine kernel; proposals,
 * multinomial NUTS trees (generalised U-turn criterion, depth <= max_depth, divergence threshold max_dh), accept/reject
 * and the warm-up (step-size search, dual averaging to `delta`, Stan's windowed diagonal-metric adaptation) stay on the
 * device.  Random numbers: Philox4x32-10 keyed by `seed`, counter = (chain_offset + b, transition, index, purpose): a
 * chain's draws do not depend on B, on the device count or on the execution order (oracle/nuts_ref.py replays them).
 * Outputs, one record per kept transition (n_rec = n_samples, + n_adapt when record_warmup), chain-major:
 *   theta (p x n_rec x B), lp (n_rec x B; constrained-space log joint, Turing's `lp` column read by select --chains),
 *   q (dim x n_rec x B, optional), accept, eps (n_rec x B), depth, n_leapfrog, divergent (n_rec x B ints),
 *   status (B): 0 ok, 1 the initial point q0 has zero density, 2 not finished.  n_grad_evals: log-density + gradient
 *   evaluations launched (one per active slot and leapfrog step; finished chains are compacted out of the batch). */
typedef struct gpl_mcmc_opts {
    int32_t n_samples;     /* N of `--samples` (CLI/src/main.jl:65-71, default 200) */
    int32_t n_adapt;       /* warm-up transitions; < 0: min(1000, N / 2) as Turing's NUTS(0.65) */
    int32_t max_depth;     /* <= 10; 0: 10 */
    int32_t latent;        /* 1: the reference's model */
    int32_t search_eps;    /* 1: find the initial step size (doubling heuristic) starting from eps0; 0: use eps0 */
    int32_t adapt_mass;    /* 1: windowed diagonal-metric adaptation */
    int32_t record_warmup; /* 1: also return the warm-up transitions */
    int32_t chain_offset;  /* global index of chain 0 (sharding chains over devices keeps every chain's random stream) */
    double delta;          /* target acceptance (0.65) */
    double max_dh;         /* divergence threshold (1000) */
    double obs_sd;         /* 1 */
    double eps0;           /* 0.1 */
    uint64_t seed;
} gpl_mcmc_opts;
int gpl_mcmc_nuts(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, int x_batched, const double *Y,
                  int y_batched, int p, const double *lo, const double *hi, const double *sigma2, int sigma2_batched,
                  double jitter, int B, const double *q0, const gpl_mcmc_opts *opts, double *theta, double *lp, double *q,
                  double *accept, double *eps, int *depth, int *n_leapfrog, int *divergent, int *status,
                  long long *n_grad_evals);

/* ---- prior sample: replaces rand(gp(X, sigma2)) (CLI/src/sample.jl:25) --------------------------------
 * out (n x S) = U' Z with caller-supplied standard normals Z (n x S): the RNG stays in the host. */
int gpl_sample(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, const double *theta, int p,
               double sigma2, double jitter, const double *Z, int S, double *out);

/* ---- single large-n factorisation (BASELINE config 5): cholesky(Symmetric(A)) + logdet -----------------
 * A (n x n, column-major, symmetric; only the lower triangle is read) is overwritten by the upper factor U
 * when `want_factor` != 0; logdet = 2 sum log U_ii.  *info = 0 or the failing pivot; -1 (with GPL_ERR_CUDA from the host
 * entry points) if a hand-off between the concurrently running kernels of the factorisation timed out (bounded waits:
 * a device kept busy by other work yields an error, never a hang). */
int gpl_chol_logdet(gpl_ctx *ctx, int n, double *A, int want_factor, double *logdet, int *info);
int gpl_chol_logdet_dev(gpl_ctx *ctx, int n, double *dA, int want_factor, double *dlogdet, int *dinfo, void *stream);
/* fused: build K_y from the program on the device, factor it, return logdet and lml without K ever
 * crossing PCIe (the large-n model of config 5). */
int gpl_lml_large(gpl_ctx *ctx, const gpl_prog *prog, int n, int d, const double *X, const double *y,
                  const double *theta, int p, double sigma2, double jitter, double *lml, double *logdet, int *info);

/* ---- several GPUs behind one call (SURVEY.md 8(b) "device(s)", 8(e) "one ccall") ------------------------------------------
 * A gpl_multi owns one context per listed device (devices = NULL: ordinals 0 .. n_devices-1; a device may be listed
 * twice).  The batched calls split the B independent items into contiguous blocks - the first B mod R blocks one item
 * longer, the rule of gaplac_b200/shard.py - run every block concurrently on its device and let each device write its
 * slice of the outputs straight into the caller's host buffers: the gather IS the device-to-host copy, there is no
 * collective on the data path.  Semantics, layouts and per-item error reporting are those of gpl_lml_batched /
 * gpl_mcmc_nuts; results do not depend on the number of devices (chains keep their random streams: chain_offset).
 * A gpl_prog is device-independent and can be shared (gpl_program_create accepts ctx = NULL).  The reference has no
 * counterpart: one Julia task evaluates one model at a time (CLI/src/mcmc.jl:41, CLI/src/select.jl:49-50). */
typedef struct gpl_multi gpl_multi;
int gpl_multi_init(const int *devices, int n_devices, gpl_multi **out);
int gpl_multi_destroy(gpl_multi *m);
int gpl_multi_device_count(const gpl_multi *m);
gpl_ctx *gpl_multi_context(gpl_multi *m, int part); /* the context of part r (options, timing); owned by m */
const char *gpl_multi_last_error(gpl_multi *m);
int gpl_multi_lml_batched(gpl_multi *m, const gpl_prog *prog, int n, int d, const double *X, int x_batched, const double *Y,
                          int y_batched, const double *Theta, int p, const double *sigma2, int sigma2_batched, double jitter,
                          int B, double *lml, double *dtheta, double *dy, int *info);
int gpl_multi_mcmc_nuts(gpl_multi *m, const gpl_prog *prog, int n, int d, const double *X, int x_batched, const double *Y,
                        int y_batched, int p, const double *lo, const double *hi, const double *sigma2, int sigma2_batched,
                        double jitter, int B, const double *q0, const gpl_mcmc_opts *opts, double *theta, double *lp, double *q,
                        double *accept, double *eps, int *depth, int *n_leapfrog, int *divergent, int *status,
                        long long *n_grad_evals);

#ifdef __cplusplus
}
#endif
#endif /* GAPLAC_B200_H */
