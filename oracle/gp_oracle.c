/*
 * CPU oracle (plain C, FP64) for GaPLAC's GP marginal-likelihood / posterior hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  The product (gaplac_b200/, libgaplac_b200.so) never links, loads
 * or calls this file.  It is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs, as the checker and as the CPU yardstick.
 *
 * It restates, in the reference's own call order (citations into /root/reference; [upstream] =
 * un-vendored Julia packages pinned in Manifest.toml: KernelFunctions 0.10.38, AbstractGPs 0.5.12,
 * Distances 0.10.7, LinearAlgebra -> OpenBLAS 0.3.20 dpotrf/dtrtrs):
 *
 *   kernelmatrix        one n x n temporary per node; leaves src/abstractgp_translations.jl:8-15,
 *                       src/gp_parts.jl:11-13; sum/product :21-35; column binding :45-71
 *   FiniteGP            K + sigma2 I                     CLI/src/mcmc.jl:35, CLI/src/select.jl:43,47
 *   logpdf              cholesky (upper), U' \ y, logdet CLI/src/select.jl:49-50  [upstream AbstractGPs]
 *   posterior           alpha = K^-1 y                   CLI/src/select.jl:51-52, src/plotting.jl:8
 *   mean_and_var        K*' alpha ; k** - colsumsq(U'\K*) src/plotting.jl:12
 *   rand                U' z                             CLI/src/sample.jl:25
 *
 * Parity pin: tests/test_oracle_golden.py checks gpo_lml against the 200 known lml values recovered
 * from the reference's legacy fixtures and against oracle/gp_oracle.py (SciPy LAPACK).
 *
 * The factorisation is a self-contained blocked upper Cholesky (no BLAS needed).  For the CPU
 * *baseline timing* the caller may hand in OpenBLAS's dpotrf/dtrsv via gpo_set_lapack(), so that the
 * yardstick is the reference's own LAPACK path rather than this file's plain loops.
 *
 * All matrices are column-major (Julia / LAPACK layout).  X is n x d with leading dimension n.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { GPO_SQEXP = 0, GPO_OU, GPO_LINEAR, GPO_CAT, GPO_CONSTANT, GPO_NOISE, GPO_ADD, GPO_MUL };
enum { GPO_DIRECT = 0, GPO_GEMM = 1 };

typedef struct {
    int32_t kind, col, theta_slot, var_slot;
    double value, var;
} gpo_op; /* same 32-byte layout as gpl_op in include/gaplac_b200.h */

#define LOG2PI 1.8378770664093454835606594728112

/* optional LAPACK hooks (Fortran ABI) for the baseline timing */
typedef void (*dpotrf_fn)(const char *, const int *, double *, const int *, int *);
typedef void (*dtrtrs_fn)(const char *, const char *, const char *, const int *, const int *, const double *,
                          const int *, double *, const int *, int *);
static dpotrf_fn g_dpotrf = 0;
static dtrtrs_fn g_dtrtrs = 0;
void gpo_set_lapack(void *dpotrf, void *dtrtrs) {
    g_dpotrf = (dpotrf_fn)dpotrf;
    g_dtrtrs = (dtrtrs_fn)dtrtrs;
}

/* ------------------------------------------------------------------ kernel program ---------- */
static double sqdist(double a, double b, int same_idx, int mode) {
    if (mode == GPO_DIRECT) {
        double d = a - b;
        return d * d;
    }
    if (same_idx) return 0.0; /* [upstream Distances] exact-zero diagonal */
    double v = a * a + b * b - 2.0 * (a * b);
    return v > 0.0 ? v : 0.0;
}

/* fill out[na*nb] (col-major, ld = na) with one leaf */
static void leaf_matrix(const gpo_op *op, const double *Xa, int na, const double *Xb, int nb, const double *theta,
                        int same, int mode, double *out) {
    double h = op->theta_slot >= 0 ? theta[op->theta_slot] : op->value;
    const double *a = Xa + (size_t)op->col * na, *b = Xb + (size_t)op->col * nb;
    for (int j = 0; j < nb; ++j)
        for (int i = 0; i < na; ++i) {
            double v;
            switch (op->kind) {
            case GPO_SQEXP: v = exp(-sqdist(a[i], b[j], same && i == j, mode) / (2.0 * h * h)); break;
            case GPO_OU: v = exp(-sqrt(sqdist(a[i], b[j], same && i == j, mode)) / h); break;
            case GPO_LINEAR: v = a[i] * b[j] + h; break;
            case GPO_CAT: v = (a[i] == b[j]) ? 1.0 : 0.0; break;
            case GPO_CONSTANT: v = h; break;
            case GPO_NOISE: v = (same && i == j) ? 1.0 : 0.0; break;
            default: v = NAN;
            }
            out[(size_t)j * na + i] = v;
        }
}

/* K(Xa, Xb) (na x nb, col-major).  Returns 0, or -1 on a malformed program. */
int gpo_eval_program(const gpo_op *ops, int n_ops, const double *Xa, int na, const double *Xb, int nb,
                     const double *theta, int same, int mode, double *K) {
    size_t sz = (size_t)na * nb;
    double **stack = (double **)calloc((size_t)n_ops + 1, sizeof(double *));
    int sp = 0, rc = 0;
    for (int t = 0; t < n_ops && rc == 0; ++t) {
        const gpo_op *op = &ops[t];
        double *cur;
        if (op->kind == GPO_ADD || op->kind == GPO_MUL) {
            if (sp < 2) { rc = -1; break; }
            double *rhs = stack[--sp];
            cur = stack[sp - 1];
            if (op->kind == GPO_ADD)
                for (size_t e = 0; e < sz; ++e) cur[e] += rhs[e];
            else
                for (size_t e = 0; e < sz; ++e) cur[e] *= rhs[e];
            free(rhs);
        } else {
            cur = (double *)malloc(sz * sizeof(double));
            leaf_matrix(op, Xa, na, Xb, nb, theta, same, mode, cur);
            stack[sp++] = cur;
        }
        double v = op->var_slot >= 0 ? theta[op->var_slot] : op->var;
        if (v != 1.0)
            for (size_t e = 0; e < sz; ++e) cur[e] *= v;
    }
    if (rc == 0 && sp != 1) rc = -1;
    if (rc == 0) memcpy(K, stack[0], sz * sizeof(double));
    for (int i = 0; i < sp; ++i) free(stack[i]);
    free(stack);
    return rc;
}

/* ------------------------------------------------------------------ dense kernels ----------- */
/* Upper Cholesky A = U'U in place (upper triangle of col-major A), blocked left-looking by columns.
 * Returns 0 or the 1-based index of the failing pivot (LAPACK convention). */
static int chol_upper(double *A, int n) {
    if (g_dpotrf) {
        int info = 0;
        g_dpotrf("U", &n, A, &n, &info);
        return info;
    }
    for (int j = 0; j < n; ++j) {
        double *cj = A + (size_t)j * n;
        /* U[0:j, j] = U[0:j,0:j]' \ A[0:j, j]  (forward substitution down column j) */
        for (int i = 0; i < j; ++i) {
            const double *ci = A + (size_t)i * n;
            double s = cj[i];
            for (int k = 0; k < i; ++k) s -= ci[k] * cj[k];
            cj[i] = s / ci[i];
        }
        double s = cj[j];
        for (int k = 0; k < j; ++k) s -= cj[k] * cj[k];
        if (!(s > 0.0)) return j + 1;
        cj[j] = sqrt(s);
    }
    return 0;
}

/* z = U' \ y  (U upper, col-major), nrhs columns */
static void solve_ut(const double *U, int n, double *Y, int nrhs) {
    if (g_dtrtrs) {
        int info = 0;
        g_dtrtrs("U", "T", "N", &n, &nrhs, U, &n, Y, &n, &info);
        return;
    }
    for (int r = 0; r < nrhs; ++r) {
        double *y = Y + (size_t)r * n;
        for (int i = 0; i < n; ++i) {
            const double *ci = U + (size_t)i * n;
            double s = y[i];
            for (int k = 0; k < i; ++k) s -= ci[k] * y[k];
            y[i] = s / ci[i];
        }
    }
}

/* x = U \ z */
static void solve_u(const double *U, int n, double *z) {
    for (int i = n - 1; i >= 0; --i) {
        double s = z[i] / U[(size_t)i * n + i];
        z[i] = s;
        const double *ci = U + (size_t)i * n;
        for (int k = 0; k < i; ++k) z[k] -= ci[k] * s;
    }
}

static int build_cov(const gpo_op *ops, int n_ops, const double *X, int n, const double *theta, double sigma2,
                     double jitter, int mode, double *K) {
    int rc = gpo_eval_program(ops, n_ops, X, n, X, n, theta, 1, mode, K);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) K[(size_t)i * n + i] += sigma2 + jitter;
    return 0;
}

/* logpdf(FiniteGP, y).  info = 0 or failing pivot; value = -inf when not positive definite. */
double gpo_lml(const gpo_op *ops, int n_ops, int n, int d, const double *X, const double *y, const double *theta,
               double sigma2, double jitter, int mode, int *info) {
    (void)d;
    double *K = (double *)malloc((size_t)n * n * sizeof(double));
    double *z = (double *)malloc((size_t)n * sizeof(double));
    double out = -INFINITY;
    int inf = -1;
    if (build_cov(ops, n_ops, X, n, theta, sigma2, jitter, mode, K) == 0) {
        inf = chol_upper(K, n);
        if (inf == 0) {
            memcpy(z, y, (size_t)n * sizeof(double));
            solve_ut(K, n, z, 1);
            double ld = 0.0, q = 0.0;
            for (int i = 0; i < n; ++i) {
                ld += log(K[(size_t)i * n + i]);
                q += z[i] * z[i];
            }
            out = -0.5 * (n * LOG2PI + 2.0 * ld + q);
        }
    }
    if (info) *info = inf;
    free(K);
    free(z);
    return out;
}

/* Batched: item b uses X + b*x_stride, Y + b*y_stride (strides in doubles; 0 = shared), Theta + b*p.
 * `threads` OpenMP threads, one independent evaluation per thread at a time. */
void gpo_lml_batched(const gpo_op *ops, int n_ops, int n, int d, const double *X, long x_stride, const double *Y,
                     long y_stride, const double *Theta, int p, const double *sigma2, long sigma2_stride,
                     double jitter, int mode, int B, int threads, double *out, int *info) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < B; ++b) {
        int inf = 0;
        out[b] = gpo_lml(ops, n_ops, n, d, X + (size_t)b * x_stride, Y + (size_t)b * y_stride,
                         Theta + (size_t)b * p, sigma2[(size_t)b * sigma2_stride], jitter, mode, &inf);
        if (info) info[b] = inf;
    }
}

/* posterior(FiniteGP, y): U (n x n upper, col-major, lower part left as garbage) and alpha. */
int gpo_posterior(const gpo_op *ops, int n_ops, int n, int d, const double *X, const double *y, const double *theta,
                  double sigma2, double jitter, int mode, double *U, double *alpha) {
    (void)d;
    if (build_cov(ops, n_ops, X, n, theta, sigma2, jitter, mode, U)) return -1;
    int inf = chol_upper(U, n);
    if (inf) return inf;
    memcpy(alpha, y, (size_t)n * sizeof(double));
    solve_ut(U, n, alpha, 1);
    solve_u(U, n, alpha);
    return 0;
}

/* mean_and_var(PosteriorGP, X*) for m test points (Xs m x d col-major). var may be NULL. */
int gpo_mean_var(const gpo_op *ops, int n_ops, int n, int d, const double *X, const double *U, const double *alpha,
                 const double *theta, int m, const double *Xs, int mode, double *mean, double *var) {
    double *Ks = (double *)malloc((size_t)n * m * sizeof(double));
    int rc = gpo_eval_program(ops, n_ops, X, n, Xs, m, theta, 0, mode, Ks);
    if (rc) { free(Ks); return rc; }
    for (int j = 0; j < m; ++j) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += Ks[(size_t)j * n + i] * alpha[i];
        mean[j] = s;
    }
    if (var) {
        solve_ut(U, n, Ks, m);
        double *row = (double *)malloc((size_t)d * sizeof(double));
        for (int j = 0; j < m; ++j) {
            double kss, q = 0.0;
            for (int c = 0; c < d; ++c) row[c] = Xs[(size_t)c * m + j];
            gpo_eval_program(ops, n_ops, row, 1, row, 1, theta, 0, mode, &kss);
            for (int i = 0; i < n; ++i) q += Ks[(size_t)j * n + i] * Ks[(size_t)j * n + i];
            var[j] = kss - q;
        }
        free(row);
    }
    free(Ks);
    return 0;
}

/* rand(FiniteGP) = U' z for S host-supplied normal vectors (Z n x S) */
int gpo_sample(const gpo_op *ops, int n_ops, int n, int d, const double *X, const double *theta, double sigma2,
               double jitter, int mode, const double *Z, int S, double *out) {
    (void)d;
    double *K = (double *)malloc((size_t)n * n * sizeof(double));
    if (build_cov(ops, n_ops, X, n, theta, sigma2, jitter, mode, K)) { free(K); return -1; }
    int inf = chol_upper(K, n);
    if (inf) { free(K); return inf; }
    for (int s = 0; s < S; ++s)
        for (int i = 0; i < n; ++i) {
            double acc = 0.0; /* (U' z)_i = sum_{k<=i} U[k,i] z_k */
            for (int k = 0; k <= i; ++k) acc += K[(size_t)i * n + k] * Z[(size_t)s * n + k];
            out[(size_t)s * n + i] = acc;
        }
    free(K);
    return 0;
}

/* Cholesky + logdet of a given SPD matrix (config C5 entry). A is overwritten by U. */
int gpo_chol_logdet(int n, double *A, double *logdet) {
    int inf = chol_upper(A, n);
    if (inf) return inf;
    double ld = 0.0;
    for (int i = 0; i < n; ++i) ld += log(A[(size_t)i * n + i]);
    *logdet = 2.0 * ld;
    return 0;
}

/* K_y itself, for element-wise checks of the covariance-construction kernel */
int gpo_cov(const gpo_op *ops, int n_ops, int n, int d, const double *X, const double *theta, double sigma2,
            double jitter, int mode, double *K) {
    (void)d;
    return build_cov(ops, n_ops, X, n, theta, sigma2, jitter, mode, K);
}
