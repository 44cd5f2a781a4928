// One host call, several GPUs: independent items of a batch (hyperparameter proposals, per-feature models, chains) are
// split into contiguous blocks, one per device, evaluated concurrently on that device's own context, and every device
// writes its slice straight into the caller's host buffers (SURVEY.md 8(b) "context = device(s)", 8(e) "one ccall").
// There is no data-path collective: the only exchange is the per-item results coming home.  The reference has no
// counterpart (one Julia task, one chain: CLI/src/mcmc.jl:41); a multi-process host would use one context per rank and
// NCCL for the gather instead (gaplac_b200/shard.py).  Built on the public single-device entry points only.
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gaplac_b200.h"

struct gpl_multi {
    std::vector<gpl_ctx *> ctx;
    std::vector<int> device;
    std::string err;
};

namespace {
thread_local std::string g_multi_error;

int mfail(gpl_multi *m, int code, const std::string &msg) {
    g_multi_error = msg;
    if (m) m->err = msg;
    return code;
}

// contiguous block [lo, hi) of part r of R; the first B % R parts get one extra item (same rule as gaplac_b200/shard.py)
void block_of(int B, int r, int R, int &lo, int &hi) {
    const int q = B / R, rem = B % R;
    lo = r * q + (r < rem ? r : rem);
    hi = lo + q + (r < rem ? 1 : 0);
}

template <class F>
int run_parts(gpl_multi *m, int B, F part) {
    const int R = (int)m->ctx.size();
    std::vector<int> rc(R, GPL_OK);
    std::vector<std::thread> th;
    for (int r = 0; r < R; ++r) {
        int lo, hi;
        block_of(B, r, R, lo, hi);
        if (hi <= lo) continue;
        th.emplace_back([&, r, lo, hi]() { rc[r] = part(r, lo, hi); });
    }
    for (auto &t : th) t.join();
    for (int r = 0; r < R; ++r)
        if (rc[r] != GPL_OK) {
            char buf[640];
            snprintf(buf, sizeof(buf), "device %d (part %d of %d): %s", m->device[r], r, R, gpl_last_error(m->ctx[r]));
            return mfail(m, rc[r], buf);
        }
    return GPL_OK;
}
}  // namespace

extern "C" {

int gpl_multi_init(const int *devices, int n_devices, gpl_multi **out) {
    if (!out) return mfail(nullptr, GPL_ERR_ARG, "gpl_multi_init: out is NULL");
    *out = nullptr;
    if (n_devices <= 0) return mfail(nullptr, GPL_ERR_ARG, "gpl_multi_init: n_devices must be positive");
    gpl_multi *m = new (std::nothrow) gpl_multi();
    if (!m) return mfail(nullptr, GPL_ERR_ARG, "gpl_multi_init: out of host memory");
    for (int r = 0; r < n_devices; ++r) {
        const int dev = devices ? devices[r] : r;
        gpl_ctx *c = nullptr;
        const int rc = gpl_init(dev, &c);
        if (rc != GPL_OK) {
            const std::string msg = std::string("gpl_multi_init: ") + gpl_last_error(nullptr);
            for (gpl_ctx *x : m->ctx) gpl_destroy(x);
            delete m;
            return mfail(nullptr, rc, msg);
        }
        m->ctx.push_back(c);
        m->device.push_back(dev);
    }
    *out = m;
    return GPL_OK;
}

int gpl_multi_destroy(gpl_multi *m) {
    if (!m) return GPL_OK;
    for (gpl_ctx *c : m->ctx) gpl_destroy(c);
    delete m;
    return GPL_OK;
}

int gpl_multi_device_count(const gpl_multi *m) { return m ? (int)m->ctx.size() : GPL_ERR_ARG; }

gpl_ctx *gpl_multi_context(gpl_multi *m, int part) {
    return (m && part >= 0 && part < (int)m->ctx.size()) ? m->ctx[part] : nullptr;
}

const char *gpl_multi_last_error(gpl_multi *m) { return m ? m->err.c_str() : g_multi_error.c_str(); }

int gpl_multi_lml_batched(gpl_multi *m, const gpl_prog *prog, int n, int d, const double *X, int x_batched, const double *Y,
                          int y_batched, const double *Theta, int p, const double *sigma2, int sigma2_batched, double jitter,
                          int B, double *lml, double *dtheta, double *dy, int *info) {
    if (!m || !prog) return mfail(m, GPL_ERR_ARG, "gpl_multi_lml_batched: null handle");
    if (B <= 0 || n <= 0 || d <= 0 || !lml) return mfail(m, GPL_ERR_ARG, "gpl_multi_lml_batched: bad argument");
    return run_parts(m, B, [&](int r, int lo, int hi) {
        const size_t o = (size_t)lo;
        return gpl_lml_batched(m->ctx[r], prog, n, d, x_batched ? X + o * n * d : X, x_batched, y_batched ? Y + o * n : Y, y_batched,
                               Theta ? Theta + o * p : nullptr, p, sigma2_batched ? sigma2 + o : sigma2, sigma2_batched, jitter,
                               hi - lo, lml + o, dtheta ? dtheta + o * p : nullptr, dy ? dy + o * n : nullptr,
                               info ? info + o : nullptr);
    });
}

int gpl_multi_mcmc_nuts(gpl_multi *m, const gpl_prog *prog, int n, int d, const double *X, int x_batched, const double *Y,
                        int y_batched, int p, const double *lo_b, const double *hi_b, const double *sigma2, int sigma2_batched,
                        double jitter, int B, const double *q0, const gpl_mcmc_opts *opts, double *theta, double *lp, double *q,
                        double *accept, double *eps, int *depth, int *n_leapfrog, int *divergent, int *status,
                        long long *n_grad_evals) {
    if (!m || !prog || !opts) return mfail(m, GPL_ERR_ARG, "gpl_multi_mcmc_nuts: null handle");
    if (B <= 0 || n <= 0) return mfail(m, GPL_ERR_ARG, "gpl_multi_mcmc_nuts: bad argument");
    const int n_adapt = opts->n_adapt >= 0 ? opts->n_adapt : (opts->n_samples / 2 < 1000 ? opts->n_samples / 2 : 1000);
    const size_t n_rec = (size_t)opts->n_samples + (opts->record_warmup ? n_adapt : 0), dim = (size_t)p + (opts->latent ? n : 0);
    std::vector<long long> evals(m->ctx.size(), 0);
    const int rc = run_parts(m, B, [&](int r, int lo, int hi) {
        const size_t o = (size_t)lo;
        gpl_mcmc_opts part = *opts;
        part.chain_offset = opts->chain_offset + lo;  // every chain keeps its own random stream wherever it runs
        return gpl_mcmc_nuts(m->ctx[r], prog, n, d, x_batched ? X + o * n * d : X, x_batched, y_batched ? Y + o * n : Y, y_batched, p,
                             lo_b, hi_b, sigma2_batched ? sigma2 + o : sigma2, sigma2_batched, jitter, hi - lo, q0 + o * dim, &part,
                             theta + o * n_rec * p, lp + o * n_rec, q ? q + o * n_rec * dim : nullptr,
                             accept ? accept + o * n_rec : nullptr, eps ? eps + o * n_rec : nullptr,
                             depth ? depth + o * n_rec : nullptr, n_leapfrog ? n_leapfrog + o * n_rec : nullptr,
                             divergent ? divergent + o * n_rec : nullptr, status ? status + o : nullptr, &evals[r]);
    });
    if (n_grad_evals) {
        *n_grad_evals = 0;
        for (long long e : evals) *n_grad_evals += e;
    }
    return rc;
}

}  // extern "C"
