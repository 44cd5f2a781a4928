// Per-chain state machine of the batched on-device sampler (NUTS with Stan-style warm-up), written once for host and
// device: the CUDA driver (mcmc.cu) runs it with one warp per chain, tests/test_mcmc_core.py compiles the same header for
// the host and checks it transition by transition against the independent NumPy reference sampler of the test suite.
//
// Replaces `sample(m, NUTS(0.65), N)` on the model body of CLI/src/mcmc.jl:31-41 [upstream Turing 0.21.1 /
// AdvancedHMC 0.3.5]: multinomial NUTS, generalised U-turn criterion, max depth 10, divergence threshold 1000, diagonal
// metric, dual averaging to acceptance 0.65, windowed variance adaptation.  The reference runs ONE chain serially and
// gets every gradient from ForwardDiff; here B chains advance in lockstep: one batched log-density + analytic-gradient
// evaluation per leapfrog step for all chains (lml_lockstep.cu + lml_grad_lockstep.cu), then one call of
// chain_advance() per chain, which consumes the gradient, updates the tree / adaptation state and emits the next point
// to evaluate.  Chains are independent: each is at its own transition, depth and leaf.
//
// Randomness: Philox4x32-10, key = seed, counter = (chain, transition, index, purpose) - a pure function of what the
// draw is for, so host reference and device agree draw by draw whatever the execution order.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define MC_HD __host__ __device__ __forceinline__
#define MC_M __host__ __device__ __forceinline__
#else
#define MC_HD static inline
#define MC_M inline
#endif

namespace gpl {

constexpr int MC_MAX_DEPTH = 10;
constexpr int MC_MAX_P = 16;
enum : int { MC_INIT = 0, MC_FINDEPS = 1, MC_TREE = 2, MC_DONE = 3, MC_FAILED = 4 };
enum : int { MC_TAG_MOMENTUM = 0, MC_TAG_DIRECTION = 1, MC_TAG_LEAF = 2, MC_TAG_MERGE = 3, MC_TAG_FINDEPS = 4 };

struct McmcConfig {
    int n, p, dim;       // observations, hyperparameter slots, dimension of the position (p + n if latent else p)
    int latent;          // 1: the reference's model (latent fx, Y ~ N(fx, obs_sd^2)); 0: Y ~ N(0, K + sigma2 I) directly
    int max_depth;       // <= MC_MAX_DEPTH
    int n_samples, n_adapt;
    int search_eps, adapt_mass, record_warmup, record_q;
    int w_enabled, w_init_buffer, w_term_buffer, w_base;  // windowed variance adaptation (Stan)
    double delta, max_dh, obs_sd, eps0;
    unsigned long long seed;
    double lo[MC_MAX_P], hi[MC_MAX_P];  // Uniform prior bounds per slot
};

struct ChainState {
    int phase, t, depth, i, v, n_leap, fe_dir, fe_k, diverged, da_counter, w_counter, w_next, w_size, w_n, n_rec, pad;
    double eps, h0, w_tree, w_sub, sum_acc;
    double logp_cur, lp_cur, logp_prop, lp_prop, logp_sub, lp_sub, fe_h0;
    double da_mu, da_sbar, da_xbar;
};

// Stan's windowed-adaptation schedule: 75-step initial buffer, doubling windows from 25, 50-step terminal buffer; shorter
// warm-ups split 15 % / 75 % / 10 %; below 20 steps the metric is not adapted.
MC_HD void mc_setup_windows(McmcConfig &c) {
    c.w_init_buffer = 75;
    c.w_term_buffer = 50;
    c.w_base = 25;
    c.w_enabled = 1;
    if (c.n_adapt < 20) {
        c.w_enabled = c.w_init_buffer = c.w_term_buffer = c.w_base = 0;
    } else if (c.w_init_buffer + c.w_base + c.w_term_buffer > c.n_adapt) {
        c.w_init_buffer = (int)(0.15 * c.n_adapt);
        c.w_term_buffer = (int)(0.1 * c.n_adapt);
        c.w_base = c.n_adapt - (c.w_init_buffer + c.w_term_buffer);
    }
}

// vectors of one chain (each `dim` doubles), in this order, followed by r_ck[D] and rs_ck[D]
enum : int { V_QCUR = 0, V_GCUR, V_QL, V_RL, V_GL, V_QR, V_RR, V_GR, V_RHO, V_QPROP, V_GPROP, V_QSUB, V_GSUB, V_RHOSUB, V_MINV,
             V_WMEAN, V_WM2, V_NFIXED };
MC_HD long long mc_vectors_per_chain(int max_depth) { return V_NFIXED + 2 * max_depth; }

// per-transition outputs of one chain
struct ChainOut {
    double *theta;   // n_rec x p
    double *lp;      // n_rec   constrained-space log joint (Turing's `lp` column)
    double *accept;  // n_rec
    double *eps;     // n_rec   step size the transition used
    int *depth, *n_leap, *divergent;  // n_rec each
    double *q;       // n_rec x dim or NULL
};

// ---- Philox4x32-10 ------------------------------------------------------------------------------------------------------
MC_HD void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}
MC_HD double mc_u53(uint32_t hi, uint32_t lo) { return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6) + 0.5) / 9007199254740992.0; }
MC_HD double mc_uniform(unsigned long long seed, int chain, int trans, int idx, int tag) {
    uint32_t o[4];
    philox4x32((uint32_t)chain, (uint32_t)trans, (uint32_t)idx, (uint32_t)tag, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    return mc_u53(o[0], o[1]);
}
MC_HD void mc_normal_pair(unsigned long long seed, int chain, int trans, int idx, int tag, double &z0, double &z1) {
    uint32_t o[4];
    philox4x32((uint32_t)chain, (uint32_t)trans, (uint32_t)idx, (uint32_t)tag, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    const double u1 = mc_u53(o[0], o[1]), u2 = mc_u53(o[2], o[3]);
    const double r = sqrt(-2.0 * log(u1)), a = 6.283185307179586476925286766559 * u2;
    z0 = r * cos(a);
    z1 = r * sin(a);
}

// ---- small scalar helpers ---------------------------------------------------------------------------------------------------
MC_HD double mc_logaddexp(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    const double m = a > b ? a : b;
    return m + log1p(exp(-fabs(a - b)));
}
MC_HD double mc_log_sigmoid(double u) { return u > 0.0 ? -log1p(exp(-u)) : u - log1p(exp(u)); }
MC_HD void mc_ckpt_idxs(int i, int &lo, int &hi) {
    int pc = 0;
    for (int x = i >> 1; x; x &= x - 1) ++pc;
    int t = 0;
    while ((i >> t) & 1) ++t;
    hi = pc;
    lo = pc - t + 1;
}

// The state machine.  Team: how the `dim`-long vector operations of one chain are shared out (one warp on the device,
// one thread on the host); every scalar is computed identically by all members, so control flow is uniform.
template <class Team>
struct ChainMachine {
    const McmcConfig &cfg;
    ChainState s;
    double *vec;     // this chain's vectors
    int chain;
    const double *Y;     // observations (n)
    double *theta_eval;  // p   : where the next evaluation's hyperparameters go
    double *y_eval;      // n   : ... and its latent vector (latent model only)
    ChainOut out;
    int dim, p, n;

    MC_M ChainMachine(const McmcConfig &c, const ChainState &st, double *v, int ch, const double *Yobs, double *th_e, double *y_e,
                       const ChainOut &o)
        : cfg(c), s(st), vec(v), chain(ch), Y(Yobs), theta_eval(th_e), y_eval(y_e), out(o), dim(c.dim), p(c.p), n(c.n) {}

    MC_M double *V(int k) const { return vec + (long long)k * dim; }
    MC_M double *RCK(int k) const { return vec + (long long)(V_NFIXED + k) * dim; }
    MC_M double *RSCK(int k) const { return vec + (long long)(V_NFIXED + cfg.max_depth + k) * dim; }

    // ---- vector helpers -----------------------------------------------------------------------------------------------------
    MC_M void copy(double *d, const double *a) const {
        for (int k = Team::lane(); k < dim; k += Team::size()) d[k] = a[k];
    }
    MC_M double dot_minv(const double *a, const double *b) const {  // sum minv a b
        const double *mi = V(V_MINV);
        double acc = 0.0;
        for (int k = Team::lane(); k < dim; k += Team::size()) acc = fma(mi[k] * a[k], b[k], acc);
        return Team::sum(acc);
    }
    MC_M double energy(double logp, const double *r) const {
        const double h = -logp + 0.5 * dot_minv(r, r);
        return isfinite(h) ? h : INFINITY;
    }
    MC_M void momentum(double *r, int trans, int tag) const {  // r = z / sqrt(minv)
        const double *mi = V(V_MINV);
        Team::sync();  // the pairs below are dealt to the team differently from the strided loops: order against earlier accesses
        for (int m = Team::lane(); 2 * m < dim; m += Team::size()) {
            double z0, z1;
            mc_normal_pair(cfg.seed, chain, trans, m, tag, z0, z1);
            r[2 * m] = z0 / sqrt(mi[2 * m]);
            if (2 * m + 1 < dim) r[2 * m + 1] = z1 / sqrt(mi[2 * m + 1]);
        }
        Team::sync();
    }
    MC_M bool turning(const double *rl, const double *rr, const double *rho) const {
        return !(dot_minv(rl, rho) > 0.0 && dot_minv(rr, rho) > 0.0);
    }

    // ---- model: (lml, dtheta, dy) at the evaluated point q -> log target, Turing's lp, gradient g ------------------------------
    MC_M void consume(const double *q, double lml, int info, const double *dth, const double *dy, double *g, double &logp,
                       double &lp) const {
        const bool ok = info == 0 && isfinite(lml);
        double acc = 0.0, lw = 0.0;
        for (int k = Team::lane(); k < p; k += Team::size()) {
            const double u = q[k], sg = 1.0 / (1.0 + exp(-u)), w = cfg.hi[k] - cfg.lo[k];
            acc += mc_log_sigmoid(u) + mc_log_sigmoid(-u);
            lw += log(w);
            g[k] = ok ? (1.0 - 2.0 * sg) + dth[k] * w * sg * (1.0 - sg) : 0.0;
        }
        double rr = 0.0;
        if (cfg.latent) {
            const double iv = 1.0 / (cfg.obs_sd * cfg.obs_sd);
            for (int k = Team::lane(); k < n; k += Team::size()) {
                const double r = Y[k] - q[p + k];
                rr = fma(r, r, rr);
                g[p + k] = ok ? dy[k] + r * iv : 0.0;
            }
        }
        acc = Team::sum(acc);
        lw = Team::sum(lw);
        rr = Team::sum(rr);
        if (!ok) {
            logp = lp = -INFINITY;
            return;
        }
        double ll = 0.0;
        if (cfg.latent) ll = -0.5 * n * 1.8378770664093454835606594728112 - n * log(cfg.obs_sd) - 0.5 * rr / (cfg.obs_sd * cfg.obs_sd);
        logp = lml + acc + ll;
        lp = lml - lw + ll;
    }
    MC_M void emit(const double *q) const {  // the next point to evaluate
        for (int k = Team::lane(); k < p; k += Team::size())
            theta_eval[k] = cfg.lo[k] + (cfg.hi[k] - cfg.lo[k]) * (1.0 / (1.0 + exp(-q[k])));
        if (cfg.latent)
            for (int k = Team::lane(); k < n; k += Team::size()) y_eval[k] = q[p + k];
    }
    // first half of a leapfrog step from the end (q, r, g): r += e/2 g ; q += e minv r ; then the gradient at q is needed
    MC_M void half_step(double *q, double *r, const double *g, double e) const {
        const double *mi = V(V_MINV);
        for (int k = Team::lane(); k < dim; k += Team::size()) {
            const double rk = fma(0.5 * e, g[k], r[k]);
            r[k] = rk;
            q[k] = fma(e * mi[k], rk, q[k]);
        }
        emit(q);
    }
    MC_M void finish_step(double *r, const double *g, double e) const {
        for (int k = Team::lane(); k < dim; k += Team::size()) r[k] = fma(0.5 * e, g[k], r[k]);
    }

    // ---- phases ---------------------------------------------------------------------------------------------------------------
    MC_M void start_findeps() {
        copy(V(V_QL), V(V_QCUR));
        copy(V(V_GL), V(V_GCUR));
        momentum(V(V_RL), s.fe_k, MC_TAG_FINDEPS);
        s.fe_h0 = energy(s.logp_cur, V(V_RL));
        half_step(V(V_QL), V(V_RL), V(V_GL), s.eps);
        s.phase = MC_FINDEPS;
    }
    MC_M void da_init() {
        s.da_mu = log(10.0 * s.eps);
        s.da_counter = 0;
        s.da_sbar = s.da_xbar = 0.0;
    }
    MC_M void new_subtree() {
        s.v = mc_uniform(cfg.seed, chain, s.t, s.depth, MC_TAG_DIRECTION) < 0.5 ? 1 : -1;
        s.i = 0;
        s.w_sub = -INFINITY;
        if (s.v > 0) half_step(V(V_QR), V(V_RR), V(V_GR), s.eps);
        else half_step(V(V_QL), V(V_RL), V(V_GL), -s.eps);
    }
    MC_M void begin_transition() {
        momentum(V(V_RL), s.t, MC_TAG_MOMENTUM);
        copy(V(V_RR), V(V_RL));
        copy(V(V_RHO), V(V_RL));
        copy(V(V_QL), V(V_QCUR));
        copy(V(V_QR), V(V_QCUR));
        copy(V(V_GL), V(V_GCUR));
        copy(V(V_GR), V(V_GCUR));
        copy(V(V_QPROP), V(V_QCUR));
        copy(V(V_GPROP), V(V_GCUR));
        s.logp_prop = s.logp_cur;
        s.lp_prop = s.lp_cur;
        s.h0 = energy(s.logp_cur, V(V_RL));
        s.w_tree = 0.0;
        s.depth = 0;
        s.n_leap = 0;
        s.sum_acc = 0.0;
        s.diverged = 0;
        s.phase = MC_TREE;
        new_subtree();
    }
    MC_M void window_next() {
        const int last = cfg.n_adapt - cfg.w_term_buffer - 1;
        if (s.w_next == last) return;
        s.w_size *= 2;
        s.w_next = s.w_counter + s.w_size;
        if (s.w_next == last) return;
        if (s.w_next + 2 * s.w_size >= cfg.n_adapt - cfg.w_term_buffer) s.w_next = last;
    }
    MC_M bool window_learn() {  // Welford update with the new sample; true when minv was replaced
        if (!cfg.w_enabled) {
            s.w_counter++;
            return false;
        }
        const int c = s.w_counter;
        double *mean = V(V_WMEAN), *m2 = V(V_WM2), *mi = V(V_MINV);
        const double *q = V(V_QCUR);
        if (c >= cfg.w_init_buffer && c < cfg.n_adapt - cfg.w_term_buffer && c != cfg.n_adapt) {
            s.w_n++;
            for (int k = Team::lane(); k < dim; k += Team::size()) {
                const double d = q[k] - mean[k];
                mean[k] += d / s.w_n;
                m2[k] += (q[k] - mean[k]) * d;
            }
        }
        bool updated = false;
        if (c == s.w_next && c != cfg.n_adapt) {
            window_next();
            const double nn = s.w_n;
            for (int k = Team::lane(); k < dim; k += Team::size()) {
                const double var = m2[k] / (nn - 1.0);
                mi[k] = (nn / (nn + 5.0)) * var + 1e-3 * (5.0 / (nn + 5.0));
                mean[k] = 0.0;
                m2[k] = 0.0;
            }
            s.w_n = 0;
            updated = true;
        }
        s.w_counter++;
        return updated;
    }
    MC_M void end_transition() {
        copy(V(V_QCUR), V(V_QPROP));
        copy(V(V_GCUR), V(V_GPROP));
        s.logp_cur = s.logp_prop;
        s.lp_cur = s.lp_prop;
        const double stat = s.sum_acc / s.n_leap;
        if (s.t >= cfg.n_adapt || cfg.record_warmup) {
            const int r = s.n_rec;
            const double *q = V(V_QCUR);
            for (int k = Team::lane(); k < p; k += Team::size())
                out.theta[(long long)r * p + k] = cfg.lo[k] + (cfg.hi[k] - cfg.lo[k]) * (1.0 / (1.0 + exp(-q[k])));
            if (out.q)
                for (int k = Team::lane(); k < dim; k += Team::size()) out.q[(long long)r * dim + k] = q[k];
            if (Team::lane() == 0) {
                out.lp[r] = s.lp_cur;
                out.accept[r] = stat;
                out.eps[r] = s.eps;
                out.depth[r] = s.depth;
                out.n_leap[r] = s.n_leap;
                out.divergent[r] = s.diverged;
            }
            s.n_rec++;
        }
        if (s.t < cfg.n_adapt) {
            // dual averaging (Stan): learn_stepsize
            s.da_counter++;
            const double st = stat > 1.0 ? 1.0 : stat, eta = 1.0 / (s.da_counter + 10.0);
            s.da_sbar = (1.0 - eta) * s.da_sbar + eta * (cfg.delta - st);
            const double x = s.da_mu - s.da_sbar * sqrt((double)s.da_counter) / 0.05;
            const double x_eta = pow((double)s.da_counter, -0.75);
            s.da_xbar = (1.0 - x_eta) * s.da_xbar + x_eta * x;
            s.eps = exp(x);
            if (cfg.adapt_mass && window_learn()) da_init();
            if (s.t == cfg.n_adapt - 1) s.eps = exp(s.da_xbar);
        }
        Team::sync();
        s.t++;
        if (s.t >= cfg.n_adapt + cfg.n_samples) s.phase = MC_DONE;
        else begin_transition();
    }

    // One call per gradient evaluation: (lml, info, dth, dy) belong to the point emitted by the previous call.
    MC_M void advance(double lml, int info, const double *dth, const double *dy) {
        if (s.phase == MC_INIT) {
            double *mi = V(V_MINV), *mean = V(V_WMEAN), *m2 = V(V_WM2);
            for (int k = Team::lane(); k < dim; k += Team::size()) {
                mi[k] = 1.0;
                mean[k] = 0.0;
                m2[k] = 0.0;
            }
            Team::sync();
            consume(V(V_QCUR), lml, info, dth, dy, V(V_GCUR), s.logp_cur, s.lp_cur);
            if (!isfinite(s.logp_cur)) {
                s.phase = MC_FAILED;  // the initial point has zero density
                return;
            }
            s.eps = cfg.eps0;
            s.t = 0;
            s.n_rec = 0;
            s.w_counter = 0;
            s.w_size = cfg.w_base;
            s.w_next = cfg.w_init_buffer + cfg.w_base - 1;
            s.w_n = 0;
            s.fe_k = 0;
            s.fe_dir = 0;
            if (cfg.search_eps) {
                start_findeps();
            } else {
                da_init();
                begin_transition();
            }
        } else if (s.phase == MC_FINDEPS) {
            double logp1, lp1;
            consume(V(V_QL), lml, info, dth, dy, V(V_GL), logp1, lp1);
            finish_step(V(V_RL), V(V_GL), s.eps);
            const double dh = s.fe_h0 - energy(logp1, V(V_RL));
            s.fe_k++;
            if (s.fe_dir == 0) s.fe_dir = dh > -0.22314355131420976 ? 1 : -1;  // log 0.8
            const bool stop = (s.fe_dir == 1 && !(dh > -0.22314355131420976)) || (s.fe_dir == -1 && !(dh < -0.22314355131420976)) ||
                              s.fe_k >= 60;
            if (!stop) {
                s.eps = s.fe_dir == 1 ? 2.0 * s.eps : 0.5 * s.eps;
                start_findeps();
            } else {
                da_init();
                begin_transition();
            }
        } else if (s.phase == MC_TREE) {
            double *qE = s.v > 0 ? V(V_QR) : V(V_QL), *rE = s.v > 0 ? V(V_RR) : V(V_RL), *gE = s.v > 0 ? V(V_GR) : V(V_GL);
            double logp2, lp2;
            consume(qE, lml, info, dth, dy, gE, logp2, lp2);
            finish_step(rE, gE, s.v * s.eps);
            s.n_leap++;
            double dh = energy(logp2, rE) - s.h0;
            if (isnan(dh)) dh = INFINITY;
            const double a = exp(-dh);
            s.sum_acc += a < 1.0 ? a : 1.0;
            double *rho_sub = V(V_RHOSUB);
            bool take;
            if (s.i == 0) {
                s.w_sub = -dh;
                take = true;
                copy(rho_sub, rE);
            } else {
                const double w_new = mc_logaddexp(s.w_sub, -dh);
                take = mc_uniform(cfg.seed, chain, s.t, s.n_leap, MC_TAG_LEAF) < exp(-dh - w_new);
                s.w_sub = w_new;
                for (int k = Team::lane(); k < dim; k += Team::size()) rho_sub[k] += rE[k];
            }
            if (take) {
                copy(V(V_QSUB), qE);
                copy(V(V_GSUB), gE);
                s.logp_sub = logp2;
                s.lp_sub = lp2;
            }
            bool sub_div = dh > cfg.max_dh, sub_turn = false;
            if (!sub_div) {
                int lo, hi;
                mc_ckpt_idxs(s.i, lo, hi);
                if ((s.i & 1) == 0) {
                    copy(RCK(hi), rE);
                    copy(RSCK(hi), rho_sub);
                } else {
                    for (int k = hi; k >= lo && !sub_turn; --k) {
                        // subtree momentum sum rho_sub - rs_ck[k] + r_ck[k], dotted with minv r at both of its ends
                        const double *mi = V(V_MINV), *rc = RCK(k), *rs = RSCK(k);
                        double d0 = 0.0, d1 = 0.0;
                        for (int e = Team::lane(); e < dim; e += Team::size()) {
                            const double sr = rho_sub[e] - rs[e] + rc[e];
                            d0 = fma(mi[e] * rc[e], sr, d0);
                            d1 = fma(mi[e] * rE[e], sr, d1);
                        }
                        d0 = Team::sum(d0);
                        d1 = Team::sum(d1);
                        sub_turn = !(d0 > 0.0 && d1 > 0.0);
                    }
                }
            }
            if (sub_div) s.diverged = 1;
            if (sub_div || sub_turn) {
                end_transition();
                return;
            }
            if (s.i + 1 < (1 << s.depth)) {  // next leaf of the same subtree
                s.i++;
                half_step(qE, rE, gE, s.v * s.eps);
                return;
            }
            // the subtree is complete: merge it into the tree
            const double pr = exp(s.w_sub - s.w_tree);
            if (mc_uniform(cfg.seed, chain, s.t, s.depth, MC_TAG_MERGE) < (pr < 1.0 ? pr : 1.0)) {
                copy(V(V_QPROP), V(V_QSUB));
                copy(V(V_GPROP), V(V_GSUB));
                s.logp_prop = s.logp_sub;
                s.lp_prop = s.lp_sub;
            }
            s.w_tree = mc_logaddexp(s.w_tree, s.w_sub);
            double *rho = V(V_RHO);
            for (int k = Team::lane(); k < dim; k += Team::size()) rho[k] += rho_sub[k];
            s.depth++;
            const bool turn = turning(V(V_RL), V(V_RR), rho);
            if (s.depth < cfg.max_depth && !turn) new_subtree();
            else end_transition();
        }
    }
};

}  // namespace gpl
