// Host build of the sampler's per-chain state machine (gaplac_b200/csrc/mcmc_core.h) for the CPU test suite:
// the SAME header the CUDA driver runs with one warp per chain, here with a one-thread team and the gradient supplied
// by a callback (the test passes the CPU oracle).  Test harness only; not part of libgaplac_b200.so.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../gaplac_b200/csrc/mcmc_core.h"

using namespace gpl;

struct TeamSerial {
    static int lane() { return 0; }
    static int size() { return 1; }
    static double sum(double v) { return v; }
    static void sync() {}
};

// grad(theta[p], y[n], &lml, &info, dth[p], dy[n])
typedef void (*grad_fn)(const double *theta, const double *y, double *lml, int *info, double *dth, double *dy);

extern "C" int mc_host_config_size() { return (int)sizeof(McmcConfig); }
extern "C" void mc_host_setup_windows(McmcConfig *c) { mc_setup_windows(*c); }

extern "C" int mc_host_run(const McmcConfig *cfg, int chain, const double *Y, const double *q0, grad_fn grad, double *theta_out,
                           double *lp, double *accept, double *eps, int *depth, int *n_leap, int *divergent, double *q_out,
                           long long *n_evals) {
    const int dim = cfg->dim, p = cfg->p, n = cfg->n;
    std::vector<double> vec((size_t)mc_vectors_per_chain(cfg->max_depth) * dim, 0.0), th(p > 0 ? p : 1), ye(n), dth(p > 0 ? p : 1), dy(n);
    std::memcpy(ye.data(), Y, sizeof(double) * n);
    ChainState st;
    std::memset(&st, 0, sizeof(st));
    st.phase = MC_INIT;
    ChainOut out{theta_out, lp, accept, eps, depth, n_leap, divergent, q_out};
    std::memcpy(vec.data() + (size_t)V_QCUR * dim, q0, sizeof(double) * dim);
    ChainMachine<TeamSerial> m(*cfg, st, vec.data(), chain, Y, th.data(), ye.data(), out);
    m.emit(m.V(V_QCUR));
    long long evals = 0;
    while (m.s.phase != MC_DONE && m.s.phase != MC_FAILED) {
        double lml = 0.0;
        int info = 0;
        grad(th.data(), ye.data(), &lml, &info, dth.data(), dy.data());
        ++evals;
        m.advance(lml, info, dth.data(), dy.data());
    }
    if (n_evals) *n_evals = evals;
    return m.s.phase == MC_DONE ? m.s.n_rec : -1;
}
