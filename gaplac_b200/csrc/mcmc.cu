// Batched on-device sampler: kernels around the shared per-chain state machine (mcmc_core.h).
//
// Replaces the serial loop `sample(m, NUTS(0.65), N)` of CLI/src/mcmc.jl:39-41 [upstream Turing / AdvancedHMC] for B
// independent chains at once.  One warp per chain runs ChainMachine::advance after every batched
// log-density + gradient evaluation (lml_lockstep.cu + lml_grad_lockstep.cu): it finishes the leapfrog step with the new
// gradient, updates the NUTS tree (multinomial proposal, checkpointed U-turn checks), at the end of a transition records
// the draw and adapts step size / diagonal metric, and writes the next point to evaluate (hyperparameters -> theta_eval,
// latent vector -> y_eval) for the next batched evaluation.  Proposals, accept/reject decisions and adaptation never
// leave the device; the host replays a captured graph of (evaluation, advance) and polls `done`.
#include "kernels.h"
#include "mcmc_core.h"

namespace gpl {

namespace {
struct TeamWarp {
    __device__ __forceinline__ static int lane() { return threadIdx.x & 31; }
    __device__ __forceinline__ static int size() { return 32; }
    __device__ __forceinline__ static double sum(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ static void sync() { __syncwarp(); }
};

__device__ __forceinline__ ChainOut chain_out(const McmcDevParams &prm, int b) {
    const long long r = (long long)b * prm.n_rec;
    ChainOut o;
    o.theta = prm.theta_out + r * prm.cfg.p;
    o.lp = prm.lp_out + r;
    o.accept = prm.accept_out + r;
    o.eps = prm.eps_out + r;
    o.depth = prm.depth_out + r;
    o.n_leap = prm.nleap_out + r;
    o.divergent = prm.div_out + r;
    o.q = prm.q_out ? prm.q_out + r * prm.cfg.dim : nullptr;
    return o;
}
}  // namespace

// q0 -> the chain's current position; emits the first evaluation point
__global__ void __launch_bounds__(128) mcmc_init_kernel(const __grid_constant__ McmcDevParams prm) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= prm.B) return;
    const int dim = prm.cfg.dim;
    double *vec = prm.vec + (long long)b * prm.vec_stride;
    for (int k = TeamWarp::lane(); k < dim; k += 32) vec[(long long)V_QCUR * dim + k] = prm.q0[(long long)b * dim + k];
    __syncwarp();
    ChainState st;
    memset(&st, 0, sizeof(st));
    st.phase = MC_INIT;
    ChainMachine<TeamWarp> m(prm.cfg, st, vec, b + prm.chain_offset, prm.Y + (long long)b * prm.y_stride,
                             prm.theta_eval + (long long)b * prm.cfg.p, prm.y_eval + (long long)b * prm.cfg.n, chain_out(prm, b));
    m.emit(m.V(V_QCUR));
    if (TeamWarp::lane() == 0) {
        prm.state[b] = st;
        prm.slot_chain[b] = b;
    }
}

__global__ void __launch_bounds__(128) mcmc_advance_kernel(const __grid_constant__ McmcDevParams prm) {
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= prm.n_slots) return;
    const int b = prm.slot_chain[s];
    const ChainState st = prm.state[b];
    if (st.phase == MC_DONE || st.phase == MC_FAILED) return;
    ChainMachine<TeamWarp> m(prm.cfg, st, prm.vec + (long long)b * prm.vec_stride, b + prm.chain_offset,
                             prm.Y + (long long)b * prm.y_stride, prm.theta_eval + (long long)s * prm.cfg.p,
                             prm.y_eval + (long long)s * prm.cfg.n, chain_out(prm, b));
    m.advance(prm.lml[s], prm.info[s], prm.dtheta + (long long)s * prm.cfg.p, prm.dy + (long long)s * prm.cfg.n);
    __syncwarp();
    if (TeamWarp::lane() == 0) {
        prm.state[b] = m.s;
        if (m.s.phase == MC_DONE || m.s.phase == MC_FAILED) atomicAdd(prm.done, 1u);
    }
}

// Active chains, in chain order, into the first slots (one CTA; the order is fixed, and a chain's results do not depend on
// its slot: the batched evaluation is position-independent bit for bit).
__global__ void __launch_bounds__(256) mcmc_compact_kernel(const __grid_constant__ McmcDevParams prm) {
    __shared__ int base, cnt[256];
    const int tid = threadIdx.x;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int b0 = 0; b0 < prm.B; b0 += 256) {
        const int b = b0 + tid;
        const int act = b < prm.B && prm.state[b].phase != MC_DONE && prm.state[b].phase != MC_FAILED;
        cnt[tid] = act;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {  // inclusive scan
            const int v = tid >= o ? cnt[tid - o] : 0;
            __syncthreads();
            cnt[tid] += v;
            __syncthreads();
        }
        if (act) prm.slot_chain[base + cnt[tid] - 1] = b;
        __syncthreads();
        if (tid == 255) base += cnt[255];
        __syncthreads();
    }
    if (tid == 0) *prm.n_active = base;
}

// after a compaction every active chain writes the point it is waiting on into its new slot
__global__ void __launch_bounds__(128) mcmc_reemit_kernel(const __grid_constant__ McmcDevParams prm) {
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= prm.n_slots) return;
    const int b = prm.slot_chain[s];
    const ChainState st = prm.state[b];
    ChainMachine<TeamWarp> m(prm.cfg, st, prm.vec + (long long)b * prm.vec_stride, b + prm.chain_offset,
                             prm.Y + (long long)b * prm.y_stride, prm.theta_eval + (long long)s * prm.cfg.p,
                             prm.y_eval + (long long)s * prm.cfg.n, chain_out(prm, b));
    const int which = st.phase == MC_INIT ? V_QCUR : (st.phase == MC_FINDEPS ? V_QL : (st.v > 0 ? V_QR : V_QL));
    m.emit(m.V(which));
}

__global__ void mcmc_gather_kernel(const double *src, double *dst, const int *slot_chain, int n_slots, long long width) {
    const long long total = (long long)n_slots * width;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long s = e / width, k = e - s * width;
        dst[e] = src[(long long)slot_chain[s] * width + k];
    }
}

// status[b] = 0 (chain complete) / 1 (initial point has zero density) / 2 (not finished: iteration cap)
__global__ void mcmc_status_kernel(const ChainState *state, int B, int *status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) status[b] = state[b].phase == MC_DONE ? 0 : (state[b].phase == MC_FAILED ? 1 : 2);
}

size_t mcmc_state_bytes() { return sizeof(ChainState); }

}  // namespace gpl
