"""Host reference sampler for the `mcmc` command's model (CLI/src/mcmc.jl:31-41).  TEST INFRASTRUCTURE ONLY.

The reference samples with Turing's `NUTS(0.65)` [upstream Turing 0.21.1 / AdvancedHMC 0.3.5, Manifest.toml:1468-1472,
42-46]: multinomial No-U-Turn sampling with the generalised U-turn criterion, maximum tree depth 10, divergence
threshold 1000, diagonal mass matrix, Stan-style windowed warm-up with dual averaging of the step size (target
acceptance 0.65), n_adapt = min(1000, N / 2).  Those packages are not vendored and Julia is absent, and no seed is fixed
anywhere in the reference: chain parity with Turing itself is UNPINNED.  What this module pins is the
device sampler (gaplac_b200/csrc/mcmc*.{h,cu}): both consume the same counter-based random stream (Philox4x32-10 keyed
by the seed, counter = (chain, transition, index, purpose)), so the device chains must reproduce these chains draw by
draw (tests/test_gpu_mcmc.py: first K transitions to 1e-8, then distributional agreement).

Model, exactly as the reference's model body (position q = (u, fx) in unconstrained space):
    theta_k = lo_k + (hi_k - lo_k) sigmoid(u_k)          l ~ Uniform(0, 20)                    mcmc.jl:32 (+ logit bijector)
    fx ~ N(0, K(theta) + sigma2 I)                        fx ~ FiniteGP(GP(k), RowVecs(X), 0.1) mcmc.jl:35
    Y_i ~ N(fx_i, obs_sd^2)                               Y .~ Normal.(fx, 1)                   mcmc.jl:36
`latent=False` drops the latent layer (Y ~ N(0, K + sigma2 I) directly, hyperparameters only: the legacy sampler behind the
golden chains).  The log-density and its gradient come from oracle.gp_oracle.lml_grad.

Algorithm: iterative NUTS (Hoffman & Gelman 2014 with multinomial sampling, Betancourt 2017; iterative tree building with
checkpointed U-turn checks, Phan et al. 2019), biased progressive sampling between subtrees, uniform progressive sampling
inside a subtree; step size found by the doubling heuristic of Stan's init_stepsize, adapted by Nesterov dual averaging
(gamma 0.05, t0 10, kappa 0.75); diagonal inverse mass matrix from Stan's windowed variance estimator (init buffer 75,
base window 25, terminal buffer 50, regularisation (n/(n+5)) var + 1e-3 (5/(n+5))).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import gp_oracle as O

M32 = 0xFFFFFFFF
LOG08 = float(np.log(0.8))

TAG_MOMENTUM, TAG_DIRECTION, TAG_LEAF, TAG_MERGE, TAG_FINDEPS = 0, 1, 2, 3, 4


# ------------------------------------------------------------------------------------------------ Philox4x32-10
def philox4x32(counter, key):
    """Philox4x32-10 (Salmon et al. 2011): counter 4 x u32, key 2 x u32 -> 4 x u32."""
    c0, c1, c2, c3 = (int(c) & M32 for c in counter)
    k0, k1 = (int(k) & M32 for k in key)
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def _u53(hi, lo):
    """Two 32-bit words -> a double in (0, 1) with 53 random bits."""
    return ((hi >> 5) * 67108864.0 + (lo >> 6) + 0.5) / 9007199254740992.0


def uniform(seed, chain, trans, idx, tag):
    o = philox4x32((chain, trans, idx, tag), (seed & M32, (seed >> 32) & M32))
    return _u53(o[0], o[1])


def normal_pair(seed, chain, trans, idx, tag):
    """Box-Muller: two standard normals from one Philox block."""
    o = philox4x32((chain, trans, idx, tag), (seed & M32, (seed >> 32) & M32))
    u1, u2 = _u53(o[0], o[1]), _u53(o[2], o[3])
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(2.0 * np.pi * u2), r * np.sin(2.0 * np.pi * u2)


def normals(seed, chain, trans, tag, dim):
    z = np.empty(dim + (dim & 1))
    for m in range((dim + 1) // 2):
        z[2 * m], z[2 * m + 1] = normal_pair(seed, chain, trans, m, tag)
    return z[:dim]


# ------------------------------------------------------------------------------------------------ model
@dataclass
class Model:
    ops: list
    X: np.ndarray
    Y: np.ndarray               # observations (n)
    lo: np.ndarray              # Uniform prior bounds per hyperparameter slot (p)
    hi: np.ndarray
    sigma2: float = 0.1
    jitter: float = 0.0
    latent: bool = True
    obs_sd: float = 1.0
    evals: int = field(default=0)

    @property
    def p(self):
        return len(self.lo)

    @property
    def dim(self):
        return self.p + (len(self.Y) if self.latent else 0)

    def theta(self, q):
        s = 1.0 / (1.0 + np.exp(-q[: self.p]))
        return self.lo + (self.hi - self.lo) * s

    def logp_grad(self, q):
        """log target in unconstrained space and its gradient; (-inf, zeros) where the covariance is not PD."""
        self.evals += 1
        p = self.p
        u = q[:p]
        s = 1.0 / (1.0 + np.exp(-u))
        th = self.lo + (self.hi - self.lo) * s
        f = q[p:] if self.latent else self.Y
        try:
            val, dth, dy = O.lml_grad(self.ops, self.X, f, th, self.sigma2, self.jitter)
        except Exception:
            return -np.inf, np.zeros_like(q), -np.inf
        if not np.isfinite(val):
            return -np.inf, np.zeros_like(q), -np.inf
        logjac = np.sum(-np.logaddexp(0.0, -u) - np.logaddexp(0.0, u))      # log s + log(1 - s); the Uniform density cancels
        g = np.empty_like(q)
        g[:p] = (1.0 - 2.0 * s) + dth * (self.hi - self.lo) * s * (1.0 - s)
        lp = val - np.sum(np.log(self.hi - self.lo))                          # constrained-space log joint (Turing's `lp`)
        if self.latent:
            r = self.Y - f
            ll = -0.5 * len(f) * O.LOG2PI - len(f) * np.log(self.obs_sd) - 0.5 * float(r @ r) / self.obs_sd ** 2
            g[p:] = dy + r / self.obs_sd ** 2
            return val + logjac + ll, g, lp + ll
        return val + logjac, g, lp


# ------------------------------------------------------------------------------------------------ adaptation (Stan)
class DualAveraging:
    def __init__(self, eps, delta):
        self.delta, self.gamma, self.t0, self.kappa = delta, 0.05, 10.0, 0.75
        self.set_mu(eps)
        self.restart()

    def set_mu(self, eps):
        self.mu = np.log(10.0 * eps)

    def restart(self):
        self.counter, self.s_bar, self.x_bar = 0, 0.0, 0.0

    def learn(self, stat):
        self.counter += 1
        stat = min(1.0, stat)
        eta = 1.0 / (self.counter + self.t0)
        self.s_bar = (1.0 - eta) * self.s_bar + eta * (self.delta - stat)
        x = self.mu - self.s_bar * np.sqrt(self.counter) / self.gamma
        x_eta = self.counter ** (-self.kappa)
        self.x_bar = (1.0 - x_eta) * self.x_bar + x_eta * x
        return float(np.exp(x))

    def complete(self):
        return float(np.exp(self.x_bar))


class WindowedVariance:
    def __init__(self, n_warmup, dim):
        self.W, self.counter = n_warmup, 0
        self.init_buffer, self.term_buffer, self.base = 75, 50, 25
        if n_warmup < 20:
            self.init_buffer = self.term_buffer = self.base = 0          # no metric adaptation
            self.enabled = False
        else:
            self.enabled = True
            if self.init_buffer + self.base + self.term_buffer > n_warmup:
                self.init_buffer = int(0.15 * n_warmup)
                self.term_buffer = int(0.1 * n_warmup)
                self.base = n_warmup - (self.init_buffer + self.term_buffer)
        self.size = self.base
        self.next = self.init_buffer + self.size - 1
        self.n, self.mean, self.m2 = 0, np.zeros(dim), np.zeros(dim)

    def _next_window(self):
        if self.next == self.W - self.term_buffer - 1:
            return
        self.size *= 2
        self.next = self.counter + self.size
        if self.next == self.W - self.term_buffer - 1:
            return
        if self.next + 2 * self.size >= self.W - self.term_buffer:
            self.next = self.W - self.term_buffer - 1

    def learn(self, q):
        """-> new inverse mass (variance) vector at the end of a window, else None."""
        if not self.enabled:
            self.counter += 1
            return None
        c = self.counter
        if c >= self.init_buffer and c < self.W - self.term_buffer and c != self.W:
            self.n += 1
            d = q - self.mean
            self.mean = self.mean + d / self.n
            self.m2 = self.m2 + (q - self.mean) * d
        if c == self.next and c != self.W:
            self._next_window()
            n = self.n
            var = self.m2 / (n - 1.0)
            var = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0))
            self.n, self.mean, self.m2 = 0, np.zeros_like(self.mean), np.zeros_like(self.m2)
            self.counter += 1
            return var
        self.counter += 1
        return None


# ------------------------------------------------------------------------------------------------ NUTS
def ckpt_idxs(i):
    idx_max = bin(i >> 1).count("1")
    t = 0
    while (i >> t) & 1:
        t += 1
    return idx_max - t + 1, idx_max


def is_turning(minv, rl, rr, rho):
    return not (float(np.dot(minv * rl, rho)) > 0.0 and float(np.dot(minv * rr, rho)) > 0.0)


def leapfrog(model, q, r, g, eps, minv):
    r = r + 0.5 * eps * g
    q = q + eps * (minv * r)
    logp, g, lp = model.logp_grad(q)
    r = r + 0.5 * eps * g
    return q, r, g, logp, lp


def energy(logp, r, minv):
    h = -logp + 0.5 * float(np.dot(minv * r, r))
    return h if np.isfinite(h) else np.inf


def find_eps(model, q, logp, g, eps, minv, seed, chain, max_tries=60):
    """Stan's init_stepsize: double / halve until the one-step acceptance ratio crosses 0.8."""
    direction, k = 0, 0
    while k < max_tries:
        r0 = normals(seed, chain, k, TAG_FINDEPS, len(q)) / np.sqrt(minv)
        h0 = energy(logp, r0, minv)
        _, r1, _, logp1, _ = leapfrog(model, q, r0, g, eps, minv)
        dh = h0 - energy(logp1, r1, minv)
        k += 1
        if direction == 0:
            direction = 1 if dh > LOG08 else -1
        if direction == 1 and not (dh > LOG08):
            break
        if direction == -1 and not (dh < LOG08):
            break
        eps = 2.0 * eps if direction == 1 else 0.5 * eps
    return eps


def nuts_transition(model, q, logp, lp, g, eps, minv, seed, chain, t, max_depth=10, max_dh=1000.0):
    dim = len(q)
    r0 = normals(seed, chain, t, TAG_MOMENTUM, dim) / np.sqrt(minv)
    h0 = energy(logp, r0, minv)
    qL, rL, gL = q.copy(), r0.copy(), g.copy()
    qR, rR, gR = q.copy(), r0.copy(), g.copy()
    rho = r0.copy()
    prop = (q, logp, lp, g)
    w_tree, depth, n_leap, sum_acc = 0.0, 0, 0, 0.0
    turning = diverging = False
    while depth < max_depth and not turning and not diverging:
        v = 1 if uniform(seed, chain, t, depth, TAG_DIRECTION) < 0.5 else -1
        zq, zr, zg = (qR, rR, gR) if v > 0 else (qL, rL, gL)
        rho_sub, w_sub, sub_prop = np.zeros(dim), -np.inf, None
        sub_turn = sub_div = False
        r_ck, rs_ck = np.zeros((max_depth, dim)), np.zeros((max_depth, dim))
        for i in range(1 << depth):
            zq, zr, zg, logp2, lp2 = leapfrog(model, zq, zr, zg, v * eps, minv)
            n_leap += 1
            dh = energy(logp2, zr, minv) - h0
            if np.isnan(dh):
                dh = np.inf
            sum_acc += min(1.0, float(np.exp(-dh)))
            if i == 0:
                w_sub, sub_prop, rho_sub = -dh, (zq, logp2, lp2, zg), zr.copy()
            else:
                w_new = np.logaddexp(w_sub, -dh)
                if uniform(seed, chain, t, n_leap, TAG_LEAF) < np.exp(-dh - w_new):
                    sub_prop = (zq, logp2, lp2, zg)
                w_sub, rho_sub = w_new, rho_sub + zr
            if dh > max_dh:
                sub_div = True
                break
            lo_i, hi_i = ckpt_idxs(i)
            if i % 2 == 0:
                r_ck[hi_i], rs_ck[hi_i] = zr, rho_sub
            else:
                for k in range(hi_i, lo_i - 1, -1):
                    if is_turning(minv, r_ck[k], zr, rho_sub - rs_ck[k] + r_ck[k]):
                        sub_turn = True
                        break
                if sub_turn:
                    break
        if v > 0:
            qR, rR, gR = zq, zr, zg
        else:
            qL, rL, gL = zq, zr, zg
        if sub_div:
            diverging = True
            break
        if sub_turn:
            turning = True
            break
        if uniform(seed, chain, t, depth, TAG_MERGE) < min(1.0, float(np.exp(w_sub - w_tree))):
            prop = sub_prop
        w_tree = float(np.logaddexp(w_tree, w_sub))
        rho = rho + rho_sub
        depth += 1
        turning = is_turning(minv, rL, rR, rho)
    return prop, sum_acc / n_leap, depth, n_leap, diverging


def sample_chain(model, q0, n_samples, n_adapt, seed, chain=0, eps0=0.1, search_eps=True, delta=0.65, adapt_mass=True,
                 max_depth=10, max_dh=1000.0):
    """One chain.  Returns a dict of per-transition records (warm-up included, in order): q, theta, lp, eps (the step size
    the transition used), accept, depth, n_leapfrog, divergent."""
    q = np.asarray(q0, dtype=np.float64).copy()
    dim = len(q)
    minv = np.ones(dim)
    logp, g, lp = model.logp_grad(q)
    if not np.isfinite(logp):
        raise ValueError("initial point has zero density")
    eps = find_eps(model, q, logp, g, eps0, minv, seed, chain) if search_eps else eps0
    da = DualAveraging(eps, delta)
    wv = WindowedVariance(n_adapt, dim) if adapt_mass else None
    T = n_adapt + n_samples
    rec = dict(q=np.empty((T, dim)), theta=np.empty((T, model.p)), lp=np.empty(T), eps=np.empty(T), accept=np.empty(T),
               depth=np.empty(T, dtype=int), n_leapfrog=np.empty(T, dtype=int), divergent=np.zeros(T, dtype=bool), eps_init=eps)
    for t in range(T):
        (q, logp, lp, g), stat, depth, n_leap, div = nuts_transition(model, q, logp, lp, g, eps, minv, seed, chain, t,
                                                                     max_depth, max_dh)
        rec["q"][t], rec["theta"][t], rec["lp"][t], rec["eps"][t] = q, model.theta(q), lp, eps
        rec["accept"][t], rec["depth"][t], rec["n_leapfrog"][t], rec["divergent"][t] = stat, depth, n_leap, div
        if t < n_adapt:
            eps = da.learn(stat)
            if wv is not None:
                var = wv.learn(q)
                if var is not None:
                    minv = var
                    da.set_mu(eps)
                    da.restart()
            if t == n_adapt - 1:
                eps = da.complete()
    return rec
