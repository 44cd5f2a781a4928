"""CPU oracle (NumPy/SciPy) for GaPLAC's GP marginal-likelihood / posterior hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gaplac_b200/`` may import this module; it is used by
``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs as the checker and the CPU yardstick, never as the product path.

What it restates (citations are into /root/reference, or [upstream] for the un-vendored Julia
packages pinned in Manifest.toml: KernelFunctions 0.10.38, AbstractGPs 0.5.12, Distances 0.10.7,
LinearAlgebra/OpenBLAS 0.3.20):

* leaf kernels                      src/abstractgp_translations.jl:8-15, src/gp_parts.jl:11-13
* sum / product composition         src/abstractgp_translations.jl:21-35, 45-71 (i-th leaf reads its column)
* FiniteGP(gp, X, sigma2)           CLI/src/mcmc.jl:35, CLI/src/select.jl:43,47, CLI/src/sample.jl:25,
                                    src/plotting.jl:6   → N(0, K + sigma2 I)
* logpdf / posterior / mean_and_var / rand   [upstream AbstractGPs]; call sites CLI/src/select.jl:49-52,
                                    src/plotting.jl:8,12, CLI/src/sample.jl:25
* the call sequence of the reference: materialise K (one temporary per node), add the diagonal,
  ``cholesky(Symmetric(K))`` (upper factor, LAPACK dpotrf), ``U' \\ y`` (dtrtrs), ``2 sum(log(diag U))``.

Parity pin: the two legacy fixtures of SURVEY.md §8(c) (tests/golden/, 200 known lml values at n=923)
— see tests/test_oracle_golden.py.  The Julia packages themselves cannot run here (no julia binary),
so the [upstream] formulas are pinned only through those fixtures and through mpmath spot checks.

Kernel program (the representation shared with include/gaplac_b200.h): a postfix list of
``Op(kind, col, theta_slot, var_slot, value, var)``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np
import scipy.linalg as sla

SQEXP, OU, LINEAR, CAT, CONSTANT, NOISE, ADD, MUL = range(8)
KIND_NAMES = ["SQEXP", "OU", "LINEAR", "CAT", "CONSTANT", "NOISE", "ADD", "MUL"]
LOG2PI = float(np.log(2.0 * np.pi))


@dataclass
class Op:
    kind: int
    col: int = 0
    theta_slot: int = -1
    var_slot: int = -1
    value: float = 1.0
    var: float = 1.0


# ---------------------------------------------------------------------------------------------
# pairwise distances
# ---------------------------------------------------------------------------------------------
def _sqdist(a: np.ndarray, b: np.ndarray, same: bool, mode: str) -> np.ndarray:
    """Squared distance of two real columns.

    mode="direct": (a_i - b_j)^2 — what the GPU kernels compute.
    mode="gemm":   ||a||^2 + ||b||^2 - 2ab clamped at 0 with an exactly-zero diagonal — the
                   [upstream Distances 0.10.7] pairwise(SqEuclidean()) expansion used for RowVecs.
    """
    if mode == "direct":
        d = a[:, None] - b[None, :]
        return d * d
    r = np.outer(a, b)
    d2 = np.maximum(a[:, None] ** 2 + b[None, :] ** 2 - 2.0 * r, 0.0)
    if same:
        np.fill_diagonal(d2, 0.0)
    return d2


def _leaf(op: Op, Xa, Xb, theta, same, mode):
    """Value of one leaf on all pairs, and its derivative w.r.t. its own hyperparameter."""
    na, nb = Xa.shape[0], Xb.shape[0]
    h = theta[op.theta_slot] if op.theta_slot >= 0 else op.value
    if op.kind == SQEXP:      # exp(-(x-x')^2 / (2 l^2))   src/abstractgp_translations.jl:8,12
        d2 = _sqdist(Xa[:, op.col], Xb[:, op.col], same, mode)
        k = np.exp(-d2 / (2.0 * h * h))
        dk = k * d2 / (h * h * h)
    elif op.kind == OU:       # exp(-|x-x'| / l)            src/abstractgp_translations.jl:9,13
        d = np.sqrt(_sqdist(Xa[:, op.col], Xb[:, op.col], same, mode))
        k = np.exp(-d / h)
        dk = k * d / (h * h)
    elif op.kind == LINEAR:   # x x' + c                    src/abstractgp_translations.jl:10,14
        k = np.outer(Xa[:, op.col], Xb[:, op.col]) + h
        dk = np.ones((na, nb))
    elif op.kind == CAT:      # [x == x']                   src/gp_parts.jl:11-13
        k = (Xa[:, op.col][:, None] == Xb[:, op.col][None, :]).astype(np.float64)
        dk = np.zeros((na, nb))
    elif op.kind == CONSTANT:  # c for every pair           SURVEY.md §A.2
        k = np.full((na, nb), float(h))
        dk = np.ones((na, nb))
    elif op.kind == NOISE:    # delta_ij by row index, only on K(X, X)   SURVEY.md §A.2
        k = np.eye(na) if same else np.zeros((na, nb))
        dk = np.zeros((na, nb))
    else:
        raise ValueError(f"not a leaf: {op.kind}")
    return k, dk


def eval_program(ops: Sequence[Op], Xa, Xb, theta, same: bool, mode: str = "direct",
                 n_theta: int | None = None, want_grad: bool = False):
    """Evaluate the postfix kernel-program on all pairs (rows of Xa) x (rows of Xb).

    Returns K, or (K, dK) with dK[s] = dK/dtheta_s when want_grad.  One temporary per node, as
    the reference's kernelmatrix does [upstream KernelFunctions KernelSum/KernelProduct].
    """
    Xa = np.asarray(Xa, dtype=np.float64).reshape(len(Xa), -1)
    Xb = np.asarray(Xb, dtype=np.float64).reshape(len(Xb), -1)
    theta = np.asarray(theta, dtype=np.float64).ravel()
    p = len(theta) if n_theta is None else n_theta
    stack = []
    for op in ops:
        if op.kind in (ADD, MUL):
            (kb, gb), (ka, ga) = stack.pop(), stack.pop()
            if op.kind == ADD:
                k = ka + kb
                g = [x + y for x, y in zip(ga, gb)] if want_grad else None
            else:
                k = ka * kb
                g = [x * kb + ka * y for x, y in zip(ga, gb)] if want_grad else None
        else:
            k, dk = _leaf(op, Xa, Xb, theta, same, mode)
            g = None
            if want_grad:
                g = [np.zeros_like(k) for _ in range(p)]
                if op.theta_slot >= 0:
                    g[op.theta_slot] = g[op.theta_slot] + dk
        v = theta[op.var_slot] if op.var_slot >= 0 else op.var
        if want_grad:
            g = [v * x for x in g]
            if op.var_slot >= 0:
                g[op.var_slot] = g[op.var_slot] + k
        k = v * k
        stack.append((k, g))
    if len(stack) != 1:
        raise ValueError("malformed postfix program")
    K, G = stack[0]
    return (K, G) if want_grad else K


def prior_diag(ops, Xs, theta, mode="direct"):
    """diag K(X*, X*) for the latent function: Noise contributes 0."""
    Xs = np.asarray(Xs, dtype=np.float64).reshape(len(Xs), -1)
    out = np.empty(len(Xs))
    for i in range(len(Xs)):
        out[i] = eval_program(ops, Xs[i:i + 1], Xs[i:i + 1], theta, same=False, mode=mode)[0, 0]
    return out


def cov(ops, X, theta, sigma2, jitter=0.0, mode="direct"):
    """K_y = K(X,X) + (sigma2 + jitter) I   [upstream AbstractGPs cov(::FiniteGP)]; call sites a7."""
    K = eval_program(ops, X, X, theta, same=True, mode=mode)
    K[np.diag_indices_from(K)] += sigma2 + jitter
    return K


# ---------------------------------------------------------------------------------------------
# the AbstractGPs-facing calls
# ---------------------------------------------------------------------------------------------
def lml(ops, X, y, theta, sigma2, jitter=0.0, mode="direct"):
    """logpdf(FiniteGP, y) = -1/2 (n log 2pi + logdet K_y + y' K_y^-1 y)   [upstream AbstractGPs];
    call sites CLI/src/select.jl:49-50, CLI/src/mcmc.jl:35.  Upper Cholesky as Julia's cholesky()."""
    y = np.asarray(y, dtype=np.float64)
    K = cov(ops, X, theta, sigma2, jitter, mode)
    try:
        U = sla.cholesky(K, lower=False, check_finite=False)
    except sla.LinAlgError:
        return -np.inf
    z = sla.solve_triangular(U, y, trans="T", lower=False, check_finite=False)
    logdet = 2.0 * np.sum(np.log(np.diag(U)))
    return -0.5 * (len(y) * LOG2PI + logdet + float(z @ z))


def lml_grad(ops, X, y, theta, sigma2, jitter=0.0, mode="direct"):
    """(lml, dlml/dtheta, dlml/dy) with the analytic gradient of SURVEY.md §A.3:
    dlml/dtheta_s = -1/2 sum_ij (K^-1 - alpha alpha')_ij dK_ij/dtheta_s ;  dlml/dy = -alpha."""
    y = np.asarray(y, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64).ravel()
    K, G = eval_program(ops, X, X, theta, same=True, mode=mode, want_grad=True)
    K[np.diag_indices_from(K)] += sigma2 + jitter
    U = sla.cholesky(K, lower=False, check_finite=False)
    z = sla.solve_triangular(U, y, trans="T", lower=False, check_finite=False)
    alpha = sla.solve_triangular(U, z, lower=False, check_finite=False)
    Kinv = sla.cho_solve((U, False), np.eye(len(y)), check_finite=False)
    W = Kinv - np.outer(alpha, alpha)
    val = -0.5 * (len(y) * LOG2PI + 2.0 * np.sum(np.log(np.diag(U))) + float(z @ z))
    dtheta = np.array([-0.5 * np.sum(W * g) for g in G])
    return val, dtheta, -alpha


def posterior(ops, X, y, theta, sigma2, jitter=0.0, mode="direct"):
    """posterior(FiniteGP, y): C = chol(K_y) (upper U), alpha = C \\ y   [upstream AbstractGPs];
    call sites CLI/src/select.jl:51-52, src/plotting.jl:8."""
    K = cov(ops, X, theta, sigma2, jitter, mode)
    U = sla.cholesky(K, lower=False, check_finite=False)
    alpha = sla.cho_solve((U, False), np.asarray(y, dtype=np.float64), check_finite=False)
    return U, alpha


def mean_and_var(ops, X, U, alpha, Xs, theta, mode="direct"):
    """mean_and_var(PosteriorGP, X*): mean = K(X*,X) alpha ; var = diag K(X*,X*) - colsumsq(U' \\ K(X,X*))
    (latent-function variance, sigma2 NOT added back)   [upstream AbstractGPs]; call site src/plotting.jl:12."""
    Ks = eval_program(ops, X, Xs, theta, same=False, mode=mode)          # n x m
    mean = Ks.T @ alpha
    V = sla.solve_triangular(U, Ks, trans="T", lower=False, check_finite=False)
    var = prior_diag(ops, Xs, theta, mode) - np.sum(V * V, axis=0)
    return mean, var


def sample(ops, X, theta, sigma2, z, jitter=0.0, mode="direct"):
    """rand(FiniteGP) = U' z with host-supplied normals z   [upstream AbstractGPs]; call site CLI/src/sample.jl:25."""
    K = cov(ops, X, theta, sigma2, jitter, mode)
    U = sla.cholesky(K, lower=False, check_finite=False)
    return U.T @ np.asarray(z, dtype=np.float64)


def mcmc_logjoint(ops, X, Y, ell_slot_theta, fx, sigma2=0.1):
    """The log-joint of CLI/src/mcmc.jl:31-37 (constrained space, no logit Jacobian):
    log U(l;0,20) + logpdf(N(0, K_l + sigma2 I); fx) + sum_i log N(Y_i; fx_i, 1), with its gradient
    w.r.t. (theta, fx)."""
    theta = np.asarray(ell_slot_theta, dtype=np.float64).ravel()
    val, dth, dfx = lml_grad(ops, X, fx, theta, sigma2)
    Y = np.asarray(Y, dtype=np.float64)
    fx = np.asarray(fx, dtype=np.float64)
    r = Y - fx
    val = val - np.log(20.0) - 0.5 * len(Y) * LOG2PI - 0.5 * float(r @ r)
    return val, dth, dfx + r


# ---------------------------------------------------------------------------------------------
# golden fixtures: programs and closed-form legacy priors
# ---------------------------------------------------------------------------------------------
def golden_program(tag: str):
    """Kernel-programs of Appendix D for the two legacy fixtures; X columns = [PersonID, StoolPairs, nutrient]."""
    ops = [Op(CAT, col=0), Op(CAT, col=1), Op(MUL, var_slot=0),
           Op(CAT, col=0, var_slot=1), Op(ADD),
           Op(LINEAR, col=2, value=0.0, var_slot=2), Op(ADD)]
    if tag == "3206":
        ops += [Op(NOISE, var_slot=3), Op(ADD)]
    return ops


def golden_prior(tag: str, row: dict) -> float:
    """Closed-form log-prior of the deleted legacy sampler (not part of the backend)."""
    if tag == "3206":
        return sum(2 * np.log(2.0) + 2 * np.log(row[k]) - 2 * row[k] for k in ("var1", "var2", "var3", "var4"))
    s = sum(2 * np.log(2.0) + 2 * np.log(row[k]) - 2 * row[k] for k in ("var1", "var2", "var3"))
    return s + np.log(row["eta"]) - row["eta"] + 0.5 * LOG2PI


def load_golden(tag: str, golden_dir: str):
    """→ X (n x 3: PersonID, StoolPairs, nutrient), y (bug), Theta (rows x p), sigma2 (rows), lpi (rows), prior (rows)."""
    import csv
    import os
    with open(os.path.join(golden_dir, f"input_pair_{tag}.csv")) as f:
        rows = list(csv.DictReader(f))
    X = np.array([[float(r["PersonID"]), float(r["StoolPairs"]), float(r["nutrient"])] for r in rows])
    y = np.array([float(r["bug"]) for r in rows])
    with open(os.path.join(golden_dir, f"mcmc_{tag}.csv")) as f:
        chain = [{k: float(v) for k, v in r.items()} for r in csv.DictReader(f)]
    if tag == "3206":
        Theta = np.array([[c["var1"], c["var2"], c["var3"], c["var4"]] for c in chain])
        sigma2 = np.zeros(len(chain))
    else:
        Theta = np.array([[c["var1"], c["var2"], c["var3"]] for c in chain])
        sigma2 = np.array([c["eta"] ** 2 for c in chain])
    lpi = np.array([c["lpi"] for c in chain])
    prior = np.array([golden_prior(tag, c) for c in chain])
    return X, y, Theta, sigma2, lpi, prior


GOLDEN_JITTER = 1e-9
