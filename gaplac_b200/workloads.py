"""Synthetic workloads of BASELINE.json's configs, shared by bench.py and the tests.

Each generator is deterministic (NumPy default_rng(seed)) and returns plain arrays plus the postfix
kernel-program of SURVEY.md Appendix D.  Drawing y from the model needs a Cholesky on the host; that is
workload generation (NumPy), not the product path.
"""
from __future__ import annotations

import numpy as np

from .formula import Op
from ._lib import ADD, CAT, LINEAR, MUL, NOISE, OU, SQEXP


def prog_c1():
    """C1  y ~| SqExp(:x)   theta = (l); sigma2 = 0.1 fixed (CLI/src/mcmc.jl:35)."""
    return [Op(SQEXP, col=0, theta_slot=0)]


def prog_c2():
    """C2  SqExp(:x) + OU(:x) + Noise   theta = (l_se, l_ou, s2_noise)."""
    return [Op(SQEXP, col=0, theta_slot=0), Op(OU, col=0, theta_slot=1), Op(ADD), Op(NOISE, var_slot=2), Op(ADD)]


def prog_c3():
    """C3  Cat(:subject) * SqExp(:time) + Noise   theta = (l, s2_noise)."""
    return [Op(CAT, col=0), Op(SQEXP, col=1, theta_slot=0), Op(MUL), Op(NOISE, var_slot=1), Op(ADD)]


def prog_c4():
    """C4  SqExp(:x) + Linear(:z) + Noise   theta = (l, c, s2_noise)."""
    return [Op(SQEXP, col=0, theta_slot=0), Op(LINEAR, col=1, theta_slot=1), Op(ADD), Op(NOISE, var_slot=2), Op(ADD)]


def prog_c5():
    """C5  SqExp(:x) + Noise   theta = (l, s2_noise)."""
    return [Op(SQEXP, col=0, theta_slot=0), Op(NOISE, var_slot=1), Op(ADD)]


def _k_sqexp(x, l):
    d = x[:, None] - x[None, :]
    return np.exp(-d * d / (2 * l * l))


def _k_ou(x, l):
    return np.exp(-np.abs(x[:, None] - x[None, :]) / l)


def make_c1(seed: int = 1, n: int = 50):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-5, 5, n)
    K = _k_sqexp(x, 1.5) + 0.1 * np.eye(n)
    y = np.linalg.cholesky(K) @ rng.standard_normal(n)
    return dict(X=x.reshape(-1, 1), y=y, sigma2=0.1, ops=prog_c1())


def make_c2(seed: int = 2, n: int = 512, B: int = 4096):
    """Shared X, y; B proposals theta_b = (l_se, l_ou, s2) with l ~ U(0.2, 5), s2 ~ U(0.05, 0.5)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-5, 5, n)
    K = _k_sqexp(x, 1.5) + _k_ou(x, 1.0) + 0.1 * np.eye(n)
    y = np.linalg.cholesky(K) @ rng.standard_normal(n)
    Theta = np.column_stack([rng.uniform(0.2, 5, B), rng.uniform(0.2, 5, B), rng.uniform(0.05, 0.5, B)])
    return dict(X=x.reshape(-1, 1), y=y, Theta=Theta, sigma2=0.0, ops=prog_c2())


def make_c3(seed: int = 3, subjects: int = 30, per_subject: int = 10, features: int = 2000):
    """30 subjects x 10 time points; per-feature responses drawn from Cat*SqExp(l=30)+Noise(0.1), then
    rank-based inverse-normal transformed (src/utils.jl:16-28)."""
    from scipy.special import ndtri
    rng = np.random.default_rng(seed)
    n = subjects * per_subject
    subj = np.repeat(np.arange(1, subjects + 1), per_subject).astype(np.float64)
    time = np.concatenate([np.sort(rng.uniform(0, 365, per_subject)) for _ in range(subjects)])
    K = (subj[:, None] == subj[None, :]) * _k_sqexp(time, 30.0) + 0.1 * np.eye(n)
    L = np.linalg.cholesky(K)
    Y = (L @ rng.standard_normal((n, features))).T          # (features, n)
    ranks = np.argsort(np.argsort(Y, axis=1), axis=1) + 1.0
    Y = ndtri((ranks - 0.5) / n)
    Theta = np.column_stack([rng.uniform(5, 80, features), rng.uniform(0.05, 0.5, features)])
    return dict(X=np.column_stack([subj, time]), Y=Y, Theta=Theta, sigma2=0.0, ops=prog_c3())


def make_c4(seed: int = 4, n: int = 2048, m: int = 20000):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-5, 5, n)
    z = rng.uniform(-3, 3, n)
    theta = np.array([1.0, 0.5, 0.1])
    K = _k_sqexp(x, theta[0]) + np.outer(z, z) + theta[1] + theta[2] * np.eye(n)
    y = np.linalg.cholesky(K) @ rng.standard_normal(n)
    g = int(np.ceil(np.sqrt(m)))
    gx, gz = np.meshgrid(np.linspace(-5, 5, g), np.linspace(-3, 3, g), indexing="ij")
    Xs = np.column_stack([gx.ravel(), gz.ravel()])[:m]
    return dict(X=np.column_stack([x, z]), y=y, theta=theta, sigma2=0.0, Xs=Xs, ops=prog_c4())


def make_c5(seed: int = 5, n: int = 8192):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-50, 50, n)
    y = rng.standard_normal(n)
    return dict(X=x.reshape(-1, 1), y=y, theta=np.array([1.0, 0.1]), sigma2=0.0, ops=prog_c5())


# ---- the reference's legacy fixtures (tests/golden/, extracted by tools/make_golden.py from test/testin/*.tsv) ----------
GOLDEN_JITTER = 1e-9


def prog_golden(tag: str = "3206"):
    """Cat(PersonID)*Cat(StoolPairs) + Cat(PersonID) + Linear(nutrient) [+ Noise], one variance slot per term
    (generating command test/pred.jl:22).  X columns = [PersonID, StoolPairs, nutrient]."""
    ops = [Op(CAT, col=0), Op(CAT, col=1), Op(MUL, var_slot=0), Op(CAT, col=0, var_slot=1), Op(ADD),
           Op(LINEAR, col=2, value=0.0, var_slot=2), Op(ADD)]
    if tag == "3206":
        ops += [Op(NOISE, var_slot=3), Op(ADD)]
    return ops


def make_golden(golden_dir: str, tag: str = "3206"):
    """X (n x 3), y (bug), Theta (chain rows x p), sigma2 (rows), and the fixture's own answers: lml = l_pi - legacy prior."""
    import csv
    import os
    with open(os.path.join(golden_dir, f"input_pair_{tag}.csv")) as f:
        rows = list(csv.DictReader(f))
    X = np.array([[float(r["PersonID"]), float(r["StoolPairs"]), float(r["nutrient"])] for r in rows])
    y = np.array([float(r["bug"]) for r in rows])
    with open(os.path.join(golden_dir, f"mcmc_{tag}.csv")) as f:
        chain = [{k: float(v) for k, v in r.items()} for r in csv.DictReader(f)]
    names = ("var1", "var2", "var3", "var4") if tag == "3206" else ("var1", "var2", "var3")
    Theta = np.array([[c[k] for k in names] for c in chain])
    sigma2 = np.zeros(len(chain)) if tag == "3206" else np.array([c["eta"] ** 2 for c in chain])
    prior = np.array([sum(2 * np.log(2.0) + 2 * np.log(c[k]) - 2 * c[k] for k in names) for c in chain])
    if tag != "3206":
        prior = prior + np.array([np.log(c["eta"]) - c["eta"] + 0.5 * np.log(2 * np.pi) for c in chain])
    lml_known = np.array([c["lpi"] for c in chain]) - prior
    return dict(X=X, y=y, Theta=Theta, sigma2=sigma2, jitter=GOLDEN_JITTER, lml_known=lml_known, ops=prog_golden(tag))
