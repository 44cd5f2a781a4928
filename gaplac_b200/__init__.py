"""gaplac_b200 — B200 (sm_100a) backend for GaPLAC's GP marginal-likelihood / posterior hot path.

The product is libgaplac_b200.so (C ABI in include/gaplac_b200.h, CUDA kernels in gaplac_b200/csrc/).
This package is the host-side mirror of the reference's interface for that path: the formula language
(formula.py), the AbstractGPs-facing calls (gp.py), and the ctypes binding that stands in for Julia's
`ccall` (_lib.py).  There is no CPU fallback.
"""
from .formula import (Cat, Constant, GPComponent, GPOperation, Gaussian, KernelProgram, Linear, Noise, Op, OU, Slot,
                      Spec, SqExp, formula, gp_spec, kernel, likelihood, make_gp, parse_formula, response, varnames)
from .gp import GP, FiniteGP, PosteriorGP, default_context, logpdf, logpdf_batched, mean, mean_and_var, posterior, rand
from ._lib import Context, GaplacError, MultiContext, PosDefException, Program
from . import mcmc
from .chain import (log2_harmmean_exp2, log_bayes_factor, log_evidence_harmonic, mixture_summary, predict_chain,
                    select_chains_log2_bayes, select_formulae_log2_bayes)

__all__ = [
    "Cat", "Constant", "GPComponent", "GPOperation", "Gaussian", "KernelProgram", "Linear", "Noise", "Op", "OU",
    "Slot", "Spec", "SqExp", "formula", "gp_spec", "kernel", "likelihood", "make_gp", "parse_formula", "response",
    "varnames", "GP", "FiniteGP", "PosteriorGP", "default_context", "logpdf", "logpdf_batched", "mean",
    "mean_and_var", "posterior", "rand", "Context", "GaplacError", "MultiContext", "PosDefException", "Program", "mcmc",
    "predict_chain", "mixture_summary", "log_evidence_harmonic", "log_bayes_factor", "log2_harmmean_exp2",
    "select_chains_log2_bayes", "select_formulae_log2_bayes",
]
