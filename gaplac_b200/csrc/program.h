// Kernel-program: host-side compilation of the postfix gpl_op list (include/gaplac_b200.h) into the flat
// sum-of-products form the CUDA kernels evaluate per matrix entry.
//
// Replaces the object graph `kernel()` builds on every log-density call in the reference
// (src/abstractgp_translations.jl:31-35 _convert2eq fold, :45-71 column binding; rebuilt per call at
// CLI/src/mcmc.jl:33): K_ij = sum_t coef_t * prod_{f in t} leaf_f(x_i[col_f], x_j[col_f]).
#pragma once
#include <stdint.h>

#include "../../include/gaplac_b200.h"

namespace gpl {

// device factor kinds (GPL_CONSTANT with a fixed value folds into the coefficient; with a slot it
// becomes F_PARAM, as does every variance slot)
enum : int { F_SQEXP = 0, F_OU = 1, F_LINEAR = 2, F_CAT = 3, F_NOISE = 4, F_PARAM = 5 };

struct DevFactor {
    int32_t kind;
    int32_t col;
    int32_t slot;  // hyperparameter slot or -1
    int32_t pad;
    double value;  // fixed hyperparameter when slot < 0
};

struct DevProgram {
    int32_t n_terms;
    int32_t n_factors;
    int32_t n_theta;  // slots referenced
    int32_t n_cols;   // columns referenced
    int32_t term_begin[GPL_MAX_TERMS + 1];
    int32_t leaf_begin[GPL_MAX_TERMS];  // factors [term_begin, leaf_begin) of a term are F_PARAM, the rest leaves
    int32_t has_noise;  // bit t set: term t contains an F_NOISE factor (zero off the diagonal and on cross-covariances)
    double coef[GPL_MAX_TERMS];
    DevFactor f[GPL_MAX_FACTORS];
};

// Returns GPL_OK or a negative status; msg (>= 160 bytes) receives the reason.
int compile_program(const gpl_op *ops, int n_ops, DevProgram *out, char *msg);

}  // namespace gpl
