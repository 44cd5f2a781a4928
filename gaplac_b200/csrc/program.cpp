// Host-side compilation of the postfix kernel-program into a sum of products (see program.h).
#include "program.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace gpl {
namespace {

struct Term {
    double coef = 1.0;
    std::vector<DevFactor> fs;
};
using Poly = std::vector<Term>;

void scale(Poly &p, const gpl_op &op) {
    // variance multiplier of a node: a slot becomes a PARAM factor, a fixed value folds into coef
    for (Term &t : p) {
        if (op.var_slot >= 0) {
            DevFactor f{};
            f.kind = F_PARAM;
            f.col = 0;
            f.slot = op.var_slot;
            f.value = 0.0;
            t.fs.push_back(f);
        } else {
            t.coef *= op.var;
        }
    }
}

}  // namespace

int compile_program(const gpl_op *ops, int n_ops, DevProgram *out, char *msg) {
    msg[0] = 0;
    if (!ops || n_ops <= 0 || n_ops > GPL_MAX_OPS) {
        snprintf(msg, 160, "kernel-program: n_ops=%d outside 1..%d", n_ops, GPL_MAX_OPS);
        return GPL_ERR_ARG;
    }
    std::vector<Poly> stack;
    int max_slot = -1, max_col = -1;
    for (int i = 0; i < n_ops; ++i) {
        const gpl_op &op = ops[i];
        if (op.theta_slot >= GPL_MAX_THETA || op.var_slot >= GPL_MAX_THETA) {
            snprintf(msg, 160, "kernel-program: op %d uses a slot >= %d", i, GPL_MAX_THETA);
            return GPL_ERR_LIMIT;
        }
        if (op.var_slot > max_slot) max_slot = op.var_slot;
        Poly cur;
        switch (op.kind) {
        case GPL_ADD:
        case GPL_MUL: {
            if (stack.size() < 2) {
                snprintf(msg, 160, "kernel-program: op %d (%s) needs two operands", i, op.kind == GPL_ADD ? "ADD" : "MUL");
                return GPL_ERR_ARG;
            }
            Poly rhs = std::move(stack.back());
            stack.pop_back();
            Poly lhs = std::move(stack.back());
            stack.pop_back();
            if (op.kind == GPL_ADD) {
                cur = std::move(lhs);
                cur.insert(cur.end(), rhs.begin(), rhs.end());
            } else {
                for (const Term &a : lhs)
                    for (const Term &b : rhs) {
                        Term t;
                        t.coef = a.coef * b.coef;
                        t.fs = a.fs;
                        t.fs.insert(t.fs.end(), b.fs.begin(), b.fs.end());
                        cur.push_back(std::move(t));
                    }
            }
            break;
        }
        case GPL_SQEXP:
        case GPL_OU:
        case GPL_LINEAR:
        case GPL_CAT: {
            if (op.col < 0 || op.col >= GPL_MAX_COLS) {
                snprintf(msg, 160, "kernel-program: op %d column %d outside 0..%d", i, op.col, GPL_MAX_COLS - 1);
                return GPL_ERR_LIMIT;
            }
            if (op.col > max_col) max_col = op.col;
            Term t;
            DevFactor f{};
            f.kind = op.kind == GPL_SQEXP ? F_SQEXP : op.kind == GPL_OU ? F_OU : op.kind == GPL_LINEAR ? F_LINEAR : F_CAT;
            f.col = op.col;
            f.slot = op.kind == GPL_CAT ? -1 : op.theta_slot;
            f.value = op.value;
            if (f.slot > max_slot) max_slot = f.slot;
            if ((op.kind == GPL_SQEXP || op.kind == GPL_OU) && f.slot < 0 && !(op.value > 0.0)) {
                snprintf(msg, 160, "kernel-program: op %d lengthscale must be > 0", i);
                return GPL_ERR_ARG;
            }
            t.fs.push_back(f);
            cur.push_back(std::move(t));
            break;
        }
        case GPL_CONSTANT: {
            Term t;
            if (op.theta_slot >= 0) {
                DevFactor f{};
                f.kind = F_PARAM;
                f.slot = op.theta_slot;
                if (f.slot > max_slot) max_slot = f.slot;
                t.fs.push_back(f);
            } else {
                t.coef = op.value;
            }
            cur.push_back(std::move(t));
            break;
        }
        case GPL_NOISE: {
            Term t;
            DevFactor f{};
            f.kind = F_NOISE;
            f.slot = -1;
            t.fs.push_back(f);
            cur.push_back(std::move(t));
            break;
        }
        default:
            snprintf(msg, 160, "kernel-program: op %d has unknown kind %d", i, op.kind);
            return GPL_ERR_ARG;
        }
        scale(cur, op);
        if ((int)cur.size() > GPL_MAX_TERMS) {
            snprintf(msg, 160, "kernel-program: expands to more than %d additive terms", GPL_MAX_TERMS);
            return GPL_ERR_LIMIT;
        }
        stack.push_back(std::move(cur));
    }
    if (stack.size() != 1) {
        snprintf(msg, 160, "kernel-program: malformed postfix (%zu values left on the stack)", stack.size());
        return GPL_ERR_ARG;
    }
    const Poly &p = stack[0];
    std::memset(out, 0, sizeof(*out));
    int nf = 0;
    out->n_terms = (int)p.size();
    for (int t = 0; t < (int)p.size(); ++t) {
        out->term_begin[t] = nf;
        out->coef[t] = p[t].coef;
        // per-item scalar factors first: the kernels fold them into one coefficient per term (kfun.cuh)
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) out->leaf_begin[t] = nf;
            for (const DevFactor &f : p[t].fs) {
                if ((f.kind == F_PARAM) != (pass == 0)) continue;
                if (nf >= GPL_MAX_FACTORS) {
                    snprintf(msg, 160, "kernel-program: more than %d factors after expansion", GPL_MAX_FACTORS);
                    return GPL_ERR_LIMIT;
                }
                if (f.kind == F_NOISE) out->has_noise |= 1 << t;
                out->f[nf++] = f;
            }
        }
    }
    out->term_begin[p.size()] = nf;
    out->n_factors = nf;
    out->n_theta = max_slot + 1;
    out->n_cols = max_col + 1;
    return GPL_OK;
}

}  // namespace gpl
