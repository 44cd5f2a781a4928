"""One small call of every kernel family, for compute-sanitizer (tools/sanitize.py runs this under memcheck / racecheck /
synccheck).  Sizes are tiny: the tools slow kernels down 10-100x.  `--big worker` includes the look-ahead path with the
persistent worker CTA (kernels that wait on one another through flags): only safe under tools that keep kernels
concurrent; the default exercises the look-ahead streams without the worker (chol_variant 3) and the serial variant."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from gaplac_b200 import _lib, mcmc, workloads as W
    from gaplac_b200.formula import Op
    from gaplac_b200._lib import ADD, CAT, CONSTANT, LINEAR, MUL, NOISE, OU, SQEXP
    big = sys.argv[sys.argv.index("--big") + 1] if "--big" in sys.argv else "streams"
    ops = [Op(SQEXP, col=0, theta_slot=0, var_slot=3), Op(OU, col=1, theta_slot=1), Op(MUL), Op(LINEAR, col=2, theta_slot=2),
           Op(CAT, col=3), Op(MUL, var=0.7), Op(ADD), Op(CONSTANT, value=0.3), Op(ADD), Op(NOISE, var_slot=4), Op(ADD)]
    th = np.array([1.3, 0.8, 0.4, 1.7, 0.2])
    rng = np.random.default_rng(0)

    def data(n):
        return (np.column_stack([rng.uniform(-3, 3, n), rng.uniform(0, 5, n), rng.standard_normal(n),
                                 rng.integers(0, 4, n).astype(float)]), rng.standard_normal(n))

    ctx = _lib.Context(0)
    prog = ctx.program(ops)
    X, y = data(150)
    Th = np.vstack([th, th * 1.1, th * 0.9])
    done = []
    ctx.cov(prog, X, th, 0.1)                                              # cov_dense_kernel
    ctx.cross_cov(prog, X, data(70)[0], th)
    done.append("cov_dense")
    ctx.lml_batched(prog, X, y, Th, 0.1)                                   # lk_diag / lk_potrf_warp / lk_below
    done.append("lockstep lml")
    ctx.lml_batched(prog, X, y, Th, 0.1, grad=True)                        # + lk_winv / lk_minv / lk_alpha / lk_gradc / lk_gradsum
    done.append("lockstep gradient")
    ctx.set_option("lml_variant", 1)
    ctx.lml_batched(prog, X, y, Th, 0.1, grad=True)                        # fused per-item kernels
    ctx.lml_batched(prog, X, y, Th, 0.1)
    ctx.set_option("lml_variant", 0)
    done.append("fused lml / gradient")
    post = ctx.posterior_fit(prog, X, y, th, 0.1)                          # keep path + predict (three slab widths)
    for m in (33, 9472 // 8, 12000 // 8):
        post.mean_and_var(data(m)[0])
    post.factor()
    post.free()
    ctx.predict_batched(prog, X, y, Th, 0.1, data(40)[0])                  # lk_post_kernel + batched predict
    ctx.sample(prog, X, th, 0.1, rng.standard_normal((150, 2)))            # sample_kernel
    done.append("posterior / predict / sample")
    d1 = W.make_c1(n=20)
    mcmc.nuts(ctx, ctx.program(d1["ops"]), d1["X"], d1["y"], [0.0], [20.0], sigma2=0.1, n_samples=3, n_adapt=3, seed=1, chains=2)
    done.append("sampler (graph replay)")
    Xb, yb = data(600)                                                     # nt = 10: three panels
    for variant in ([2, 3] + ([1] if big == "worker" else [])):
        ctx.set_option("chol_variant", variant)                            # 2: one stream; 3: look-ahead streams, no worker; 1: worker
        p2 = ctx.posterior_fit(prog, Xb, yb, th, 0.1)                      # cov_tiles + big_* + backward
        p2.mean_and_var(data(50)[0])
        p2.free()
        A = np.cov(rng.standard_normal((300, 900))) + np.eye(300)
        ctx.chol_logdet(A)                                                 # dense_to_tiles / tiles_to_upper
        done.append(f"large-n path variant {variant}")
    ctx.set_option("chol_variant", 0)
    ctx.close()
    print("sanitize target ok:", "; ".join(done))


if __name__ == "__main__":
    main()
