"""Host logic: formula parsing and flattening (mirrors the reference's inline testsets, src/interface.jl:69-87,
which check types only) plus the column-binding / hyperparameter-override semantics of
src/abstractgp_translations.jl:8-15,45-71."""
import pytest

import gaplac_b200 as G
from gaplac_b200._lib import ADD, CAT, CONSTANT, LINEAR, MUL, NOISE, OU, SQEXP


def test_formula_parsing_reference_cases():          # src/interface.jl:70-86
    s1 = G.gp_spec("y ~| SqExp(:t)")
    assert isinstance(G.likelihood(s1), G.Gaussian) and G.response(s1) == "y"
    assert isinstance(G.formula(s1), G.GPComponent) and isinstance(G.formula(s1), G.SqExp)
    s2 = G.gp_spec("bug ~| SqExp(:t) + Linear(:x)")
    assert G.response(s2) == "bug" and isinstance(G.formula(s2), G.GPOperation)
    s3 = G.gp_spec("bug ~| SqExp(:t) * Cat(:g) + Linear(:x)")
    assert isinstance(G.formula(s3), G.GPOperation) and G.formula(s3).op == "add"
    assert G.formula(s3).lhs.op == "multiply"        # Julia precedence: (A*B) + C


def test_readme_and_legacy_syntax():
    s = G.gp_spec('y :~| SqExp(:x; l=1.5)')           # README.md:48
    assert s.formula.l == 1.5 and s.response == "y"
    assert G.gp_spec("y ~| SqExp(:x, l=2)").formula.l == 2.0     # README.md:101
    s = G.gp_spec("bug :~| Cat(PersonID) * Cat(StoolPairs) + Cat(PersonID) + Linear(nutrient) + Noise")  # test/pred.jl:3
    assert G.varnames(s.formula) == ["PersonID", "StoolPairs", "PersonID", "nutrient"]
    assert isinstance(G.gp_spec("y ~| Constant(1) + OU(:t; l=3)").formula.lhs, G.Constant)


@pytest.mark.parametrize("bad", ["y SqExp(:t)", "y ~ SqExp(:t)", " ~| SqExp(:t)", "y ~| Foo(:t)", "y ~| SqExp(:t) +",
                                 "y ~| SqExp(:t))", "y : Poisson ~| SqExp(:t)", "y ~| __import__('os')"])
def test_invalid_specifications_raise(bad):           # ArgumentError in the reference (src/interface.jl:15,17,23)
    with pytest.raises(ValueError):
        G.gp_spec(bad)


def test_column_binding_follows_leaf_order():         # src/abstractgp_translations.jl:45-71
    kp, vs = G.kernel(G.gp_spec("y ~| SqExp(:a) * Cat(:b) + Linear(:c)").formula)
    assert vs == ["a", "b", "c"]
    assert [(o.kind, o.col) for o in kp.ops] == [(SQEXP, 0), (CAT, 1), (MUL, 0), (LINEAR, 2), (ADD, 0)]
    kp, vs = G.kernel(G.gp_spec("y ~| SqExp(:x) + OU(:x) + Noise").formula)
    assert vs == ["x", "x"] and [o.col for o in kp.ops if o.kind in (SQEXP, OU)] == [0, 1]
    kp, vs = G.kernel(G.gp_spec("y ~| SqExp(:x) + OU(:x) + Noise").formula, unique_columns=True)
    assert vs == ["x"] and [o.col for o in kp.ops if o.kind in (SQEXP, OU)] == [0, 0]
    assert kp.ops[-2].kind == NOISE and kp.ops[-1].kind == ADD


def test_hyperparameter_override_and_slots():         # makekernel(c, hyperparams[varname(c)]) :13-15, :33
    f = G.gp_spec("y ~| SqExp(:x; l=1.5) + Linear(:z; c=2) + OU(:x)").formula
    kp, _ = G.kernel(f, hyperparams={"x": 4.0})
    assert [o.value for o in kp.ops if o.kind in (SQEXP, OU)] == [4.0, 4.0]     # both :x leaves share it
    assert [o.value for o in kp.ops if o.kind == LINEAR] == [2.0]
    kp, _ = G.kernel(f, hyperparams={"x": G.Slot(0), "z": G.Slot(1)})
    assert kp.n_theta == 2 and [o.theta_slot for o in kp.ops if o.kind in (SQEXP, LINEAR, OU)] == [0, 1, 0]
    with pytest.raises(TypeError):                    # Cat has no 2-argument makekernel -> MethodError
        G.kernel(G.gp_spec("y ~| Cat(:g)").formula, hyperparams={"g": 1.0})


def test_make_gp_counts_variables():                  # src/interface.jl:36-41
    gp, vs = G.make_gp(G.gp_spec("y ~| SqExp(:t) * Cat(:g) + Linear(:x)"))
    assert vs == ["t", "g", "x"] and isinstance(gp, G.GP)


def test_variance_multipliers_and_constant():
    f = (G.Cat("P") * G.Cat("S")).scaled(G.Slot(0)) + G.Cat("P").scaled(G.Slot(1)) + G.Noise(var=G.Slot(2))
    kp, vs = G.kernel(f)
    assert kp.n_theta == 3 and vs == ["P", "S", "P"]
    assert [o.var_slot for o in kp.ops] == [-1, -1, 0, 1, -1, 2, -1]
    kp, _ = G.kernel(G.Constant(c=G.Slot(0)) + G.SqExp("x"))
    assert kp.ops[0].kind == CONSTANT and kp.ops[0].theta_slot == 0
