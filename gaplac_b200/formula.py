"""GaPLAC formula language: AST, closed-grammar parser, and flattening to the backend's kernel-program.

Host-side mirror of the reference's formula layer (no arithmetic happens here):

* AST                     src/gp_parts.jl:3-9 (GPOperation), :21-47 (SqExp, Linear, OU, Cat), :51-59 (varnames, +, *)
* gp_spec("y :lik ~| f")  src/interface.jl:12-34 — the reference `eval`s the formula text as Julia code; this
                          parser accepts the same surface syntax through a closed grammar instead
* kernel(formula; hyperparams) / make_gp   src/abstractgp_translations.jl:31-35, 45-71; src/interface.jl:36-41

`Constant` and `Noise` are the two components the README and the legacy fixtures name but the current src/
lacks; per-node variance multipliers (`var=`) reproduce the legacy θc[σ2] parameters.
A hyperparameter may be a number (fixed) or `Slot(k)`: the k-th entry of the per-item hyperparameter vector
of a batched evaluation.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Union

from ._lib import ADD, CAT, CONSTANT, LINEAR, MUL, NOISE, OU as K_OU, SQEXP


@dataclass(frozen=True)
class Slot:
    """Reference to entry `index` of the per-item hyperparameter vector theta."""
    index: int


Hyper = Union[float, int, Slot]


class GPComponent:
    """abstract type GPCompnent (src/gp_parts.jl:3)."""
    var: Hyper = 1.0

    def __add__(self, other: "GPComponent") -> "GPOperation":   # src/gp_parts.jl:55
        return GPOperation("add", self, other)

    def __mul__(self, other: "GPComponent") -> "GPOperation":   # src/gp_parts.jl:59
        return GPOperation("multiply", self, other)

    def scaled(self, var: Hyper):
        """Same node with a variance multiplier (number or Slot)."""
        import copy
        c = copy.copy(self)
        object.__setattr__(c, "var", var)
        return c


@dataclass
class GPOperation(GPComponent):       # src/gp_parts.jl:5-9
    op: str
    lhs: GPComponent
    rhs: GPComponent
    var: Hyper = 1.0


@dataclass
class SqExp(GPComponent):             # src/gp_parts.jl:21-27 ; keyword l (default 1)
    varname: str
    l: Hyper = 1.0
    var: Hyper = 1.0


@dataclass
class Linear(GPComponent):            # src/gp_parts.jl:29-35 ; keyword c (default 0)
    varname: str
    c: Hyper = 0.0
    var: Hyper = 1.0


@dataclass
class OU(GPComponent):                # src/gp_parts.jl:37-43
    varname: str
    l: Hyper = 1.0
    var: Hyper = 1.0


@dataclass
class Cat(GPComponent):               # src/gp_parts.jl:45-47
    varname: str
    var: Hyper = 1.0


@dataclass
class Constant(GPComponent):          # SURVEY.md §A.2 (legacy `Constant(1)`, test/oldtests.jl:11)
    c: Hyper = 1.0
    var: Hyper = 1.0


@dataclass
class Noise(GPComponent):             # SURVEY.md §A.2 (legacy fixtures: θ4·I)
    var: Hyper = 1.0


def varnames(c: GPComponent) -> list[str]:
    """Left-to-right leaf order (src/gp_parts.jl:51-53). Constant/Noise read no variable."""
    if isinstance(c, GPOperation):
        return varnames(c.lhs) + varnames(c.rhs)
    return [c.varname] if hasattr(c, "varname") else []


# ---------------------------------------------------------------------------------------------------------
# Spec and parser (src/interface.jl:1-34)
# ---------------------------------------------------------------------------------------------------------
class Gaussian:                        # src/liklihoods.jl
    def __repr__(self):
        return "Gaussian()"


@dataclass
class Spec:
    response: str
    lik: object
    formula: GPComponent


def response(s: Spec) -> str:
    return s.response


def likelihood(s: Spec):
    return s.lik


def formula(s: Spec) -> GPComponent:
    return s.formula


_TOKEN = re.compile(r"\s*(?:(?P<num>[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?))"
                    r"|(?P<sym>:[A-Za-z_][A-Za-z_0-9]*)|(?P<name>[A-Za-z_][A-Za-z_0-9]*)"
                    r"|(?P<str>\"[^\"]*\")|(?P<op>[()+*;,=]))")

_LEAVES = {"SqExp": SqExp, "OU": OU, "Linear": Linear, "Cat": Cat, "Constant": Constant, "Noise": Noise}


class _Parser:
    def __init__(self, text: str):
        self.toks = []
        pos = 0
        text = text.strip()
        while pos < len(text):
            m = _TOKEN.match(text, pos)
            if not m or m.end() == pos:
                raise ValueError(f"Invalid formula specification near {text[pos:pos + 12]!r}")
            kind = m.lastgroup
            self.toks.append((kind, m.group(kind)))
            pos = m.end()
        self.i = 0

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else (None, None)

    def take(self, value=None):
        k, v = self.peek()
        if k is None or (value is not None and v != value):
            raise ValueError(f"Invalid formula specification: expected {value!r}, got {v!r}")
        self.i += 1
        return k, v

    def expr(self) -> GPComponent:
        node = self.term()
        while self.peek()[1] == "+":
            self.take()
            node = node + self.term()
        return node

    def term(self) -> GPComponent:
        node = self.factor()
        while self.peek()[1] == "*":
            self.take()
            node = node * self.factor()
        return node

    def value(self) -> Hyper:
        k, v = self.take()
        if k == "num":
            return float(v)
        if k == "name" and v == "Slot":
            self.take("(")
            _, n = self.take()
            self.take(")")
            return Slot(int(n))
        raise ValueError(f"Invalid formula specification: expected a number, got {v!r}")

    def factor(self) -> GPComponent:
        k, v = self.take()
        if v == "(":
            node = self.expr()
            self.take(")")
            return node
        if k != "name" or v not in _LEAVES:
            raise ValueError(f"Invalid formula specification: unknown component {v!r}")
        cls = _LEAVES[v]
        pos, kw = [], {}
        if self.peek()[1] == "(":
            self.take("(")
            while self.peek()[1] != ")":
                k2, v2 = self.peek()
                if k2 == "name" and self.i + 1 < len(self.toks) and self.toks[self.i + 1][1] == "=":
                    self.take()
                    self.take("=")
                    kw[v2] = self.value()
                elif k2 == "sym":
                    self.take()
                    pos.append(v2[1:])
                elif k2 == "str":
                    self.take()
                    pos.append(v2[1:-1])
                elif k2 == "name" and v2 != "Slot":
                    self.take()
                    pos.append(v2)          # legacy bare names: Cat(PersonID) (test/pred.jl:3)
                else:
                    pos.append(self.value())
                if self.peek()[1] in (",", ";"):
                    self.take()
            self.take(")")
        try:
            return cls(*pos, **kw)
        except TypeError as e:
            raise ValueError(f"Invalid formula specification: {v}: {e}") from None


def parse_formula(text: str) -> GPComponent:
    p = _Parser(text)
    node = p.expr()
    if p.peek()[0] is not None:
        raise ValueError(f"Invalid formula specification: trailing {p.peek()[1]!r}")
    return node


def gp_spec(text: str) -> Spec:
    """`resp [: lik] ~| formula` (src/interface.jl:12-34); raises ValueError where the reference throws ArgumentError."""
    tilde = text.find("~")
    if tilde < 0 or tilde + 1 >= len(text) or text[tilde + 1] != "|":
        raise ValueError("Invalid formula specification")
    colon = text.find(":")
    if colon < 0 or colon > tilde:
        lik, resp = Gaussian(), text[:tilde].strip()
    else:
        lik_txt = text[colon + 1:tilde].strip()
        if lik_txt not in ("", "Gaussian", "Gaussian()"):
            raise ValueError(f"Invalid formula specification: likelihood {lik_txt!r} not supported (Gaussian only)")
        lik, resp = Gaussian(), text[:colon].strip()
    if not resp:
        raise ValueError("Invalid formula specification")
    return Spec(resp, lik, parse_formula(text[tilde + 2:]))


# ---------------------------------------------------------------------------------------------------------
# AST -> postfix kernel-program (src/abstractgp_translations.jl:8-15, 31-35, 45-71)
# ---------------------------------------------------------------------------------------------------------
@dataclass
class Op:
    kind: int
    col: int = 0
    theta_slot: int = -1
    var_slot: int = -1
    value: float = 1.0
    var: float = 1.0


@dataclass
class KernelProgram:
    """What `kernel()` returns in place of a KernelFunctions kernel object."""
    ops: list = field(default_factory=list)
    vars: list = field(default_factory=list)      # variable of each input column
    n_theta: int = 0


def _hyper(h: Hyper):
    if isinstance(h, Slot):
        return h.index, 1.0
    return -1, float(h)


def kernel(f: GPComponent, hyperparams: dict | None = None, unique_columns: bool = False):
    """Flatten the AST.  i-th variable-bearing leaf reads the i-th column (src/abstractgp_translations.jl:45-71);
    with unique_columns=True a variable used by several leaves is stored once.  hyperparams[varname] overrides the
    leaf's own hyperparameter exactly as makekernel(c, hyperparams[varname(c)]) does (:13-15, :33); Cat has no
    hyperparameter there (MethodError -> TypeError here).  Returns (KernelProgram, vars)."""
    hyperparams = hyperparams or {}
    vs = varnames(f)
    cols = list(dict.fromkeys(vs)) if unique_columns else list(vs)
    ops: list[Op] = []
    counter = [0]
    max_slot = [-1]

    def slot_of(h):
        s, v = _hyper(h)
        max_slot[0] = max(max_slot[0], s)
        return s, v

    def walk(c: GPComponent):
        vslot, vval = slot_of(c.var)
        if isinstance(c, GPOperation):
            walk(c.lhs)
            walk(c.rhs)
            if c.op not in ("add", "multiply"):
                raise ValueError(f"Operation {c.op} not yet supported")     # _convertop, :21-29
            ops.append(Op(ADD if c.op == "add" else MUL, var_slot=vslot, var=vval))
            return
        if isinstance(c, Noise):
            ops.append(Op(NOISE, var_slot=vslot, var=vval))
            return
        if isinstance(c, Constant):
            s, v = slot_of(c.c)
            ops.append(Op(CONSTANT, theta_slot=s, value=v, var_slot=vslot, var=vval))
            return
        col = cols.index(c.varname) if unique_columns else counter[0]
        counter[0] += 1
        if isinstance(c, Cat):
            if c.varname in hyperparams:
                raise TypeError("no method matching makekernel(::Cat, hyperparameter)")
            ops.append(Op(CAT, col=col, var_slot=vslot, var=vval))
            return
        h = hyperparams.get(c.varname, c.l if isinstance(c, (SqExp, OU)) else c.c)
        s, v = slot_of(h)
        kind = SQEXP if isinstance(c, SqExp) else K_OU if isinstance(c, OU) else LINEAR
        ops.append(Op(kind, col=col, theta_slot=s, value=v, var_slot=vslot, var=vval))

    walk(f)
    return KernelProgram(ops=ops, vars=cols, n_theta=max_slot[0] + 1), cols


def make_gp(spec: Spec, hyperparams: dict | None = None, unique_columns: bool = False):
    """(GP(kern), vars) — src/interface.jl:36-41."""
    from .gp import GP
    kp, vs = kernel(spec.formula, hyperparams, unique_columns)
    n_leaves = sum(1 for o in kp.ops if o.kind in (SQEXP, K_OU, LINEAR, CAT))
    if not unique_columns and len(vs) != n_leaves:
        raise RuntimeError("Something went wrong with equation parsing, number of variables should == number of kernels")
    return GP(kp), vs
