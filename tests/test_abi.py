"""The C-ABI library loads and exports every symbol include/gaplac_b200.h declares; host-side logic that needs no
GPU (program compilation and its error paths) works; and without a GPU the product fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from gaplac_b200 import _lib
from gaplac_b200.formula import Op
from gaplac_b200._lib import ADD, CAT, MUL, NOISE, SQEXP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "gaplac_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gpl_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gaplac_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names
    assert lib.gpl_abi_version() == 2


def test_gpl_op_layout_is_32_bytes():
    assert C.sizeof(_lib.GplOp) == 32


def _create(ops):
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.gpl_program_create(None, _lib.ops_array(ops), len(ops), C.byref(h))
    return rc, h


def test_program_compiles_without_a_device():
    rc, h = _create([Op(CAT, col=0), Op(CAT, col=1), Op(MUL, var_slot=0), Op(CAT, col=0, var_slot=1), Op(ADD),
                     Op(NOISE, var_slot=3), Op(ADD)])
    assert rc == 0
    lib = _lib.load()
    assert lib.gpl_program_n_theta(h) == 4 and lib.gpl_program_n_cols(h) == 2
    lib.gpl_program_destroy(h)


@pytest.mark.parametrize("ops,status", [
    ([Op(ADD)], _lib.GPL_ERR_ARG),                                     # operator without operands
    ([Op(SQEXP, col=0), Op(SQEXP, col=0)], _lib.GPL_ERR_ARG),          # two values left on the stack
    ([Op(SQEXP, col=0, value=0.0)], _lib.GPL_ERR_ARG),                 # lengthscale must be > 0
    ([Op(SQEXP, col=99)], _lib.GPL_ERR_LIMIT),
    ([Op(SQEXP, col=0, theta_slot=40)], _lib.GPL_ERR_LIMIT),
    ([Op(42)], _lib.GPL_ERR_ARG),
])
def test_malformed_programs_are_rejected(ops, status):
    rc, _ = _create(ops)
    assert rc == status
    assert _lib.load().gpl_last_error(None)


def test_product_expansion_limit():
    ops = [Op(SQEXP, col=0), Op(CAT, col=1), Op(ADD)]
    for _ in range(4):                       # ((a+b) * (a+b)) ... : 2, 4, 8, 16, 32 terms
        ops = ops + ops + [Op(MUL)]
    rc, _ = _create(ops) if len(ops) <= 64 else (_lib.GPL_ERR_ARG, None)
    assert rc in (_lib.GPL_ERR_LIMIT, _lib.GPL_ERR_ARG)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.GaplacError) as e:
        _lib.Context(0)
    assert e.value.status == _lib.GPL_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "gaplac_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"
