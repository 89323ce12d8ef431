"""The C-ABI boundary: libs4fgpu.so loads and exports every symbol include/s4fgpu.h declares, the ctypes
mirrors of the parameter structs have the C layout, and the product path fails loudly without a GPU
(no CPU fallback).  No compute calls: this file runs on the CPU-only box."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "s4fgpu.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s4fgpu_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def gpu_lib():
    from solids4foam_b200 import build
    build.build()
    from solids4foam_b200._lib import lib
    return lib()


def test_header_declares_the_documented_surface():
    fns = header_functions()
    assert len(fns) >= 28
    for must in ("s4fgpu_create", "s4fgpu_set_mesh", "s4fgpu_set_geometry", "s4fgpu_set_law", "s4fgpu_set_controls",
                 "s4fgpu_set_bc", "s4fgpu_evolve", "s4fgpu_outer_iteration", "s4fgpu_update_total_fields",
                 "s4fgpu_comm_init", "s4fgpu_op_amul", "s4fgpu_op_solve"):
        assert must in fns


def test_library_exports_every_declared_symbol(gpu_lib):
    from solids4foam_b200._lib import EXPORTS
    fns = header_functions()
    missing = [f for f in fns if not hasattr(gpu_lib, f)]
    assert not missing, missing
    assert sorted(EXPORTS) == fns, (set(fns) ^ set(EXPORTS))


def test_no_torch_or_cxx_types_in_the_exported_signatures():
    """extern "C", plain pointers and sizes only."""
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "solids4foam_b200", "libs4fgpu.so")],
                         capture_output=True, text=True, check=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if " T " in l]
    api = [n for n in names if n.startswith("s4fgpu_")]
    assert len(api) == len(header_functions())          # unmangled => extern "C"
    assert not [n for n in names if "torch" in n or "at6Tensor" in n]


def test_struct_layouts_match_the_c_header(tmp_path):
    """Compile a probe against include/s4fgpu.h and compare sizeof/offsetof with the ctypes mirrors."""
    from solids4foam_b200 import case as K
    probe = tmp_path / "probe.c"
    probe.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "s4fgpu.h"
int main(void) {
  printf("%zu %zu %zu\n", sizeof(s4fgpu_law), sizeof(s4fgpu_controls), sizeof(s4fgpu_stats));
  printf("%zu %zu %zu %zu\n", offsetof(s4fgpu_law, sigma0), offsetof(s4fgpu_law, tableSigY), offsetof(s4fgpu_law, updateBEbarConsistent), offsetof(s4fgpu_law, DEpsilonPRelax));
  printf("%zu %zu %zu %zu\n", offsetof(s4fgpu_controls, fieldRelaxD), offsetof(s4fgpu_controls, tolerance), offsetof(s4fgpu_controls, g), offsetof(s4fgpu_controls, checkEvery));
  printf("%zu %zu %zu\n", offsetof(s4fgpu_stats, nIterations), offsetof(s4fgpu_stats, relResidual), offsetof(s4fgpu_stats, totalInnerIterations));
  return 0; }''')
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)])
    rows = [[int(x) for x in l.split()] for l in subprocess.check_output([str(exe)], text=True).splitlines()]
    assert rows[0] == [C.sizeof(K.Law), C.sizeof(K.Controls), C.sizeof(K.Stats)]
    assert rows[1] == [K.Law.sigma0.offset, K.Law.tableSigY.offset, K.Law.updateBEbarConsistent.offset, K.Law.DEpsilonPRelax.offset]
    assert rows[2] == [K.Controls.fieldRelaxD.offset, K.Controls.tolerance.offset, K.Controls.g.offset, K.Controls.checkEvery.offset]
    assert rows[3] == [K.Stats.nIterations.offset, K.Stats.relResidual.offset, K.Stats.totalInnerIterations.offset]


def test_enum_values_match_the_header():
    from solids4foam_b200 import case as K
    src = open(HEADER).read()
    vals = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"\b(S4F_[A-Z0-9_]+)\s*=\s*(\d+)", src))
    assert vals["S4F_BC_SOLID_TRACTION"] == K.BC_SOLID_TRACTION and vals["S4F_BC_SOLID_SYMMETRY"] == K.BC_SOLID_SYMMETRY
    assert vals["S4F_MODEL_NONLIN_TL_TOTAL_DISP"] == K.MODEL_NONLIN_TL_TOTAL_DISP and vals["S4F_MODEL_NONLIN_UL"] == K.MODEL_NONLIN_UL
    assert vals["S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC"] == K.LAW_NEO_HOOKEAN_MISES_PLASTIC
    assert vals["S4F_PRECOND_DIC"] == K.PRECOND_DIC and vals["S4F_PRECOND_CHEBYSHEV"] == K.PRECOND_CHEBYSHEV
    for name, idx in K.FIELD.items():
        key = "S4F_FIELD_" + re.sub(r"(?<!^)(?=[A-Z])", "_", name).upper().replace("__", "_")
        key = {"S4F_FIELD_D_OLD_OLD": "S4F_FIELD_D_OLDOLD", "S4F_FIELD_GRAD_D_OLD": "S4F_FIELD_GRAD_D_OLD",
               "S4F_FIELD_EPSILON_P_EQ": "S4F_FIELD_EPSILON_P_EQ", "S4F_FIELD_B_EBAR": "S4F_FIELD_BEBAR",
               "S4F_FIELD_D_LAMBDA": "S4F_FIELD_DLAMBDA", "S4F_FIELD_D_EPSILON_P": "S4F_FIELD_DEPSILON_P",
               "S4F_FIELD_TRACTION_GRADIENT_B": "S4F_FIELD_TRACTION_GRADIENT_B",
               "S4F_FIELD_D_D": "S4F_FIELD_DD", "S4F_FIELD_GRAD_D_D": "S4F_FIELD_GRAD_DD",
               "S4F_FIELD_D_D_B": "S4F_FIELD_DD_B", "S4F_FIELD_SIGMAF": "S4F_FIELD_SIGMA_F", "S4F_FIELD_GRAD_DF": "S4F_FIELD_GRAD_D_F"}.get(key, key)
        assert vals[key] == idx, (name, key)


def test_product_path_fails_loudly_without_a_gpu(gpu_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = gpu_lib.s4fgpu_create(C.byref(h), 0)
    assert rc != 0 and not h
    msg = gpu_lib.s4fgpu_last_error(None).decode()
    assert "no CPU fallback" in msg
    from solids4foam_b200 import cases
    from solids4foam_b200.solid_model import SolidModel
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SolidModel(cases.cantilever(4, 2, 2))


def test_product_package_does_not_import_the_oracle():
    code = ("import sys; import solids4foam_b200, solids4foam_b200.solid_model, solids4foam_b200._lib; "
            "assert not [m for m in sys.modules if m.startswith('oracle')], 'oracle imported by the product'")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dp, _, fs in os.walk(os.path.join(ROOT, "solids4foam_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "liboracle" not in txt and "s4fo_" not in txt and "import oracle" not in txt, f
