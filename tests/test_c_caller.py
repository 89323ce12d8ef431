"""A plain-C program (tests/c_abi_driver.c, compiled with gcc against include/s4fgpu.h and linked to libs4fgpu.so)
drives a small cantilever end to end through the C-ABI: the boundary is usable from compiled host code, not only through
ctypes.  CPU part: it compiles, links and fails loudly without a GPU.  GPU part: its result equals the oracle's."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi_driver.c")
LIBDIR = os.path.join(ROOT, "solids4foam_b200")


def build_driver(tmpdir) -> str:
    from solids4foam_b200 import build as b
    b.build()
    exe = os.path.join(str(tmpdir), "c_abi_driver")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), SRC, "-L" + LIBDIR, "-ls4fgpu", "-lm",
           "-Wl,-rpath," + LIBDIR, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_c_driver_compiles_links_and_refuses_to_run_without_a_gpu(tmp_path):
    exe = build_driver(tmp_path)
    if have_gpu():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, "4", "2", "2", str(tmp_path / "D.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_driver_matches_the_oracle(tmp_path):
    from oracle.binding import OracleSolid
    from solids4foam_b200 import case as K
    from solids4foam_b200 import cases
    exe = build_driver(tmp_path)
    out = str(tmp_path / "D.bin")
    r = subprocess.run([exe, "6", "3", "3", out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    D = np.fromfile(out).reshape(-1, 3)
    o = OracleSolid(cases.cantilever(6, 3, 3, L=2.0, fieldRelaxD=0.9, nCorrectors=20000, solutionTolerance=1e-11,
                                     alternativeTolerance=1e-11, tolerance=1e-13, preconditioner=K.PRECOND_DIAGONAL))
    so = o.evolve()
    assert so["converged"]
    Do = o.get("D")
    assert D.shape == Do.shape
    assert np.linalg.norm(D - Do) / np.linalg.norm(Do) < 1e-8, r.stdout
