"""solids4foam_b200: the B200-resident hot path of solids4foam's segregated finite-volume solid solver.

Layout: ``csrc/`` (CUDA kernels + the C-ABI of include/s4fgpu.h), ``solid_model.py`` (host mirror of the
solidModel interface), ``case.py``/``cases.py``/``mesh.py`` (case description and synthetic meshes).
"""
from . import case, cases, mesh  # noqa: F401

__all__ = ["case", "cases", "mesh"]
