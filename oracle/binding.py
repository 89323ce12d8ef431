"""ctypes binding of the CPU oracle -- TEST INFRASTRUCTURE ONLY (see s4f_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It reuses the case plumbing of the package (apply_case) because the oracle mirrors the
C-ABI of include/s4fgpu.h with the prefix s4fo_.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from solids4foam_b200 import case as K

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "s4f_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "s4fgpu.h")
    stale = (not os.path.exists(so)) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.s4fo_create.restype = C.c_void_p
        L.s4fo_destroy.argtypes = [C.c_void_p]
        L.s4fo_last_error.argtypes = [C.c_void_p]
        L.s4fo_last_error.restype = C.c_char_p
        K.declare_api(L, "s4fo_", C.c_void_p)
        L.s4fo_get_ls_vectors.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.s4fo_table_lookup.argtypes = [C.POINTER(K.Law), C.c_double]
        L.s4fo_table_lookup.restype = C.c_double
        L.s4fo_uns_grad_from_points.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.s4fo_set_cpu_gamg.argtypes = [C.c_void_p, C.c_int]
        L.s4fo_cpu_gamg_levels.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
        _LIB = L
    return _LIB


class OracleSolid:
    """CPU oracle with the same python surface as solids4foam_b200.solid_model.SolidModel."""

    def __init__(self, case: K.SolidCase):
        self.L = lib()
        self.case = case
        self.h = C.c_void_p(self.L.s4fo_create())
        K.apply_case(self.L, "s4fo_", self.h, case, self._check)
        self._check(self.L.s4fo_initialise(self.h))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.L.s4fo_last_error(self.h).decode())

    def set_cpu_gamg(self, on: bool = True) -> None:
        """preconditioner GAMG = the oracle's own CPU multigrid (the GPU path's preconditioner family) instead of DIC"""
        self.L.s4fo_set_cpu_gamg(self.h, 1 if on else 0)

    def cpu_gamg_levels(self):
        sizes = (C.c_int * 16)()
        n = self.L.s4fo_cpu_gamg_levels(self.h, sizes, 16)
        return [int(sizes[i]) for i in range(n)]

    def __del__(self):
        try:
            self.L.s4fo_destroy(self.h)
        except Exception:
            pass

    def get(self, name: str) -> np.ndarray:
        out = np.zeros(K.field_size(self.case.mesh, name))
        self._check(self.L.s4fo_download(self.h, K.FIELD[name], K._dptr(out)))
        return out

    def set(self, name: str, a) -> None:
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(K.field_size(self.case.mesh, name))
        self._check(self.L.s4fo_upload(self.h, K.FIELD[name], K._dptr(a)))

    def set_controls(self, controls: K.Controls) -> None:
        self.case.controls = controls
        self._check(self.L.s4fo_set_controls(self.h, C.byref(controls)))

    def set_bc(self, patch_name: str, bc: K.BC) -> None:
        m = self.case.mesh
        ip = [p.name for p in m.patches].index(patch_name)
        p = m.patches[ip]
        val = None if bc.value is None else np.ascontiguousarray(np.broadcast_to(bc.value, (p.size, 3)), dtype=np.float64)
        pr = None if bc.pressure is None else np.ascontiguousarray(np.broadcast_to(bc.pressure, (p.size,)), dtype=np.float64)
        self._check(self.L.s4fo_set_bc(self.h, ip, bc.kind, None if val is None else K._dptr(val),
                                       None if pr is None else K._dptr(pr)))

    def initialise(self):
        self._check(self.L.s4fo_initialise(self.h))

    def new_timestep(self, deltaT: float = 1.0):
        self._check(self.L.s4fo_new_timestep(self.h, deltaT))

    def outer_iteration(self) -> dict:
        st = K.Stats()
        self._check(self.L.s4fo_outer_iteration(self.h, C.byref(st)))
        return st.as_dict()

    def evolve(self) -> dict:
        st = K.Stats()
        self._check(self.L.s4fo_evolve(self.h, C.byref(st)))
        return st.as_dict()

    def interpolate_to_points(self, name: str = "D", with_gradient: bool = False) -> np.ndarray:
        out = np.empty((self.case.mesh.points.shape[0], 3))
        mode = K.POINT_INTERP_GRAD if with_gradient else K.POINT_INTERP_PATCH
        self._check(self.L.s4fo_interpolate_to_points(self.h, K.FIELD[name], mode, K._dptr(out)))
        return out

    def update_total_fields(self):
        if self.case.controls.solidModel in K.MOVING_MESH_MODELS:
            pointDD = self.interpolate_to_points("DD")
            self._check(self.L.s4fo_update_total_fields(self.h))
            K.move_mesh(self.L, "s4fo_", self.h, self.case, pointDD, self._check)
            self.pointDD = pointDD
            return
        self._check(self.L.s4fo_update_total_fields(self.h))

    updateTotalFields = update_total_fields

    def op_grad(self):
        self._check(self.L.s4fo_op_grad(self.h))

    def op_correct(self):
        self._check(self.L.s4fo_op_correct(self.h))

    def op_assemble(self):
        self._check(self.L.s4fo_op_assemble(self.h))

    def op_amul(self, cmpt: int, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        self._check(self.L.s4fo_op_amul(self.h, cmpt, K._dptr(x), K._dptr(y)))
        return y

    def op_solve(self, psi: np.ndarray, source: np.ndarray):
        psi = np.ascontiguousarray(psi, dtype=np.float64).copy()
        source = np.ascontiguousarray(source, dtype=np.float64)
        st = K.Stats()
        self._check(self.L.s4fo_op_solve(self.h, K._dptr(psi), K._dptr(source), C.byref(st)))
        return psi, st.as_dict()

    def uns_grad_from_points(self, pointD: np.ndarray) -> None:
        """unsLinGeomSolid gradients (fvcGradf.C) from given vertex displacements, skipping the vol->point interpolation."""
        pd = np.ascontiguousarray(pointD, dtype=np.float64)
        self._check(self.L.s4fo_uns_grad_from_points(self.h, K._dptr(pd)))

    def ls_vectors(self):
        m = self.case.mesh
        lsP = np.zeros((m.nInternalFaces + m.nBoundaryFaces, 3))
        lsN = np.zeros((m.nInternalFaces, 3))
        self._check(self.L.s4fo_get_ls_vectors(self.h, K._dptr(lsP), K._dptr(lsN)))
        return lsP, lsN
