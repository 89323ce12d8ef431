"""Host-side mirror of the reference's solidModel interface for the GPU path.

``SolidModel.New(case)`` plays ``solidModel::New`` (SM/solidModel/solidModel.C:1673-1749): it looks
the model name up in a table and constructs it; the constructor mirrors mesh, geometry, law and
boundary conditions into the device once (``s4fgpu_set_*``) and performs the constructor's consistent
start (linGeomTotalDispSolid.C:82-84).  ``evolve()`` / ``updateTotalFields()`` / ``D()`` / ``sigma()``
/ ``gradD()`` / ``setTraction()`` keep the reference names (solidModel.H:570-798).
Everything numerical happens behind the C-ABI in CUDA; this file is plumbing only.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import case as K
from ._lib import lib

KERNELS = dict(spmv1=0, spmv3=1, pcg_iter=2, grad=3, law=4, rhs=5, pcg_p=6, pcg_xr=7, gamg_vcycle=8, gamg_step0=9, halo3=10, dot_reduce=11)


def _euler_d2dt2(vf, vf_o, vf_oo, dt: float, dt0: float):
    """[OF-ext] EulerD2dt2Scheme::fvcD2dt2 on a fixed mesh: rDeltaT2 (c vf - (c + c00) vf.o + c00 vf.oo)."""
    c, c00 = (dt + dt0) / (2.0 * dt), (dt + dt0) / (2.0 * dt0)
    return 4.0 / (dt + dt0) ** 2 * (c * vf - (c + c00) * vf_o + c00 * vf_oo)


class SolidModel:
    # run-time selection table: solidProperties "solidModel" -> enum behind the C-ABI.  The plugin
    # registers the same names with the prefix "gpu" (foam_plugin/).
    _table: Dict[str, int] = {
        "gpuLinearGeometryTotalDisplacement": K.MODEL_LIN_GEOM_TOTAL_DISP,
        "linearGeometryTotalDisplacement": K.MODEL_LIN_GEOM_TOTAL_DISP,
        "gpuNonLinearGeometryTotalLagrangianTotalDisplacement": K.MODEL_NONLIN_TL_TOTAL_DISP,
        "nonLinearGeometryTotalLagrangianTotalDisplacement": K.MODEL_NONLIN_TL_TOTAL_DISP,
        "gpuNonLinearGeometryTotalLagrangian": K.MODEL_NONLIN_TL,
        "nonLinearGeometryTotalLagrangian": K.MODEL_NONLIN_TL,
        "gpuUnsLinearGeometry": K.MODEL_UNS_LIN_GEOM,
        "unsLinearGeometry": K.MODEL_UNS_LIN_GEOM,
        "gpuUnsNonLinearGeometryTotalLagrangian": K.MODEL_UNS_NONLIN_TL,
        "unsNonLinearGeometryTotalLagrangian": K.MODEL_UNS_NONLIN_TL,
        "gpuUnsNonLinearGeometryUpdatedLagrangian": K.MODEL_UNS_NONLIN_UL,
        "unsNonLinearGeometryUpdatedLagrangian": K.MODEL_UNS_NONLIN_UL,
        "gpuNonLinearGeometryUpdatedLagrangian": K.MODEL_NONLIN_UL,
        "nonLinearGeometryUpdatedLagrangian": K.MODEL_NONLIN_UL,
    }

    @classmethod
    def New(cls, case: K.SolidCase, solidModel: Optional[str] = None, device: int = 0, comm=None) -> "SolidModel":
        if solidModel is not None:
            if solidModel not in cls._table:
                raise KeyError(f"Unknown solidModel type {solidModel}\nValid solidModel types are: {sorted(cls._table)}")
            case.controls.solidModel = cls._table[solidModel]
        return cls(case, device=device, comm=comm)

    def __init__(self, case: K.SolidCase, device: int = 0, comm=None):
        self.L = lib()
        self.case = case
        self.h = C.c_void_p()
        rc = self.L.s4fgpu_create(C.byref(self.h), device)
        if rc != 0:
            raise RuntimeError("s4fgpu_create: " + self.L.s4fgpu_last_error(None).decode())
        if comm is not None:
            nRanks, rank, uid = comm
            self._check(self.L.s4fgpu_comm_init(self.h, nRanks, rank, uid))
        K.apply_case(self.L, "s4fgpu_", self.h, case, self._check)
        self._check(self.L.s4fgpu_initialise(self.h))
        # mesh motion of the updated-Lagrangian model on the device (s4fgpu_move_points); 2-D meshes (empty patches) keep the
        # host route: set_geometry / set_points with the geometry the host recomputed
        self.device_mesh_motion = bool(np.all(np.asarray(case.mesh.solutionD) != 0))

    # ---- plumbing ----
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError("libs4fgpu: " + self.L.s4fgpu_last_error(self.h).decode())

    def close(self) -> None:
        if getattr(self, "h", None) is not None and self.h:
            self.L.s4fgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get(self, name: str, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Download a field into host memory (``out``: caller-owned buffer, e.g. pinned, of the field's shape)."""
        shape = K.field_size(self.case.mesh, name)
        if out is None:
            out = np.empty(shape)
        elif out.shape != tuple(shape) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous float64 array of shape {shape}")
        self._check(self.L.s4fgpu_download(self.h, K.FIELD[name], K._dptr(out)))
        return out

    def set(self, name: str, a) -> None:
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(K.field_size(self.case.mesh, name))
        self._check(self.L.s4fgpu_upload(self.h, K.FIELD[name], K._dptr(a)))

    def set_controls(self, controls: K.Controls) -> None:
        self.case.controls = controls
        self._check(self.L.s4fgpu_set_controls(self.h, C.byref(controls)))

    def set_bc(self, patch_name: str, bc: K.BC) -> None:
        m = self.case.mesh
        ip = [p.name for p in m.patches].index(patch_name)
        p = m.patches[ip]
        val = None if bc.value is None else np.ascontiguousarray(np.broadcast_to(bc.value, (p.size, 3)), dtype=np.float64)
        pr = None if bc.pressure is None else np.ascontiguousarray(np.broadcast_to(bc.pressure, (p.size,)), dtype=np.float64)
        self._check(self.L.s4fgpu_set_bc(self.h, ip, bc.kind, None if val is None else K._dptr(val),
                                         None if pr is None else K._dptr(pr)))

    # ---- reference-named interface ----
    def _traction_patch(self, patch_name: str) -> K.BC:
        bc = self.case.bcs.get(patch_name)
        if bc is None or bc.kind != K.BC_SOLID_TRACTION:      # the FatalError of solidModel.C:1784-1800
            raise RuntimeError(f"Boundary condition of patch {patch_name} should instead be type solidTraction")
        return bc

    def setTraction(self, patch_name: str, traction, pressure=None) -> None:
        """solidModel::setTraction (solidModel.C:1752-1817): new traction on a solidTraction patch (the pressure is kept
        unless given)."""
        bc = self._traction_patch(patch_name)
        new = K.solidTraction(traction, bc.pressure if pressure is None else pressure)
        self.case.bcs[patch_name] = new
        self.set_bc(patch_name, new)

    def setPressure(self, patch_name: str, pressure) -> None:
        """solidModel::setPressure (solidModel.C:1820-1890): new pressure on a solidTraction patch, the traction is kept."""
        bc = self._traction_patch(patch_name)
        new = K.solidTraction(bc.value if bc.value is not None else (0.0, 0.0, 0.0), pressure)
        self.case.bcs[patch_name] = new
        self.set_bc(patch_name, new)

    # -- what the FSI coupler reads (fluidSolidInterface::updateDisplacement / AitkenCouplingInterface.C:67-136) --
    def enable_interface_fields(self) -> None:
        """Keep pointD.oldTime() and the boundary values of D at the two old time levels: what faceZonePointDisplacementOld
        and faceZoneAcceleration need.  Costs one vol->point interpolation and one boundary download per time step."""
        if getattr(self, "_iface", None) is None:
            Db = self.get("D_b")
            self._iface = dict(pointD_old=self.pointD(), Db_o=Db.copy(), Db_oo=Db.copy(), dt=None, dt0=None)

    def patchMeshPoints(self, patch_name: str) -> np.ndarray:
        """PrimitivePatch::meshPoints() of a patch: its points in the order the faces meet them."""
        m = self.case.mesh
        sl = m.patch_slice(patch_name)
        verts = np.asarray(m.faces)[m.nInternalFaces + sl.start:m.nInternalFaces + sl.stop].ravel()
        _, first = np.unique(verts, return_index=True)
        return verts[np.sort(first)]

    def faceZonePointDisplacementIncrement(self, patch_name: str) -> np.ndarray:
        """solidModel::faceZonePointDisplacementIncrement (solidModel.C:1579-1593): pointDD at the interface's meshPoints."""
        self.enable_interface_fields()
        ids = self.patchMeshPoints(patch_name)
        if self.case.controls.solidModel in K.INCREMENTAL_MODELS:
            return self.interpolate_to_points("DD", with_gradient=True)[ids]
        return (self.pointD() - self._iface["pointD_old"])[ids]

    def faceZonePointDisplacementOld(self, patch_name: str) -> np.ndarray:
        """solidModel::faceZonePointDisplacementOld (solidModel.C:1596-1610): pointD.oldTime() at the interface's meshPoints."""
        self.enable_interface_fields()
        return self._iface["pointD_old"][self.patchMeshPoints(patch_name)]

    def faceZoneAcceleration(self, patch_name: str) -> np.ndarray:
        """solidModel::faceZoneAcceleration (solidModel.C:1613-1625): the patch field of fvc::d2dt2(D) with the case's d2dt2
        scheme (Euler: EulerD2dt2Scheme, variable step; steadyState: zero)."""
        self.enable_interface_fields()
        sl = self.case.mesh.patch_slice(patch_name)
        it, sch = self._iface, self.case.controls.d2dt2Scheme
        Db = self.get("D_b")[sl]
        if sch == K.D2DT2_STEADY_STATE or it["dt"] is None:
            return np.zeros_like(Db)
        if sch != K.D2DT2_EULER:
            raise RuntimeError("faceZoneAcceleration: available with the Euler and steadyState d2dt2 schemes")
        dt, dt0 = it["dt"], it["dt0"] if it["dt0"] else it["dt"]
        return _euler_d2dt2(Db, it["Db_o"][sl], it["Db_oo"][sl], dt, dt0)

    def initialise(self) -> None:
        self._check(self.L.s4fgpu_initialise(self.h))

    def new_timestep(self, deltaT: float = 1.0) -> None:
        it = getattr(self, "_iface", None)
        if it is not None:          # oldTime() of the interface fields, before the device shifts its own time levels
            it["pointD_old"] = self.pointD()
            it["Db_oo"], it["Db_o"] = it["Db_o"], self.get("D_b")
            it["dt0"], it["dt"] = it["dt"], float(deltaT)
        self._check(self.L.s4fgpu_new_timestep(self.h, deltaT))

    def outer_iteration(self) -> dict:
        st = K.Stats()
        self._check(self.L.s4fgpu_outer_iteration(self.h, C.byref(st)))
        return st.as_dict()

    def evolve(self) -> dict:
        st = K.Stats()
        self._check(self.L.s4fgpu_evolve(self.h, C.byref(st)))
        return st.as_dict()

    def interpolate_to_points(self, name: str = "D", with_gradient: bool = False) -> np.ndarray:
        """mechanicalModel::interpolate(D, pointD) (mechanicalModel.C:786-826) or, ``with_gradient``, interpolate(D, gradD,
        pointD) (:829-877): vol -> point interpolation on the device."""
        out = np.empty((self.case.mesh.points.shape[0], 3))
        mode = K.POINT_INTERP_GRAD if with_gradient else K.POINT_INTERP_PATCH
        self._check(self.L.s4fgpu_interpolate_to_points(self.h, K.FIELD[name], mode, K._dptr(out)))
        return out

    def pointD(self) -> np.ndarray:
        """pointD as the solid models update it after the loop, e.g. linGeomTotalDispSolid.C:212."""
        return self.interpolate_to_points("D", with_gradient=True)

    def movingMesh(self) -> bool:
        """solidModel::movingMesh(): the updated-Lagrangian model moves the mesh at the end of every step."""
        return self.case.controls.solidModel in K.MOVING_MESH_MODELS

    def updateTotalFields(self) -> None:
        """solidModel::updateTotalFields; for the updated-Lagrangian model nonLinGeomUpdatedLagSolid::updateTotalFields
        (nonLinGeomUpdatedLagSolid.C:360-374): density update, moveMesh(oldPoints, DD, pointDD), law history."""
        if self.movingMesh():
            pointDD = self.interpolate_to_points("DD")          # stays on the device as well
            self._check(self.L.s4fgpu_update_total_fields(self.h))
            if self.device_mesh_motion:
                self._check(self.L.s4fgpu_move_points(self.h, None))      # points, geometry, weights, GAMG coefficients: all on the device
                K.move_mesh(self.L, "s4fgpu_", self.h, self.case, pointDD, self._check, mirror=False)   # the host's own polyMesh follows
            else:
                K.move_mesh(self.L, "s4fgpu_", self.h, self.case, pointDD, self._check)
            self.pointDD = pointDD
            return
        self._check(self.L.s4fgpu_update_total_fields(self.h))

    update_total_fields = updateTotalFields

    def D(self) -> np.ndarray:
        return self.get("D")

    def gradD(self) -> np.ndarray:
        return self.get("gradD")

    def sigma(self) -> np.ndarray:
        return self.get("sigma")

    # ---- single operators ----
    def op_grad(self) -> None:
        self._check(self.L.s4fgpu_op_grad(self.h))

    def op_correct(self) -> None:
        self._check(self.L.s4fgpu_op_correct(self.h))

    def op_assemble(self) -> None:
        self._check(self.L.s4fgpu_op_assemble(self.h))

    def op_amul(self, cmpt: int, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        self._check(self.L.s4fgpu_op_amul(self.h, cmpt, K._dptr(x), K._dptr(y)))
        return y

    def op_solve(self, psi: np.ndarray, source: np.ndarray):
        psi = np.ascontiguousarray(psi, dtype=np.float64).copy()
        source = np.ascontiguousarray(source, dtype=np.float64)
        st = K.Stats()
        self._check(self.L.s4fgpu_op_solve(self.h, K._dptr(psi), K._dptr(source), C.byref(st)))
        return psi, st.as_dict()

    def time_kernel(self, kernel: str, reps: int = 20, flush_l2: bool = True):
        ms, by = C.c_double(), C.c_double()
        self._check(self.L.s4fgpu_time_kernel(self.h, KERNELS[kernel], reps, 1 if flush_l2 else 0, C.byref(ms), C.byref(by)))
        return ms.value, by.value

    def timer_start(self) -> None:
        self._check(self.L.s4fgpu_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        self._check(self.L.s4fgpu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def synchronize(self) -> None:
        self._check(self.L.s4fgpu_synchronize(self.h))

    def gamg_info(self) -> dict:
        n, sizes, by, st = C.c_int(), (C.c_int * 16)(), C.c_double(), C.c_double()
        self._check(self.L.s4fgpu_gamg_info(self.h, C.byref(n), sizes, 16, C.byref(by), C.byref(st)))
        return dict(levels=[sizes[i] for i in range(n.value)], bytes_per_vcycle=by.value, setup_seconds=st.value,
                    distributed_levels=int(self.L.s4fgpu_gamg_distributed_levels(self.h)))

    def launch_count(self) -> int:
        return int(self.L.s4fgpu_launch_count(self.h))


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib().s4fgpu_get_unique_id(buf)
    if rc != 0:
        raise RuntimeError("ncclGetUniqueId failed")
    return buf.raw
