"""Case description: what the OpenFOAM dictionaries of a solids4foam case hold for this path.

``constant/solidProperties``  -> :class:`Controls` (solidModel + <model>Coeffs)
``constant/mechanicalProperties`` -> :func:`mechanical_law` (one law; ``laws.size()==1`` branch,
mechanicalModel.C:476-483)
``system/fvSchemes`` / ``system/fvSolution`` -> :class:`Controls`
``0/D`` boundaryField -> :class:`BC` per patch

The ctypes structures mirror ``include/s4fgpu.h`` field by field.  The same description drives the
CUDA library (``solid_model.py``) and, in the tests only, the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence

import numpy as np

from .mesh import FvMesh, PROCESSOR

# ---- enums (include/s4fgpu.h) ---------------------------------------------------------------
BC_FIXED_DISPLACEMENT, BC_SOLID_TRACTION, BC_SOLID_SYMMETRY, BC_PROCESSOR = 0, 1, 2, 3
MODEL_LIN_GEOM_TOTAL_DISP, MODEL_NONLIN_TL_TOTAL_DISP, MODEL_NONLIN_TL, MODEL_NONLIN_UL, MODEL_UNS_LIN_GEOM, MODEL_UNS_NONLIN_TL, MODEL_UNS_NONLIN_UL = 0, 1, 2, 3, 4, 5, 6
LAW_LINEAR_ELASTIC, LAW_NEO_HOOKEAN_ELASTIC, LAW_NEO_HOOKEAN_MISES_PLASTIC, LAW_LINEAR_ELASTIC_MISES_PLASTIC = 0, 1, 2, 3
GRAD_LEAST_SQUARES, GRAD_GAUSS_LINEAR, GRAD_POINT_CELLS_LEAST_SQUARES = 0, 1, 2
D2DT2_STEADY_STATE, D2DT2_EULER, D2DT2_BACKWARD = 0, 1, 2
STAB_NONE, STAB_RHIE_CHOW = 0, 1
RELAX_FIXED, RELAX_AITKEN = 0, 1
SOLVER_PCG, SOLVER_PBICGSTAB = 0, 1
PRECOND_NONE, PRECOND_DIAGONAL, PRECOND_DIC, PRECOND_CHEBYSHEV, PRECOND_GAMG = 0, 1, 2, 3, 4
POINT_INTERP_PATCH, POINT_INTERP_GRAD = 0, 1

FIELD = dict(D=0, D_old=1, D_oldOld=2, gradD=3, sigma=4, D_b=5, gradD_b=6, sigma_b=7, source=8, diag=9,
             upper=10, epsilonPEq=11, sigmaY=12, bEbar=13, DLambda=14, J=15, F=16, gradD_old=17,
             DEpsilonP=18, tractionGradient_b=19, epsilonP=20, DD=21, gradDD=22, rho=23, DD_b=24, sigmaHyd=25, gradSigmaHyd=26, sigmaf=27, gradDf=28)
# (ncomp, 'N' | 'B' | 'F')
FIELD_SHAPE = dict(D=(3, "N"), D_old=(3, "N"), D_oldOld=(3, "N"), gradD=(9, "N"), sigma=(6, "N"), D_b=(3, "B"),
                   gradD_b=(9, "B"), sigma_b=(6, "B"), source=(3, "N"), diag=(3, "N"), upper=(1, "F"),
                   epsilonPEq=(1, "N"), sigmaY=(1, "N"), bEbar=(6, "N"), DLambda=(1, "N"), J=(1, "N"), F=(9, "N"),
                   gradD_old=(9, "N"), DEpsilonP=(6, "N"), tractionGradient_b=(3, "B"), epsilonP=(6, "N"),
                   DD=(3, "N"), gradDD=(9, "N"), rho=(1, "N"), DD_b=(3, "B"), sigmaHyd=(1, "N"), gradSigmaHyd=(3, "N"), sigmaf=(6, "FB"), gradDf=(9, "FB"))

MODEL_NAMES = {
    # reference TypeName -> (gpu TypeName registered by the plugin, enum)
    "linearGeometryTotalDisplacement": MODEL_LIN_GEOM_TOTAL_DISP,
    "nonLinearGeometryTotalLagrangianTotalDisplacement": MODEL_NONLIN_TL_TOTAL_DISP,
    "nonLinearGeometryTotalLagrangian": MODEL_NONLIN_TL,
    "nonLinearGeometryUpdatedLagrangian": MODEL_NONLIN_UL,
    "unsLinearGeometry": MODEL_UNS_LIN_GEOM,
    "unsNonLinearGeometryTotalLagrangian": MODEL_UNS_NONLIN_TL,
    "unsNonLinearGeometryUpdatedLagrangian": MODEL_UNS_NONLIN_UL,
}
INCREMENTAL_MODELS = (MODEL_NONLIN_TL, MODEL_NONLIN_UL, MODEL_UNS_NONLIN_UL)      # the models that solve for DD
MOVING_MESH_MODELS = (MODEL_NONLIN_UL, MODEL_UNS_NONLIN_UL)
LAW_NAMES = {
    "linearElastic": LAW_LINEAR_ELASTIC,
    "neoHookeanElastic": LAW_NEO_HOOKEAN_ELASTIC,
    "neoHookeanElasticMisesPlastic": LAW_NEO_HOOKEAN_MISES_PLASTIC,
    "linearElasticMisesPlastic": LAW_LINEAR_ELASTIC_MISES_PLASTIC,
}


class Law(C.Structure):
    _fields_ = [("kind", C.c_int), ("rho", C.c_double), ("mu", C.c_double), ("K", C.c_double),
                ("lambda_", C.c_double), ("sigma0", C.c_double * 6), ("nTable", C.c_int),
                ("tableEps", C.c_double * 64), ("tableSigY", C.c_double * 64),
                ("updateBEbarConsistent", C.c_int), ("DEpsilonPRelax", C.c_double),
                ("solvePressureEqn", C.c_int), ("pressureSmoothingScaleFactor", C.c_double),
                ("sigmaHydTolerance", C.c_double), ("sigmaHydRelTol", C.c_double), ("sigmaHydMaxIter", C.c_int),
                ("sigmaHydRelax", C.c_double)]


class Controls(C.Structure):
    _fields_ = [("solidModel", C.c_int), ("gradScheme", C.c_int), ("d2dt2Scheme", C.c_int),
                ("stabilisation", C.c_int), ("stabScaleFactor", C.c_double), ("relaxationMethod", C.c_int),
                ("fieldRelaxD", C.c_double), ("solver", C.c_int), ("preconditioner", C.c_int),
                ("tolerance", C.c_double), ("relTol", C.c_double), ("maxIter", C.c_int),
                ("nCorrectors", C.c_int), ("solutionTolerance", C.c_double),
                ("alternativeTolerance", C.c_double), ("materialTolerance", C.c_double),
                ("g", C.c_double * 3), ("deltaT", C.c_double), ("deltaT0", C.c_double),
                ("chebyshevDegree", C.c_int), ("checkEvery", C.c_int),
                ("gamgSinglePrecision", C.c_int), ("gamgOverCorrection", C.c_double),
                ("gamgSmootherDegree", C.c_int), ("gamgCycle", C.c_int), ("gamgSmootherRatio", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("nCorr", C.c_int), ("converged", C.c_int), ("initialResidual", C.c_double * 3),
                ("finalResidual", C.c_double * 3), ("nIterations", C.c_int * 3),
                ("solverPerfInitRes", C.c_double), ("relResidual", C.c_double),
                ("materialResidual", C.c_double), ("totalInnerIterations", C.c_longlong)]

    def as_dict(self):
        return dict(nCorr=self.nCorr, converged=bool(self.converged), initialResidual=list(self.initialResidual),
                    finalResidual=list(self.finalResidual), nIterations=list(self.nIterations),
                    solverPerfInitRes=self.solverPerfInitRes, relResidual=self.relResidual,
                    materialResidual=self.materialResidual, totalInnerIterations=self.totalInnerIterations)


def default_controls(**kw) -> Controls:
    """Defaults of the reference: solidModel.C:1147-1167 (tolerances, nCorrectors), :1307-1315 (RhieChow
    0.1), tutorials' fvSolution (PCG, tolerance 1e-9, relTol 0.1), [OF-ext] maxIter 1000."""
    c = Controls()
    c.solidModel = MODEL_LIN_GEOM_TOTAL_DISP
    c.gradScheme = GRAD_LEAST_SQUARES
    c.d2dt2Scheme = D2DT2_STEADY_STATE
    c.stabilisation = STAB_RHIE_CHOW
    c.stabScaleFactor = 0.1
    c.relaxationMethod = RELAX_FIXED
    c.fieldRelaxD = 1.0
    c.solver = SOLVER_PCG
    c.preconditioner = PRECOND_DIAGONAL
    c.tolerance = 1e-9
    c.relTol = 0.1
    c.maxIter = 1000
    c.nCorrectors = 10000
    c.solutionTolerance = 1e-6
    c.alternativeTolerance = 1e-7
    c.materialTolerance = 1e-5
    c.deltaT = 1.0
    c.deltaT0 = 1.0
    c.chebyshevDegree = 4
    c.checkEvery = 4
    c.gamgSinglePrecision = 0
    c.gamgOverCorrection = 2.2
    c.gamgSmootherDegree = 3
    c.gamgCycle = 2          # K-cycle on level 1 (also on decomposed meshes: its dot products are all-reduced in the kernel)
    c.gamgSmootherRatio = 0.3
    for k, v in kw.items():
        if k == "g":
            for i in range(3):
                c.g[i] = v[i]
        elif k == "solidModel" and isinstance(v, str):
            c.solidModel = MODEL_NAMES[v]
        else:
            if not hasattr(c, k):
                raise KeyError(k)
            setattr(c, k, v)
    return c


def mechanical_law(type: str, rho: float = 0.0, E: Optional[float] = None, nu: Optional[float] = None,
                   mu: Optional[float] = None, K: Optional[float] = None, planeStress: bool = False,
                   sigma0: Optional[Sequence[float]] = None, table: Optional[Sequence[Sequence[float]]] = None,
                   updateBEbarConsistent: bool = True, DEpsilonPRelax: float = 1.0, solvePressureEqn: bool = False,
                   pressureSmoothingScaleFactor: float = 100.0, sigmaHydTolerance: float = 0.0, sigmaHydRelTol: float = 0.0,
                   sigmaHydMaxIter: int = 0, sigmaHydRelax: float = 1.0) -> Law:
    """The mechanicalProperties entry -> POD parameters, with the reference constructors' formulas:
    linearElastic.C:62-133, neoHookeanElastic.C:51-85, neoHookeanElasticMisesPlastic.C:868-930,
    linearElasticMisesPlastic (same E,nu -> mu,K as linearElastic)."""
    kind = LAW_NAMES[type]
    L = Law()
    L.kind = kind
    L.rho = rho
    lam = 0.0
    if solvePressureEqn and planeStress and kind in (LAW_LINEAR_ELASTIC, LAW_LINEAR_ELASTIC_MISES_PLASTIC):
        raise ValueError("planeStress must be 'off' when solvePressureEqn is enabled")       # linearElastic.C:112-119
    if kind in (LAW_LINEAR_ELASTIC, LAW_LINEAR_ELASTIC_MISES_PLASTIC):
        if E is not None and nu is not None:
            if nu < -1.0 or nu > 0.5:
                raise ValueError("Unphysical Poisson's ratio: nu should be >= -1.0 and <= 0.5")
            mu_ = E / (2.0 * (1.0 + nu))
            if nu < 0.5:
                if planeStress:
                    lam = nu * E / ((1.0 + nu) * (1.0 - nu))
                    K_ = E / (3.0 * (1.0 - nu))
                else:
                    lam = nu * E / ((1.0 + nu) * (1.0 - 2.0 * nu))
                    K_ = E / (3.0 * (1.0 - 2.0 * nu))
            else:
                lam = K_ = 1e15   # GREAT
        elif mu is not None and K is not None:
            mu_, K_ = mu, K
            E_ = 9.0 * K_ * mu_ / (3.0 * K_ + mu_)
            nu_ = (3.0 * K_ - 2.0 * mu_) / (2.0 * (3.0 * K_ + mu_))
            if nu_ >= 0.5:
                lam = K_ = 1e15
            elif planeStress:          # linearElastic.C:107-113: lambda AND K are reset for plane stress
                lam = nu_ * E_ / ((1.0 + nu_) * (1.0 - nu_))
                K_ = E_ / (3.0 * (1.0 - nu_))
            else:
                lam = nu_ * E_ / ((1.0 + nu_) * (1.0 - 2.0 * nu_))
        else:
            raise ValueError("Either E and nu or mu and K elastic parameters should be specified")
    else:
        if E is not None and nu is not None and mu is None and K is None:
            mu_ = E / (2.0 * (1.0 + nu))
            if planeStress:
                K_ = (nu * E / ((1.0 + nu) * (1.0 - nu))) + (2.0 / 3.0) * mu_
            else:
                K_ = (nu * E / ((1.0 + nu) * (1.0 - 2.0 * nu))) + (2.0 / 3.0) * mu_
        elif mu is not None and K is not None and E is None and nu is None:
            mu_, K_ = mu, K
        else:
            raise ValueError("Either E and nu or mu and K should be specified")
        lam = K_ - (2.0 / 3.0) * mu_
    L.mu, L.K, L.lambda_ = mu_, K_, lam
    for i in range(6):
        L.sigma0[i] = 0.0 if sigma0 is None else float(sigma0[i])
    L.nTable = 0
    if kind in (LAW_NEO_HOOKEAN_MISES_PLASTIC, LAW_LINEAR_ELASTIC_MISES_PLASTIC):
        if table is None or len(table) < 1:
            raise ValueError("plasticity law needs the (epsilonP sigmaY) table")
        if len(table) > 64:
            raise ValueError("table too long (max 64 points)")
        L.nTable = len(table)
        for i, (e, s) in enumerate(table):
            L.tableEps[i] = e
            L.tableSigY[i] = s
    L.updateBEbarConsistent = 1 if updateBEbarConsistent else 0
    L.DEpsilonPRelax = DEpsilonPRelax
    L.solvePressureEqn = 1 if solvePressureEqn else 0        # mechanicalLaw.C:1525-1532
    L.pressureSmoothingScaleFactor = pressureSmoothingScaleFactor
    L.sigmaHydTolerance, L.sigmaHydRelTol, L.sigmaHydMaxIter, L.sigmaHydRelax = sigmaHydTolerance, sigmaHydRelTol, sigmaHydMaxIter, sigmaHydRelax
    return L


# tutorials/solids/elastoplasticity/neckingBar/constant/plasticStrainVsYieldStress
NECKING_BAR_TABLE = [(0.000, 0.451e9), (0.006, 0.476e9), (0.019, 0.525e9), (0.038, 0.583e9),
                     (0.066, 0.642e9), (0.147, 0.710e9), (0.500, 0.777e9), (1.000, 0.831e9)]


@dataclass
class BC:
    kind: int
    value: Optional[np.ndarray] = None      # [size,3] displacement / traction (None = zero)
    pressure: Optional[np.ndarray] = None   # [size]
    # time series of the reference's patch fields: displacementSeries (fixedDisplacement...C:258-294), tractionSeries /
    # pressureSeries (solidTraction...C:120-170): s4f's interpolationTable, piece-wise linear, outOfBounds clamp.
    # [(t, (x y z))] / [(t, p)]; at(t) gives the BC of that time (the driver calls set_bc with it every time step)
    value_series: Optional[list] = None
    pressure_series: Optional[list] = None

    def at(self, t: float) -> "BC":
        if self.value_series is None and self.pressure_series is None:
            return self
        v, pr = self.value, self.pressure
        if self.value_series is not None:
            v = np.asarray(interpolate_series(self.value_series, t), dtype=np.float64)
        if self.pressure_series is not None:
            pr = np.asarray(float(interpolate_series(self.pressure_series, t)), dtype=np.float64)
        return BC(self.kind, v, pr)


def interpolate_series(series, t: float):
    """numerics/interpolationTable/interpolationTable.C:493-632 with outOfBounds clamp: first / last ordinate outside the
    abscissa range, linear in between."""
    ts = [float(a) for a, _ in series]
    ys = [np.asarray(b, dtype=np.float64) for _, b in series]
    if t <= ts[0]:
        return ys[0]
    if t >= ts[-1]:
        return ys[-1]
    for i in range(len(ts) - 1):
        if ts[i] <= t <= ts[i + 1]:
            w = (t - ts[i]) / (ts[i + 1] - ts[i])
            return (1.0 - w) * ys[i] + w * ys[i + 1]
    return ys[-1]


def fixedDisplacement(value=(0.0, 0.0, 0.0)) -> BC:
    return BC(BC_FIXED_DISPLACEMENT, np.asarray(value, dtype=np.float64))


def solidTraction(traction=(0.0, 0.0, 0.0), pressure=None) -> BC:
    return BC(BC_SOLID_TRACTION, np.asarray(traction, dtype=np.float64),
              None if pressure is None else np.asarray(pressure, dtype=np.float64))


def solidSymmetry() -> BC:
    return BC(BC_SOLID_SYMMETRY)


@dataclass
class SolidCase:
    mesh: FvMesh
    bcs: Dict[str, BC]
    law: Law
    controls: Controls
    name: str = "case"


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def apply_case(lib, prefix: str, handle, case: SolidCase, check) -> None:
    """Mirror mesh, geometry, law, controls and BCs through the C-ABI (``prefix`` = ``s4fgpu_``; the
    tests reuse this plumbing, with another prefix, for their CPU checker's mirror of the same interface)."""
    m = case.mesh
    f = lambda n: getattr(lib, prefix + n)
    own = np.ascontiguousarray(m.owner, dtype=np.int32)
    nei = np.ascontiguousarray(m.neighbour, dtype=np.int32)
    fc = np.ascontiguousarray(m.faceCells, dtype=np.int32)
    nP = len(m.patches)
    pStart = np.array([p.start for p in m.patches], dtype=np.int32)
    pSize = np.array([p.size for p in m.patches], dtype=np.int32)
    pKind = np.array([p.kind for p in m.patches], dtype=np.int32)
    pNbr = np.array([p.nbr_rank for p in m.patches], dtype=np.int32)
    solD = np.ascontiguousarray(m.solutionD, dtype=np.int32)
    check(f("set_mesh")(handle, m.nCells, m.nInternalFaces, _iptr(own), _iptr(nei), nP, _iptr(pStart),
                        _iptr(pSize), _iptr(pKind), _iptr(pNbr), _iptr(fc), _iptr(solD)))
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
            (m.C, m.V, m.Sf, m.magSf, m.Cf, m.weights, m.nonOrthDeltaCoeffs, m.nonOrthCorrVec, m.CnbrB)]
    has_points = m.points is not None and m.faces is not None
    if has_points and m.nRanks > 1:      # decomposed: the points first -- set_geometry lays out the point-neighbour ghosts with the rows
        set_points(lib, prefix, handle, m, check)
    check(f("set_geometry")(handle, *[_dptr(a) for a in arrs]))
    if has_points and m.nRanks == 1:
        set_points(lib, prefix, handle, m, check)
    check(f("set_controls")(handle, C.byref(case.controls)))
    check(f("set_law")(handle, C.byref(case.law)))
    for ip, p in enumerate(m.patches):
        if p.kind == PROCESSOR:
            check(f("set_bc")(handle, ip, BC_PROCESSOR, None, None))
            continue
        bc = case.bcs[p.name]
        val = None
        if bc.value is not None:
            val = np.ascontiguousarray(np.broadcast_to(bc.value, (p.size, 3)), dtype=np.float64)
        pr = None
        if bc.pressure is not None:
            pr = np.ascontiguousarray(np.broadcast_to(bc.pressure, (p.size,)), dtype=np.float64)
        check(f("set_bc")(handle, ip, bc.kind, None if val is None else _dptr(val),
                          None if pr is None else _dptr(pr)))


def set_geometry(lib, prefix: str, handle, m: FvMesh, check) -> None:
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in
            (m.C, m.V, m.Sf, m.magSf, m.Cf, m.weights, m.nonOrthDeltaCoeffs, m.nonOrthCorrVec, m.CnbrB)]
    check(getattr(lib, prefix + "set_geometry")(handle, *[_dptr(a) for a in arrs]))


def set_points(lib, prefix: str, handle, m: FvMesh, check) -> None:
    pts = np.ascontiguousarray(m.points, dtype=np.float64)
    fv = np.ascontiguousarray(m.faces, dtype=np.int32)
    ptr = np.arange(0, fv.size + 1, fv.shape[1], dtype=np.int32)
    check(getattr(lib, prefix + "set_points")(handle, pts.shape[0], _dptr(pts), _iptr(ptr), _iptr(fv)))


def move_mesh(lib, prefix: str, handle, case: "SolidCase", pointDD: np.ndarray, check, mirror: bool = True) -> None:
    """solidModel::moveMesh (SM/solidModel/solidModel.C:2008-2148) on the host side of the boundary: symmetryPlane
    points keep their plane (:2040-2080), newPoints = oldPoints + pointDD, mesh.movePoints(newPoints); the new geometry
    is mirrored again (set_geometry keeps fields, boundary data and history; set_points refreshes the weights)."""
    from . import mesh as M
    m = case.mesh
    pdd = np.array(pointDD, dtype=np.float64, copy=True)
    F = m.nInternalFaces
    for p in m.patches:
        if p.kind != M.SYMMETRY_PLANE or p.size == 0:
            continue
        sl = slice(F + p.start, F + p.start + p.size)
        nrm = m.Sf[sl] / m.magSf[sl, None]
        pn = np.zeros((m.points.shape[0], 3)); cnt = np.zeros(m.points.shape[0])
        for j in range(m.faces.shape[1]):                 # pointNormals: average of the adjacent face normals
            np.add.at(pn, m.faces[sl][:, j], nrm); np.add.at(cnt, m.faces[sl][:, j], 1.0)
        mp = np.nonzero(cnt > 0)[0]
        pnn = pn[mp] / np.linalg.norm(pn[mp], axis=1)[:, None]
        avgN = pnn.mean(axis=0)
        for ax in range(3):
            if abs(avgN[ax]) > 0.95:
                pdd[mp, ax] = 0.0
                break
    case.mesh = M.move_points(m, m.points + pdd)
    if mirror:          # mirror=False: the device has moved its own copy (s4fgpu_move_points); only the host polyMesh follows
        set_geometry(lib, prefix, handle, case.mesh, check)
        set_points(lib, prefix, handle, case.mesh, check)


def field_size(mesh: FvMesh, name: str) -> tuple:
    nc, where = FIELD_SHAPE[name]
    n = dict(N=mesh.nCells, B=mesh.nBoundaryFaces, F=mesh.nInternalFaces, FB=mesh.nInternalFaces + mesh.nBoundaryFaces)[where]
    return (n, nc) if nc > 1 else (n,)


def declare_api(lib, prefix: str, handle_t) -> None:
    """ctypes prototypes of the shared part of the interface."""
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    f = lambda n: getattr(lib, prefix + n)
    f("set_mesh").argtypes = [handle_t, C.c_int, C.c_int, ip, ip, C.c_int, ip, ip, ip, ip, ip, ip]
    f("set_geometry").argtypes = [handle_t] + [dp] * 9
    f("set_law").argtypes = [handle_t, C.POINTER(Law)]
    f("set_controls").argtypes = [handle_t, C.POINTER(Controls)]
    f("set_bc").argtypes = [handle_t, C.c_int, C.c_int, dp, dp]
    f("upload").argtypes = [handle_t, C.c_int, dp]
    f("download").argtypes = [handle_t, C.c_int, dp]
    f("initialise").argtypes = [handle_t]
    f("new_timestep").argtypes = [handle_t, C.c_double]
    f("outer_iteration").argtypes = [handle_t, C.POINTER(Stats)]
    f("evolve").argtypes = [handle_t, C.POINTER(Stats)]
    f("update_total_fields").argtypes = [handle_t]
    f("op_grad").argtypes = [handle_t]
    f("op_correct").argtypes = [handle_t]
    f("op_assemble").argtypes = [handle_t]
    f("op_amul").argtypes = [handle_t, C.c_int, dp, dp]
    f("op_solve").argtypes = [handle_t, dp, dp, C.POINTER(Stats)]
    f("set_points").argtypes = [handle_t, C.c_int, dp, ip, ip]
    f("interpolate_to_points").argtypes = [handle_t, C.c_int, C.c_int, dp]
    for n in ("set_mesh", "set_geometry", "set_law", "set_controls", "set_bc", "upload", "download", "initialise",
              "new_timestep", "outer_iteration", "evolve", "update_total_fields", "op_grad", "op_correct",
              "op_assemble", "op_amul", "op_solve", "set_points", "interpolate_to_points"):
        f(n).restype = C.c_int
