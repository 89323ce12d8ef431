"""Synthetic cases of BASELINE.json / SURVEY.md section 8d, expressed as :class:`SolidCase`.

C1 plateHole      tutorials/solids/linearElasticity/plateHole (0/D, mechanicalProperties, fvSolution)
C2 cantilever     synthetic 3-D hex cantilever, linearElastic, 8 x 1 x 1 m
C3 neo-Hookean    same beam, neoHookeanElastic, nonLinearGeometryTotalLagrangianTotalDisplacement
C4 notched bar    neoHookeanElasticMisesPlastic, cosine notch by point scaling
patch test        tutorials/solids/linearElasticity/patchTest (linear displacement on distorted cells)
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import case as K
from . import mesh as M


# ---- C2: hex cantilever --------------------------------------------------------------------------
def cantilever(nx: int, ny: int, nz: int, rank: int = 0, nRanks: int = 1, L: float = 8.0, H: float = 1.0,
               W: float = 1.0, traction=(0.0, -1e6, 0.0), E: float = 200e9, nu: float = 0.3,
               rho: float = 7800.0, general: bool = False, **ctl) -> K.SolidCase:
    """x=0 fixedDisplacement (0 0 0); x=L solidTraction (0 -1e6 0); other faces traction free;
    linearElastic E 200e9 nu 0.3; steadyState; leastSquares; RhieChow 0.1; PCG relTol 0.1."""
    names = ("fixed", "loaded", "yMin", "yMax", "zMin", "zMax")
    if nRanks > 1:
        mesh = M.hex_box_decomposed(nx, ny, nz, L, H, W, rank, nRanks, names=names)
    elif general:      # general builder: keeps points()/faces() (vol->point interpolation, mesh motion)
        mesh = M.hex_box_general(nx, ny, nz, L, H, W, names=names)
    else:
        mesh = M.hex_box(nx, ny, nz, L, H, W, names=names)
    bcs = {}
    for p in mesh.patches:
        if p.kind == M.PROCESSOR:
            continue
        if p.name == "fixed":
            bcs[p.name] = K.fixedDisplacement((0.0, 0.0, 0.0))
        elif p.name == "loaded":
            bcs[p.name] = K.solidTraction(traction)
        else:
            bcs[p.name] = K.solidTraction((0.0, 0.0, 0.0))
    law = K.mechanical_law("linearElastic", rho=rho, E=E, nu=nu)
    return K.SolidCase(mesh, bcs, law, K.default_controls(**ctl), name=f"cantilever_{nx}x{ny}x{nz}")


# ---- C1: plate with a hole -------------------------------------------------------------------------
def kirsch_stress(C: np.ndarray, T: float = 1e6, a: float = 0.5) -> np.ndarray:
    """Cartesian Kirsch stress (symmTensor order) at points C: polar components rotated to x-y,
    analyticalPlateHoleTractionFvPatchVectorField.C:36-88."""
    x, y = C[:, 0], C[:, 1]
    r = np.hypot(x, y)
    th = np.arctan2(y, x)
    srr = T * (1 - a**2 / r**2) / 2 + T * (1 + 3 * a**4 / r**4 - 4 * a**2 / r**2) * np.cos(2 * th) / 2
    srt = -T * (1 - 3 * a**4 / r**4 + 2 * a**2 / r**2) * np.sin(2 * th) / 2
    stt = T * (1 + a**2 / r**2) / 2 - T * (1 + 3 * a**4 / r**4) * np.cos(2 * th) / 2
    c, s = np.cos(th), np.sin(th)
    sxx = c * c * srr - 2 * s * c * srt + s * s * stt
    syy = s * s * srr + 2 * s * c * srt + c * c * stt
    sxy = s * c * (srr - stt) + (c * c - s * s) * srt
    out = np.zeros((C.shape[0], 6))
    out[:, 0], out[:, 1], out[:, 3] = sxx, sxy, syy
    return out


def kirsch_displacement(C: np.ndarray, T: float = 1e6, a: float = 0.5, E: float = 200e9, nu: float = 0.3) -> np.ndarray:
    """plateHoleAnalyticalSolution.C:91-122 (plane strain kappa = 3 - 4 nu)."""
    mu = E / (2 * (1 + nu))
    kappa = 3 - 4 * nu
    r = np.hypot(C[:, 0], C[:, 1])
    th = np.arctan2(C[:, 1], C[:, 0])
    f = a * T / (8 * mu)
    ux = f * ((r / a) * (kappa + 1) * np.cos(th) + (2 * a / r) * ((1 + kappa) * np.cos(th) + np.cos(3 * th))
              - (2 * a**3 / r**3) * np.cos(3 * th))
    uy = f * ((r / a) * (kappa - 3) * np.sin(th) + (2 * a / r) * ((1 - kappa) * np.sin(th) + np.sin(3 * th))
              - (2 * a**3 / r**3) * np.sin(3 * th))
    return np.stack([ux, uy, np.zeros_like(ux)], axis=1)


def plate_hole(refine: int = 1, cell_perm_seed: Optional[int] = None, T: float = 1e6, **ctl) -> K.SolidCase:
    """left/down solidSymmetry, right/up analyticalPlateHoleTraction (traction = n & sigma_Kirsch(Cf),
    ...FvPatchVectorField.C:176-210), hole traction free; E 200e9 nu 0.3 plane strain; D relax 0.7."""
    mesh = M.plate_hole(refine=refine, cell_perm_seed=cell_perm_seed)
    F = mesh.nInternalFaces
    nrm = mesh.boundary_normals()
    bcs = {"left": K.solidSymmetry(), "down": K.solidSymmetry(), "hole": K.solidTraction((0.0, 0.0, 0.0))}
    for name in ("right", "up"):
        sl = mesh.patch_slice(name)
        cf = mesh.Cf[F:][sl].copy()
        cf[:, 2] = 0.0
        sg = kirsch_stress(cf, T)
        n = nrm[sl]
        t = np.stack([n[:, 0] * sg[:, 0] + n[:, 1] * sg[:, 1], n[:, 0] * sg[:, 1] + n[:, 1] * sg[:, 3],
                      np.zeros(n.shape[0])], axis=1)
        bcs[name] = K.solidTraction(t)
    law = K.mechanical_law("linearElastic", rho=7854.0, E=200e9, nu=0.3)
    c = dict(fieldRelaxD=0.7, nCorrectors=1000)
    c.update(ctl)
    return K.SolidCase(mesh, bcs, law, K.default_controls(**c), name="plateHole")


# ---- patch test --------------------------------------------------------------------------------
PATCH_TEST_A = np.array([1e-6, 4e-6, 0.0])
PATCH_TEST_B = np.array([[2e-6, 3e-6, 0.0], [5e-6, 6e-6, 0.0], [0.0, 0.0, 0.0]])   # D_i = a_i + B_ij x_j


def patch_test(n: int = 4, distort: float = 0.25, seed: int = 7, **ctl) -> K.SolidCase:
    """Linear displacement d = a + B x prescribed on all sides of a distorted 2-D quad mesh
    (tutorials/solids/linearElasticity/patchTest/README.md:40-62): strain must be exactly constant."""
    rng = np.random.default_rng(seed)

    def pmap(p):
        q = p.copy()
        h = 1.0 / n
        interior = (p[:, 0] > 1e-9) & (p[:, 0] < 1 - 1e-9) & (p[:, 1] > 1e-9) & (p[:, 1] < 1 - 1e-9)
        # same in-plane shift on the two z-layers
        key = np.round(p[:, :2] / h).astype(np.int64)
        shift = rng.uniform(-distort * h, distort * h, size=(n + 1, n + 1, 2))
        q[:, :2] += np.where(interior[:, None], shift[key[:, 0], key[:, 1]], 0.0)
        # skew the whole patch so that boundary faces are non-orthogonal too
        q[:, 0] += 0.2 * q[:, 1]
        return q
    names = ("left", "right", "bottom", "top", "front", "back")
    kinds = (M.PATCH, M.PATCH, M.PATCH, M.PATCH, M.EMPTY, M.EMPTY)
    mesh = M.hex_box_general(n, n, 1, 1.0, 1.0, 0.1, point_map=pmap, names=names, kinds=kinds)
    F = mesh.nInternalFaces
    bcs = {}
    for p in mesh.patches:
        cf = mesh.Cf[F:][p.start:p.start + p.size]
        bcs[p.name] = K.fixedDisplacement(PATCH_TEST_A[None, :] + cf @ PATCH_TEST_B.T)
    law = K.mechanical_law("linearElastic", rho=7854.0, E=200e9, nu=0.3)
    c = dict(tolerance=1e-15, relTol=0.0, solutionTolerance=1e-14, alternativeTolerance=1e-14, nCorrectors=2000)
    c.update(ctl)
    return K.SolidCase(mesh, bcs, law, K.default_controls(**c), name="patchTest")


# ---- C3 / C4 -----------------------------------------------------------------------------------
def neo_hookean_cantilever(nx, ny, nz, traction=(0.0, -50.0, 0.0), E=3e6, nu=0.3, rho=1000.0,
                           L=8.0, H=1.0, W=1.0, general=False, **ctl) -> K.SolidCase:
    """C3: tutorials/solids/hyperelasticity/cantileverBeam material, total-Lagrangian total displacement."""
    base = cantilever(nx, ny, nz, L=L, H=H, W=W, traction=traction, general=general)
    law = K.mechanical_law("neoHookeanElastic", rho=rho, E=E, nu=nu)
    c = dict(solidModel=K.MODEL_NONLIN_TL_TOTAL_DISP)
    c.update(ctl)
    return K.SolidCase(base.mesh, base.bcs, law, K.default_controls(**c), name=f"neoHookeanCantilever_{nx}x{ny}x{nz}")


def notch_map(L=8.0, H=1.0, W=1.0, depth=0.2, width=0.25):
    def pmap(p):
        q = p.copy()
        s = 1.0 - depth * np.exp(-((p[:, 0] - 0.5 * L) / width) ** 2)
        q[:, 1] = 0.5 * H + (p[:, 1] - 0.5 * H) * s
        q[:, 2] = 0.5 * W + (p[:, 2] - 0.5 * W) * s
        return q
    return pmap


def notched_bar(nx, ny, nz, elongation=0.0016, L=8.0, H=1.0, W=1.0, rank=0, nRanks=1, **ctl) -> K.SolidCase:
    """C4: neoHookeanElasticMisesPlastic E 200e9 nu 0.3 + neckingBar table; x=0 fixed, x=L displaced."""
    names = ("fixed", "pulled", "yMin", "yMax", "zMin", "zMax")
    if nRanks > 1:
        mesh = M.hex_box_decomposed(nx, ny, nz, L, H, W, rank, nRanks, names=names, point_map=notch_map(L, H, W))
    else:
        mesh = M.hex_box(nx, ny, nz, L, H, W, names=names, point_map=notch_map(L, H, W))
    bcs = {}
    for p in mesh.patches:
        if p.kind == M.PROCESSOR:
            continue
        if p.name == "fixed":
            bcs[p.name] = K.fixedDisplacement((0.0, 0.0, 0.0))
        elif p.name == "pulled":
            bcs[p.name] = K.fixedDisplacement((elongation * L, 0.0, 0.0))
        else:
            bcs[p.name] = K.solidTraction((0.0, 0.0, 0.0))
    law = K.mechanical_law("neoHookeanElasticMisesPlastic", rho=7833.0, E=200e9, nu=0.3, table=K.NECKING_BAR_TABLE)
    c = dict(solidModel=K.MODEL_NONLIN_TL_TOTAL_DISP)
    c.update(ctl)
    return K.SolidCase(mesh, bcs, law, K.default_controls(**c), name=f"notchedBar_{nx}x{ny}x{nz}")


# ---- C5: solid side of fluidSolidInteraction/beamInCrossFlow -----------------------------------------
def beam_in_cross_flow(refine: int = 1, pressure: float = 50.0, deltaT: float = 0.1, **ctl) -> K.SolidCase:
    """C5: tutorials/fluidSolidInteraction/beamInCrossFlow/constant/solid/polyMesh/blockMeshDict
    (0.1 x 0.2 x 0.2 m beam standing on y = 0, 4 x 8 x 8 cells x refine), neoHookeanElastic E 1e4 nu 0.4 rho 1000
    (constant/solid/mechanicalProperties), nonLinearGeometryUpdatedLagrangian, backward d2dt2, leastSquares
    (system/solid/fvSchemes), relaxation 0.9 (fvSolution); bottom fixed, z = 0 symmetry plane, the other faces are
    the FSI interface, here loaded by a PRESCRIBED uniform pressure on the upstream (x-) face (stand-in for the
    fluid load; the caller ramps it)."""
    nx, ny, nz = 4 * refine, 8 * refine, 8 * refine
    names = ("upstream", "downstream", "bottom", "top", "side", "symmetry")      # symmetry plane at zMax, as in the tutorial
    kinds = (M.PATCH, M.PATCH, M.PATCH, M.PATCH, M.PATCH, M.SYMMETRY_PLANE)
    mesh = M.hex_box_general(nx, ny, nz, 0.1, 0.2, 0.2, names=names, kinds=kinds)
    bcs = {}
    for p in mesh.patches:
        if p.name == "bottom":
            bcs[p.name] = K.fixedDisplacement((0.0, 0.0, 0.0))
        elif p.name == "symmetry":
            bcs[p.name] = K.solidSymmetry()
        elif p.name == "upstream":
            bcs[p.name] = K.solidTraction((0.0, 0.0, 0.0), pressure=np.full(p.size, pressure))
        else:
            bcs[p.name] = K.solidTraction((0.0, 0.0, 0.0))
    law = K.mechanical_law("neoHookeanElastic", rho=1000.0, E=1e4, nu=0.4)
    c = dict(solidModel=K.MODEL_NONLIN_UL, d2dt2Scheme=K.D2DT2_BACKWARD, deltaT=deltaT, deltaT0=deltaT, fieldRelaxD=0.9,
             nCorrectors=1000)
    c.update(ctl)
    return K.SolidCase(mesh, bcs, law, K.default_controls(**c), name=f"beamInCrossFlow_solid_x{refine}")
