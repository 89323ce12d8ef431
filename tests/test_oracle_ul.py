"""CPU tests of the oracle's updated-Lagrangian path: nonLinGeomUpdatedLagSolid
(SM/nonLinGeomUpdatedLagSolid/nonLinGeomUpdatedLagSolid.C:159-273, :360-374), the vol->point interpolation that moves the
mesh (enhancedVolPointInterpolation, solidModel::moveMesh SM/solidModel/solidModel.C:2008-2148) and the rho-field inertia
terms (NUM/backwardD2dt2Scheme/backwardD2dt2Scheme.C:149-222, :391-470).

The reference ships no golden fields for this model; the pins are physical identities: exactness for a homogeneous
deformation, agreement with the total-Lagrangian formulations of the same problem, mass conservation, and the small-load
limit of the dynamics.
"""
import numpy as np

from oracle.binding import OracleSolid
from solids4foam_b200 import case as K
from solids4foam_b200 import cases
from solids4foam_b200 import mesh as M
from s4f_testutil import rel_l2

NAMES = ("xMin", "xMax", "yMin", "yMax", "zMin", "zMax")
TIGHT = dict(solutionTolerance=1e-10, alternativeTolerance=1e-10, tolerance=1e-14, relTol=1e-3, nCorrectors=3000,
             preconditioner=K.PRECOND_DIC)


def test_move_points_affine_map_scales_geometry():
    mesh = M.hex_box_general(3, 4, 5, 1.0, 2.0, 1.5, names=NAMES)
    A = np.eye(3) + np.array([[0.1, 0.05, 0.0], [0.0, -0.08, 0.02], [0.03, 0.0, 0.06]])
    moved = M.move_points(mesh, mesh.points @ A.T)
    assert np.allclose(moved.V, np.linalg.det(A) * mesh.V, rtol=1e-12)
    assert np.allclose(moved.C, mesh.C @ A.T, atol=1e-13)
    assert np.allclose(moved.Cf, mesh.Cf @ A.T, atol=1e-13)
    # Nanson: Sf' = det(A) A^-T Sf
    assert np.allclose(moved.Sf, np.linalg.det(A) * mesh.Sf @ np.linalg.inv(A), atol=1e-13)
    assert moved.owner is mesh.owner or np.array_equal(moved.owner, mesh.owner)


def test_vol_to_point_interpolation_of_a_linear_field():
    """Inverse-distance weights reproduce a linear field at points whose donors are placed symmetrically: internal points
    of a uniform hex mesh (8 cells) and points inside a flat patch (4 faces); at patch edges the donors are one-sided and
    the value is that of the donors' centroid (enhancedVolPointInterpolation.C:165-245)."""
    case = cases.neo_hookean_cantilever(6, 4, 4, general=True, solidModel=K.MODEL_NONLIN_TL_TOTAL_DISP, L=3.0, H=2.0, W=2.0)
    o = OracleSolid(case)
    m = case.mesh
    G = np.array([[0.01, 0.02, -0.01], [0.0, 0.03, 0.01], [0.02, -0.01, 0.005]])
    lin = lambda X: 0.1 + X @ G.T
    F = m.nInternalFaces
    o.set("D", lin(m.C)); o.set("D_b", lin(m.Cf[F:]))
    pD = o.interpolate_to_points("D")
    x, y, z = m.points.T
    on = lambda a, hi: (np.abs(a) < 1e-12) | (np.abs(a - hi) < 1e-12)
    nb = on(x, 3.0).astype(int) + on(y, 2.0).astype(int) + on(z, 2.0).astype(int)
    exact = lin(m.points)
    assert np.abs(pD - exact)[nb == 0].max() < 1e-14          # internal points
    assert np.abs(pD - exact)[nb == 1].max() < 1e-14          # inside a patch
    assert np.abs(pD - exact)[nb >= 2].max() > 1e-4           # edges/corners: one-sided donors
    # an edge point along x between yMin and zMin sees 2 faces of each patch: the donors' centroid is (x, h/4, h/4)
    h = 0.5
    e = np.nonzero((nb == 2) & on(y, 2.0) & on(z, 2.0) & (np.abs(y) < 1e-12) & (np.abs(z) < 1e-12))[0][0]
    assert np.allclose(pD[e], lin(m.points[e] + np.array([0.0, h / 4, h / 4])), atol=1e-14)


def test_gradient_extrapolated_vol_to_point_interpolation_is_exact_for_linear_fields():
    """interpolate(vf, gradVf, pf) (enhancedVolPointInterpolate.C:351-418): every donor is extrapolated to the point with its
    own gradient, so a linear field is reproduced at ALL points, one-sided boundary points included, on a distorted mesh."""
    case = cases.patch_test(n=5)
    if case.mesh.points is None:
        case = cases.neo_hookean_cantilever(6, 4, 4, general=True, solidModel=K.MODEL_NONLIN_TL_TOTAL_DISP, L=3.0, H=2.0, W=2.0)
    o = OracleSolid(case)
    m = case.mesh
    G = np.array([[0.01, 0.02, -0.01], [0.0, 0.03, 0.01], [0.02, -0.01, 0.005]])
    if m.solutionD[2] == 0:
        G[2, :] = 0; G[:, 2] = 0
    lin = lambda X: 0.1 + X @ G.T
    F = m.nInternalFaces
    o.set("D", lin(m.C)); o.set("D_b", lin(m.Cf[F:]))
    o.set("gradD", np.tile(G.T.reshape(1, 9), (m.nCells, 1)))       # gradD_ij = d_i D_j
    pD = o.interpolate_to_points("D", with_gradient=True)
    assert np.abs(pD - lin(m.points)).max() < 1e-15 + 1e-13 * np.abs(pD).max()


def test_homogeneous_deformation_in_two_updated_lagrangian_steps():
    """A prescribed affine motion x = A X in two increments: the first increment is exact (least-squares gradient, Gauss
    divergence of a constant flux tensor and the Rhie-Chow term all vanish identically for a linear field), F = A1; the
    second works on the moved mesh, F = relF & F.old = A2 A1 up to the one-sided vol->point weights at the patch edges;
    rho J stays rho0 (updateTotalFields :362)."""
    mesh = M.hex_box_general(4, 4, 4, 1, 1, 1, names=NAMES)
    A1 = np.eye(3) + np.array([[0.04, 0.02, 0], [0.0, -0.01, 0.01], [0.005, 0, 0.02]])
    A2 = np.eye(3) + np.array([[0.03, -0.01, 0.01], [0.01, 0.02, 0.0], [0.0, 0.01, -0.02]])
    law = K.mechanical_law("neoHookeanElastic", rho=1000.0, E=3e6, nu=0.3)
    ctl = K.default_controls(solidModel=K.MODEL_NONLIN_UL, **TIGHT)
    F0 = mesh.nInternalFaces
    Xb = {p.name: mesh.Cf[F0 + p.start:F0 + p.start + p.size].copy() for p in mesh.patches}
    bcs = {n: K.fixedDisplacement(np.zeros((len(Xb[n]), 3))) for n in NAMES}
    c = K.SolidCase(mesh, bcs, law, ctl)
    o = OracleSolid(c)
    V0 = mesh.V.sum()
    for step, A in enumerate((A1, A2 @ A1)):
        o.new_timestep(1.0)
        for n in NAMES:
            o.set_bc(n, K.fixedDisplacement(Xb[n] @ (A - np.eye(3)).T))
        st = o.evolve()
        assert st["converged"]
        err = np.abs(o.get("F").reshape(-1, 3, 3) - A).max()
        assert err < (1e-10 if step == 0 else 5e-4), err
        # neoHookeanElastic.C:275-303 closed form of the Cauchy stress for F = A
        J = np.linalg.det(A); b = A @ A.T * J ** (-2.0 / 3.0)
        s = (law.mu * (b - np.trace(b) / 3 * np.eye(3)) + 0.5 * law.K * (J * J - 1) * np.eye(3)) / J
        sig = o.get("sigma")
        ref = np.array([s[0, 0], s[0, 1], s[0, 2], s[1, 1], s[1, 2], s[2, 2]])
        assert np.abs(sig - ref).max() / np.abs(ref).max() < (1e-9 if step == 0 else 2e-2)
        o.update_total_fields()
        assert np.allclose(o.get("rho") * o.get("J"), 1000.0, rtol=1e-12)
        assert abs(c.mesh.V.sum() / V0 - J) < 5e-3


def test_updated_and_total_lagrangian_reach_the_same_equilibria():
    """nonLinearGeometryUpdatedLagrangian (mesh moved every step, relF from grad(DD) on the moved mesh) against
    nonLinearGeometryTotalLagrangian (DD on the reference mesh): the first step is the same discrete problem; later steps
    differ by discretisation error only."""
    res = {}
    for model in (K.MODEL_NONLIN_TL, K.MODEL_NONLIN_UL):
        c = cases.neo_hookean_cantilever(12, 3, 3, traction=(0, 0, 0), solidModel=model, general=True, L=4.0,
                                         **dict(TIGHT, solutionTolerance=1e-8, alternativeTolerance=1e-8, nCorrectors=20000))
        o = OracleSolid(c)
        n = c.mesh.patch("loaded").size
        steps = []
        for t in (-4e3, -8e3, -12e3):
            tr = np.zeros((n, 3)); tr[:, 1] = t
            o.new_timestep(1.0)
            o.set_bc("loaded", K.solidTraction(tr))
            st = o.evolve()
            assert st["converged"], st
            o.update_total_fields()
            steps.append((o.get("D"), o.get("sigma")))
        res[model] = steps
    tl, ul = res[K.MODEL_NONLIN_TL], res[K.MODEL_NONLIN_UL]
    assert rel_l2(ul[0][0], tl[0][0]) < 1e-12 and rel_l2(ul[0][1], tl[0][1]) < 1e-12
    assert np.abs(tl[2][0]).max() > 0.2           # a finite deflection (5 % of the span)
    assert rel_l2(ul[2][0], tl[2][0]) < 5e-3
    assert rel_l2(ul[2][1], tl[2][1]) < 2e-2


def test_updated_lagrangian_dynamics_small_load_limit():
    """fvm::d2dt2(rho, DD) + fvc::d2dt2(rho, D.oldTime()) with the density field against rho*fvm::d2dt2(D) of the
    total-displacement model: for a load small enough that relJ = 1 and the mesh motion is negligible the two
    discretise the same equation, except that the GREAT-deltaT0 start-up of the explicit chain lasts one step longer
    (backwardD2dt2Scheme.C:48-68); with the same start-up in both the two agree to solver tolerance (checked while writing
    the oracle).  Euler has no start-up: agreement to solver tolerance.  Rhie-Chow is off: it smooths D in one model and
    DD in the other."""
    for scheme, tol in ((K.D2DT2_EULER, 1e-6), (K.D2DT2_BACKWARD, 0.2)):
        out = {}
        for model in (K.MODEL_NONLIN_TL_TOTAL_DISP, K.MODEL_NONLIN_UL):
            c = cases.neo_hookean_cantilever(10, 3, 3, traction=(0, -0.5, 0), solidModel=model, general=True, L=4.0,
                                             d2dt2Scheme=scheme, deltaT=2e-3, deltaT0=2e-3, stabilisation=K.STAB_NONE,
                                             **dict(TIGHT, solutionTolerance=1e-8, alternativeTolerance=1e-8))
            o = OracleSolid(c)
            hist = []
            for _ in range(5):
                o.new_timestep(2e-3)
                st = o.evolve()
                assert st["converged"]
                o.update_total_fields()
                hist.append(o.get("D"))
            out[model] = hist
        a, b = out[K.MODEL_NONLIN_TL_TOTAL_DISP], out[K.MODEL_NONLIN_UL]
        assert np.abs(a[-1]).max() > 0
        assert rel_l2(b[0], a[0]) < 1e-6            # first step: same start-up coefficients
        assert rel_l2(b[-1], a[-1]) < tol, (scheme, rel_l2(b[-1], a[-1]))
        # the beam is still accelerating (period >> 5 dt): the tip moves monotonically
        tip = [np.abs(h[:, 1]).max() for h in b]
        assert all(t1 > t0 for t0, t1 in zip(tip, tip[1:]))


def test_beam_in_cross_flow_solid_side_runs_and_conserves_mass():
    """C5: solid side of tutorials/fluidSolidInteraction/beamInCrossFlow with a prescribed pressure ramp."""
    c = cases.beam_in_cross_flow(refine=1, solutionTolerance=1e-7, alternativeTolerance=1e-7, preconditioner=K.PRECOND_DIC)
    o = OracleSolid(c)
    m0 = (c.mesh.V * 1000.0).sum()
    npatch = c.mesh.patch("upstream").size
    for step in range(1, 4):
        o.new_timestep(0.1)
        o.set_bc("upstream", K.solidTraction(np.zeros((npatch, 3)), pressure=np.full(npatch, 50.0 * min(0.1 * step, 1.0))))
        st = o.evolve()
        assert st["converged"], st
        o.update_total_fields()
    D = o.get("D")
    assert D[:, 0].max() > 1e-5 and D[:, 0].max() > 10 * abs(D[:, 0].min())        # pushed downstream
    pD = o.interpolate_to_points("D")
    sym = np.abs(c.mesh.points[:, 2] - c.mesh.points[:, 2].max()) < 1e-9
    assert np.abs(pD[sym, 2]).max() < 1e-15                                          # points stay in the symmetry plane
    assert np.abs(c.mesh.points[sym, 2] - 0.2).max() < 1e-12
    assert abs((c.mesh.V * o.get("rho")).sum() / m0 - 1) < 1e-3                       # rho V = rho0 V0 up to the vol->point error


# ---------------------------------------------------------------------------------------------
# pressure smoothing: mechanicalLaw::updateSigmaHyd with solvePressureEqn
# ---------------------------------------------------------------------------------------------
def _beam_with_pressure_eqn(n, solve, law="linearElastic", model=K.MODEL_LIN_GEOM_TOTAL_DISP, **kw):
    tight = dict(solutionTolerance=1e-9, alternativeTolerance=1e-9, tolerance=1e-13, relTol=1e-3, nCorrectors=20000,
                 preconditioner=K.PRECOND_DIC, solidModel=model)
    tight.update(kw)
    c = cases.cantilever(2 * n, n, n, L=2.0, **tight)
    if law == "linearElastic":
        c.law = K.mechanical_law(law, rho=7800.0, E=200e9, nu=0.3, solvePressureEqn=solve)
    else:
        c.law = K.mechanical_law(law, rho=7800.0, E=200e9, nu=0.3, solvePressureEqn=solve)
    return c


def test_pressure_equation_is_a_consistent_smoothing():
    """mechanicalLaw.C:1442-1453: 'the fvm and fvc laplacian terms cancel at convergence and the laplacian - div(grad) term
    produce a smoothing/diffusion': the converged fields with and without solvePressureEqn differ by a term that vanishes
    under mesh refinement, tr(sigma)/3 is the solved sigmaHyd, and the hydrostatic stress field is smoother."""
    diff = {}
    for n in (4, 8):
        res = {}
        for solve in (False, True):
            o = OracleSolid(_beam_with_pressure_eqn(n, solve))
            st = o.evolve()
            assert st["converged"]
            res[solve] = o
        sa, sb = res[False].get("sigma"), res[True].get("sigma")
        tr = lambda s: (s[:, 0] + s[:, 3] + s[:, 5]) / 3.0
        assert np.allclose(tr(sb), res[True].get("sigmaHyd"), rtol=0, atol=1e-9 * np.abs(sb).max())
        rough = lambda s: np.std(np.diff(tr(s).reshape(n, n, 2 * n), axis=2))
        assert rough(sb) < rough(sa)
        diff[n] = rel_l2(res[True].get("D"), res[False].get("D"))
    assert diff[8] < 0.25 * diff[4] and diff[8] < 0.05, diff


def test_pressure_equation_with_zero_scale_factor_returns_the_explicit_cell_values():
    """pressureSmoothingScaleFactor 0: rDAf = 0 and the equation degenerates to sigmaHyd = sigmaHydExplicit = K tr(epsilon) in
    the cells (linearElastic.C:337); the patch values are zeroGradient (mechanicalLaw.C:452-458), not K tr(epsilon_b)."""
    c = _beam_with_pressure_eqn(4, True)
    c.law.pressureSmoothingScaleFactor = 0.0
    o = OracleSolid(c)
    assert o.evolve()["converged"]
    g = o.get("gradD")
    tr_eps = g[:, 0] + g[:, 4] + g[:, 8]
    assert np.allclose(o.get("sigmaHyd"), c.law.K * tr_eps, rtol=1e-8, atol=1e-9 * np.abs(c.law.K * tr_eps).max())
    sb = o.get("sigma_b")
    trb = (sb[:, 0] + sb[:, 3] + sb[:, 5]) / 3.0
    assert np.allclose(trb, o.get("sigmaHyd")[c.mesh.faceCells], rtol=1e-9, atol=1e-9 * np.abs(trb).max())


# ---------------------------------------------------------------------------------------------
# pointCellsLeastSquares gradient (first half)
# ---------------------------------------------------------------------------------------------
def test_point_cells_least_squares_is_exact_for_linear_fields_and_passes_the_patch_test():
    """[OF-ext] LeastSquaresGrad over the cell-point-cell stencil: exact for linear fields on distorted meshes (2-D with an
    empty direction and 3-D); with it the patch test still returns the constant strain of patchTest/README.md:102-113."""
    G = np.array([[0.01, 0.02, -0.01], [0.03, -0.01, 0.02], [0.005, 0.0, 0.01]])
    pmap = lambda p: p + 0.04 * np.sin(3.0 * p[:, [1, 2, 0]])
    c3 = cases.neo_hookean_cantilever(5, 4, 3, general=True, L=2.0, solidModel=K.MODEL_NONLIN_TL_TOTAL_DISP,
                                      gradScheme=K.GRAD_POINT_CELLS_LEAST_SQUARES)
    c3.mesh = M.hex_box_general(5, 4, 3, 2.0, 1.0, 1.0, names=("fixed", "loaded", "yMin", "yMax", "zMin", "zMax"), point_map=pmap)
    for c, g in ((cases.patch_test(n=5, gradScheme=K.GRAD_POINT_CELLS_LEAST_SQUARES), G * np.array([[1, 1, 0], [1, 1, 0], [0, 0, 0]])), (c3, G)):
        o = OracleSolid(c)
        m = c.mesh
        F = m.nInternalFaces
        lin = lambda X: 0.1 + X @ g.T
        o.set("D", lin(m.C)); o.set("D_b", lin(m.Cf[F:]))
        o.op_grad()
        assert np.abs(o.get("gradD").reshape(-1, 3, 3) - g.T).max() < 1e-15
    o = OracleSolid(cases.patch_test(n=4, gradScheme=K.GRAD_POINT_CELLS_LEAST_SQUARES))
    assert o.evolve()["converged"]
    gD = o.get("gradD").reshape(-1, 3, 3)
    eps = 0.5 * (gD + gD.transpose(0, 2, 1))
    assert np.abs(eps[:, 0, 0] - 2e-6).max() < 1e-13 and np.abs(eps[:, 1, 1] - 6e-6).max() < 1e-13 and np.abs(eps[:, 0, 1] - 4e-6).max() < 1e-13


def test_point_cells_stencil_size_and_plate_hole_accuracy():
    """An interior hex cell sees 26 point neighbours; on the plate-hole case (C1) the scheme reaches the Kirsch solution at the
    discretisation level the face-neighbour scheme does."""
    c = cases.neo_hookean_cantilever(5, 5, 5, general=True, L=1.0, solidModel=K.MODEL_NONLIN_TL_TOTAL_DISP, gradScheme=K.GRAD_POINT_CELLS_LEAST_SQUARES)
    o = OracleSolid(c)
    # stencil size through the gradient of an indicator (the boundary values stay as they are: difference of two gradients)
    centre = int(np.argmin(np.linalg.norm(c.mesh.C - 0.5, axis=1)))
    D = np.zeros((c.mesh.nCells, 3))
    o.set("D", D); o.op_grad(); g0 = o.get("gradD")
    D[centre, 0] = 1.0
    o.set("D", D); o.op_grad()
    touched = np.nonzero(np.abs(o.get("gradD") - g0).sum(axis=1) > 0)[0]
    assert len(touched) == 27                                   # the cell itself and its 26 point neighbours
    errs = {}
    for scheme in (K.GRAD_LEAST_SQUARES, K.GRAD_POINT_CELLS_LEAST_SQUARES):
        cp = cases.plate_hole(gradScheme=scheme)
        op = OracleSolid(cp)
        assert op.evolve()["converged"]
        sa = cases.kirsch_stress(cp.mesh.C)
        errs[scheme] = rel_l2(op.get("sigma")[:, [0, 1, 3]], sa[:, [0, 1, 3]])
    assert errs[K.GRAD_POINT_CELLS_LEAST_SQUARES] < 1.5 * errs[K.GRAD_LEAST_SQUARES] and errs[K.GRAD_POINT_CELLS_LEAST_SQUARES] < 0.1, errs


# ---------------------------------------------------------------------------------------------
# unsLinGeomSolid: face stresses from vertex-based face gradients (second half)
# ---------------------------------------------------------------------------------------------
def test_uns_gradient_formulas_are_exact_for_linear_vertex_fields():
    """fvcGradf.C: with exact vertex values the Gauss cell gradient (:442-680), the in-plane face gradient (:123-298) plus the
    corrected snGrad (:104-107) reproduce a linear field on distorted 2-D and 3-D meshes, and the face stress is Hooke's law of it."""
    G3 = np.array([[0.01, 0.02, -0.01], [0.03, -0.01, 0.02], [0.005, 0.0, 0.01]])
    pmap = lambda p: p + 0.04 * np.sin(3.0 * p[:, [1, 2, 0]])
    c3 = cases.cantilever(5, 4, 3, general=True, L=2.0, solidModel=K.MODEL_UNS_LIN_GEOM)
    c3.mesh = M.hex_box_general(5, 4, 3, 2.0, 1.0, 1.0, names=("fixed", "loaded", "yMin", "yMax", "zMin", "zMax"), point_map=pmap)
    for c, G in ((cases.patch_test(n=5, solidModel=K.MODEL_UNS_LIN_GEOM), G3 * np.array([[1, 1, 0], [1, 1, 0], [0, 0, 0]])), (c3, G3)):
        o = OracleSolid(c)
        m = c.mesh
        F = m.nInternalFaces
        lin = lambda X: 0.1 + X @ G.T
        o.set("D", lin(m.C)); o.set("D_b", lin(m.Cf[F:]))
        o.uns_grad_from_points(lin(m.points))
        assert np.abs(o.get("gradD").reshape(-1, 3, 3) - G.T).max() < 1e-14
        gf = o.get("gradDf").reshape(-1, 3, 3)
        assert np.abs(gf[:F] - G.T).max() < 1e-14
        eps = 0.5 * (G + G.T)
        sig = 2 * c.law.mu * eps + c.law.lambda_ * np.trace(eps) * np.eye(3)
        ref = np.array([sig[0, 0], sig[0, 1], sig[0, 2], sig[1, 1], sig[1, 2], sig[2, 2]])
        assert np.abs(o.get("sigmaf")[:F] - ref).max() < 1e-12 * np.abs(ref).max()


def test_uns_model_reaches_the_kirsch_solution_and_agrees_with_the_cell_centred_model():
    """unsLinGeomSolid (unsLinGeomSolid.C:100-175) on the plate-hole case (C1): discretisation-level agreement with the Kirsch
    closed form; on a beam it converges to the Euler-Bernoulli deflection under mesh refinement."""
    cp = cases.plate_hole(solidModel=K.MODEL_UNS_LIN_GEOM)
    op = OracleSolid(cp)
    assert op.evolve()["converged"]
    assert rel_l2(op.get("sigma")[:, [0, 1, 3]], cases.kirsch_stress(cp.mesh.C)[:, [0, 1, 3]]) < 0.06
    assert rel_l2(op.get("D")[:, :2], cases.kirsch_displacement(cp.mesh.C)[:, :2]) < 0.02
    # beam: 8 x 1 x 1 tutorial-like cantilever, tip deflection against Euler-Bernoulli P L^3/(3 E I) = 1.28e-3.  The one-sided
    # inverse-distance vertex values on the patches (enhancedVolPointInterpolation, the OpenFOAM.com flavour) make the scheme
    # stiff on coarse meshes; it converges to the beam solution under refinement (0.43, 0.75, 0.88 of it for 4, 8, 12 cells
    # across), where the cell-centred model is at 0.93, 0.99, 1.01.
    ratio = {}
    for n in (4, 8):
        c = cases.cantilever(4 * n, n, n, L=4.0, general=True, solidModel=K.MODEL_UNS_LIN_GEOM, solutionTolerance=1e-7,
                             alternativeTolerance=1e-7, tolerance=1e-12, nCorrectors=40000, preconditioner=K.PRECOND_DIC)
        o = OracleSolid(c)
        assert o.evolve()["converged"]
        ratio[n] = np.abs(o.get("D")[:, 1]).max() / 1.28e-3
    assert 0.35 < ratio[4] < ratio[8] < 1.0 and (1 - ratio[8]) < 0.55 * (1 - ratio[4]), ratio


# ---------------------------------------------------------------------------------------------
# independent numpy restatements (plain loops on small meshes) of the point-based operators, on RANDOM fields
# ---------------------------------------------------------------------------------------------
def _point_cells(mesh):
    pc = [set() for _ in range(mesh.points.shape[0])]
    F = mesh.nInternalFaces
    cells = np.concatenate([mesh.owner, mesh.faceCells])
    for f, verts in enumerate(mesh.faces):
        for v in verts:
            pc[v].add(int(cells[f]))
            if f < F:
                pc[v].add(int(mesh.neighbour[f]))
    return pc


def test_vol_to_point_weights_against_a_numpy_restatement():
    """enhancedVolPointInterpolation.C:165-245 restated with python loops: w = 1/|x_p - C| over pointCells for points off the
    patches, w = 1/|x_p - Cf| over the patch faces at the point otherwise (random cell and patch values, distorted mesh)."""
    pmap = lambda p: p + 0.03 * np.sin(4.0 * p[:, [2, 0, 1]])
    c = cases.cantilever(4, 3, 3, L=2.0, general=True)
    c.mesh = M.hex_box_general(4, 3, 3, 2.0, 1.0, 1.0, names=("fixed", "loaded", "yMin", "yMax", "zMin", "zMax"), point_map=pmap)
    o = OracleSolid(c)
    m = c.mesh
    F = m.nInternalFaces
    rng = np.random.default_rng(2)
    D, Db = rng.standard_normal((m.nCells, 3)), rng.standard_normal((m.nBoundaryFaces, 3))
    o.set("D", D); o.set("D_b", Db)
    got = o.interpolate_to_points("D")
    pc = _point_cells(m)
    pb = [[] for _ in range(m.points.shape[0])]
    for b in range(m.nBoundaryFaces):
        for v in m.faces[F + b]:
            pb[v].append(b)
    ref = np.zeros_like(got)
    for p, x in enumerate(m.points):
        if pb[p]:
            w = np.array([1.0 / np.linalg.norm(x - m.Cf[F + b]) for b in pb[p]])
            ref[p] = (w[:, None] * Db[pb[p]]).sum(axis=0) / w.sum()
        else:
            cl = sorted(pc[p])
            w = np.array([1.0 / np.linalg.norm(x - m.C[i]) for i in cl])
            ref[p] = (w[:, None] * D[cl]).sum(axis=0) / w.sum()
    assert np.abs(got - ref).max() < 1e-13


def test_point_cells_least_squares_against_a_numpy_restatement():
    """LeastSquaresVectors over the cell-point-cell stencil restated with python loops (stencil: cells sharing a point, boundary
    faces at the cell's points; dd = sum d d/|d|^2; ls = inv(dd) d/|d|^2) on a random field."""
    pmap = lambda p: p + 0.03 * np.sin(4.0 * p[:, [2, 0, 1]])
    c = cases.cantilever(4, 3, 3, L=2.0, general=True, gradScheme=K.GRAD_POINT_CELLS_LEAST_SQUARES)
    c.mesh = M.hex_box_general(4, 3, 3, 2.0, 1.0, 1.0, names=("fixed", "loaded", "yMin", "yMax", "zMin", "zMax"), point_map=pmap)
    o = OracleSolid(c)
    m = c.mesh
    F = m.nInternalFaces
    rng = np.random.default_rng(4)
    D, Db = rng.standard_normal((m.nCells, 3)), rng.standard_normal((m.nBoundaryFaces, 3))
    o.set("D", D); o.set("D_b", Db)
    o.op_grad()
    got = o.get("gradD").reshape(-1, 3, 3)
    pc = _point_cells(m)
    cell_pts = [set() for _ in range(m.nCells)]
    for p, cl in enumerate(pc):
        for i in cl:
            cell_pts[i].add(p)
    pb = [[] for _ in range(m.points.shape[0])]
    for b in range(m.nBoundaryFaces):
        for v in m.faces[F + b]:
            pb[v].append(b)
    for i in range(m.nCells):
        nb_cells = sorted({j for p in cell_pts[i] for j in pc[p]} - {i})
        nb_faces = sorted({b for p in cell_pts[i] for b in pb[p]})
        X = np.concatenate([m.C[nb_cells], m.Cf[F:][nb_faces]]) if nb_faces else m.C[nb_cells]
        U = np.concatenate([D[nb_cells], Db[nb_faces]]) if nb_faces else D[nb_cells]
        d = X - m.C[i]
        r2 = (d * d).sum(axis=1)
        dd = np.einsum("ka,kb->ab", d / r2[:, None], d)
        ls = (np.linalg.inv(dd) @ (d / r2[:, None]).T).T            # [k,3]
        g = np.einsum("ka,kj->aj", ls, U - D[i])                    # grad_aj = d_a D_j
        assert np.abs(got[i] - g).max() < 1e-10 * max(1.0, np.abs(g).max()), i


def test_uns_gradients_against_a_numpy_restatement():
    """fvcGradf.C restated with python loops on random vertex values: the in-plane face gradient (edge loop), the Gauss cell
    gradient (triangle fans, volume from the same fans) and gradDf = fsGrad + n snGrad on internal faces."""
    pmap = lambda p: p + 0.03 * np.sin(4.0 * p[:, [2, 0, 1]])
    c = cases.cantilever(3, 3, 2, L=1.5, general=True, solidModel=K.MODEL_UNS_LIN_GEOM)
    c.mesh = M.hex_box_general(3, 3, 2, 1.5, 1.0, 1.0, names=("fixed", "loaded", "yMin", "yMax", "zMin", "zMax"), point_map=pmap)
    o = OracleSolid(c)
    m = c.mesh
    F, nF = m.nInternalFaces, m.nInternalFaces + m.nBoundaryFaces
    rng = np.random.default_rng(6)
    D, pD = rng.standard_normal((m.nCells, 3)), rng.standard_normal((m.points.shape[0], 3))
    o.set("D", D)
    o.uns_grad_from_points(pD)
    gD, gDf = o.get("gradD").reshape(-1, 3, 3), o.get("gradDf").reshape(-1, 3, 3)
    cells = np.concatenate([m.owner, m.faceCells])
    G = np.zeros((m.nCells, 3, 3)); V3 = np.zeros(m.nCells)
    # non-orthogonal part of snGrad(D) needs fvc::grad(D) of the gradScheme (least squares): take it from a second oracle
    o2 = OracleSolid(cases.cantilever(3, 3, 2, L=1.5, general=True))
    o2.case.mesh = m
    for f in range(nF):
        verts = m.faces[f]
        P = m.points[verts]; U = pD[verts]
        n = m.Sf[f] / m.magSf[f]
        T = np.zeros((3, 3))
        cp, cf = P.mean(axis=0), U.mean(axis=0)
        Gf = np.zeros((3, 3)); Vf = 0.0
        for i in range(4):
            p0, p1, u0, u1 = P[i], P[(i + 1) % 4], U[i], U[(i + 1) % 4]
            e = p1 - p0
            e = e - n * (n @ e)
            T += np.outer(np.cross(e, n), 0.5 * (u0 + u1))
            St = 0.5 * np.cross(p0 - cp, p1 - cp)
            Gf += np.outer(St, (u0 + u1 + cf) / 3.0)
            Vf += St @ ((cp + p0 + p1) / 3.0)
        T /= m.magSf[f]
        G[cells[f]] += Gf; V3[cells[f]] += Vf
        if f < F:
            G[m.neighbour[f]] -= Gf; V3[m.neighbour[f]] -= Vf
            # orthogonal part of the corrected snGrad; the correction-vector part is checked through exactness elsewhere
            sn = m.nonOrthDeltaCoeffs[f] * (D[m.neighbour[f]] - D[m.owner[f]])
            resid = gDf[f] - T - np.outer(n, sn)
            # what is left must be n (x) (corr & grad_f): normal in its first index
            assert np.abs(resid - np.outer(n, n @ resid)).max() < 1e-10
    ref = G / (V3 / 3.0)[:, None, None]
    assert np.abs(gD - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())
    assert np.allclose(V3 / 3.0, m.V, rtol=2e-2)          # the fan volume is the cell volume up to face warpage


def test_updated_lagrangian_inertia_is_exact_for_quadratic_rigid_motion():
    """fvm::d2dt2(rho, DD) + fvc::d2dt2(rho, D.oldTime()) (nonLinGeomUpdatedLagSolid.C:173-176) with the backward scheme is
    backward ddt applied twice to D = D.old + DD: exact for D(t) = a t^2 once every old-time level of both chains holds the
    quadratic (DD.o .. DD.oooo, D.o .. D.ooooo: from the seventh step), A DD - source = rho V 2a for a traction-free body that
    is translated rigidly, with the mesh moved after every step.  Euler: first order, the same identity with an O(dt) defect
    only in the start-up."""
    import scipy.sparse as sp
    dt, a = 1e-3, np.array([3.0, -2.0, 0.5])
    for scheme in (K.D2DT2_BACKWARD, K.D2DT2_EULER):
        mesh = M.hex_box_general(4, 3, 3, 2.0, 1.0, 1.0, names=("a", "b", "c", "d", "e", "f"))
        bcs = {p.name: K.solidTraction((0.0, 0.0, 0.0)) for p in mesh.patches}
        law = K.mechanical_law("neoHookeanElastic", rho=1000.0, E=3e6, nu=0.3)
        c = K.SolidCase(mesh, bcs, law, K.default_controls(solidModel=K.MODEL_NONLIN_UL, d2dt2Scheme=scheme, deltaT=dt, deltaT0=dt,
                                                           stabilisation=K.STAB_NONE))
        o = OracleSolid(c)
        N = mesh.nCells
        rigid = lambda t: np.tile(a * t * t, (N, 1))
        ok_steps = 0
        for step in range(1, 10):
            o.new_timestep(dt)
            o.set("DD", rigid(step * dt) - rigid((step - 1) * dt))
            o.set("D", rigid(step * dt))
            o.initialise()                     # boundary values and gradient of the rigid increment (zero strain)
            o.op_assemble()
            m = c.mesh
            src, diag, upper = o.get("source"), o.get("diag"), o.get("upper")
            DD = o.get("DD")
            res = np.zeros((N, 3))
            for q in range(3):
                A = sp.coo_matrix((np.concatenate([diag[:, q], upper, upper]),
                                   (np.concatenate([np.arange(N), m.owner, m.neighbour]), np.concatenate([np.arange(N), m.neighbour, m.owner]))),
                                  shape=(N, N)).tocsr()
                res[:, q] = A @ DD[:, q] - src[:, q]
            exact = 1000.0 * m.V[:, None] * 2.0 * a[None, :]
            if step >= 7 or (scheme == K.D2DT2_EULER and step >= 4):
                assert np.abs(res - exact).max() < 1e-6 * np.abs(exact).max(), (scheme, step)
                ok_steps += 1
            o.update_total_fields()            # rho / relJ (relJ = 1) and the rigid mesh motion
            assert np.allclose(o.get("rho"), 1000.0, rtol=1e-12)
        assert ok_steps >= 3
        assert np.allclose(c.mesh.C - mesh.C, a * (9 * dt) ** 2, atol=1e-12)       # the mesh followed the body


def test_pressure_equation_against_a_scipy_restatement():
    """The first pressure solve (grad(sigmaHyd) still zero) restated with scipy: (V + sum a) p_P - sum a p_N = V p_explicit with
    a = scale * interpolate(impK/DEqnA) magSf nonOrthDeltaCoeffs, DEqnA = component-averaged diagonal / V (mechanicalLaw.C:1432-1453),
    solved directly; then grad(sigmaHyd) of that field is the gradScheme's gradient with zero-gradient patches."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    c = _beam_with_pressure_eqn(4, True, tolerance=1e-14, relTol=0.0)
    o = OracleSolid(c)
    m = c.mesh
    N, F = m.nCells, m.nInternalFaces
    rng = np.random.default_rng(8)
    g = 1e-4 * rng.standard_normal((N, 9))
    o.set("gradD", g)
    o.op_assemble()                                   # the momentum diagonal DEqnA is built from
    diag = o.get("diag")
    o.op_correct()
    p = o.get("sigmaHyd")
    pExp = c.law.K * (g[:, 0] + g[:, 4] + g[:, 8])
    impK = 2 * c.law.mu + c.law.lambda_
    r = impK / (diag.mean(axis=1) / m.V)
    w = m.weights[:F]
    a = c.law.pressureSmoothingScaleFactor * (w * r[m.owner] + (1 - w) * r[m.neighbour]) * m.magSf[:F] * m.nonOrthDeltaCoeffs[:F]
    d = m.V + np.bincount(m.owner, weights=a, minlength=N) + np.bincount(m.neighbour, weights=a, minlength=N)
    A = sp.coo_matrix((np.concatenate([d, -a, -a]), (np.concatenate([np.arange(N), m.owner, m.neighbour]),
                                                    np.concatenate([np.arange(N), m.neighbour, m.owner]))), shape=(N, N)).tocsc()
    ref = spla.spsolve(A, m.V * pExp)
    assert rel_l2(p, ref) < 1e-10
    # sigma carries the solved hydrostatic stress; the smoothing is active
    s = o.get("sigma")
    assert np.allclose((s[:, 0] + s[:, 3] + s[:, 5]) / 3.0, p, rtol=0, atol=1e-9 * np.abs(p).max())
    assert rel_l2(p, pExp) > 1e-2
    # grad(sigmaHyd): least squares with zero-gradient patch values (p_b = p_P)
    lsP, lsN = o.ls_vectors()
    gp = np.zeros((N, 3))
    dp = p[m.neighbour] - p[m.owner]
    for q in range(3):
        gp[:, q] = np.bincount(m.owner, weights=lsP[:F, q] * dp, minlength=N) - np.bincount(m.neighbour, weights=lsN[:, q] * dp, minlength=N)
    assert rel_l2(o.get("gradSigmaHyd"), gp) < 1e-10


# ---------------------------------------------------------------------------------------------
# unsNonLinGeomTotalLagSolid (finite-strain face stresses)
# ---------------------------------------------------------------------------------------------
def test_uns_total_lagrangian_face_stress_is_exact_for_a_uniform_deformation_gradient():
    """With exact vertex values of a linear displacement field on a distorted mesh the face gradient is the exact gradient
    (fvcGradf.C), so Ff = I + gradDf.T() is uniform and sigmaf must equal the closed-form compressible neo-Hookean Cauchy stress
    (neoHookeanElastic.C:338-352) of that F on every internal face."""
    G = np.array([[0.05, 0.02, -0.01], [0.03, -0.04, 0.02], [0.01, 0.02, 0.06]])
    pmap = lambda p: p + 0.04 * np.sin(3.0 * p[:, [1, 2, 0]])
    c = cases.neo_hookean_cantilever(5, 4, 3, general=True, L=2.0, solidModel=K.MODEL_UNS_NONLIN_TL)
    c.mesh = M.hex_box_general(5, 4, 3, 2.0, 1.0, 1.0, names=("fixed", "loaded", "yMin", "yMax", "zMin", "zMax"), point_map=pmap)
    o = OracleSolid(c)
    m = c.mesh
    F = m.nInternalFaces
    lin = lambda X: 0.1 + X @ G.T
    o.set("D", lin(m.C)); o.set("D_b", lin(m.Cf[F:]))
    o.uns_grad_from_points(lin(m.points))
    assert np.abs(o.get("gradDf").reshape(-1, 3, 3)[:F] - G.T).max() < 1e-14
    Fm = np.eye(3) + G                                  # F = I + gradD.T(), gradD = G.T
    J = np.linalg.det(Fm)
    b = J ** (-2.0 / 3.0) * (Fm @ Fm.T)
    sig = (0.5 * c.law.K * (J * J - 1.0) * np.eye(3) + c.law.mu * (b - np.trace(b) / 3.0 * np.eye(3))) / J
    ref = np.array([sig[0, 0], sig[0, 1], sig[0, 2], sig[1, 1], sig[1, 2], sig[2, 2]])
    assert np.abs(o.get("sigmaf")[:F] - ref).max() < 1e-12 * np.abs(ref).max()


def test_uns_total_lagrangian_model_reduces_to_the_linear_uns_model_at_small_strain():
    """unsNonLinGeomTotalLagSolid with neoHookeanElastic under a vanishing load against unsLinGeomSolid with linearElastic
    (same E, nu): the neo-Hookean law linearises to Hooke's, J Finv.T() -> I, so the two discretisations must agree to O(strain);
    this pins the finite-strain face path on the (independently pinned) linear one.  Under a large load it still converges,
    with its own criterion, to a visibly non-linear state."""
    kw = dict(general=True, L=2.0, E=3e6, nu=0.3, nCorrectors=5000, tolerance=1e-12, relTol=0.01, preconditioner=K.PRECOND_DIC)
    a = OracleSolid(cases.cantilever(8, 4, 4, traction=(0.0, -1e-3, 0.0), solidModel=K.MODEL_UNS_LIN_GEOM, solutionTolerance=1e-10,
                                     alternativeTolerance=1e-10, **kw))
    b = OracleSolid(cases.neo_hookean_cantilever(8, 4, 4, traction=(0.0, -1e-3, 0.0), solidModel=K.MODEL_UNS_NONLIN_TL,
                                                 solutionTolerance=1e-8, **kw))
    for o in (a, b):
        o.new_timestep(1.0)
    sa, sb = a.evolve(), b.evolve()
    assert sa["converged"] and sb["converged"]
    assert rel_l2(b.get("D"), a.get("D")) < 1e-6
    big = OracleSolid(cases.neo_hookean_cantilever(8, 4, 4, traction=(0.0, -2e4, 0.0), solidModel=K.MODEL_UNS_NONLIN_TL,
                                                   solutionTolerance=1e-9, **kw))
    big.new_timestep(1.0)
    st = big.evolve()
    assert st["converged"] and st["nCorr"] > 1 and st["relResidual"] <= 1e-9
    lin = a.get("D") * (2e4 / 1e-3)                              # the linear answer scaled to the large load
    assert big.get("D")[:, 1].min() < -0.1
    assert 2e-3 < rel_l2(big.get("D"), lin) < 0.1                # close to, but not, the linear extrapolation


def test_uns_updated_lagrangian_first_step_equals_total_lagrangian_and_second_step_agrees():
    """unsNonLinGeomUpdatedLagSolid against unsNonLinGeomTotalLagSolid: from the reference configuration relFf = Ff and the mesh
    is the initial one, so the first load step is the same discrete problem (different convergence criteria: agreement to the
    outer tolerance); after updateTotalFields (mesh moved, Ff.oldTime() stored, density updated) the second step is
    discretised on the deformed mesh and agrees with the total-Lagrangian answer to discretisation accuracy."""
    kw = dict(general=True, L=2.0, nCorrectors=8000, tolerance=1e-12, relTol=0.01, preconditioner=K.PRECOND_DIC)
    tl = OracleSolid(cases.neo_hookean_cantilever(8, 4, 4, traction=(0.0, -1e4, 0.0), solidModel=K.MODEL_UNS_NONLIN_TL,
                                                  solutionTolerance=1e-10, **kw))
    ul = OracleSolid(cases.neo_hookean_cantilever(8, 4, 4, traction=(0.0, -1e4, 0.0), solidModel=K.MODEL_UNS_NONLIN_UL,
                                                  solutionTolerance=1e-9, alternativeTolerance=1e-9, **kw))
    for o in (tl, ul):
        o.new_timestep(1.0)
        assert o.evolve()["converged"]
    assert rel_l2(ul.get("D"), tl.get("D")) < 1e-6
    assert rel_l2(ul.get("sigmaf"), tl.get("sigmaf")) < 1e-6
    p0 = ul.case.mesh.points.copy()
    for o in (tl, ul):
        o.update_total_fields()
        o.new_timestep(1.0)
        o.set_bc("loaded", K.solidTraction((0.0, -2e4, 0.0)))
        assert o.evolve()["converged"]
    assert np.abs(ul.case.mesh.points - p0).max() > 0.05          # the updated-Lagrangian mesh did move
    assert np.abs(tl.get("D")[:, 1]).max() > 0.15
    assert rel_l2(ul.get("D"), tl.get("D")) < 2e-3
    assert rel_l2(ul.get("rho"), np.full(ul.case.mesh.nCells, ul.case.law.rho)) > 1e-4      # rho_ = rho_.oldTime() / relJ_
