"""On-disk formats: OpenFOAM dictionaries, polyMesh and vol-field files read / written by the host
side, checked against the reference's own tutorial files where this container has them and by round trips otherwise."""
import os

import numpy as np
import pytest

from solids4foam_b200 import case as K
from solids4foam_b200 import cases
from solids4foam_b200 import foam_io as IO
from solids4foam_b200 import mesh as M

REF = "/root/reference/tutorials"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only mounted in the build container")


def test_dictionary_parser_handles_the_openfoam_syntax():
    d = IO.parse_foam_dict('''
        FoamFile { version 2.0; class dictionary; }
        // comment
        planeStress     no;   /* block
        comment */
        mechanical ( steel { type linearElastic; rho rho [1 -3 0 0 0 0 0] 7854; E E [1 -1 -2 0 0 0 0] 200e+9; nu nu [0 0 0 0 0 0 0] 0.3; } );
        solvers { "D|DD" { solver PCG; preconditioner FDIC; tolerance 1e-09; relTol 0.1; } }
        value uniform (0 1.5 -2);
    ''')
    assert d["planeStress"] == "no"
    name, law = d["mechanical"][0]
    assert name == "steel" and law["type"] == "linearElastic" and IO._scalar(law["E"]) == 200e9 and IO._scalar(law["rho"]) == 7854
    assert IO._lookup(d["solvers"], "DD")["preconditioner"] == "FDIC" and IO._lookup(d["solvers"], "D")["relTol"] == 0.1
    assert d["value"] == ["uniform", [0, 1.5, -2]]


@needs_ref
def test_plate_hole_tutorial_directory_gives_the_hand_built_case():
    """tutorials/solids/linearElasticity/plateHole: constant/{solidProperties,mechanicalProperties}, system/{fvSchemes,
    fvSolution}, 0/D read from the reference's own files reproduce cases.plate_hole() field by field."""
    mesh = M.plate_hole()
    c = IO.read_case(os.path.join(REF, "solids/linearElasticity/plateHole"), mesh=mesh)
    ref = cases.plate_hole()
    for f, _ in K.Law._fields_:
        a, b = getattr(c.law, f), getattr(ref.law, f)
        assert (list(a) == list(b)) if hasattr(a, "__len__") else (a == b), f
    for f in ("solidModel", "gradScheme", "d2dt2Scheme", "stabilisation", "stabScaleFactor", "relaxationMethod", "fieldRelaxD", "solver",
              "tolerance", "relTol", "nCorrectors", "solutionTolerance", "alternativeTolerance", "materialTolerance"):
        assert getattr(c.controls, f) == getattr(ref.controls, f), f
    assert c.controls.preconditioner == K.PRECOND_DIC              # FDIC in the tutorial
    assert set(c.bcs) == set(ref.bcs)
    for name in ref.bcs:
        assert c.bcs[name].kind == ref.bcs[name].kind, name
        if ref.bcs[name].value is not None:
            p = mesh.patch(name)
            assert np.allclose(np.broadcast_to(c.bcs[name].value, (p.size, 3)), np.broadcast_to(ref.bcs[name].value, (p.size, 3)), rtol=0, atol=1e-9)


@needs_ref
def test_beam_in_cross_flow_solid_dictionaries():
    base = os.path.join(REF, "fluidSolidInteraction/beamInCrossFlow")
    # the solid region keeps its files under constant/solid and system/solid: present them as a case directory
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "constant")); os.makedirs(os.path.join(tmp, "system"))
        for sub, names in (("constant", ("solidProperties", "mechanicalProperties")), ("system", ("fvSchemes", "fvSolution"))):
            for n in names:
                os.symlink(os.path.join(base, sub, "solid", n), os.path.join(tmp, sub, n))
        law = IO.read_mechanical_law(tmp)
        ctl = IO.read_controls(tmp)
    ref = cases.beam_in_cross_flow()
    assert (law.kind, law.rho, law.mu, law.K) == (ref.law.kind, ref.law.rho, ref.law.mu, ref.law.K)
    assert ctl.solidModel == K.MODEL_NONLIN_UL and ctl.d2dt2Scheme == K.D2DT2_BACKWARD and ctl.gradScheme == K.GRAD_LEAST_SQUARES
    assert ctl.fieldRelaxD == ref.controls.fieldRelaxD == 0.9 and ctl.nCorrectors == ref.controls.nCorrectors == 1000
    assert ctl.tolerance == 1e-9 and ctl.relTol == 0.1


@needs_ref
def test_necking_bar_tutorial_law_and_table():
    """tutorials/solids/elastoplasticity/neckingBar: neoHookeanElasticMisesPlastic with the hardening table read through the
    "file|fileName" key and $FOAM_CASE; updated-Lagrangian solid model."""
    base = os.path.join(REF, "solids/elastoplasticity/neckingBar")
    law = IO.read_mechanical_law(base)
    ref = K.mechanical_law("neoHookeanElasticMisesPlastic", rho=7833.0, E=200e9, nu=0.3, table=K.NECKING_BAR_TABLE)
    assert law.kind == K.LAW_NEO_HOOKEAN_MISES_PLASTIC and law.nTable == 8
    assert (law.rho, law.mu, law.K) == (ref.rho, ref.mu, ref.K)
    assert list(law.tableEps)[:8] == list(ref.tableEps)[:8] and list(law.tableSigY)[:8] == list(ref.tableSigY)[:8]
    sp = IO.read_foam_dict(os.path.join(base, "constant", "solidProperties"))
    assert K.MODEL_NAMES[str(sp["solidModel"])] == K.MODEL_NONLIN_UL


def test_poly_mesh_round_trip(tmp_path):
    names = ("fixed", "loaded", "yMin", "yMax", "zMin", "zMax")
    kinds = (M.PATCH, M.PATCH, M.PATCH, M.PATCH, M.SYMMETRY_PLANE, M.PATCH)
    pmap = lambda p: p + 0.03 * np.sin(3.0 * p[:, [1, 2, 0]])              # a non-orthogonal mesh
    mesh = M.hex_box_general(5, 4, 3, 2.0, 1.0, 1.0, names=names, kinds=kinds, point_map=pmap, cell_perm_seed=3)
    IO.write_poly_mesh(str(tmp_path / "polyMesh"), mesh)
    back = IO.read_poly_mesh(str(tmp_path / "polyMesh"))
    assert back.nCells == mesh.nCells and np.array_equal(back.owner, mesh.owner) and np.array_equal(back.neighbour, mesh.neighbour)
    assert [(p.name, p.kind, p.start, p.size) for p in back.patches] == [(p.name, p.kind, p.start, p.size) for p in mesh.patches]
    for f in ("C", "V", "Sf", "magSf", "Cf", "weights", "nonOrthDeltaCoeffs", "nonOrthCorrVec", "CnbrB"):
        assert np.allclose(getattr(back, f), getattr(mesh, f), rtol=1e-13, atol=1e-15), f
    assert np.array_equal(back.faces, mesh.faces) and np.array_equal(back.points, mesh.points)
    moved = M.move_points(back, back.points * 1.01)                          # a read mesh can be moved (updated Lagrangian)
    assert np.allclose(moved.V, 1.01 ** 3 * mesh.V, rtol=1e-12)


def test_polygon_geometry_matches_the_quad_routine():
    mesh = M.hex_box_general(3, 3, 2, 1.0, 1.0, 1.0, point_map=lambda p: p + 0.05 * np.cos(2.0 * p[:, [2, 0, 1]]))
    fptr = np.arange(0, 4 * mesh.faces.shape[0] + 1, 4)
    c, a = IO.polygon_centres_and_areas(mesh.points, fptr, mesh.faces.reshape(-1).astype(np.int64))
    c2, a2 = M.face_centres_and_areas(mesh.points, mesh.faces)
    assert np.allclose(c, c2, atol=1e-15) and np.allclose(a, a2, atol=1e-15)
    # a triangle: centroid and half cross product
    pts = np.array([[0.0, 0, 0], [2, 0, 0], [0, 3, 0]])
    c, a = IO.polygon_centres_and_areas(pts, np.array([0, 3]), np.array([0, 1, 2]))
    assert np.allclose(c[0], pts.mean(axis=0)) and np.allclose(a[0], [0, 0, 3.0])


def test_vol_field_round_trip(tmp_path):
    mesh = M.hex_box_general(3, 2, 2)
    rng = np.random.default_rng(0)
    D = rng.standard_normal((mesh.nCells, 3)); Db = rng.standard_normal((mesh.nBoundaryFaces, 3))
    sig = rng.standard_normal((mesh.nCells, 6))
    IO.write_vol_field(str(tmp_path / "1"), "D", mesh, D, Db)
    IO.write_vol_field(str(tmp_path / "1"), "sigma", mesh, sig, dimensions="[1 -1 -2 0 0 0 0]")
    Di, Dbv = IO.read_vol_field(str(tmp_path / "1" / "D"), mesh)
    assert np.array_equal(Di, D)
    for p in mesh.patches:
        assert np.array_equal(Dbv[p.name], Db[p.start:p.start + p.size])
    si, _ = IO.read_vol_field(str(tmp_path / "1" / "sigma"), mesh)
    assert np.array_equal(si, sig)
    txt = open(tmp_path / "1" / "sigma").read()
    assert "volSymmTensorField" in txt and "List<symmTensor>" in txt


def test_case_directory_round_trip(tmp_path):
    """write_case -> read_case gives back the law, the controls the dictionaries carry and the boundary data."""
    c = cases.neo_hookean_cantilever(6, 3, 3, general=True, traction=(0.0, -2e3, 10.0), solidModel=K.MODEL_NONLIN_UL,
                                     d2dt2Scheme=K.D2DT2_BACKWARD, deltaT=0.05, fieldRelaxD=0.9, g=(0.0, -9.81, 0.0), nCorrectors=123,
                                     tolerance=1e-11, relTol=0.01, gradScheme=K.GRAD_GAUSS_LINEAR, preconditioner=K.PRECOND_GAMG)
    c.law = K.mechanical_law("neoHookeanElasticMisesPlastic", rho=1000.0, E=3e6, nu=0.3, table=K.NECKING_BAR_TABLE, solvePressureEqn=True,
                             pressureSmoothingScaleFactor=50.0)
    IO.write_case(str(tmp_path), c)
    r = IO.read_case(str(tmp_path))
    for f in ("kind", "rho", "nTable", "solvePressureEqn", "pressureSmoothingScaleFactor"):
        assert getattr(r.law, f) == getattr(c.law, f), f
    assert np.isclose(r.law.mu, c.law.mu, rtol=1e-15) and np.isclose(r.law.K, c.law.K, rtol=1e-15)
    assert list(r.law.tableSigY)[:8] == list(c.law.tableSigY)[:8]
    for f in ("solidModel", "gradScheme", "d2dt2Scheme", "stabilisation", "stabScaleFactor", "fieldRelaxD", "solver", "preconditioner",
              "tolerance", "relTol", "maxIter", "nCorrectors", "solutionTolerance", "alternativeTolerance", "deltaT"):
        assert getattr(r.controls, f) == getattr(c.controls, f), f
    assert list(r.controls.g) == [0.0, -9.81, 0.0]
    assert np.allclose(r.mesh.V, c.mesh.V, rtol=1e-13)
    for p in c.mesh.patches:
        a, b = r.bcs[p.name], c.bcs[p.name]
        assert a.kind == b.kind
        if b.value is not None:
            assert np.array_equal(np.broadcast_to(a.value, (p.size, 3)), np.broadcast_to(b.value, (p.size, 3)))


# ---- time series and refused entries (a case must not run with a silently different set-up) ----------------------------
@needs_ref
def test_necking_bar_displacement_series_is_read_and_interpolated():
    """neckingBar/0/DD: the "loading" patch is a fixedDisplacement driven by displacementSeries (constant/timeVsDisp:
    (0 (0 0 0)) (1 (0.007 0 0)), clamp) -- fixedDisplacementFvPatchVectorField.C:258-294 evaluates it at every time."""
    base = os.path.join(REF, "solids/elastoplasticity/neckingBar")
    d = IO.read_foam_dict(os.path.join(base, "0", "DD"))
    sub = IO._lookup(d["boundaryField"], "loading")
    series = IO._read_series(sub["displacementSeries"], base)
    assert series[0][0] == 0.0 and series[-1][0] == 1.0 and list(series[-1][1]) == [0.007, 0, 0]
    bc = K.fixedDisplacement((0.0, 0.0, 0.0))
    bc.value_series = series
    assert np.allclose(bc.at(0.25).value, [0.00175, 0.0, 0.0]) and np.allclose(bc.at(7.0).value, [0.007, 0.0, 0.0])      # clamp
    assert np.allclose(bc.at(-1.0).value, 0.0)


def _write_min_case(tmp_path, patch_entries: str, stab="RhieChow", extra_relax=""):
    mesh = M.hex_box_general(3, 2, 2, 1.0, 1.0, 1.0, names=("fixed", "loaded", "yMin", "yMax", "zMin", "zMax"))
    c = cases.cantilever(3, 2, 2, general=True)
    IO.write_case(str(tmp_path), c)
    p = tmp_path / "0" / "D"
    txt = p.read_text()
    import re
    txt = re.sub(r"loaded\s*\{[^}]*\}", "loaded { " + patch_entries + " }", txt, count=1)
    p.write_text(txt)
    if stab != "RhieChow":
        sp = tmp_path / "constant" / "solidProperties"
        sp.write_text(sp.read_text().replace("RhieChow", stab))
    if extra_relax:
        fs = tmp_path / "system" / "fvSolution"
        fs.write_text(fs.read_text().replace("relaxationFactors\n{", "relaxationFactors\n{\n    equations { D " + extra_relax + "; }"))
    return mesh


def test_traction_series_runs_through_the_case_reader(tmp_path):
    (tmp_path / "constant").mkdir(parents=True, exist_ok=True)
    _write_min_case(tmp_path, 'type solidTraction; tractionSeries { file "$FOAM_CASE/constant/timeVsTraction"; outOfBounds clamp; } '
                              'pressure uniform 0; value uniform (0 0 0);')
    (tmp_path / "constant" / "timeVsTraction").write_text("( (0 (0 0 0)) (2 (0 -2e6 0)) )\n")
    c = IO.read_case(str(tmp_path))
    bc = c.bcs["loaded"]
    assert bc.value_series is not None and np.allclose(bc.at(1.0).value, [0.0, -1e6, 0.0]) and np.allclose(bc.at(5.0).value, [0.0, -2e6, 0.0])


@pytest.mark.parametrize("entries,msg", [
    ("type solidTraction; traction uniform (0 -1e6 0); pressure uniform 0; secondOrder yes; value uniform (0 0 0);", "secondOrder"),
    ("type solidTraction; traction uniform (0 -1e6 0); pressure uniform 0; setEffectiveTraction yes; value uniform (0 0 0);", "setEffectiveTraction"),
    ("type solidTraction; tractionField sigmaTrac; pressure uniform 0; value uniform (0 0 0);", "tractionField"),
    ('type solidTraction; traction uniform (0 0 0); tractionSeries { file "x"; } pressure uniform 0; value uniform (0 0 0);', "exactly one"),
])
def test_unsupported_patch_entries_are_refused(tmp_path, entries, msg):
    _write_min_case(tmp_path, entries)
    with pytest.raises(ValueError, match=msg):
        IO.read_case(str(tmp_path))


def test_unsupported_stabilisation_and_equation_relaxation_are_refused(tmp_path):
    d1, d2 = tmp_path / "a", tmp_path / "b"
    d1.mkdir(); d2.mkdir()
    ok = 'type solidTraction; traction uniform (0 -1e6 0); pressure uniform 0; value uniform (0 0 0);'
    _write_min_case(d1, ok, stab="JamesonSchmidtTurkel")
    with pytest.raises(ValueError, match="stabilisation type"):
        IO.read_case(str(d1))
    _write_min_case(d2, ok, extra_relax="0.9")
    with pytest.raises(ValueError, match="equation relaxation"):
        IO.read_case(str(d2))


def test_decomposed_case_directories_round_trip(tmp_path):
    """foam_io.decompose_case restates decomposePar (simple, (P 1 1)): processorN/constant/polyMesh with processor patches whose
    faces match across the cut, cellProcAddressing, processorN/0/D.  Read back rank by rank (cell centres exchanged in memory)
    the parts reproduce the analytically decomposed box: cells, volumes, centres across the cut, weights, outward normals."""
    c = cases.cantilever(9, 3, 2, general=True)
    IO.write_case(str(tmp_path), c)
    IO.decompose_case(str(tmp_path), 3)
    sent = {}
    for r in range(3):            # first pass: what every rank would send
        def collect(send, r=r):
            for q, a in send.items():
                sent[(r, q)] = a
            return {q: a + 1.0 for q, a in send.items()}
        IO.read_poly_mesh(str(tmp_path / f"processor{r}" / "constant" / "polyMesh"), rank=r, nRanks=3, exchange=collect)
    nAll = 0
    for r in range(3):
        cs = IO.read_decomposed_case(str(tmp_path), r, 3, lambda send, r=r: {q: sent[(q, r)] for q in send})
        m = cs.mesh
        ref = M.hex_box_decomposed(9, 3, 2, 8.0, 1.0, 1.0, r, 3, names=("fixed", "loaded", "yMin", "yMax", "zMin", "zMax"))
        nAll += m.nCells
        assert np.array_equal(np.sort(m.cellGlobal), np.sort(ref.cellGlobal))
        ia, ib = np.argsort(m.cellGlobal), np.argsort(ref.cellGlobal)
        assert np.allclose(m.C[ia], ref.C[ib]) and np.allclose(m.V[ia], ref.V[ib])
        F = m.nInternalFaces
        assert (m.owner < m.neighbour).all()                                       # upper-triangular local addressing
        for p in m.patches:
            if p.kind != M.PROCESSOR:
                assert cs.bcs[p.name].kind == c.bcs[p.name].kind
                continue
            sl = slice(p.start, p.start + p.size)
            fsl = slice(F + p.start, F + p.start + p.size)
            d = m.CnbrB[sl] - m.C[m.faceCells[sl]]
            assert (np.einsum("ij,ij->i", m.Sf[fsl], d) > 0).all()                  # normals point out of this processor
            assert np.allclose(m.weights[fsl], 0.5) and np.allclose(m.nonOrthCorrVec[fsl], 0, atol=1e-12)
            assert np.allclose(np.abs(d[:, 0]), 8.0 / 9.0) and np.allclose(d[:, 1:], 0, atol=1e-12)
    assert nAll == c.mesh.nCells
    # the loaded patch lives on the last processor only, with its traction
    last = IO.read_decomposed_case(str(tmp_path), 2, 3, lambda send: {q: sent[(q, 2)] for q in send})
    assert last.mesh.patch("loaded").size == 6 and np.allclose(last.bcs["loaded"].value, [0.0, -1e6, 0.0])


def test_manual_decomposition_with_corners_matches_face_by_face(tmp_path):
    """decompose_case with a cell -> processor map as ``method manual`` would give it: blocks dealt out like a chess board on two
    processors and 2 x 2 on four (where the diagonal processors share an edge but no face).  On every pair of processors the two
    sides of the processor patch list the same faces in the same order (centres equal, area vectors opposite), the points of
    those faces are bit-identical copies (what the device-side point matching relies on), cells and volumes add up, and the
    geometry of a warped serial mesh is reproduced cell by cell."""
    nx, ny, nz = 8, 6, 3
    c = cases.cantilever(nx, ny, nz, general=True, L=2.0)
    c.mesh = M.hex_box_general(nx, ny, nz, 2.0, 1.0, 1.0, names=("fixed", "loaded", "yMin", "yMax", "zMin", "zMax"),
                               point_map=lambda p: p + 0.03 * np.sin(3.0 * p[:, [1, 2, 0]]))
    idx = np.arange(c.mesh.nCells)
    bx, by = ((idx % nx) >= nx // 2).astype(np.int64), (((idx // nx) % ny) >= ny // 2).astype(np.int64)
    for world, cell_rank in ((2, (bx + by) % 2), (4, bx + 2 * by)):
        d = tmp_path / f"w{world}"
        IO.write_case(str(d), c)
        IO.decompose_case(str(d), world, cell_rank=cell_rank)
        sent = {}
        for r in range(world):
            def collect(send, r=r):
                for q, a in send.items():
                    sent[(r, q)] = a
                return {q: a + 1.0 for q, a in send.items()}
            IO.read_poly_mesh(str(d / f"processor{r}" / "constant" / "polyMesh"), rank=r, nRanks=world, exchange=collect)
        parts = [IO.read_decomposed_case(str(d), r, world, lambda send, r=r: {q: sent[(q, r)] for q in send}).mesh for r in range(world)]
        assert sum(m.nCells for m in parts) == c.mesh.nCells
        V = np.zeros(c.mesh.nCells); C = np.zeros((c.mesh.nCells, 3))
        for m in parts:
            V[m.cellGlobal] = m.V; C[m.cellGlobal] = m.C
        assert np.allclose(V, c.mesh.V, rtol=1e-13) and np.allclose(C, c.mesh.C, atol=1e-13)
        nbrs = {r: {p.nbr_rank for p in parts[r].patches if p.kind == M.PROCESSOR} for r in range(world)}
        if world == 4:
            assert 3 not in nbrs[0] and 2 not in nbrs[1]                    # diagonal blocks: no processor patch between them ...
            shared = set(map(tuple, parts[0].points.tolist())) & set(map(tuple, parts[3].points.tolist()))
            assert len(shared) == nz + 1                                     # ... but the points of the central edge, bit for bit
        for r in range(world):
            m = parts[r]
            F = m.nInternalFaces
            for p in m.patches:
                if p.kind != M.PROCESSOR:
                    continue
                o = parts[p.nbr_rank]
                q = next(x for x in o.patches if x.kind == M.PROCESSOR and x.nbr_rank == r)
                assert q.size == p.size
                a, b = slice(F + p.start, F + p.start + p.size), slice(o.nInternalFaces + q.start, o.nInternalFaces + q.start + q.size)
                assert np.allclose(m.Cf[a], o.Cf[b], rtol=0, atol=1e-14) and np.allclose(m.Sf[a], -o.Sf[b], rtol=0, atol=1e-15)   # reversed vertex order: round-off
                assert np.allclose(m.weights[a] + o.weights[b], 1.0, atol=1e-14)
                mine = set(map(tuple, m.points[np.unique(m.faces[a])].tolist()))
                theirs = set(map(tuple, o.points[np.unique(o.faces[b])].tolist()))
                assert mine == theirs


@needs_ref
def test_hron_turek_fsi3_solid_dictionaries_select_the_uns_total_lagrangian_model():
    """tutorials/fluidSolidInteraction/HronTurekFsi3 (the FSI benchmark): its solid region runs
    unsNonLinearGeometryTotalLagrangian with neoHookeanElastic (E 5.6e6, nu 0.4, rho 1000) -- the face-stress model and the
    surface-field law of SURVEY 8f row f1.  The readers map the dictionaries onto the corresponding enums and parameters."""
    base = os.path.join(REF, "fluidSolidInteraction/HronTurekFsi3")
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "constant")); os.makedirs(os.path.join(tmp, "system"))
        for sub, names in (("constant", ("solidProperties", "mechanicalProperties")), ("system", ("fvSchemes", "fvSolution"))):
            for n in names:
                os.symlink(os.path.join(base, sub, "solid", n), os.path.join(tmp, sub, n))
        law = IO.read_mechanical_law(tmp)
        ctl = IO.read_controls(tmp)
    ref = K.mechanical_law("neoHookeanElastic", rho=1000.0, E=5.6e6, nu=0.4)
    assert (law.kind, law.rho, law.mu, law.K) == (K.LAW_NEO_HOOKEAN_ELASTIC, ref.rho, ref.mu, ref.K)
    assert ctl.solidModel == K.MODEL_UNS_NONLIN_TL
    assert ctl.nCorrectors == 1000 and ctl.solutionTolerance == 1e-7


def test_point_neighbours_across_processors_are_recovered_from_the_point_coordinates(tmp_path):
    """The host logic behind the point-neighbour ghosts of decomposed meshes (s4f_build_point_ghosts), restated in numpy: every
    processor publishes, per point, its coordinates and the cells around it; identifying points across processors by their
    coordinates BIT FOR BIT gives every processor, for each of its points, exactly the cells the serial mesh has around that
    point -- also the cells of a processor that shares only an edge with it (2 x 2 blocks), with no tolerance involved."""
    nx, ny, nz = 8, 6, 3
    c = cases.cantilever(nx, ny, nz, general=True, L=2.0)
    c.mesh = M.hex_box_general(nx, ny, nz, 2.0, 1.0, 1.0, names=("fixed", "loaded", "yMin", "yMax", "zMin", "zMax"),
                               point_map=lambda p: p + 0.03 * np.sin(3.0 * p[:, [1, 2, 0]]))
    idx = np.arange(c.mesh.nCells)
    bx, by = ((idx % nx) >= nx // 2).astype(np.int64), (((idx // nx) % ny) >= ny // 2).astype(np.int64)

    def point_cells(mesh):          # point -> set of (global) cells, from the fv faces (the hex faces cover every cell-point pair)
        F = mesh.nInternalFaces
        cg = mesh.cellGlobal if mesh.cellGlobal is not None else np.arange(mesh.nCells)
        out = [set() for _ in range(mesh.points.shape[0])]
        for f in range(mesh.faces.shape[0]):
            cells = [mesh.owner[f], mesh.neighbour[f]] if f < F else [mesh.faceCells[f - F]]
            for v in mesh.faces[f]:
                out[v].update(int(cg[x]) for x in cells)
        return out

    serial = {tuple(x): s for x, s in zip(c.mesh.points.tolist(), point_cells(c.mesh))}
    world, cell_rank = 4, bx + 2 * by
    IO.write_case(str(tmp_path), c)
    IO.decompose_case(str(tmp_path), world, cell_rank=cell_rank)
    sent = {}
    for r in range(world):
        def collect(send, r=r):
            for q, a in send.items():
                sent[(r, q)] = a
            return {q: a + 1.0 for q, a in send.items()}
        IO.read_poly_mesh(str(tmp_path / f"processor{r}" / "constant" / "polyMesh"), rank=r, nRanks=world, exchange=collect)
    parts = [IO.read_decomposed_case(str(tmp_path), r, world, lambda send, r=r: {q: sent[(q, r)] for q in send}).mesh for r in range(world)]
    published = []                                          # what every processor publishes: coordinates -> its cells around the point
    for m in parts:
        published.append({tuple(x): s for x, s in zip(m.points.tolist(), point_cells(m))})
    nRemote = 0
    for r, mine in enumerate(published):
        for xyz, local in mine.items():
            found = set(local)
            for q, theirs in enumerate(published):
                if q != r and xyz in theirs:                # the same point on another processor: bit-identical coordinates
                    found |= theirs[xyz]; nRemote += len(theirs[xyz])
            assert found == serial[xyz], (r, xyz)
    assert nRemote > 0
    # processors 0 and 3 share no face, yet the points of the central edge see each other's cells
    edge = set(published[0]) & set(published[3])
    assert len(edge) == nz + 1 and all(published[3][x] - published[0][x] for x in edge)
