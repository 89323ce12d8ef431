"""Host-side logic (CPU): the synthetic fvMesh generators, the decomposePar-style slab decomposition and the
law-shell parameter formulas.  What is checked is what OpenFOAM guarantees about an fvMesh and what the
C-ABI (s4fgpu_set_mesh / s4fgpu_set_geometry) relies on."""
import numpy as np
import pytest

from solids4foam_b200 import case as K
from solids4foam_b200 import cases
from solids4foam_b200 import mesh as M


def _check_fvmesh(m: M.FvMesh):
    F, B, N = m.nInternalFaces, m.nBoundaryFaces, m.nCells
    # lduAddressing: upper-triangular order (owner < neighbour; sorted by owner, then neighbour)
    assert (m.owner < m.neighbour).all()
    key = m.owner.astype(np.int64) * N + m.neighbour
    assert (np.diff(key) > 0).all()
    assert m.Sf.shape == (F + B, 3) and m.weights.shape == (F + B,)
    # closed cells: sum of outward face-area vectors vanishes
    s = np.zeros((N, 3))
    np.add.at(s, m.owner, m.Sf[:F])
    np.add.at(s, m.neighbour, -m.Sf[:F])
    np.add.at(s, m.faceCells, m.Sf[F:])
    scale = m.magSf.max()
    solved = m.solutionD.astype(bool)
    assert np.abs(s[:, solved]).max() < 1e-12 * scale
    # divergence theorem: V = (1/3) sum_f Cf . Sf
    v = np.zeros(N)
    np.add.at(v, m.owner, (m.Cf[:F] * m.Sf[:F]).sum(1))
    np.add.at(v, m.neighbour, -(m.Cf[:F] * m.Sf[:F]).sum(1))
    np.add.at(v, m.faceCells, (m.Cf[F:] * m.Sf[F:]).sum(1))
    if solved.all():
        assert np.allclose(v / 3, m.V, rtol=1e-10)
    assert (m.V > 0).all()
    # interpolation weights in (0,1); face normal points owner -> neighbour
    assert ((m.weights[:F] > 0) & (m.weights[:F] < 1)).all()
    d = m.C[m.neighbour] - m.C[m.owner]
    assert ((d * m.Sf[:F]).sum(1) > 0).all()
    assert (m.nonOrthDeltaCoeffs > 0).all()
    # patches tile the boundary-face range
    pos = 0
    for p in m.patches:
        assert p.start == pos
        pos += p.size
    assert pos == B


@pytest.mark.parametrize("dims", [(5, 3, 2), (1, 1, 1), (7, 1, 3)])
def test_hex_box_is_a_valid_fvmesh(dims):
    m = M.hex_box(*dims, 2.0, 1.0, 0.5)
    _check_fvmesh(m)
    assert m.nCells == dims[0] * dims[1] * dims[2]
    assert m.V.sum() == pytest.approx(1.0)
    assert m.is_orthogonal()
    # blockMesh numbering: cell i + nx (j + ny k)
    nx = dims[0]
    assert np.allclose(m.C[:nx, 0], (np.arange(nx) + 0.5) * 2.0 / nx)


def test_rectilinear_fast_path_equals_general_builder():
    a = M.hex_box(6, 4, 3, 3.0, 2.0, 1.0)
    b = M.hex_box_general(6, 4, 3, 3.0, 2.0, 1.0)
    assert (a.owner == b.owner).all() and (a.neighbour == b.neighbour).all() and (a.faceCells == b.faceCells).all()
    for f in ("C", "V", "Sf", "magSf", "Cf", "weights", "nonOrthDeltaCoeffs"):
        assert np.allclose(getattr(a, f), getattr(b, f), rtol=1e-12, atol=1e-14), f


def test_plate_hole_mesh_matches_blockmeshdict():
    m = M.plate_hole()
    _check_fvmesh(m)
    assert m.nCells == 1000                       # 10x10 + 10x10 + 20x10 + 20x20 + 10x20
    assert list(m.solutionD) == [1, 1, 0]
    names = {p.name for p in m.patches}
    assert {"left", "right", "down", "up", "hole"} <= names
    assert not m.is_orthogonal()
    area = 4.0 - np.pi * 0.25 / 4                  # quarter plate minus quarter hole (thickness 0.5)
    assert m.V.sum() / 0.5 == pytest.approx(area, rel=2e-3)
    F = m.nInternalFaces
    hole = m.patch_slice("hole")
    r = np.hypot(m.Cf[F:][hole, 0], m.Cf[F:][hole, 1])
    assert np.allclose(r, 0.5, rtol=5e-3)


def test_notched_bar_mesh_is_non_orthogonal_but_valid():
    c = cases.notched_bar(12, 4, 4)
    _check_fvmesh(c.mesh)
    assert not c.mesh.is_orthogonal()


@pytest.mark.parametrize("nRanks", [2, 3, 4, 8])
def test_slab_decomposition_matches_whole_mesh(nRanks):
    nx, ny, nz = 10, 3, 2
    whole = M.hex_box(nx, ny, nz, 8.0, 1.0, 1.0)
    parts = [M.hex_box_decomposed(nx, ny, nz, 8.0, 1.0, 1.0, r, nRanks) for r in range(nRanks)]
    assert sum(p.nCells for p in parts) == whole.nCells
    seen = np.concatenate([p.cellGlobal for p in parts])
    assert sorted(seen) == list(range(whole.nCells))
    nProcFaces = 0
    for r, p in enumerate(parts):
        _check_fvmesh(p)
        assert np.allclose(p.C, whole.C[p.cellGlobal]) and np.allclose(p.V, whole.V[p.cellGlobal])
        F = p.nInternalFaces
        for pa in p.patches:
            if pa.kind != M.PROCESSOR:
                continue
            nProcFaces += pa.size
            q = parts[pa.nbr_rank]
            pb = [x for x in q.patches if x.kind == M.PROCESSOR and x.nbr_rank == r][0]
            assert pa.size == pb.size == ny * nz
            sa, sb = slice(pa.start, pa.start + pa.size), slice(pb.start, pb.start + pb.size)
            # same faces in the same order on both sides: equal centres, opposite area vectors
            assert np.allclose(p.Cf[F:][sa], q.Cf[q.nInternalFaces:][sb])
            assert np.allclose(p.Sf[F:][sa], -q.Sf[q.nInternalFaces:][sb])
            # CnbrB = the neighbour rank's cell centres (patchNeighbourField of C)
            assert np.allclose(p.CnbrB[sa], q.C[q.faceCells[sb]])
            assert np.allclose(p.weights[F:][sa] + q.weights[q.nInternalFaces:][sb], 1.0)
    # internal faces are conserved: cut faces appear once on each side
    assert sum(p.nInternalFaces for p in parts) + nProcFaces // 2 == whole.nInternalFaces


def test_slab_ranges_cover_and_balance():
    for nx, P in ((800, 8), (10, 3), (7, 7)):
        r = M.slab_ranges(nx, P)
        assert r[0][0] == 0 and r[-1][1] == nx
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_mechanical_law_parameter_formulas():
    # linearElastic.C:62-133
    L = K.mechanical_law("linearElastic", E=200e9, nu=0.3)
    assert L.mu == pytest.approx(200e9 / 2.6)
    assert L.K == pytest.approx(200e9 / (3 * 0.4))
    with pytest.raises(ValueError):
        K.mechanical_law("linearElastic", E=1.0, nu=0.6)
    with pytest.raises(ValueError):
        K.mechanical_law("linearElastic", E=1.0)
    L2 = K.mechanical_law("linearElastic", mu=L.mu, K=L.K)
    assert L2.lambda_ == pytest.approx(L.lambda_, rel=1e-12)
    # neoHookeanElastic.C:51-85: K = lambda + 2/3 mu
    L3 = K.mechanical_law("neoHookeanElastic", E=3e6, nu=0.3)
    assert L3.K == pytest.approx(0.3 * 3e6 / (1.3 * 0.4) + 2 / 3 * L3.mu)
    with pytest.raises(ValueError):
        K.mechanical_law("neoHookeanElasticMisesPlastic", E=1.0, nu=0.3)       # no table
    with pytest.raises(KeyError):
        K.mechanical_law("noSuchLaw", E=1.0, nu=0.3)


def test_solid_model_selection_table_rejects_unknown_names():
    from solids4foam_b200.solid_model import SolidModel
    with pytest.raises(KeyError, match="Unknown solidModel type"):
        SolidModel.New(cases.cantilever(2, 1, 1), solidModel="vertexCentredLinearGeometry")


def test_default_controls_are_the_reference_defaults():
    c = K.default_controls()
    assert (c.nCorrectors, c.solutionTolerance, c.alternativeTolerance, c.materialTolerance) == (10000, 1e-6, 1e-7, 1e-5)
    assert c.stabilisation == K.STAB_RHIE_CHOW and c.stabScaleFactor == 0.1
    assert (c.tolerance, c.relTol, c.maxIter) == (1e-9, 0.1, 1000)
    with pytest.raises(KeyError):
        K.default_controls(noSuchKey=1)


def test_fvmesh_geometry_under_an_affine_map_follows_the_exact_transformation_rules():
    """A check of mesh.py's geometry that shares no formula with it (the oracle and the GPU path both take their geometry from
    mesh.py, so parity cannot see an error there): under x -> A x + b a box mesh keeps planar faces, and continuum kinematics
    gives every quantity of the mapped mesh exactly from the unmapped one -- volumes scale with det A, centroids map like points,
    area vectors follow Nanson's formula det(A) A^-T Sf, a mid-way face keeps the weight 1/2, and the delta coefficient is
    1 / (n . A d0)."""
    A = np.array([[1.1, 0.35, -0.2], [0.15, 0.9, 0.25], [-0.1, 0.3, 1.2]])
    b = np.array([0.3, -0.2, 0.5])
    m0 = M.hex_box_general(5, 4, 3, 2.0, 1.0, 1.5)
    m1 = M.hex_box_general(5, 4, 3, 2.0, 1.0, 1.5, point_map=lambda p: p @ A.T + b)
    detA = np.linalg.det(A)
    AinvT = np.linalg.inv(A).T
    assert detA > 0 and not m1.is_orthogonal()
    assert np.abs(m1.V - detA * m0.V).max() < 1e-13
    assert np.abs(m1.C - (m0.C @ A.T + b)).max() < 1e-13
    assert np.abs(m1.Cf - (m0.Cf @ A.T + b)).max() < 1e-13
    Sf = detA * (m0.Sf @ AinvT.T)                                   # Nanson: Sf = det(A) A^-T Sf0
    assert np.abs(m1.Sf - Sf).max() < 1e-13
    assert np.abs(m1.magSf - np.linalg.norm(Sf, axis=1)).max() < 1e-13
    F = m1.nInternalFaces
    assert np.abs(m1.weights[:F] - 0.5).max() < 1e-13               # uniform spacing: the face is mid-way, affine maps keep ratios
    n = Sf / np.linalg.norm(Sf, axis=1)[:, None]
    d0 = np.concatenate([m0.C[m0.neighbour] - m0.C[m0.owner], m0.Cf[F:] - m0.C[m0.faceCells]])
    d = d0 @ A.T
    nd = np.einsum("ij,ij->i", n, d)
    assert nd.min() > 0.05 * np.linalg.norm(d, axis=1).max()         # the limiter of the delta coefficients is not active here
    assert np.abs(m1.nonOrthDeltaCoeffs - 1.0 / nd).max() < 1e-12 * (1.0 / nd).max()
    corr = n[:F] - d[:F] / nd[:F, None]
    assert np.abs(m1.nonOrthCorrVec[:F] - corr).max() < 1e-12
    assert np.abs(np.einsum("ij,ij->i", m1.nonOrthCorrVec[:F], n[:F])).max() < 1e-12     # n . corr = 1 - (n . d) / (n . d) = 0


@pytest.mark.parametrize("which", ["warped", "plateHole", "notchedBar"])
def test_fvmesh_cells_are_closed_and_volumes_add_up(which):
    """Formula-independent identities on non-planar meshes: the outward area vectors of every cell sum to zero, and the cell
    volumes sum to the volume enclosed by the boundary faces, (1/3) sum_b Sf_b . Cf_b (divergence theorem on the triangulated
    boundary surface)."""
    if which == "warped":
        m = M.hex_box_general(6, 5, 4, 2.0, 1.0, 1.0, point_map=lambda p: p + 0.05 * np.sin(3.0 * p[:, [1, 2, 0]]))
    elif which == "plateHole":
        m = M.plate_hole(refine=1)
    else:
        m = cases.notched_bar(12, 4, 4).mesh
    F = m.nInternalFaces
    tot = np.zeros((m.nCells, 3))
    np.add.at(tot, m.owner, m.Sf[:F]); np.add.at(tot, m.neighbour, -m.Sf[:F])
    if which != "plateHole":                  # 2-D case: the faces of the empty patches are not part of the fvMesh
        np.add.at(tot, m.faceCells, m.Sf[F:])
        assert np.abs(tot).max() < 1e-14 * m.magSf.max() * 10
        vol = np.einsum("ij,ij->", m.Sf[F:], m.Cf[F:]) / 3.0
        assert abs(m.V.sum() - vol) < 1e-3 * vol       # non-planar faces: Cf is the area-magnitude-weighted centroid, not the exact first moment
    else:
        np.add.at(tot, m.faceCells, m.Sf[F:])
        assert np.abs(tot[:, :2]).max() < 1e-13        # closed in the plane; the z faces are the empty patches
        assert abs(m.V.sum() - 0.5 * (4.0 - np.pi * 0.25 / 4.0)) < 1e-3 * m.V.sum()    # quarter plate 2 x 2 minus a quarter disc r = 0.5, thickness 0.5 (polygonal hole)


def test_move_points_keeps_the_processor_identity_of_a_decomposed_mesh(tmp_path):
    """mesh.move_points on one processor's part of a decomposed mesh (the host mirror that follows the device-side mesh
    motion): rank, number of ranks, global cell numbers and processor patches survive, the geometry is that of the moved points."""
    from solids4foam_b200 import foam_io as IO
    c = cases.cantilever(6, 3, 2, general=True, L=2.0)
    IO.write_case(str(tmp_path), c)
    IO.decompose_case(str(tmp_path), 2)
    sent = {}
    for r in range(2):
        def collect(send, r=r):
            for q, a in send.items():
                sent[(r, q)] = a
            return {q: a + 1.0 for q, a in send.items()}     # placeholder centres for this first pass (it only collects)
        IO.read_poly_mesh(str(tmp_path / f"processor{r}" / "constant" / "polyMesh"), rank=r, nRanks=2, exchange=collect)
    m = IO.read_decomposed_case(str(tmp_path), 0, 2, lambda send: {q: sent[(q, 0)] for q in send}).mesh
    assert m.nRanks == 2 and any(p.kind == M.PROCESSOR for p in m.patches)
    shift = np.array([0.1, -0.2, 0.05])
    moved = M.move_points(m, m.points + shift)
    assert moved.rank == 0 and moved.nRanks == 2 and np.array_equal(moved.cellGlobal, m.cellGlobal)
    assert [(p.name, p.kind, p.size, p.nbr_rank) for p in moved.patches] == [(p.name, p.kind, p.size, p.nbr_rank) for p in m.patches]
    assert np.allclose(moved.C, m.C + shift) and np.allclose(moved.V, m.V) and np.allclose(moved.Sf, m.Sf)
