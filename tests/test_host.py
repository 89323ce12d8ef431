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
