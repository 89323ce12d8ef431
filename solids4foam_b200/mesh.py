"""Host-side finite-volume mesh in OpenFOAM (polyMesh / fvMesh) conventions.

This module plays the role that ``blockMesh`` + ``fvMesh`` geometry + ``decomposePar``
play for the reference: it produces, for synthetic cases, exactly the arrays the
solids4foam plugin would read off an ``fvMesh`` and hand to the C-ABI
(``include/s4fgpu.h``): ``owner/neighbour`` in upper-triangular order, patch
``faceCells``, and the geometric fields ``C, V, Sf, magSf, Cf, weights,
nonOrthDeltaCoeffs, nonOrthCorrectionVectors``.

Conventions (all [OF-ext], i.e. OpenFOAM library behaviour restated from its
published algorithms):

* cells of a structured block are numbered ``i + nx*(j + ny*k)``;
* internal faces are sorted by owner then neighbour, owner < neighbour, and the
  face area vector ``Sf`` points from owner to neighbour; boundary faces follow,
  patch by patch, ``Sf`` pointing out of the domain;
* face centres / areas by triangle decomposition about the vertex average, cell
  centres / volumes by pyramid decomposition about the face-centre average
  (``primitiveMesh::makeFaceCentresAndAreas / makeCellCentresAndVols``);
* ``weights = |Sf.(C_N-Cf)| / (|Sf.(Cf-C_P)| + |Sf.(C_N-Cf)|)``,
  ``nonOrthDeltaCoeffs = 1/max(n.d, 0.05|d|)``,
  ``nonOrthCorrectionVectors = n - d*nonOrthDeltaCoeffs`` (``surfaceInterpolation``);
* on non-coupled boundary faces ``d = Cf - C_P`` and the patch ``deltaCoeffs`` is
  ``1/max(n.d, 0.05|d|)``, the value for which the solids4foam boundary
  non-orthogonal corrections (``patchCorrectionVectors.C:24-36``) are consistent;
* ``empty`` patches (2-D cases) carry no fv faces: they are used for the cell
  geometry and then dropped, and their direction is flagged in ``solutionD``.

Processor patches (``decomposePar`` analogue) keep outward ``Sf`` on both sides, the
same face ordering on both sides, and carry the neighbour cell centres so that a
processor face is treated exactly like an internal face whose neighbour is remote.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

# patch kinds (mesh level)
PATCH = 0
EMPTY = 1
SYMMETRY_PLANE = 2
PROCESSOR = 3


@dataclass
class PatchInfo:
    name: str
    kind: int
    start: int          # offset into the boundary-face arrays (0-based, after internal faces)
    size: int
    nbr_rank: int = -1  # for PROCESSOR patches


@dataclass
class FvMesh:
    nCells: int
    owner: np.ndarray            # [F] int32 (internal faces)
    neighbour: np.ndarray        # [F] int32
    faceCells: np.ndarray        # [B] int32
    patches: List[PatchInfo]
    C: np.ndarray                # [N,3]
    V: np.ndarray                # [N]
    Sf: np.ndarray               # [F+B,3]
    magSf: np.ndarray            # [F+B]
    Cf: np.ndarray               # [F+B,3]
    weights: np.ndarray          # [F+B]  (1 on non-coupled boundary faces)
    nonOrthDeltaCoeffs: np.ndarray   # [F+B]
    nonOrthCorrVec: np.ndarray       # [F+B,3] (0 on non-coupled boundary faces)
    CnbrB: np.ndarray            # [B,3] neighbour cell centre on processor faces, Cf elsewhere
    solutionD: np.ndarray        # [3] int32, 1 = solved direction, 0 = empty direction
    cellGlobal: Optional[np.ndarray] = None   # [N] global cell ids for decomposed meshes
    rank: int = 0
    nRanks: int = 1
    meta: Dict = field(default_factory=dict)
    # polyMesh points() / faces() (quads) of the fv faces above, in the same order (internal faces, then the kept
    # boundary faces); present for meshes made by the general builder.  ``topo`` keeps what a geometry rebuild after
    # mesh motion needs (all faces incl. those of empty patches).
    points: Optional[np.ndarray] = None      # [nPoints,3]
    faces: Optional[np.ndarray] = None       # [F+B,4] int32
    topo: Optional[Dict] = None

    @property
    def nInternalFaces(self) -> int:
        return int(self.owner.shape[0])

    @property
    def nBoundaryFaces(self) -> int:
        return int(self.faceCells.shape[0])

    def patch(self, name: str) -> PatchInfo:
        for p in self.patches:
            if p.name == name:
                return p
        raise KeyError(name)

    def patch_slice(self, name: str) -> slice:
        p = self.patch(name)
        return slice(p.start, p.start + p.size)

    def boundary_normals(self) -> np.ndarray:
        F = self.nInternalFaces
        return self.Sf[F:] / self.magSf[F:, None]

    def is_orthogonal(self, tol: float = 1e-12) -> bool:
        return bool(np.all(np.abs(self.nonOrthCorrVec) < tol))


# ----------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------

def _bincount3(idx: np.ndarray, w: np.ndarray, n: int) -> np.ndarray:
    out = np.empty((n, 3))
    for c in range(3):
        out[:, c] = np.bincount(idx, weights=w[:, c], minlength=n)
    return out


def face_centres_and_areas(points: np.ndarray, faces: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Quad-face centres and area vectors, triangle decomposition about the vertex average
    ([OF-ext] ``primitiveMesh::makeFaceCentresAndAreas``)."""
    p = points[faces]                        # [nF,4,3]
    fc = p.mean(axis=1)
    sumN = np.zeros_like(fc)
    sumA = np.zeros(fc.shape[0])
    sumAc = np.zeros_like(fc)
    for e in range(4):
        a0 = p[:, e]
        a1 = p[:, (e + 1) % 4]
        c = a0 + a1 + fc
        n = np.cross(a1 - a0, fc - a0)
        a = np.sqrt(np.einsum("ij,ij->i", n, n))
        sumN += n
        sumA += a
        sumAc += a[:, None] * c
    ctr = np.where(sumA[:, None] > 1e-300, sumAc / (3.0 * np.maximum(sumA, 1e-300))[:, None], fc)
    return ctr, 0.5 * sumN


def cell_centres_and_volumes(nCells: int, fCtrs: np.ndarray, fAreas: np.ndarray,
                             own_all: np.ndarray, nei: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """[OF-ext] ``primitiveMesh::makeCellCentresAndVols``: pyramids about the face-centre average."""
    F = nei.shape[0]
    cEst = _bincount3(own_all, fCtrs, nCells) + _bincount3(nei, fCtrs[:F], nCells)
    nFaces = np.bincount(own_all, minlength=nCells) + np.bincount(nei, minlength=nCells)
    cEst /= nFaces[:, None]
    pyrO = np.einsum("ij,ij->i", fAreas, fCtrs - cEst[own_all])
    pcO = 0.75 * fCtrs + 0.25 * cEst[own_all]
    pyrN = np.einsum("ij,ij->i", fAreas[:F], cEst[nei] - fCtrs[:F])
    pcN = 0.75 * fCtrs[:F] + 0.25 * cEst[nei]
    vol = np.bincount(own_all, weights=pyrO, minlength=nCells) + np.bincount(nei, weights=pyrN, minlength=nCells)
    ctr = _bincount3(own_all, pyrO[:, None] * pcO, nCells) + _bincount3(nei, pyrN[:, None] * pcN, nCells)
    ctr /= vol[:, None]
    return ctr, vol / 3.0


def _finish_mesh(nCells, owner_int, neighbour, faceCells, patches, C, V, Sf, Cf, CnbrB_proc,
                 solutionD, **kw) -> FvMesh:
    """Interpolation geometry from (C, Sf, Cf): weights, nonOrthDeltaCoeffs, correction vectors."""
    F = owner_int.shape[0]
    B = faceCells.shape[0]
    magSf = np.sqrt(np.einsum("ij,ij->i", Sf, Sf))
    n = Sf / magSf[:, None]
    weights = np.ones(F + B)
    nod = np.empty(F + B)
    corr = np.zeros((F + B, 3))
    CnbrB = Cf[F:].copy()
    coupled = np.zeros(B, dtype=bool)
    for p in patches:
        if p.kind == PROCESSOR:
            coupled[p.start:p.start + p.size] = True
    if CnbrB_proc is not None:
        CnbrB[coupled] = CnbrB_proc[coupled]
    # internal faces
    own_all = np.concatenate([owner_int, faceCells])
    Cn = np.concatenate([C[neighbour], CnbrB])      # "other side" point: neighbour centre / Cf
    Cp = C[own_all]
    d = Cn - Cp
    # weights on internal + coupled faces
    sel = np.concatenate([np.ones(F, bool), coupled])
    SfdOwn = np.abs(np.einsum("ij,ij->i", Sf, Cf - Cp))
    SfdNei = np.abs(np.einsum("ij,ij->i", Sf, Cn - Cf))
    with np.errstate(invalid="ignore", divide="ignore"):
        w = SfdNei / (SfdOwn + SfdNei)
    weights[sel] = w[sel]
    magd = np.sqrt(np.einsum("ij,ij->i", d, d))
    nd = np.einsum("ij,ij->i", n, d)
    nod[:] = 1.0 / np.maximum(nd, 0.05 * magd)
    corr[sel] = (n - d * nod[:, None])[sel]
    return FvMesh(nCells=nCells, owner=owner_int.astype(np.int32), neighbour=neighbour.astype(np.int32),
                  faceCells=faceCells.astype(np.int32), patches=patches, C=C, V=V, Sf=Sf, magSf=magSf,
                  Cf=Cf, weights=weights, nonOrthDeltaCoeffs=nod, nonOrthCorrVec=corr, CnbrB=CnbrB,
                  solutionD=np.asarray(solutionD, dtype=np.int32), **kw)


# ----------------------------------------------------------------------------
# general builder: hexes -> polyMesh (used for multi-block / unstructured-numbered cases)
# ----------------------------------------------------------------------------

# OpenFOAM hex cell model: vertex order 0-3 bottom (counter-clockwise seen from top), 4-7 top.
# Faces with outward-pointing normals:
_HEX_FACES = np.array([
    [0, 4, 7, 3],   # x-
    [1, 2, 6, 5],   # x+
    [0, 1, 5, 4],   # y-
    [3, 7, 6, 2],   # y+
    [0, 3, 2, 1],   # z-
    [4, 5, 6, 7],   # z+
])


def poly_mesh_from_hexes(points: np.ndarray, hexes: np.ndarray,
                         classify: Callable[[np.ndarray, np.ndarray], np.ndarray],
                         patch_defs: Sequence[Tuple[str, int]],
                         cell_perm: Optional[np.ndarray] = None) -> FvMesh:
    """Build an fvMesh from a hex connectivity list.

    ``classify(Cf, n)`` returns, for every boundary face, the index into ``patch_defs``
    (name, kind).  ``cell_perm`` optionally renumbers the cells (new = perm[old]) to exercise
    unstructured numbering.
    """
    points = np.asarray(points, dtype=np.float64)
    hexes = np.asarray(hexes, dtype=np.int64)
    if cell_perm is not None:
        inv = np.empty_like(cell_perm)
        inv[cell_perm] = np.arange(cell_perm.size)
        hexes = hexes[inv]
    nCells = hexes.shape[0]
    allf = hexes[:, _HEX_FACES].reshape(-1, 4)              # [6N,4] outward from its cell
    cellOf = np.repeat(np.arange(nCells), 6)
    key = np.sort(allf, axis=1)
    _, inv_idx, counts = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    inv_idx = inv_idx.reshape(-1)
    order = np.argsort(inv_idx, kind="stable")
    sorted_ids = inv_idx[order]
    first = np.r_[True, sorted_ids[1:] != sorted_ids[:-1]]
    starts = np.nonzero(first)[0]
    cnt = counts
    # internal: faces shared by two cells
    int_u = np.nonzero(cnt == 2)[0]
    a = order[starts[int_u]]
    b = order[starts[int_u] + 1]
    ca, cb = cellOf[a], cellOf[b]
    own_is_a = ca < cb
    own_f = np.where(own_is_a, a, b)
    own = cellOf[own_f]
    nei = np.where(own_is_a, cb, ca)
    srt = np.lexsort((nei, own))
    own, nei, own_f = own[srt], nei[srt], own_f[srt]
    faces_int = allf[own_f]
    # boundary
    bnd_u = np.nonzero(cnt == 1)[0]
    bf = order[starts[bnd_u]]
    faces_b = allf[bf]
    cells_b = cellOf[bf]
    fc_b, fa_b = face_centres_and_areas(points, faces_b)
    nb = fa_b / np.linalg.norm(fa_b, axis=1)[:, None]
    pid = np.asarray(classify(fc_b, nb), dtype=np.int64)
    if np.any(pid < 0):
        raise ValueError("unclassified boundary faces")
    srtb = np.lexsort((cells_b, pid))
    faces_b, cells_b, pid = faces_b[srtb], cells_b[srtb], pid[srtb]
    faces_all = np.concatenate([faces_int, faces_b])
    fCtrs, fAreas = face_centres_and_areas(points, faces_all)
    own_all = np.concatenate([own, cells_b])
    C, V = cell_centres_and_volumes(nCells, fCtrs, fAreas, own_all, nei)
    # drop empty patches from the fv boundary
    F = own.shape[0]
    keep = np.ones(cells_b.shape[0], bool)
    solutionD = np.ones(3, dtype=np.int32)
    patches: List[PatchInfo] = []
    start = 0
    for ip, (name, kind) in enumerate(patch_defs):
        m = pid == ip
        if kind == EMPTY:
            keep[m] = False
            if m.any():
                nrm = np.abs(nb[srtb][m]).mean(axis=0)
                solutionD[int(np.argmax(nrm))] = 0
            continue
        sz = int(m.sum())
        patches.append(PatchInfo(name, kind, start, sz))
        start += sz
    kb = np.nonzero(keep)[0]
    Sf = np.concatenate([fAreas[:F], fAreas[F:][kb]])
    Cf = np.concatenate([fCtrs[:F], fCtrs[F:][kb]])
    mesh = _finish_mesh(nCells, own, nei, cells_b[kb], patches, C, V, Sf, Cf, None, solutionD,
                        cellGlobal=np.arange(nCells, dtype=np.int64))
    mesh.points = points.copy()
    mesh.faces = np.concatenate([faces_int, faces_b[kb]]).astype(np.int32)
    mesh.topo = dict(faces_all=faces_all, own_all=own_all, kb=kb)
    return mesh


def move_points(mesh: FvMesh, new_points: np.ndarray) -> FvMesh:
    """``fvMesh::movePoints`` for a mesh of the general builder: same topology, geometry recomputed from the new
    points ([OF-ext] primitiveMesh face/cell geometry + surfaceInterpolation weights, deltaCoeffs, correction
    vectors), as solidModel::moveMesh does at the end of an updated-Lagrangian step (solidModel.C:2008-2148)."""
    if mesh.points is None or mesh.topo is None:
        raise ValueError("move_points needs a mesh with points/faces (general builder)")
    t = mesh.topo
    pts = np.asarray(new_points, dtype=np.float64)
    F = mesh.nInternalFaces
    fCtrs, fAreas = face_centres_and_areas(pts, t["faces_all"])
    C, V = cell_centres_and_volumes(mesh.nCells, fCtrs, fAreas, t["own_all"], mesh.neighbour.astype(np.int64))
    Sf = np.concatenate([fAreas[:F], fAreas[F:][t["kb"]]])
    Cf = np.concatenate([fCtrs[:F], fCtrs[F:][t["kb"]]])
    # a processor's part of a decomposed mesh: this host mirror (file output, FSI points) keeps the neighbour cell centres of
    # before the move for the interpolation geometry of its processor faces; the device computes its own from a halo exchange
    new = _finish_mesh(mesh.nCells, mesh.owner.astype(np.int64), mesh.neighbour.astype(np.int64), mesh.faceCells, mesh.patches,
                       C, V, Sf, Cf, mesh.CnbrB if mesh.nRanks > 1 else None, mesh.solutionD, cellGlobal=mesh.cellGlobal,
                       rank=mesh.rank, nRanks=mesh.nRanks)
    new.points = pts.copy(); new.faces = mesh.faces; new.topo = t; new.meta = dict(mesh.meta)
    return new


# ----------------------------------------------------------------------------
# structured hex box (fast path, any size; optional x-slab of a decomposed box)
# ----------------------------------------------------------------------------

def hex_box(nx: int, ny: int, nz: int, L: float = 1.0, H: float = 1.0, W: float = 1.0,
            i0: int = 0, i1: Optional[int] = None, rank: int = 0, nRanks: int = 1,
            point_map: Optional[Callable[[np.ndarray], np.ndarray]] = None,
            names: Sequence[str] = ("xMin", "xMax", "yMin", "yMax", "zMin", "zMax"),
            kinds: Sequence[int] = (PATCH,) * 6) -> FvMesh:
    """``blockMesh`` analogue for one hex block ``nx x ny x nz`` over ``[0,L]x[0,H]x[0,W]``.

    With ``i0,i1`` given, only the x-slab of cells ``i0 <= i < i1`` is generated and the cuts
    become ``processor`` patches (``decomposePar`` with ``simple (P 1 1)``), with the neighbour
    cell centres filled in.  ``point_map`` moves the points (e.g. the notched bar).
    """
    if i1 is None:
        i1 = nx
    if point_map is None:
        return _hex_box_rectilinear(nx, ny, nz, L, H, W, i0, i1, rank, nRanks, names, kinds)
    mx = i1 - i0
    has_lo = i0 > 0
    has_hi = i1 < nx
    # points of the slab plus one extra layer of cells on cut sides (for neighbour centres)
    e0 = i0 - (1 if has_lo else 0)
    e1 = i1 + (1 if has_hi else 0)
    ex = e1 - e0
    xs = np.linspace(0.0, L, nx + 1)[e0:e1 + 1]
    ys = np.linspace(0.0, H, ny + 1)
    zs = np.linspace(0.0, W, nz + 1)
    px, py, pz = ex + 1, ny + 1, nz + 1
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    points = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    if point_map is not None:
        points = point_map(points)

    def pid(i, j, k):
        return (i + px * (j + py * k)).astype(np.int64)

    # extended cell grid (ex x ny x nz), then restrict
    def cell_grid(nxx):
        k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nxx), indexing="ij")
        return i.ravel(), j.ravel(), k.ravel()

    off = i0 - e0
    ci, cj, ck = cell_grid(mx)              # local cells, numbered i + mx*(j + ny*k)
    nCells = mx * ny * nz
    cid = np.arange(nCells, dtype=np.int64)
    gi = ci + off                           # i index in the extended point grid
    hasx = ci < mx - 1
    hasy = cj < ny - 1
    hasz = ck < nz - 1
    cnt = hasx.astype(np.int64) + hasy + hasz
    start = np.cumsum(cnt) - cnt
    F = int(cnt.sum())
    owner = np.empty(F, dtype=np.int64)
    neighbour = np.empty(F, dtype=np.int64)
    faces = np.empty((F, 4), dtype=np.int64)
    # x+ faces
    sel = hasx
    f = start[sel]
    owner[f] = cid[sel]
    neighbour[f] = cid[sel] + 1
    i, j, k = gi[sel] + 1, cj[sel], ck[sel]
    faces[f] = np.stack([pid(i, j, k), pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i, j, k + 1)], axis=1)
    sel = hasy
    f = start[sel] + hasx[sel]
    owner[f] = cid[sel]
    neighbour[f] = cid[sel] + mx
    i, j, k = gi[sel], cj[sel] + 1, ck[sel]
    faces[f] = np.stack([pid(i, j, k), pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j, k)], axis=1)
    sel = hasz
    f = start[sel] + hasx[sel] + hasy[sel]
    owner[f] = cid[sel]
    neighbour[f] = cid[sel] + mx * ny
    i, j, k = gi[sel], cj[sel], ck[sel] + 1
    faces[f] = np.stack([pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k)], axis=1)

    # boundary faces, outward normals
    def side(which):
        if which == 0:      # x-min of slab
            k, j = np.meshgrid(np.arange(nz), np.arange(ny), indexing="ij")
            j, k = j.ravel(), k.ravel()
            i = np.full_like(j, off)
            cells = 0 + mx * (j + ny * k)
            fv = np.stack([pid(i, j, k), pid(i, j, k + 1), pid(i, j + 1, k + 1), pid(i, j + 1, k)], axis=1)
        elif which == 1:    # x-max
            k, j = np.meshgrid(np.arange(nz), np.arange(ny), indexing="ij")
            j, k = j.ravel(), k.ravel()
            i = np.full_like(j, off + mx)
            cells = (mx - 1) + mx * (j + ny * k)
            fv = np.stack([pid(i, j, k), pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i, j, k + 1)], axis=1)
        elif which == 2:    # y-min
            k, i = np.meshgrid(np.arange(nz), np.arange(mx), indexing="ij")
            i, k = i.ravel(), k.ravel()
            j = np.zeros_like(i)
            cells = i + mx * (0 + ny * k)
            ii = i + off
            fv = np.stack([pid(ii, j, k), pid(ii + 1, j, k), pid(ii + 1, j, k + 1), pid(ii, j, k + 1)], axis=1)
        elif which == 3:    # y-max
            k, i = np.meshgrid(np.arange(nz), np.arange(mx), indexing="ij")
            i, k = i.ravel(), k.ravel()
            j = np.full_like(i, ny)
            cells = i + mx * ((ny - 1) + ny * k)
            ii = i + off
            fv = np.stack([pid(ii, j, k), pid(ii, j, k + 1), pid(ii + 1, j, k + 1), pid(ii + 1, j, k)], axis=1)
        elif which == 4:    # z-min
            j, i = np.meshgrid(np.arange(ny), np.arange(mx), indexing="ij")
            i, j = i.ravel(), j.ravel()
            k = np.zeros_like(i)
            cells = i + mx * (j + ny * 0)
            ii = i + off
            fv = np.stack([pid(ii, j, k), pid(ii, j + 1, k), pid(ii + 1, j + 1, k), pid(ii + 1, j, k)], axis=1)
        else:               # z-max
            j, i = np.meshgrid(np.arange(ny), np.arange(mx), indexing="ij")
            i, j = i.ravel(), j.ravel()
            k = np.full_like(i, nz)
            cells = i + mx * (j + ny * (nz - 1))
            ii = i + off
            fv = np.stack([pid(ii, j, k), pid(ii + 1, j, k), pid(ii + 1, j + 1, k), pid(ii, j + 1, k)], axis=1)
        return cells.astype(np.int64), fv

    side_cells = []
    side_faces = []
    for s in range(6):
        c, fv = side(s)
        side_cells.append(c)
        side_faces.append(fv)

    # geometry on all faces (incl. empties) -> cell centres / volumes
    faces_all = np.concatenate([faces] + side_faces)
    own_all = np.concatenate([owner] + side_cells)
    fCtrs, fAreas = face_centres_and_areas(points, faces_all)
    C, V = cell_centres_and_volumes(nCells, fCtrs, fAreas, own_all, neighbour)

    # neighbour-slab cell centres for the processor patches: geometry of the extra cell layer
    def layer_centres(i_ext):
        """centres of the cells with extended-grid index i_ext (all j,k), from their 6 faces."""
        k, j = np.meshgrid(np.arange(nz), np.arange(ny), indexing="ij")
        j, k = j.ravel(), k.ravel()
        i = np.full_like(j, i_ext)
        v = [pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k),
             pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j + 1, k + 1), pid(i, j + 1, k + 1)]
        hexes = np.stack(v, axis=1)
        ff = hexes[:, _HEX_FACES].reshape(-1, 4)
        fc, fa = face_centres_and_areas(points, ff)
        oc = np.repeat(np.arange(hexes.shape[0]), 6)
        cc, _ = cell_centres_and_volumes(hexes.shape[0], fc, fa, oc, np.zeros(0, dtype=np.int64))
        return cc

    # assemble fv boundary: physical patches first (in the order given), then processor patches
    patches: List[PatchInfo] = []
    b_cells, b_Sf, b_Cf, b_Cnbr = [], [], [], []
    solutionD = np.ones(3, dtype=np.int32)
    offs = np.cumsum([0] + [sf.shape[0] for sf in side_faces])
    startB = 0

    def add(name, kind, s, nbr_rank=-1, cnbr=None):
        nonlocal startB
        sl = slice(F + offs[s], F + offs[s + 1])
        n = side_cells[s].shape[0]
        patches.append(PatchInfo(name, kind, startB, n, nbr_rank))
        b_cells.append(side_cells[s])
        b_Sf.append(fAreas[sl])
        b_Cf.append(fCtrs[sl])
        b_Cnbr.append(cnbr if cnbr is not None else fCtrs[sl])
        startB += n

    for s in range(6):
        if s == 0 and has_lo:
            continue
        if s == 1 and has_hi:
            continue
        if kinds[s] == EMPTY:
            solutionD[s // 2] = 0
            continue
        add(names[s], kinds[s], s)
    if has_lo:
        add(f"procBoundary{rank}to{rank - 1}", PROCESSOR, 0, rank - 1, layer_centres(0))
    if has_hi:
        add(f"procBoundary{rank}to{rank + 1}", PROCESSOR, 1, rank + 1, layer_centres(ex - 1))

    faceCells = np.concatenate(b_cells) if b_cells else np.zeros(0, dtype=np.int64)
    Sf = np.concatenate([fAreas[:F]] + b_Sf)
    Cf = np.concatenate([fCtrs[:F]] + b_Cf)
    Cnbr = np.concatenate(b_Cnbr) if b_Cnbr else np.zeros((0, 3))
    # global ids
    gk, gj, gI = ck, cj, ci + i0
    cellGlobal = (gI + nx * (gj + ny * gk)).astype(np.int64)
    return _finish_mesh(nCells, owner, neighbour, faceCells, patches, C, V, Sf, Cf, Cnbr, solutionD,
                        cellGlobal=cellGlobal, rank=rank, nRanks=nRanks,
                        meta=dict(nx=nx, ny=ny, nz=nz, L=L, H=H, W=W, i0=i0, i1=i1))


def _hex_box_rectilinear(nx, ny, nz, L, H, W, i0, i1, rank, nRanks, names, kinds) -> FvMesh:
    """Closed-form geometry of an undistorted box (same numbering as ``hex_box``; used for the large
    benchmark meshes where building points and faces would dominate the run)."""
    mx = i1 - i0
    dx, dy, dz = L / nx, H / ny, W / nz
    has_lo, has_hi = i0 > 0, i1 < nx
    nCells = mx * ny * nz
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(mx), indexing="ij")
    ci, cj, ck = i.ravel(), j.ravel(), k.ravel()
    del i, j, k
    C = np.stack([(ci + i0 + 0.5) * dx, (cj + 0.5) * dy, (ck + 0.5) * dz], axis=1)
    V = np.full(nCells, dx * dy * dz)
    hasx, hasy, hasz = ci < mx - 1, cj < ny - 1, ck < nz - 1
    cnt = hasx.astype(np.int64) + hasy + hasz
    start = np.cumsum(cnt) - cnt
    F = int(cnt.sum())
    owner = np.empty(F, dtype=np.int32)
    neighbour = np.empty(F, dtype=np.int32)
    Sf = np.zeros((F, 3))
    Cf = np.empty((F, 3))
    cid = np.arange(nCells, dtype=np.int64)
    for d, (sel, off, stride, area) in enumerate(((hasx, 0, 1, dy * dz), (hasy, hasx, mx, dx * dz),
                                                  (hasz, hasx.astype(np.int64) + hasy, mx * ny, dx * dy))):
        f = start[sel] + (off[sel] if not isinstance(off, int) else off)
        owner[f] = cid[sel]
        neighbour[f] = cid[sel] + stride
        Sf[f, d] = area
        cf = C[sel].copy()
        cf[:, d] += 0.5 * (dx, dy, dz)[d]
        Cf[f] = cf
    # boundary sides
    def side(s):
        d, hi = s // 2, s % 2
        if d == 0:
            kk, jj = np.meshgrid(np.arange(nz), np.arange(ny), indexing="ij")
            jj, kk = jj.ravel(), kk.ravel()
            cells = ((mx - 1) if hi else 0) + mx * (jj + ny * kk)
        elif d == 1:
            kk, ii = np.meshgrid(np.arange(nz), np.arange(mx), indexing="ij")
            ii, kk = ii.ravel(), kk.ravel()
            cells = ii + mx * (((ny - 1) if hi else 0) + ny * kk)
        else:
            jj, ii = np.meshgrid(np.arange(ny), np.arange(mx), indexing="ij")
            ii, jj = ii.ravel(), jj.ravel()
            cells = ii + mx * (jj + ny * ((nz - 1) if hi else 0))
        area = (dy * dz, dx * dz, dx * dy)[d]
        sf = np.zeros((cells.size, 3))
        sf[:, d] = area if hi else -area
        cf = C[cells].copy()
        cf[:, d] += (0.5 if hi else -0.5) * (dx, dy, dz)[d]
        cn = C[cells].copy()
        cn[:, d] += (1.0 if hi else -1.0) * (dx, dy, dz)[d]
        return cells.astype(np.int64), sf, cf, cn
    patches: List[PatchInfo] = []
    b_cells, b_Sf, b_Cf, b_Cn = [], [], [], []
    solutionD = np.ones(3, dtype=np.int32)
    startB = 0

    def add(name, kind, s, nbr=-1):
        nonlocal startB
        cells, sf, cf, cn = side(s)
        patches.append(PatchInfo(name, kind, startB, cells.size, nbr))
        b_cells.append(cells); b_Sf.append(sf); b_Cf.append(cf)
        b_Cn.append(cn if kind == PROCESSOR else cf)
        startB += cells.size
    for s in range(6):
        if (s == 0 and has_lo) or (s == 1 and has_hi):
            continue
        if kinds[s] == EMPTY:
            solutionD[s // 2] = 0
            continue
        add(names[s], kinds[s], s)
    if has_lo:
        add(f"procBoundary{rank}to{rank - 1}", PROCESSOR, 0, rank - 1)
    if has_hi:
        add(f"procBoundary{rank}to{rank + 1}", PROCESSOR, 1, rank + 1)
    faceCells = np.concatenate(b_cells)
    SfA = np.concatenate([Sf] + b_Sf)
    CfA = np.concatenate([Cf] + b_Cf)
    Cnbr = np.concatenate(b_Cn)
    cellGlobal = ((ci + i0) + nx * (cj + ny * ck)).astype(np.int64)
    return _finish_mesh(nCells, owner, neighbour, faceCells, patches, C, V, SfA, CfA, Cnbr, solutionD,
                        cellGlobal=cellGlobal, rank=rank, nRanks=nRanks,
                        meta=dict(nx=nx, ny=ny, nz=nz, L=L, H=H, W=W, i0=i0, i1=i1))


def slab_ranges(nx: int, nRanks: int) -> List[Tuple[int, int]]:
    """x-slab cell ranges of ``decomposePar`` ``simple (P 1 1)`` on a structured block."""
    base, rem = divmod(nx, nRanks)
    out, s = [], 0
    for r in range(nRanks):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


def hex_box_decomposed(nx, ny, nz, L, H, W, rank, nRanks, **kw) -> FvMesh:
    i0, i1 = slab_ranges(nx, nRanks)[rank]
    return hex_box(nx, ny, nz, L, H, W, i0=i0, i1=i1, rank=rank, nRanks=nRanks, **kw)


# ----------------------------------------------------------------------------
# plate with a hole (tutorials/solids/linearElasticity/plateHole/constant/polyMesh/blockMeshDict)
# ----------------------------------------------------------------------------

def plate_hole(refine: int = 1, thickness: float = 0.5, cell_perm_seed: Optional[int] = None) -> FvMesh:
    """Quarter plate (2 x 2 m) with a hole of radius 0.5 m: the five blocks of the tutorial's
    blockMeshDict (10x10, 10x10, 20x10, 20x20, 10x20 cells at ``refine=1`` -> 1000 cells), arcs
    on the hole (r=0.5) and on the r=1 ring, one cell thick with ``empty`` front and back."""
    s = np.sqrt(0.5)
    v2 = {0: (0.5, 0), 1: (1, 0), 2: (2, 0), 3: (2, s), 4: (s, s), 5: (0.5 * s, 0.5 * s),
          6: (2, 2), 7: (s, 2), 8: (0, 2), 9: (0, 1), 10: (0, 0.5)}
    arcs = {(0, 5): 0.5, (5, 10): 0.5, (1, 4): 1.0, (4, 9): 1.0}
    blocks = [((5, 4, 9, 10), 10, 10), ((0, 1, 4, 5), 10, 10), ((1, 2, 3, 4), 20, 10),
              ((4, 3, 6, 7), 20, 20), ((9, 4, 7, 8), 10, 20)]

    def edge(a, b, t):
        pa, pb = np.array(v2[a], float), np.array(v2[b], float)
        r = arcs.get((a, b)) or arcs.get((b, a))
        if r is None:
            return pa[None, :] * (1 - t)[:, None] + pb[None, :] * t[:, None]
        th_a, th_b = np.arctan2(pa[1], pa[0]), np.arctan2(pb[1], pb[0])
        th = th_a * (1 - t) + th_b * t
        return np.stack([r * np.cos(th), r * np.sin(th)], axis=1)

    pts: List[np.ndarray] = []
    hexes: List[np.ndarray] = []
    npts = 0
    for (a, b, c, d), n1, n2 in blocks:
        n1 *= refine
        n2 *= refine
        t1 = np.linspace(0, 1, n1 + 1)
        t2 = np.linspace(0, 1, n2 + 1)
        e_ab, e_dc = edge(a, b, t1), edge(d, c, t1)
        e_ad, e_bc = edge(a, d, t2), edge(b, c, t2)
        # transfinite interpolation
        P = ((1 - t2)[:, None, None] * e_ab[None] + t2[:, None, None] * e_dc[None]
             + (1 - t1)[None, :, None] * e_ad[:, None] + t1[None, :, None] * e_bc[:, None]
             - ((1 - t1)[None, :, None] * (1 - t2)[:, None, None] * e_ab[0]
                + t1[None, :, None] * (1 - t2)[:, None, None] * e_ab[-1]
                + t1[None, :, None] * t2[:, None, None] * e_dc[-1]
                + (1 - t1)[None, :, None] * t2[:, None, None] * e_dc[0]))
        P2 = P.reshape(-1, 2)
        nb = P2.shape[0]
        pts.append(np.concatenate([np.c_[P2, np.zeros(nb)], np.c_[P2, np.full(nb, thickness)]]))
        jj, ii = np.meshgrid(np.arange(n2), np.arange(n1), indexing="ij")
        ii, jj = ii.ravel(), jj.ravel()

        def q(i, j):
            return npts + i + (n1 + 1) * j
        h = np.stack([q(ii, jj), q(ii + 1, jj), q(ii + 1, jj + 1), q(ii, jj + 1),
                      q(ii, jj) + nb, q(ii + 1, jj) + nb, q(ii + 1, jj + 1) + nb, q(ii, jj + 1) + nb], axis=1)
        hexes.append(h)
        npts += 2 * nb
    points = np.concatenate(pts)
    hexes_a = np.concatenate(hexes)
    # merge coincident points
    keyp = np.round(points / 1e-9).astype(np.int64)
    _, first_idx, inv = np.unique(keyp, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    points = points[first_idx]
    hexes_a = inv[hexes_a]

    patch_defs = [("left", SYMMETRY_PLANE), ("right", PATCH), ("down", SYMMETRY_PLANE), ("up", PATCH),
                  ("hole", PATCH), ("frontAndBack", EMPTY)]

    def classify(cf, n):
        out = np.full(cf.shape[0], -1, dtype=np.int64)
        out[np.abs(n[:, 2]) > 0.9] = 5
        r = np.hypot(cf[:, 0], cf[:, 1])
        free = out < 0
        out[free & (np.abs(cf[:, 0]) < 1e-9)] = 0
        out[free & (np.abs(cf[:, 0] - 2) < 1e-9)] = 1
        out[free & (np.abs(cf[:, 1]) < 1e-9)] = 2
        out[free & (np.abs(cf[:, 1] - 2) < 1e-9)] = 3
        out[free & (out < 0) & (r < 0.5 + 1e-6)] = 4
        return out

    perm = None
    if cell_perm_seed is not None:
        perm = np.random.default_rng(cell_perm_seed).permutation(hexes_a.shape[0])
    return poly_mesh_from_hexes(points, hexes_a, classify, patch_defs, cell_perm=perm)


def hex_box_general(nx, ny, nz, L=1.0, H=1.0, W=1.0, point_map=None, cell_perm_seed=None,
                    names=("xMin", "xMax", "yMin", "yMax", "zMin", "zMax"), kinds=(PATCH,) * 6) -> FvMesh:
    """The same box through the general hex builder (cross-check of ``hex_box``; optional random
    cell renumbering / distortion for unstructured-numbering tests)."""
    px, py = nx + 1, ny + 1
    Z, Y, X = np.meshgrid(np.linspace(0, W, nz + 1), np.linspace(0, H, ny + 1), np.linspace(0, L, nx + 1),
                          indexing="ij")
    points = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    ref_points = points.copy()
    if point_map is not None:
        points = point_map(points)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()

    def pid(i, j, k):
        return i + px * (j + py * k)
    hexes = np.stack([pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k),
                      pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j + 1, k + 1), pid(i, j + 1, k + 1)], axis=1)
    # classify on the undeformed box: recompute reference face centres via nearest reference lookup
    ref_mesh_pts = ref_points

    def classify(cf, n):
        # use dominant normal direction and position relative to the deformed bounding box
        out = np.full(cf.shape[0], -1, dtype=np.int64)
        lo, hi = points.min(axis=0), points.max(axis=0)
        ax = np.argmax(np.abs(n), axis=1)
        sgn = np.sign(n[np.arange(n.shape[0]), ax])
        out[:] = 2 * ax + (sgn > 0)
        return out
    perm = None
    if cell_perm_seed is not None:
        perm = np.random.default_rng(cell_perm_seed).permutation(hexes.shape[0])
    return poly_mesh_from_hexes(points, hexes, classify, list(zip(names, kinds)), cell_perm=perm)


# ----------------------------------------------------------------------------
# decomposePar restated: a serial polyMesh -> one polyMesh per processor
# ----------------------------------------------------------------------------
def decompose_poly(points: np.ndarray, faces: Sequence[Sequence[int]], owner: np.ndarray, neighbour: np.ndarray,
                   patches_raw: Sequence[Tuple[str, Dict]], cell_rank: np.ndarray) -> List[Dict]:
    """What ``decomposePar`` writes under ``processorN/constant/polyMesh`` ([OF-ext] domainDecomposition), for a given
    cell -> processor map (``method manual`` / the result of ``simple``): per processor

    * cells in ascending global order (``cellProcAddressing``), points in ascending global order,
    * internal faces = global internal faces with both cells on the processor, in the upper-triangular order of the local
      numbering; the original patches, each with its faces on this processor in global order; then one ``processor`` patch
      per neighbour processor (ascending), its faces in global face order on BOTH sides, which is what makes the two sides
      match face by face; the side that holds the global neighbour cell stores the face reversed (normal out of its cell).

    Returns per rank: points, faces (list of vertex lists), owner, neighbour, patches_raw [(name, dict)], cellGlobal,
    pointGlobal, faceGlobal (signed: -(f+1) for a reversed face), patchSel {patch: positions of its faces in the serial patch}."""
    owner = np.asarray(owner, dtype=np.int64); neighbour = np.asarray(neighbour, dtype=np.int64)
    cell_rank = np.asarray(cell_rank, dtype=np.int64)
    nI = neighbour.size
    nRanks = int(cell_rank.max()) + 1
    out = []
    for r in range(nRanks):
        cells = np.nonzero(cell_rank == r)[0]
        loc = -np.ones(cell_rank.size, dtype=np.int64); loc[cells] = np.arange(cells.size)
        ro, rn = cell_rank[owner[:nI]], cell_rank[neighbour]
        # internal faces of this processor, sorted by (local owner, local neighbour)
        fi = np.nonzero((ro == r) & (rn == r))[0]
        lo, ln = loc[owner[fi]], loc[neighbour[fi]]
        flip_i = lo > ln                       # local numbering keeps the global order here (ascending), so never true; kept for safety
        a, b = np.where(flip_i, ln, lo), np.where(flip_i, lo, ln)
        order = np.lexsort((b, a))
        face_ids = [fi[order]]; face_flip = [flip_i[order]]
        f_owner = [a[order]]; f_neigh = b[order]
        praw = []
        patch_sel = {}
        start = fi.size
        for name, d in patches_raw:
            s0, n0 = int(d["startFace"]), int(d["nFaces"])
            pf = np.arange(s0, s0 + n0)
            pf = pf[cell_rank[owner[pf]] == r]
            patch_sel[name] = pf - s0
            dd = dict(d); dd["nFaces"] = int(pf.size); dd["startFace"] = int(start)
            praw.append((name, dd))
            face_ids.append(pf); face_flip.append(np.zeros(pf.size, dtype=bool)); f_owner.append(loc[owner[pf]])
            start += pf.size
        # processor patches: cut faces, per neighbour processor, in global face order
        cut_o = np.nonzero((ro == r) & (rn != r))[0]       # I hold the owner: keep orientation
        cut_n = np.nonzero((rn == r) & (ro != r))[0]       # I hold the neighbour: reversed
        other = np.concatenate([rn[cut_o], ro[cut_n]])
        cutf = np.concatenate([cut_o, cut_n])
        mine = np.concatenate([loc[owner[cut_o]], loc[neighbour[cut_n]]])
        flipc = np.concatenate([np.zeros(cut_o.size, dtype=bool), np.ones(cut_n.size, dtype=bool)])
        for q in sorted(set(int(x) for x in other)):
            sel = np.nonzero(other == q)[0]
            sel = sel[np.argsort(cutf[sel], kind="stable")]
            praw.append((f"procBoundary{r}to{q}", dict(type="processor", nFaces=int(sel.size), startFace=int(start), myProcNo=r, neighbProcNo=q)))
            face_ids.append(cutf[sel]); face_flip.append(flipc[sel]); f_owner.append(mine[sel])
            start += sel.size
        gids = np.concatenate(face_ids); gflip = np.concatenate(face_flip)
        used = np.unique(np.fromiter((v for f in gids for v in faces[int(f)]), dtype=np.int64))
        ploc = -np.ones(points.shape[0], dtype=np.int64); ploc[used] = np.arange(used.size)
        lfaces = []
        for f, fl in zip(gids, gflip):
            vs = [int(ploc[v]) for v in faces[int(f)]]
            lfaces.append([vs[0]] + vs[:0:-1] if fl else vs)      # reversed face keeps its first vertex ([OF-ext] face::reverseFace)
        out.append(dict(points=points[used].copy(), faces=lfaces, owner=np.concatenate(f_owner), neighbour=f_neigh, patches_raw=praw,
                        cellGlobal=cells, pointGlobal=used, faceGlobal=np.where(gflip, -(gids + 1), gids + 1), patchSel=patch_sel))
    return out
