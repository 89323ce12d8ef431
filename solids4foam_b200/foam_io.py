"""On-disk formats either side of the hot path: OpenFOAM ASCII dictionaries, ``polyMesh``
(points / faces / owner / neighbour / boundary) and vol-field files, so that a case directory of the reference can
drive the GPU path without OpenFOAM, and results can be written back as time directories.

What the reference reads where:
  constant/solidProperties       solidModel + <model>Coeffs           SM/solidModel/solidModel.C:1138-1243, :1711-1727
  constant/mechanicalProperties  planeStress + mechanical ( law {} )  mechanicalModel.C, linearElastic.C:62-133, neoHookeanElastic.C:51-85
  system/fvSchemes, fvSolution   d2dt2 / grad schemes, solver D, relaxationFactors    [OF-ext] fvSchemes / fvSolution
  0/D boundaryField              fixedDisplacement, solidTraction, solidSymmetry, analyticalPlateHoleTraction
  constant/polyMesh              [OF-ext] polyMesh files; geometry as primitiveMesh computes it (mesh.py)

Host-side plumbing only: nothing numerical on the path happens here.
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import case as K
from . import mesh as M

_HEADER = """/*--------------------------------*- C++ -*----------------------------------*\\
| written by solids4foam_b200.foam_io                                         |
\\*---------------------------------------------------------------------------*/
FoamFile
{{
    version     2.0;
    format      ascii;
    class       {cls};
    location    "{loc}";
    object      {obj};
}}
// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //

"""


# ------------------------------------------------------------------------------------------------
# dictionary parser
# ------------------------------------------------------------------------------------------------
def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def _tokenize(text: str) -> List[str]:
    return re.findall(r'"[^"]*"|[{}();\[\]]|[^\s{}();\[\]]+', _strip_comments(text))


class _Parser:
    def __init__(self, toks: List[str]):
        self.t = toks
        self.i = 0

    def peek(self) -> Optional[str]:
        return self.t[self.i] if self.i < len(self.t) else None

    def next(self) -> str:
        tok = self.t[self.i]
        self.i += 1
        return tok

    def parse_dict_body(self, until: Optional[str]) -> Dict:
        out: Dict = {}
        while self.peek() is not None and self.peek() != until:
            key = self.next().strip('"')
            if self.peek() == "{":
                self.next()
                out[key] = self.parse_dict_body("}")
                self.next()
                continue
            vals = []
            while self.peek() not in (";", None):
                vals.append(self.parse_value())
            if self.peek() == ";":
                self.next()
            out[key] = vals[0] if len(vals) == 1 else vals
        return out

    def parse_value(self):
        tok = self.next()
        if tok == "(":
            items = []
            while self.peek() != ")":
                # a list of dictionaries: name { ... }
                if self.i + 1 < len(self.t) and self.t[self.i + 1] == "{" and self.peek() not in ("(", "["):
                    name = self.next()
                    self.next()
                    items.append((name, self.parse_dict_body("}")))
                    self.next()
                else:
                    items.append(self.parse_value())
            self.next()
            return items
        if tok == "[":
            items = []
            while self.peek() != "]":
                items.append(self.next())
            self.next()
            return ("dimensions", items)
        try:
            return int(tok)
        except ValueError:
            try:
                return float(tok)
            except ValueError:
                return tok


def parse_foam_dict(text: str) -> Dict:
    """An OpenFOAM dictionary file -> nested dict; ``( a b c )`` -> list, ``name [dims] value`` -> [name, dims, value]."""
    return _Parser(_tokenize(text)).parse_dict_body(None)


def read_foam_dict(path: str) -> Dict:
    with open(path) as f:
        d = parse_foam_dict(f.read())
    d.pop("FoamFile", None)
    return d


def _scalar(v) -> float:
    """``E  E [1 -1 -2 0 0 0 0] 200e+9;`` / ``E 200e9;`` / ``[..] 200e9`` -> 200e9."""
    if isinstance(v, list):
        return float(v[-1])
    return float(v)


def _lookup(d: Dict, key: str, default=None):
    """Dictionary lookup with OpenFOAM's regular-expression keys ("D|DD")."""
    if key in d:
        return d[key]
    for k, v in d.items():
        if any(ch in k for ch in "|.*()[]") and re.fullmatch(k, key):
            return v
    return default


# ------------------------------------------------------------------------------------------------
# case directory -> law / controls / boundary conditions
# ------------------------------------------------------------------------------------------------
def read_mechanical_law(case_dir: str) -> K.Law:
    d = read_foam_dict(os.path.join(case_dir, "constant", "mechanicalProperties"))
    plane_stress = str(d.get("planeStress", "no")).lower() in ("yes", "true", "on")
    laws = d["mechanical"]
    if len(laws) != 1:
        raise ValueError("only single-law cases are on the GPU path (mechanicalModel.C:476-483, laws.size() == 1)")
    _, ld = laws[0]
    kw = dict(rho=_scalar(ld["rho"]), planeStress=plane_stress)
    for k in ("E", "nu", "mu", "K"):
        if k in ld:
            kw[k] = _scalar(ld[k])
    if "solvePressureEqn" in ld:
        kw["solvePressureEqn"] = str(ld["solvePressureEqn"]).lower() in ("yes", "true", "on")
    if "pressureSmoothingScaleFactor" in ld:
        kw["pressureSmoothingScaleFactor"] = _scalar(ld["pressureSmoothingScaleFactor"])
    if kw.get("solvePressureEqn"):          # fvSolution solvers sigmaHyd / relaxationFactors fields sigmaHyd (mechanicalLaw.C:1455-1459)
        fs = os.path.join(case_dir, "system", "fvSolution")
        if os.path.exists(fs):
            sol = read_foam_dict(fs)
            sd = _lookup(sol.get("solvers", {}), "sigmaHyd", {})
            if "tolerance" in sd:
                kw["sigmaHydTolerance"] = float(sd["tolerance"])
                kw["sigmaHydRelTol"] = float(sd.get("relTol", 0.0))
                kw["sigmaHydMaxIter"] = int(sd.get("maxIter", 1000))
            rf = _lookup(sol.get("relaxationFactors", {}).get("fields", {}), "sigmaHyd")
            if rf is not None:
                kw["sigmaHydRelax"] = float(rf)
    fname = _lookup(ld, "fileName") or _lookup(ld, "file")      # the tutorials write the key as "file|fileName"
    if fname is not None:       # plasticity: the (epsilonP sigmaY) table file, neoHookeanElasticMisesPlastic.C:868-930
        path = str(fname).strip('"').replace("$FOAM_CASE", case_dir)
        with open(path) as f:
            tbl = _Parser(_tokenize(f.read())).parse_value()
        kw["table"] = [(float(a), float(b)) for a, b in tbl]
    return K.mechanical_law(str(ld["type"]), **kw)


def read_controls(case_dir: str, **overrides) -> K.Controls:
    sp = read_foam_dict(os.path.join(case_dir, "constant", "solidProperties"))
    model = str(sp["solidModel"])
    if model not in K.MODEL_NAMES:
        raise KeyError(f"Unknown solidModel type {model}\nValid solidModel types are: {sorted(K.MODEL_NAMES)}")
    coeffs = _lookup(sp, model + "Coeffs", {})
    kw: Dict = dict(solidModel=K.MODEL_NAMES[model])
    for src, dst in (("nCorrectors", "nCorrectors"), ("solutionTolerance", "solutionTolerance"),
                     ("alternativeTolerance", "alternativeTolerance"), ("materialTolerance", "materialTolerance")):
        if src in coeffs:
            kw[dst] = type(getattr(K.default_controls(), dst))(coeffs[src])
    stab = coeffs.get("stabilisation", {})
    if stab:
        st = str(stab.get("type", "RhieChow"))
        if st not in ("none", "RhieChow"):      # momentumStabilisation.C:43-81 also knows JamesonSchmidtTurkel and Laplacian
            raise ValueError(f"stabilisation type {st} is not available on the GPU path (RhieChow, none)")
        kw["stabilisation"] = K.STAB_NONE if st == "none" else K.STAB_RHIE_CHOW
        if "scaleFactor" in stab:
            kw["stabScaleFactor"] = float(stab["scaleFactor"])
    if str(coeffs.get("relaxationMethod", "fixed")) == "Aitken":
        kw["relaxationMethod"] = K.RELAX_AITKEN
    schemes = read_foam_dict(os.path.join(case_dir, "system", "fvSchemes"))
    d2 = str(schemes.get("d2dt2Schemes", {}).get("default", "steadyState"))
    kw["d2dt2Scheme"] = dict(steadyState=K.D2DT2_STEADY_STATE, Euler=K.D2DT2_EULER, backward=K.D2DT2_BACKWARD)[d2]
    grad = schemes.get("gradSchemes", {}).get("default", "leastSquares")
    grad = grad[0] if isinstance(grad, list) else grad
    kw["gradScheme"] = {"Gauss": K.GRAD_GAUSS_LINEAR, "pointCellsLeastSquares": K.GRAD_POINT_CELLS_LEAST_SQUARES}.get(str(grad), K.GRAD_LEAST_SQUARES)
    sol = read_foam_dict(os.path.join(case_dir, "system", "fvSolution"))
    field = "DD" if kw["solidModel"] in K.INCREMENTAL_MODELS else "D"
    sd = _lookup(sol.get("solvers", {}), field, {})
    kw["solver"] = K.SOLVER_PBICGSTAB if str(sd.get("solver", "PCG")) == "PBiCGStab" else K.SOLVER_PCG
    pre = str(sd.get("preconditioner", "DIC"))
    kw["preconditioner"] = dict(DIC=K.PRECOND_DIC, FDIC=K.PRECOND_DIC, DILU=K.PRECOND_DIC, diagonal=K.PRECOND_DIAGONAL,
                                none=K.PRECOND_NONE, GAMG=K.PRECOND_GAMG).get(pre, K.PRECOND_DIC)
    if str(sd.get("solver", "PCG")) == "GAMG":
        kw["preconditioner"] = K.PRECOND_GAMG
    for src in ("tolerance", "relTol"):
        if src in sd:
            kw[src] = float(sd[src])
    if "maxIter" in sd:
        kw["maxIter"] = int(sd["maxIter"])
    rf = _lookup(sol.get("relaxationFactors", {}).get("fields", {}), field)
    if rf is not None:
        kw["fieldRelaxD"] = float(rf)
    re_ = _lookup(sol.get("relaxationFactors", {}).get("equations", {}), field)
    if re_ is not None and abs(float(re_) - 1.0) > 1e-12:      # DEqn.relax() with a factor != 1 changes the matrix diagonal
        raise ValueError(f"relaxationFactors equations {field} {re_}: equation relaxation is not available on the GPU path")
    known_pre = dict(DIC=K.PRECOND_DIC, FDIC=K.PRECOND_DIC, DILU=K.PRECOND_DIC, diagonal=K.PRECOND_DIAGONAL, none=K.PRECOND_NONE, GAMG=K.PRECOND_GAMG)
    if "preconditioner" in sd and not isinstance(sd["preconditioner"], dict) and str(sd["preconditioner"]) not in known_pre:
        raise ValueError(f"preconditioner {sd['preconditioner']} is not available on the GPU path ({', '.join(known_pre)})")
    if str(sd.get("solver", "PCG")) not in ("PCG", "PBiCGStab", "PBiCG", "GAMG"):
        raise ValueError(f"solver {sd.get('solver')} is not available on the GPU path (PCG, PBiCGStab, GAMG)")
    gpath = os.path.join(case_dir, "constant", "g")
    if os.path.exists(gpath):
        gv = read_foam_dict(gpath).get("value")
        if isinstance(gv, list):
            kw["g"] = tuple(float(x) for x in gv[-1]) if isinstance(gv[-1], list) else tuple(float(x) for x in gv)
    cd = os.path.join(case_dir, "system", "controlDict")
    if os.path.exists(cd):
        dt = read_foam_dict(cd).get("deltaT")
        if dt is not None:
            kw["deltaT"] = kw["deltaT0"] = float(dt)
    kw.update(overrides)
    return K.default_controls(**kw)


def _uniform_or_list(v, n: int, ncomp: int) -> np.ndarray:
    """``uniform (0 0 0)`` / ``uniform 0`` / ``nonuniform List<vector> N ( ... )``."""
    if isinstance(v, list) and v and v[0] == "uniform":
        x = np.asarray(v[1], dtype=np.float64)
        return np.broadcast_to(x, (n, ncomp) if ncomp > 1 else (n,)).copy()
    if isinstance(v, list) and v and v[0] == "nonuniform":
        return np.asarray(v[-1], dtype=np.float64).reshape((n, ncomp) if ncomp > 1 else (n,))
    raise ValueError(f"cannot read field entry {v!r}")


def _read_series(sub, case_dir: str):
    """A <name>Series sub-dictionary of a patch field: s4f's interpolationTable read from "file|fileName" (list of
    (t value) tuples); only outOfBounds clamp is on this path."""
    oob = str(sub.get("outOfBounds", "clamp"))
    if oob != "clamp":
        raise ValueError(f"time series with outOfBounds {oob}: only clamp is available on the GPU path")
    fname = _lookup(sub, "fileName") or _lookup(sub, "file")
    if fname is None:
        raise ValueError("time series without file / fileName")
    path = str(fname).strip('"').replace("$FOAM_CASE", case_dir)
    with open(path) as f:
        tbl = _Parser(_tokenize(f.read())).parse_value()
    return [(float(a), b) for a, b in tbl]


# what each patch-field type may carry; anything else is refused rather than silently ignored (a case with e.g.
# "secondOrder yes" would otherwise run and report a converged but different answer)
_BC_KEYS = {
    "fixedDisplacement": {"type", "value", "displacementSeries", "patchType"},
    "solidTraction": {"type", "value", "traction", "pressure", "tractionSeries", "pressureSeries", "gradient", "patchType",
                      "secondOrder", "setEffectiveTraction", "relaxationFactor", "limitCoeff"},
}
_BC_NEUTRAL = {"secondOrder": ("no", "false", "off"), "setEffectiveTraction": ("no", "false", "off"), "relaxationFactor": ("1", "1.0"),
               "limitCoeff": None}


def _check_bc_keys(patch: str, t: str, pd) -> None:
    allowed = _BC_KEYS.get(t)
    if allowed is None:
        return
    for k, v in pd.items():
        if k not in allowed:
            raise ValueError(f"patch {patch}: entry {k} of {t} is not available on the GPU path")
        neutral = _BC_NEUTRAL.get(k, ())
        if neutral and str(v).lower() not in neutral:
            raise ValueError(f"patch {patch}: {t} with {k} {v} is not available on the GPU path")


def read_boundary_conditions(case_dir: str, mesh: M.FvMesh, field: str = "D", time: str = "0") -> Dict[str, K.BC]:
    d = read_foam_dict(os.path.join(case_dir, time, field))
    out: Dict[str, K.BC] = {}
    F = mesh.nInternalFaces
    for p in mesh.patches:
        if p.kind == M.PROCESSOR:
            continue
        pd = _lookup(d["boundaryField"], p.name)
        if pd is None:
            raise KeyError(f"patch {p.name} missing in {time}/{field}")
        t = str(pd["type"])
        _check_bc_keys(p.name, t, pd)
        if t == "fixedDisplacement":
            bc = K.fixedDisplacement(_uniform_or_list(pd["value"], p.size, 3))
            if "displacementSeries" in pd:         # fixedDisplacementFvPatchVectorField.C:258-294: disp = dispSeries_(time)
                bc.value_series = _read_series(pd["displacementSeries"], case_dir)
            out[p.name] = bc
        elif t == "solidTraction":
            if ("traction" in pd) == ("tractionSeries" in pd) or ("pressure" in pd) == ("pressureSeries" in pd):
                raise ValueError(f"patch {p.name}: exactly one of traction / tractionSeries and of pressure / pressureSeries "
                                 "must be given (solidTractionFvPatchVectorField.C:104-170; tractionField / pressureField are not on the GPU path)")
            tr = _uniform_or_list(pd["traction"], p.size, 3) if "traction" in pd else np.zeros((p.size, 3))
            pr = _uniform_or_list(pd["pressure"], p.size, 1) if "pressure" in pd else np.zeros(p.size)
            bc = K.solidTraction(tr, pr)
            if "tractionSeries" in pd:
                bc.value_series = _read_series(pd["tractionSeries"], case_dir)
            if "pressureSeries" in pd:
                bc.pressure_series = _read_series(pd["pressureSeries"], case_dir)
            out[p.name] = bc
        elif t in ("solidSymmetry", "symmetryPlane", "symmetry"):
            out[p.name] = K.solidSymmetry()
        elif t == "analyticalPlateHoleTraction":
            # SM/fvPatchFields/analyticalPlateHoleTraction/...C:36-88, :176-210: traction = n & sigma_Kirsch(Cf)
            from .cases import kirsch_stress
            sl = slice(F + p.start, F + p.start + p.size)
            cf = mesh.Cf[sl].copy()
            cf[:, 2] = 0.0
            sg = kirsch_stress(cf, float(pd["farFieldTractionX"]), float(pd["holeRadius"]))
            n = mesh.Sf[sl] / mesh.magSf[sl, None]
            tr = np.stack([n[:, 0] * sg[:, 0] + n[:, 1] * sg[:, 1], n[:, 0] * sg[:, 1] + n[:, 1] * sg[:, 3], np.zeros(p.size)], axis=1)
            out[p.name] = K.solidTraction(tr)
        else:
            raise ValueError(f"patch {p.name}: boundary condition {t} is not available on the GPU path "
                             "(solidTraction, fixedDisplacement, solidSymmetry, analyticalPlateHoleTraction)")
    return out


def read_case(case_dir: str, mesh: Optional[M.FvMesh] = None, **overrides) -> K.SolidCase:
    """A solids4foam case directory -> :class:`SolidCase`.  ``mesh``: an fvMesh made elsewhere (the tutorials ship a
    blockMeshDict, not a polyMesh); otherwise ``constant/polyMesh`` is read."""
    if mesh is None:
        mesh = read_poly_mesh(os.path.join(case_dir, "constant", "polyMesh"))
    law = read_mechanical_law(case_dir)
    ctl = read_controls(case_dir, **overrides)
    field = "DD" if ctl.solidModel in K.INCREMENTAL_MODELS and os.path.exists(os.path.join(case_dir, "0", "DD")) else "D"
    bcs = read_boundary_conditions(case_dir, mesh, field)
    return K.SolidCase(mesh, bcs, law, ctl, name=os.path.basename(os.path.normpath(case_dir)))


# ------------------------------------------------------------------------------------------------
# polyMesh
# ------------------------------------------------------------------------------------------------
def _read_list_file(path: str):
    with open(path) as f:
        toks = _tokenize(f.read())
    p = _Parser(toks)
    # skip the FoamFile header dictionary
    while p.peek() is not None and p.peek() != "FoamFile":
        p.next()
    if p.peek() == "FoamFile":
        p.next(); p.next(); p.parse_dict_body("}"); p.next()
    n = p.parse_value()
    body = p.parse_value()
    return n, body


def read_poly_mesh(poly_dir: str, rank: int = 0, nRanks: int = 1, exchange=None) -> M.FvMesh:
    """[OF-ext] polyMesh ASCII files -> fvMesh with the geometry primitiveMesh / surfaceInterpolation would compute.
    ``processorN/constant/polyMesh`` of a decomposed case: pass rank, nRanks and exchange (see fv_mesh_from_poly); the
    global cell numbers come from ``cellProcAddressing`` when decomposePar wrote it."""
    _, pts = _read_list_file(os.path.join(poly_dir, "points"))
    points = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
    nF, fl = _read_list_file(os.path.join(poly_dir, "faces"))
    faces: List[List[int]] = []
    i = 0
    while i < len(fl):                      # "4(a b c d)" tokenises to 4, [a, b, c, d]
        if isinstance(fl[i], int) and i + 1 < len(fl) and isinstance(fl[i + 1], list):
            faces.append([int(x) for x in fl[i + 1]]); i += 2
        else:
            faces.append([int(x) for x in fl[i]]); i += 1
    _, own = _read_list_file(os.path.join(poly_dir, "owner"))
    _, nei = _read_list_file(os.path.join(poly_dir, "neighbour"))
    owner = np.asarray(own, dtype=np.int64)
    neighbour = np.asarray(nei, dtype=np.int64)
    _, bl = _read_list_file(os.path.join(poly_dir, "boundary"))
    patches_raw = [(name, d) for name, d in bl]
    cellGlobal = None
    cpa = os.path.join(poly_dir, "cellProcAddressing")
    if os.path.exists(cpa):
        cellGlobal = np.asarray(_read_list_file(cpa)[1], dtype=np.int64)
    return fv_mesh_from_poly(points, faces, owner, neighbour, patches_raw, rank=rank, nRanks=nRanks, exchange=exchange, cellGlobal=cellGlobal)


def polygon_centres_and_areas(points: np.ndarray, fptr: np.ndarray, fflat: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """[OF-ext] primitiveMesh::makeFaceCentresAndAreas for polygon faces given in CSR form: triangles about the vertex
    average, area-weighted centroid."""
    nF = fptr.size - 1
    cnt = np.diff(fptr)
    fid = np.repeat(np.arange(nF), cnt)
    p0 = points[fflat]
    nxt = np.arange(fflat.size) + 1
    last = fptr[1:] - 1
    nxt[last] = fptr[:-1]
    p1 = points[fflat[nxt]]
    fc = np.stack([np.bincount(fid, weights=p0[:, c], minlength=nF) for c in range(3)], axis=1) / cnt[:, None]
    fcr = fc[fid]
    n = np.cross(p1 - p0, fcr - p0)
    a = np.linalg.norm(n, axis=1)
    c3 = p0 + p1 + fcr
    sumN = np.stack([np.bincount(fid, weights=n[:, c], minlength=nF) for c in range(3)], axis=1)
    sumA = np.bincount(fid, weights=a, minlength=nF)
    sumAc = np.stack([np.bincount(fid, weights=a * c3[:, c], minlength=nF) for c in range(3)], axis=1)
    ctr = np.where(sumA[:, None] > 1e-300, sumAc / (3.0 * np.maximum(sumA, 1e-300))[:, None], fc)
    return ctr, 0.5 * sumN


def fv_mesh_from_poly(points, faces, owner, neighbour, patches_raw, rank: int = 0, nRanks: int = 1, exchange=None,
                      cellGlobal=None) -> M.FvMesh:
    """polyMesh arrays -> fvMesh.  On a decomposed mesh (``processor`` patches) the interpolation geometry of the cut faces
    needs the cell centres across the cut: ``exchange({nbr_rank: C[faceCells of the patch]}) -> {nbr_rank: their centres}``
    supplies them (torch.distributed in run_case, an in-memory swap in the tests)."""
    nI = neighbour.size
    nCells = int(owner.max()) + 1
    fptr = np.zeros(len(faces) + 1, dtype=np.int64)
    fptr[1:] = np.cumsum([len(f) for f in faces])
    fflat = np.fromiter((v for f in faces for v in f), dtype=np.int64, count=int(fptr[-1]))
    fCtrs, fAreas = polygon_centres_and_areas(points, fptr, fflat)
    C, V = M.cell_centres_and_volumes(nCells, fCtrs, fAreas, owner, neighbour)
    kind_of = dict(patch=M.PATCH, wall=M.PATCH, empty=M.EMPTY, symmetryPlane=M.SYMMETRY_PLANE, symmetry=M.SYMMETRY_PLANE,
                   processor=M.PROCESSOR)
    patches: List[M.PatchInfo] = []
    keep: List[np.ndarray] = []
    solutionD = np.ones(3, dtype=np.int32)
    start = 0
    for name, d in patches_raw:
        t = str(d["type"])
        n, s = int(d["nFaces"]), int(d["startFace"])
        kind = kind_of.get(t, M.PATCH)
        if kind == M.EMPTY:
            if n:
                nrm = np.abs(fAreas[s:s + n] / np.linalg.norm(fAreas[s:s + n], axis=1)[:, None]).mean(axis=0)
                solutionD[int(np.argmax(nrm))] = 0
            continue
        patches.append(M.PatchInfo(name, kind, start, n, nbr_rank=int(d.get("neighbProcNo", -1))))
        keep.append(np.arange(s, s + n))
        start += n
    kb = np.concatenate(keep) if keep else np.zeros(0, dtype=np.int64)
    Sf = np.concatenate([fAreas[:nI], fAreas[kb]])
    Cf = np.concatenate([fCtrs[:nI], fCtrs[kb]])
    CnbrB_proc = None
    procs = [p for p in patches if p.kind == M.PROCESSOR]
    if procs:
        if exchange is None:
            raise ValueError("a mesh with processor patches needs the neighbour cell centres: pass exchange=")
        fcB = owner[kb]
        got = exchange({p.nbr_rank: C[fcB[p.start:p.start + p.size]] for p in procs})
        CnbrB_proc = Cf[nI:].copy()
        for p in procs:
            CnbrB_proc[p.start:p.start + p.size] = got[p.nbr_rank]
    mesh = M._finish_mesh(nCells, owner[:nI], neighbour, owner[kb], patches, C, V, Sf, Cf, CnbrB_proc, solutionD,
                          cellGlobal=np.arange(nCells, dtype=np.int64) if cellGlobal is None else np.asarray(cellGlobal, dtype=np.int64),
                          rank=rank, nRanks=nRanks)
    mesh.points = points.copy()
    kept = list(range(nI)) + [int(i) for i in kb]
    if all(len(f) == 4 for f in faces):          # quads: the mesh can be moved (mesh.move_points) and its points interpolated to
        mesh.faces = np.asarray([faces[i] for i in kept], dtype=np.int32)
        mesh.topo = dict(faces_all=np.asarray(faces, dtype=np.int64), own_all=owner, kb=kb - nI)
    mesh.meta["poly"] = dict(faces=faces, owner=owner, neighbour=neighbour, patches_raw=patches_raw)
    return mesh


def write_poly_mesh(poly_dir: str, mesh: M.FvMesh) -> None:
    """Write points / faces / owner / neighbour / boundary of a mesh that carries points() and faces()."""
    if mesh.points is None or mesh.faces is None:
        raise ValueError("mesh has no points/faces (use the general builder)")
    if mesh.topo is not None and len(mesh.topo["faces_all"]) != mesh.nInternalFaces + mesh.nBoundaryFaces:
        raise ValueError("the faces of the empty patches of a 2-D mesh are not kept: cannot write a closed polyMesh")
    os.makedirs(poly_dir, exist_ok=True)
    F, B = mesh.nInternalFaces, mesh.nBoundaryFaces
    own = np.concatenate([mesh.owner, mesh.faceCells])

    def wr(name, cls, body):
        with open(os.path.join(poly_dir, name), "w") as f:
            f.write(_HEADER.format(cls=cls, loc="constant/polyMesh", obj=name))
            f.write(body)
            f.write("\n\n// ************************************************************************* //\n")
    wr("points", "vectorField", f"{mesh.points.shape[0]}\n(\n" + "\n".join(f"({p[0]!r} {p[1]!r} {p[2]!r})" for p in mesh.points.tolist()) + "\n)")
    wr("faces", "faceList", f"{F + B}\n(\n" + "\n".join(f"{len(fc)}({' '.join(str(v) for v in fc)})" for fc in mesh.faces.tolist()) + "\n)")
    wr("owner", "labelList", f"{F + B}\n(\n" + "\n".join(str(int(v)) for v in own) + "\n)")
    wr("neighbour", "labelList", f"{F}\n(\n" + "\n".join(str(int(v)) for v in mesh.neighbour) + "\n)")
    tname = {M.PATCH: "patch", M.SYMMETRY_PLANE: "symmetryPlane", M.PROCESSOR: "processor", M.EMPTY: "empty"}
    items = []
    for p in mesh.patches:
        extra = f"        myProcNo        {mesh.rank};\n        neighbProcNo    {p.nbr_rank};\n" if p.kind == M.PROCESSOR else ""
        items.append(f"    {p.name}\n    {{\n        type            {tname[p.kind]};\n        nFaces          {p.size};\n"
                     f"        startFace       {F + p.start};\n{extra}    }}")
    wr("boundary", "polyBoundaryMesh", f"{len(items)}\n(\n" + "\n".join(items) + "\n)")


def write_raw_poly_mesh(poly_dir: str, points, faces, owner, neighbour, patches_raw, rank: int = 0, cellGlobal=None) -> None:
    """points / faces / owner / neighbour / boundary (+ cellProcAddressing) from raw polyMesh arrays."""
    os.makedirs(poly_dir, exist_ok=True)

    def wr(name, cls, body):
        with open(os.path.join(poly_dir, name), "w") as f:
            f.write(_HEADER.format(cls=cls, loc="constant/polyMesh", obj=name))
            f.write(body)
            f.write("\n\n// ************************************************************************* //\n")
    wr("points", "vectorField", f"{len(points)}\n(\n" + "\n".join(f"({p[0]!r} {p[1]!r} {p[2]!r})" for p in np.asarray(points).tolist()) + "\n)")
    wr("faces", "faceList", f"{len(faces)}\n(\n" + "\n".join(f"{len(fc)}({' '.join(str(int(v)) for v in fc)})" for fc in faces) + "\n)")
    wr("owner", "labelList", f"{len(owner)}\n(\n" + "\n".join(str(int(v)) for v in owner) + "\n)")
    wr("neighbour", "labelList", f"{len(neighbour)}\n(\n" + "\n".join(str(int(v)) for v in neighbour) + "\n)")
    items = []
    for name, d in patches_raw:
        body = "".join(f"        {k:<15} {v};\n" for k, v in d.items())
        items.append(f"    {name}\n    {{\n{body}    }}")
    wr("boundary", "polyBoundaryMesh", f"{len(items)}\n(\n" + "\n".join(items) + "\n)")
    if cellGlobal is not None:
        wr("cellProcAddressing", "labelList", f"{len(cellGlobal)}\n(\n" + "\n".join(str(int(v)) for v in cellGlobal) + "\n)")


def decompose_case(case_dir: str, nRanks: int, cell_rank=None) -> None:
    """``decomposePar`` for a solid case directory (constant/polyMesh, 0/D): writes ``processorN/constant/polyMesh`` (with
    cellProcAddressing) and ``processorN/0/D``.  Default cell -> processor map: ``method simple`` with ``n (P 1 1)``, equal
    slabs of cells along x (system/decomposeParDict of the plateHole tutorial); the initial D field must be uniform."""
    case = read_case(case_dir)
    mesh = case.mesh
    meta = mesh.meta["poly"]
    if cell_rank is None:
        order = np.argsort(mesh.C[:, 0], kind="stable")
        cell_rank = np.empty(mesh.nCells, dtype=np.int64)
        cell_rank[order] = (np.arange(mesh.nCells) * nRanks) // mesh.nCells
    parts = M.decompose_poly(mesh.points, meta["faces"], np.asarray(meta["owner"]), np.asarray(meta["neighbour"]), meta["patches_raw"], cell_rank)
    kind_of = dict(patch=M.PATCH, wall=M.PATCH, empty=M.EMPTY, symmetryPlane=M.SYMMETRY_PLANE, symmetry=M.SYMMETRY_PLANE, processor=M.PROCESSOR)
    for r, part in enumerate(parts):
        pdir = os.path.join(case_dir, f"processor{r}")
        write_raw_poly_mesh(os.path.join(pdir, "constant", "polyMesh"), part["points"], part["faces"], part["owner"], part["neighbour"],
                            part["patches_raw"], rank=r, cellGlobal=part["cellGlobal"])
        patches, bcs = [], {}
        for name, d in part["patches_raw"]:
            kind = kind_of.get(str(d["type"]), M.PATCH)
            if kind == M.EMPTY:
                continue
            patches.append(M.PatchInfo(name, kind, 0, int(d["nFaces"]), nbr_rank=int(d.get("neighbProcNo", -1))))
            if kind == M.PROCESSOR:
                continue
            bc, sel = case.bcs[name], part["patchSel"][name]
            n0 = mesh.patch(name).size
            val = None if bc.value is None else np.broadcast_to(np.asarray(bc.value, dtype=np.float64), (n0, 3))[sel]
            pr = None if bc.pressure is None else np.broadcast_to(np.asarray(bc.pressure, dtype=np.float64), (n0,))[sel]
            bcs[name] = K.BC(bc.kind, val, pr)
        write_D_file(os.path.join(pdir, "0"), patches, bcs)


def read_decomposed_case(case_dir: str, rank: int, nRanks: int, exchange, **overrides) -> K.SolidCase:
    """One processor's part of a decomposed case: processor<rank>/constant/polyMesh and processor<rank>/0/<field>; the
    dictionaries under constant/ and system/ are shared."""
    pdir = os.path.join(case_dir, f"processor{rank}")
    mesh = read_poly_mesh(os.path.join(pdir, "constant", "polyMesh"), rank=rank, nRanks=nRanks, exchange=exchange)
    law = read_mechanical_law(case_dir)
    ctl = read_controls(case_dir, **overrides)
    field = "DD" if ctl.solidModel in K.INCREMENTAL_MODELS and os.path.exists(os.path.join(pdir, "0", "DD")) else "D"
    bcs = read_boundary_conditions(pdir, mesh, field)
    return K.SolidCase(mesh, bcs, law, ctl, name=os.path.basename(os.path.normpath(case_dir)) + f"/processor{rank}")


# ------------------------------------------------------------------------------------------------
# vol fields
# ------------------------------------------------------------------------------------------------
_CLS = {1: ("volScalarField", "scalar"), 3: ("volVectorField", "vector"), 6: ("volSymmTensorField", "symmTensor"),
        9: ("volTensorField", "tensor")}


def _fmt_rows(a: np.ndarray) -> str:
    if a.ndim == 1:
        return "\n".join(repr(float(v)) for v in a)
    return "\n".join("(" + " ".join(repr(float(v)) for v in row) + ")" for row in a)


def write_vol_field(time_dir: str, name: str, mesh: M.FvMesh, internal: np.ndarray, boundary: Optional[np.ndarray] = None,
                    dimensions: str = "[0 1 0 0 0 0 0]", patch_types: Optional[Dict[str, str]] = None) -> None:
    """A time-directory field file (``D``, ``sigma`` ...) with ``calculated`` patch values, what writeFields produces."""
    os.makedirs(time_dir, exist_ok=True)
    internal = np.asarray(internal, dtype=np.float64)
    nc = 1 if internal.ndim == 1 else internal.shape[1]
    cls, prim = _CLS[nc]
    with open(os.path.join(time_dir, name), "w") as f:
        f.write(_HEADER.format(cls=cls, loc=os.path.basename(os.path.normpath(time_dir)), obj=name))
        f.write(f"dimensions      {dimensions};\n\ninternalField   nonuniform List<{prim}>\n{internal.shape[0]}\n(\n{_fmt_rows(internal)}\n)\n;\n\nboundaryField\n{{\n")
        for p in mesh.patches:
            t = (patch_types or {}).get(p.name, "processor" if p.kind == M.PROCESSOR else "calculated")
            f.write(f"    {p.name}\n    {{\n        type            {t};\n")
            if boundary is not None and p.size:
                vals = np.asarray(boundary)[p.start:p.start + p.size]
                f.write(f"        value           nonuniform List<{prim}>\n{p.size}\n(\n{_fmt_rows(vals)}\n)\n;\n")
            f.write("    }\n")
        f.write("}\n\n// ************************************************************************* //\n")


def read_vol_field(path: str, mesh: M.FvMesh) -> Tuple[np.ndarray, Dict[str, np.ndarray]]:
    d = read_foam_dict(path)
    nc = None
    iv = d["internalField"]
    if iv[0] == "uniform":
        x = np.asarray(iv[1], dtype=np.float64)
        nc = 1 if x.ndim == 0 else x.size
        internal = np.broadcast_to(x, (mesh.nCells, nc) if nc > 1 else (mesh.nCells,)).copy()
    else:
        internal = np.asarray(iv[-1], dtype=np.float64)
        nc = 1 if internal.ndim == 1 else internal.shape[1]
    bvals: Dict[str, np.ndarray] = {}
    for p in mesh.patches:
        pd = _lookup(d.get("boundaryField", {}), p.name, {})
        if "value" in pd:
            bvals[p.name] = _uniform_or_list(pd["value"], p.size, nc)
    return internal, bvals


# ------------------------------------------------------------------------------------------------
# writing a case directory (the inverse of read_case, used to hand cases to the standalone driver and in the tests)
# ------------------------------------------------------------------------------------------------
_LAW_OF = {v: k for k, v in K.LAW_NAMES.items()}
_MODEL_OF = {v: k for k, v in K.MODEL_NAMES.items()}


def _dict_file(path: str, obj: str, body: str, cls: str = "dictionary") -> None:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(_HEADER.format(cls=cls, loc=os.path.basename(os.path.dirname(path)), obj=obj))
        f.write(body + "\n")


def write_D_file(zero_dir: str, patches, bcs) -> None:
    """0/D with the patch entries the reference reads (fixedDisplacement value, solidTraction traction / pressure, solidSymmetry)."""
    rows = []
    for p in patches:
        if p.kind == M.PROCESSOR:
            rows.append(f"    {p.name}\n    {{\n        type            processor;\n        value           uniform (0 0 0);\n    }}")
            continue
        bc = bcs[p.name]
        vec = lambda a: "nonuniform List<vector>\n" + f"{p.size}\n(\n" + _fmt_rows(np.broadcast_to(np.asarray(a, dtype=np.float64), (p.size, 3))) + "\n)\n"
        if bc.kind == K.BC_FIXED_DISPLACEMENT:
            rows.append(f"    {p.name}\n    {{\n        type            fixedDisplacement;\n        value           {vec(bc.value)};\n    }}")
        elif bc.kind == K.BC_SOLID_TRACTION:
            pr = np.broadcast_to(np.zeros(1) if bc.pressure is None else np.asarray(bc.pressure, dtype=np.float64), (p.size,))
            rows.append(f"    {p.name}\n    {{\n        type            solidTraction;\n        traction        {vec(bc.value)};\n"
                        f"        pressure        nonuniform List<scalar>\n{p.size}\n(\n{_fmt_rows(pr)}\n)\n;\n        value           uniform (0 0 0);\n    }}")
        else:
            rows.append(f"    {p.name}\n    {{\n        type            solidSymmetry;\n        patchType       symmetryPlane;\n        value           uniform (0 0 0);\n    }}")
    _dict_file(os.path.join(zero_dir, "D"), "D",
               "dimensions      [0 1 0 0 0 0 0];\n\ninternalField   uniform (0 0 0);\n\nboundaryField\n{\n" + "\n".join(rows) + "\n}", cls="volVectorField")


def write_case(case_dir: str, case: K.SolidCase, end_time: float = 1.0) -> None:
    """constant/{polyMesh, solidProperties, mechanicalProperties, g}, system/{controlDict, fvSchemes, fvSolution}, 0/D with the
    keys the reference reads (see the module header)."""
    m, c, L = case.mesh, case.controls, case.law
    write_poly_mesh(os.path.join(case_dir, "constant", "polyMesh"), m)
    model = _MODEL_OF[c.solidModel]
    stab = "RhieChow" if c.stabilisation == K.STAB_RHIE_CHOW else "none"
    _dict_file(os.path.join(case_dir, "constant", "solidProperties"), "solidProperties",
               f"solidModel     {model};\n\n{model}Coeffs\n{{\n    nCorrectors     {c.nCorrectors};\n    solutionTolerance {c.solutionTolerance!r};\n"
               f"    alternativeTolerance {c.alternativeTolerance!r};\n    materialTolerance {c.materialTolerance!r};\n"
               f"    relaxationMethod {'Aitken' if c.relaxationMethod == K.RELAX_AITKEN else 'fixed'};\n"
               f"    stabilisation\n    {{\n        type        {stab};\n        scaleFactor {c.stabScaleFactor!r};\n    }}\n}}")
    law = f"        type            {_LAW_OF[L.kind]};\n        rho             rho [1 -3 0 0 0 0 0] {L.rho!r};\n" \
          f"        mu              mu [1 -1 -2 0 0 0 0] {L.mu!r};\n        K               K [1 -1 -2 0 0 0 0] {L.K!r};\n"
    if L.solvePressureEqn:
        law += f"        solvePressureEqn yes;\n        pressureSmoothingScaleFactor {L.pressureSmoothingScaleFactor!r};\n"
    if L.nTable:
        law += '        fileName        "$FOAM_CASE/constant/plasticStrainVsYieldStress";\n        outOfBounds     clamp;\n'
        with open(os.path.join(case_dir, "constant", "plasticStrainVsYieldStress"), "w") as f:
            f.write("(\n" + "\n".join(f"    ({L.tableEps[i]!r} {L.tableSigY[i]!r})" for i in range(L.nTable)) + "\n)\n")
    _dict_file(os.path.join(case_dir, "constant", "mechanicalProperties"), "mechanicalProperties",
               f"planeStress     no;\n\nmechanical\n(\n    material\n    {{\n{law}    }}\n);")
    _dict_file(os.path.join(case_dir, "constant", "g"), "g",
               f"dimensions      [0 1 -2 0 0 0 0];\nvalue           ({c.g[0]!r} {c.g[1]!r} {c.g[2]!r});", cls="uniformDimensionedVectorField")
    d2 = {K.D2DT2_STEADY_STATE: "steadyState", K.D2DT2_EULER: "Euler", K.D2DT2_BACKWARD: "backward"}[c.d2dt2Scheme]
    grad = {K.GRAD_GAUSS_LINEAR: "Gauss linear", K.GRAD_POINT_CELLS_LEAST_SQUARES: "pointCellsLeastSquares"}.get(c.gradScheme, "leastSquares")
    _dict_file(os.path.join(case_dir, "system", "fvSchemes"), "fvSchemes",
               f"d2dt2Schemes\n{{\n    default {d2};\n}}\nddtSchemes\n{{\n    default {d2};\n}}\ngradSchemes\n{{\n    default {grad};\n}}\n"
               "divSchemes\n{\n    default Gauss linear;\n}\nlaplacianSchemes\n{\n    default Gauss linear corrected;\n}\n"
               "snGradSchemes\n{\n    default corrected;\n}\ninterpolationSchemes\n{\n    default linear;\n}")
    pre = {K.PRECOND_DIC: "DIC", K.PRECOND_DIAGONAL: "diagonal", K.PRECOND_NONE: "none", K.PRECOND_GAMG: "GAMG",
           K.PRECOND_CHEBYSHEV: "diagonal"}[c.preconditioner]
    _dict_file(os.path.join(case_dir, "system", "fvSolution"), "fvSolution",
               f'solvers\n{{\n    "D|DD"\n    {{\n        solver          {"PBiCGStab" if c.solver == K.SOLVER_PBICGSTAB else "PCG"};\n'
               f"        preconditioner  {pre};\n        tolerance       {c.tolerance!r};\n        relTol          {c.relTol!r};\n"
               f"        maxIter         {c.maxIter};\n    }}\n}}\n\nrelaxationFactors\n{{\n    fields\n    {{\n        \"D|DD\" {c.fieldRelaxD!r};\n    }}\n}}")
    _dict_file(os.path.join(case_dir, "system", "controlDict"), "controlDict",
               f"application     solids4Foam;\nstartTime       0;\nendTime         {end_time!r};\ndeltaT          {c.deltaT!r};\n"
               "writeControl    timeStep;\nwriteInterval   1;")
    write_D_file(os.path.join(case_dir, "0"), m.patches, case.bcs)
