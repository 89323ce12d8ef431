"""GPU parity at the sizes the benchmark runs at (the small-mesh tests never leave the first grid-stride wave, the
first SELL slices or a three-level hierarchy): 1 M cells against full oracle runs, 8 M cells (the BASELINE.json
workload, six GAMG levels) operator by operator and through the oracle's own Amul for the solve."""
import numpy as np
import pytest

from solids4foam_b200 import case as K
from solids4foam_b200 import cases
from s4f_testutil import rel_l2

pytestmark = pytest.mark.gpu

OP_TOL = 1e-12


def _analytic_D(mesh, seed=1234):
    L = mesh.C[:, 0].max() + 1e-9
    x, y, z = mesh.C[:, 0], mesh.C[:, 1], mesh.C[:, 2]
    D = 1e-3 * np.stack([np.sin(2 * np.pi * x / L), np.cos(2 * np.pi * y), x * z / L**2], axis=1)
    D += np.random.default_rng(seed).uniform(-1e-6, 1e-6, D.shape)
    return D


def _threads():
    import os
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _models(dims, threads, **kw):
    from oracle.binding import OracleSolid
    from solids4foam_b200.solid_model import SolidModel
    c = cases.cantilever(*dims, **kw)
    g, o = SolidModel(c), OracleSolid(c)
    if threads > 1:
        o.L.s4fo_set_threads(o.h, threads)      # face loops and vector kernels in parallel; same arithmetic per cell
    return g, o, c.mesh


def _check_operators(g, o, mesh, amul_components=(0, 1, 2)):
    D = _analytic_D(mesh)
    for s in (g, o):
        s.set("D", D)
        s.initialise()
    assert rel_l2(g.get("D_b"), o.get("D_b")) < OP_TOL
    assert rel_l2(g.get("gradD"), o.get("gradD")) < OP_TOL
    assert rel_l2(g.get("gradD_b"), o.get("gradD_b")) < OP_TOL
    for s in (g, o):
        s.op_correct()
    assert rel_l2(g.get("sigma"), o.get("sigma")) < OP_TOL
    for s in (g, o):
        s.op_assemble()
    assert rel_l2(g.get("diag"), o.get("diag")) < OP_TOL
    assert rel_l2(g.get("upper"), o.get("upper")) < OP_TOL
    src_g, src_o = g.get("source"), o.get("source")
    assert np.abs(src_g - src_o).max() / np.abs(src_o).max() < 1e-11
    rng = np.random.default_rng(5)
    for cmpt in amul_components:
        x = rng.standard_normal(mesh.nCells)
        assert rel_l2(g.op_amul(cmpt, x), o.op_amul(cmpt, x)) < OP_TOL
    return D, src_o


def test_operators_and_gamg_solve_at_1M_cells():
    """400x50x50: every operator to round-off, then the assembled system solved by GAMG-PCG on the device and by DIC-PCG
    in the oracle, both to 1e-11: the solutions of one linear system must agree whatever the preconditioner."""
    tight = dict(tolerance=1e-11, relTol=0.0, maxIter=5000)
    g, o, mesh = _models((400, 50, 50), _threads(), preconditioner=K.PRECOND_GAMG, **tight)
    D, src = _check_operators(g, o, mesh)
    info = g.gamg_info()
    assert len(info["levels"]) >= 5, info
    psi_g, st_g = g.op_solve(D, src)
    ctl = K.default_controls(preconditioner=K.PRECOND_DIC, **tight)
    o.set_controls(ctl)
    psi_o, st_o = o.op_solve(D, src)
    assert max(st_g["finalResidual"]) < 1e-11 and max(st_o["finalResidual"]) < 1e-11, (st_g, st_o)
    assert np.allclose(st_g["initialResidual"], st_o["initialResidual"], rtol=1e-9)
    assert rel_l2(psi_g, psi_o) < 1e-9
    assert max(st_g["nIterations"]) < 60 < min(st_o["nIterations"]), (st_g["nIterations"], st_o["nIterations"])


@pytest.mark.parametrize("pre", [K.PRECOND_DIAGONAL, K.PRECOND_DIC])
def test_first_outer_iterate_at_1M_cells(pre):
    """One whole outer iteration (right-hand side, fused PCG at relTol 0.1, boundary conditions, relaxation, gradient, law)
    with the SAME preconditioner on both sides (the oracle single-threaded: its DIC is then the exact sequential sweep that the
    device reproduces by level scheduling): equal PCG iteration counts, fields to 1e-10."""
    g, o, mesh = _models((400, 50, 50), 1, preconditioner=pre)
    sg, so = g.outer_iteration(), o.outer_iteration()
    assert sg["nIterations"] == so["nIterations"], (sg, so)
    assert np.allclose(sg["initialResidual"], so["initialResidual"], rtol=1e-9, atol=1e-30)
    assert rel_l2(g.get("D"), o.get("D")) < 1e-10
    assert rel_l2(g.get("sigma"), o.get("sigma")) < 1e-9
    assert abs(sg["relResidual"] - so["relResidual"]) <= 1e-8 * abs(so["relResidual"])


def test_operators_and_gamg_solve_at_8M_cells():
    """800x100x100, the benchmarked workload: gradient, law, matrix, right-hand side and Amul against the oracle to round-off;
    then the six-level GAMG(K-cycle)-PCG solve of the assembled system on the device, verified with the ORACLE's Amul:
    |b - A x| / |b| per component at the solver's tolerance (no oracle solve needed at this size)."""
    g, o, mesh = _models((800, 100, 100), _threads(), preconditioner=K.PRECOND_GAMG, tolerance=1e-10, relTol=0.0, maxIter=200)
    D, src = _check_operators(g, o, mesh, amul_components=(1,))
    info = g.gamg_info()
    assert len(info["levels"]) == 6, info
    psi, st = g.op_solve(D, src)
    assert max(st["nIterations"]) < 60, st
    for cmpt in range(3):
        r = src[:, cmpt] - o.op_amul(cmpt, psi[:, cmpt])
        r0 = src[:, cmpt] - o.op_amul(cmpt, D[:, cmpt])
        assert np.abs(r).sum() < 1e-8 * np.abs(r0).sum(), (cmpt, np.abs(r).sum(), np.abs(r0).sum())
