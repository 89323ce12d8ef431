"""Standalone driver: what the ``solids4Foam`` application does for a solid-only case (applications/solvers/solids4Foam/
solids4Foam.C: ``while (runTime.run()) { runTime++; solid.evolve(); solid.updateTotalFields(); solid.writeFields(runTime); }``),
with the case read from its OpenFOAM directory and the loop run on the GPU through the C-ABI.

    python -m solids4foam_b200.run_case <caseDir> [--steps N] [--device 0] [--precond GAMG]

Writes ``<time>/D``, ``<time>/sigma`` (and ``pointD`` as a plain list when the mesh carries points).  No CPU fallback."""
from __future__ import annotations

import argparse
import os

import numpy as np

from . import case as K
from . import foam_io as IO
from .solid_model import SolidModel


def run(case_dir: str, steps: int | None = None, device: int = 0, precond: str | None = None, write: bool = True, log=print):
    over = {}
    if precond:
        over["preconditioner"] = getattr(K, "PRECOND_" + precond.upper())
    case = IO.read_case(case_dir, **over)
    if case.controls.preconditioner == K.PRECOND_DIC and precond is None:
        # DIC/FDIC in fvSolution: the device has it exactly (level scheduled, iteration counts of the CPU solver) but GAMG is the
        # fast preconditioner on a GPU; --precond DIC keeps the case's own choice
        case.controls.preconditioner = K.PRECOND_GAMG
        log("fvSolution asks for the DIC/FDIC preconditioner; running PCG with the GAMG preconditioner instead "
            "(same converged solution, fewer inner iterations; --precond DIC keeps the case's own choice)")
    cd = IO.read_foam_dict(os.path.join(case_dir, "system", "controlDict"))
    dt = float(cd.get("deltaT", 1.0))
    n = steps if steps is not None else max(1, int(round((float(cd.get("endTime", dt)) - float(cd.get("startTime", 0.0))) / dt)))
    solid = SolidModel(case, device=device)
    t = float(cd.get("startTime", 0.0))
    stats = []
    timed = {name: bc for name, bc in case.bcs.items() if bc.value_series is not None or bc.pressure_series is not None}
    for _ in range(n):
        t += dt
        solid.new_timestep(dt)
        for name, bc in timed.items():          # displacementSeries / tractionSeries / pressureSeries at the new time
            solid.set_bc(name, bc.at(t))
        st = solid.evolve()
        solid.updateTotalFields()
        stats.append(st)
        log(f"Time = {t:g}\n    Corr, res, relRes, matRes, iters\n    {st['nCorr']}, {st['solverPerfInitRes']:.3e}, {st['relResidual']:.3e}, "
            f"{st['materialResidual']:.3e}, {sum(st['nIterations'])}")
        if write:
            tdir = os.path.join(case_dir, f"{t:g}")
            IO.write_vol_field(tdir, "D", solid.case.mesh, solid.get("D"), solid.get("D_b"))
            IO.write_vol_field(tdir, "sigma", solid.case.mesh, solid.get("sigma"), solid.get("sigma_b"), dimensions="[1 -1 -2 0 0 0 0]")
    return solid, stats


def main():
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("case_dir")
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--precond", default=None, choices=["GAMG", "DIC", "DIAGONAL", "CHEBYSHEV", "NONE"])
    a = ap.parse_args()
    run(a.case_dir, a.steps, a.device, a.precond)


if __name__ == "__main__":
    main()
