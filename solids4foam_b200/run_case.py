"""Standalone driver: what the ``solids4Foam`` application does for a solid-only case (applications/solvers/solids4Foam/
solids4Foam.C: ``while (runTime.run()) { runTime++; solid.evolve(); solid.updateTotalFields(); solid.writeFields(runTime); }``),
with the case read from its OpenFOAM directory and the loop run on the GPU through the C-ABI.

    python -m solids4foam_b200.run_case <caseDir> [--steps N] [--device 0] [--precond GAMG]
    python -m torch.distributed.run --nproc-per-node P --master-addr 127.0.0.1 -m solids4foam_b200.run_case <caseDir>

Writes ``<time>/D``, ``<time>/sigma``.  Under torchrun (one rank per GPU) the case must be decomposed
(``processorN/constant/polyMesh``, ``processorN/0/D`` as decomposePar writes them; ``foam_io.decompose_case`` makes them
for a serial case): every rank reads its own processor directory, runs its part of the mesh and writes
``processorN/<time>/``.  No CPU fallback."""
from __future__ import annotations

import argparse
import os

import numpy as np

from . import case as K
from . import foam_io as IO
from .solid_model import SolidModel


def dist_exchange(rank: int):
    """Cell centres across processor patches through torch.distributed: what fv_mesh_from_poly needs for the
    interpolation geometry of the cut faces."""
    import torch.distributed as dist

    def exchange(send):
        world = dist.get_world_size()
        out = [None] * world
        dist.all_gather_object(out, {(rank, q): a for q, a in send.items()})
        table = {}
        for d in out:
            table.update(d)
        return {q: table[(q, rank)] for q in send}
    return exchange


def run(case_dir: str, steps: int | None = None, device: int = 0, precond: str | None = None, write: bool = True, log=print,
        rank: int = 0, world: int = 1, comm=None):
    over = {}
    if precond:
        over["preconditioner"] = getattr(K, "PRECOND_" + precond.upper())
    if world > 1:
        if not os.path.isdir(os.path.join(case_dir, f"processor{world - 1}")) or os.path.isdir(os.path.join(case_dir, f"processor{world}")):
            raise RuntimeError(f"{case_dir} is not decomposed into {world} processor directories")
        case = IO.read_decomposed_case(case_dir, rank, world, dist_exchange(rank), **over)
        out_dir = os.path.join(case_dir, f"processor{rank}")
        if rank != 0:
            log = lambda s: None
    else:
        case = IO.read_case(case_dir, **over)
        out_dir = case_dir
    if case.controls.preconditioner == K.PRECOND_DIC and precond is None:
        # DIC/FDIC in fvSolution: the device has it exactly (level scheduled, iteration counts of the CPU solver) but GAMG is the
        # fast preconditioner on a GPU; --precond DIC keeps the case's own choice
        case.controls.preconditioner = K.PRECOND_GAMG
        log("fvSolution asks for the DIC/FDIC preconditioner; running PCG with the GAMG preconditioner instead "
            "(same converged solution, fewer inner iterations; --precond DIC keeps the case's own choice)")
    cd = IO.read_foam_dict(os.path.join(case_dir, "system", "controlDict"))
    dt = float(cd.get("deltaT", 1.0))
    n = steps if steps is not None else max(1, int(round((float(cd.get("endTime", dt)) - float(cd.get("startTime", 0.0))) / dt)))
    solid = SolidModel(case, device=device, comm=comm)
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
            tdir = os.path.join(out_dir, f"{t:g}")
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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        run(a.case_dir, a.steps, a.device, a.precond)
        return
    # one rank per GPU under torchrun: the communicator id travels over torch.distributed
    import torch
    import torch.distributed as dist
    from .solid_model import nccl_unique_id
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    run(a.case_dir, a.steps, local, a.precond, rank=rank, world=world, comm=(world, rank, bytes(uid.cpu().tolist())))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
