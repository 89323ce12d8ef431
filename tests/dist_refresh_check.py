#!/usr/bin/env python
"""GAMG coefficient refresh on a decomposed mesh, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
        tests/dist_refresh_check.py [nx ny nz]

A dynamic (Euler) cantilever in x-slabs: the first outer iteration builds the hierarchy, a new time step size re-assembles the
matrix (another diagonal) -> the hierarchy keeps its aggregates and re-sums its coefficients on the devices: couplings across
processor patches go to the other rank's aggregate, the gathered level exchanges the re-summed rows.  A random right-hand side
is then solved with the refreshed hierarchy and, in a second model with S4F_NO_AMG_REFRESH=1, with a hierarchy rebuilt from
scratch for the same matrix: same iteration counts (within one), same solution.  The sizes are chosen so that there are
distributed AND gathered levels (S4F_GAMG_REPLICATE_BELOW)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from solids4foam_b200 import case as K
    from solids4foam_b200 import cases
    from solids4foam_b200.solid_model import SolidModel, nccl_unique_id
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dims = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (96, 24, 24)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["S4F_GAMG_REPLICATE_BELOW"] = "1000"       # 55 k cells on 2 ranks: 27648 | 3456 per rank distributed, 864 gathered, 108 below it
    kw = dict(L=2.0, d2dt2Scheme=K.D2DT2_EULER, deltaT=1e-3, deltaT0=1e-3, g=(0.0, -9.81, 0.0), preconditioner=K.PRECOND_GAMG,
              tolerance=1e-11, relTol=0.0, maxIter=200)
    res = {}
    for mode in ("refresh", "rebuild"):
        if mode == "rebuild":
            os.environ["S4F_NO_AMG_REFRESH"] = "1"
        else:
            os.environ.pop("S4F_NO_AMG_REFRESH", None)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        case = cases.cantilever(*dims, rank=rank, nRanks=world, **kw)
        g = SolidModel(case, device=local, comm=(world, rank, bytes(uid.cpu().tolist())))
        g.new_timestep(1e-3)
        g.outer_iteration()                      # builds the hierarchy for this matrix
        info = g.gamg_info()
        g.new_timestep(2.5e-4)                   # another diagonal: the matrix is re-assembled
        st = g.outer_iteration()
        rng = np.random.default_rng(3 + rank)
        src = rng.standard_normal((case.mesh.nCells, 3))
        psi, sst = g.op_solve(np.zeros_like(src), src)
        res[mode] = (st["nIterations"], psi, sst["nIterations"], info)
        g.close()
    (it_a, psi_a, ss_a, ia), (it_b, psi_b, ss_b, ib) = res["refresh"], res["rebuild"]
    num = torch.tensor([float(np.sum((psi_a - psi_b) ** 2)), float(np.sum(psi_b ** 2))], dtype=torch.float64, device="cuda")
    dist.all_reduce(num)
    rel = float(torch.sqrt(num[0] / num[1]))
    ok = (ia["levels"] == ib["levels"] and ia["distributed_levels"] >= 1 and len(ia["levels"]) > ia["distributed_levels"] + 1
          and max(abs(a - b) for a, b in zip(ss_a, ss_b)) <= 1 and max(abs(a - b) for a, b in zip(it_a, it_b)) <= 1 and rel < 1e-8)
    if rank == 0:
        print(f"GAMG refresh on {world} GPUs: levels {ia['levels']} ({ia['distributed_levels']} distributed); PCG iterations of a random "
              f"right-hand side refreshed {ss_a} rebuilt {ss_b}; outer iteration {it_a} / {it_b}; relL2 of the solutions {rel:.2e}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
