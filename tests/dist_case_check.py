#!/usr/bin/env python
"""Decomposed case directory on several GPUs, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tests/dist_case_check.py <emptyDir>

Rank 0 writes a serial solids4foam case (constant/polyMesh, dictionaries, 0/D) and decomposes it the way decomposePar does
(processorN/constant/polyMesh with processor patches, processorN/0/D); every rank then reads ITS processor directory,
runs the standalone driver on its GPU and writes processorN/<time>/D.  Rank 0 reassembles D through cellProcAddressing and
compares it with the single-domain CPU oracle run of the serial case.  Exit code 0 = parity (<= 1e-6, north_star)."""
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
    from solids4foam_b200 import foam_io as IO
    from solids4foam_b200 import run_case
    from solids4foam_b200.solid_model import nccl_unique_id
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    case_dir = sys.argv[1]
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kw = dict(L=2.0, fieldRelaxD=0.9, nCorrectors=4000, solutionTolerance=1e-11, alternativeTolerance=1e-11, tolerance=1e-13,
              preconditioner=K.PRECOND_GAMG, general=True)
    if rank == 0:
        IO.write_case(case_dir, cases.cantilever(16, 6, 6, **kw), end_time=1.0)
        IO.decompose_case(case_dir, world)
    dist.barrier()
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    solid, stats = run_case.run(case_dir, steps=1, device=local, rank=rank, world=world, comm=(world, rank, bytes(uid.cpu().tolist())),
                                log=lambda s: None)
    m = solid.case.mesh
    Dfile, _ = IO.read_vol_field(os.path.join(case_dir, f"processor{rank}", "1", "D"), m)
    same_file = bool(np.array_equal(Dfile, solid.get("D")))
    out = [None] * world
    dist.all_gather_object(out, (m.cellGlobal, solid.get("D"), solid.get("sigma"), same_file))
    ok = True
    if rank == 0:
        from oracle.binding import OracleSolid
        serial = IO.read_case(case_dir, preconditioner=K.PRECOND_DIC)
        o = OracleSolid(serial)
        o.new_timestep(1.0)
        so = o.evolve()
        D = np.zeros((serial.mesh.nCells, 3)); S = np.zeros((serial.mesh.nCells, 6))
        for cg, d, s, _ in out:
            D[cg] = d; S[cg] = s
        eD = np.linalg.norm(D - o.get("D")) / np.linalg.norm(o.get("D"))
        eS = np.linalg.norm(S - o.get("sigma")) / np.linalg.norm(o.get("sigma"))
        print(f"decomposed case on {world} GPUs: converged gpu {stats[0]['converged']} ({stats[0]['nCorr']}) oracle {so['converged']} ({so['nCorr']}); "
              f"relL2 D {eD:.2e} sigma {eS:.2e}; written files equal the device fields: {[x[3] for x in out]}")
        ok = bool(stats[0]["converged"] and so["converged"] and eD < 1e-6 and eS < 1e-6 and all(x[3] for x in out))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
