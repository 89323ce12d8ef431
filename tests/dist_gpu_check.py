#!/usr/bin/env python
"""Multi-GPU parity check, one rank per GPU over NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_gpu_check.py [nx ny nz] [precond]

Each rank owns an x-slab (decomposePar simple (P 1 1)) of the hex cantilever; rank 0 gathers D and sigma and
compares them with the single-domain CPU oracle after the same number of outer iterations (first iterate:
round-off; converged: <= 1e-6).  Exit code 0 = parity."""
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
    dims = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 6, 6)
    pre = getattr(K, "PRECOND_" + (sys.argv[4] if len(sys.argv) > 4 else "DIAGONAL"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    kw = dict(L=2.0, fieldRelaxD=0.9, nCorrectors=4000, solutionTolerance=1e-11, alternativeTolerance=1e-11, tolerance=1e-13)
    case = cases.cantilever(*dims, rank=rank, nRanks=world, preconditioner=pre, **kw)
    g = SolidModel(case, device=local, comm=(world, rank, bytes(uid.cpu().tolist())))

    def gather(name, ncomp):
        loc = g.get(name)
        out = [None] * world
        dist.all_gather_object(out, (case.mesh.cellGlobal, loc))
        full = np.zeros((dims[0] * dims[1] * dims[2], ncomp))
        for cg, a in out:
            full[cg] = a
        return full

    st1 = g.outer_iteration()
    D1 = gather("D", 3)
    st = g.evolve()
    D, S = gather("D", 3), gather("sigma", 6)
    ok = True
    if rank == 0:
        from oracle.binding import OracleSolid
        o = OracleSolid(cases.cantilever(*dims, preconditioner=K.PRECOND_DIAGONAL if pre != K.PRECOND_GAMG else K.PRECOND_DIC, **kw))
        so1 = o.outer_iteration()
        e1 = np.linalg.norm(D1 - o.get("D")) / np.linalg.norm(o.get("D"))
        so = o.evolve()
        eD = np.linalg.norm(D - o.get("D")) / np.linalg.norm(o.get("D"))
        eS = np.linalg.norm(S - o.get("sigma")) / np.linalg.norm(o.get("sigma"))
        print(f"world {world} precond {pre}: first iterate relL2(D) {e1:.2e} iters gpu {st1['nIterations']} oracle {so1['nIterations']}; "
              f"converged gpu {st['converged']} ({st['nCorr']}) oracle {so['converged']} ({so['nCorr']}); relL2 D {eD:.2e} sigma {eS:.2e}")
        ok = bool(st["converged"] and so["converged"] and eD < 1e-6 and eS < 1e-6)
        if pre != K.PRECOND_GAMG:
            ok = ok and e1 < 1e-9 and st1["nIterations"] == so1["nIterations"]
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
