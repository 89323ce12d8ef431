#!/usr/bin/env python
"""The operators with a point stencil on a decomposed mesh (pointCellsLeastSquares gradient, vol->point
interpolation), one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
        tests/dist_point_check.py <emptyDir> [checker|slab] [pointCells|uns] [box|warped]

Rank 0 writes a serial cantilever case with ``gradSchemes default pointCellsLeastSquares`` on a general (points + faces)
mesh and decomposes it like decomposePar with a manual cell -> processor map: ``checker`` cuts the beam into 2 x 2 blocks in
x and y and deals them out like a chess board, so the processor boundary has corners and edges where four blocks meet --
the points there see cells and boundary faces of the other rank that no processor FACE connects them to.  Every rank reads its
processorN directory, and the checks against the single-domain CPU oracle on the serial mesh are

* grad(D) of an analytic (cubic) displacement field: the complete pointCells stencil incl. remote cells / boundary faces;
* both vol->point interpolations of that field (patch mode and gradient-extrapolated), point by point;
* the converged solution of the case (D, sigma) to north_star's 1e-6.

``ul`` / ``unsul`` run the updated-Lagrangian models (cell-centred with pointCellsLeastSquares, and the face-stress form) with
neoHookeanElastic through two load steps with the mesh moved in between on every rank's device (s4fgpu_move_points: consistent
point displacements from the point-neighbour ghosts, neighbour centres by halo exchange), against the serial oracle.

``uns`` runs the unsLinearGeometry model instead (vertex values from vol->point, face gradients and face stress; the processor
faces take the corrected snGrad of an internal face with the cell across the cut); ``warped`` distorts the mesh so that the
non-orthogonal correction (and its ghost gradients) is exercised."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def field(C):
    x, y, z = C[:, 0], C[:, 1], C[:, 2]
    return 1e-3 * np.stack([x * y + 0.3 * z * z * x, np.sin(0.7 * x) + y * z, 0.2 * x * x * y - z ** 3], axis=1)


def moving_steps(solid, case, case_dir, rank, world, IO, K, dist, model, mode, shape):
    """two load steps of an updated-Lagrangian model, mesh moved on the devices in between; rank 0 compares with the oracle"""
    loads = (-1e4, -2e4)
    res = []
    pts0 = solid.case.mesh.points.copy()
    for load in loads:
        solid.new_timestep(1.0)
        solid.set_bc("loaded", K.solidTraction((0.0, load, 0.0)))
        st = solid.evolve()
        D, S = solid.get("D"), solid.get("sigma")
        solid.updateTotalFields()
        res.append((st["converged"], st["nCorr"], D, S, solid.case.mesh.points.copy(), solid.get("rho")))
    out = [None] * world
    dist.all_gather_object(out, (solid.case.mesh.cellGlobal, res, pts0))
    if rank != 0:
        return True
    from oracle.binding import OracleSolid
    serial = IO.read_case(case_dir, preconditioner=K.PRECOND_DIC)
    o = OracleSolid(serial)
    N = serial.mesh.nCells
    ok = True
    p0 = serial.mesh.points.copy()
    msg = []
    for step, load in enumerate(loads):
        o.new_timestep(1.0)
        o.set_bc("loaded", K.solidTraction((0.0, load, 0.0)))
        so = o.evolve()
        Do, So = o.get("D"), o.get("sigma")
        o.updateTotalFields()
        D = np.zeros((N, 3)); S = np.zeros((N, 6)); R = np.zeros(N)
        conv = True
        ePts = 0.0
        key = {tuple(x): i for i, x in enumerate(np.round(p0, 12).tolist())}
        for r, (cg, rs, _) in enumerate(out):
            c_, n_, d_, s_, pts_, rho_ = rs[step]
            D[cg] = d_; S[cg] = s_; R[cg] = rho_
            conv = conv and c_
        # moved points: every rank's copy against the oracle's moved serial mesh, matched through the ORIGINAL coordinates
        for r in range(world):
            ids = np.array([key[tuple(x)] for x in np.round(out[r][2], 12).tolist()])
            ePts = max(ePts, np.abs(out[r][1][step][4] - o.case.mesh.points[ids]).max())
        eD = np.linalg.norm(D - Do) / np.linalg.norm(Do)
        eS = np.linalg.norm(S - So) / np.linalg.norm(So)
        eR = np.abs(R - o.get("rho")).max() / np.abs(o.get("rho")).max()
        msg.append(f"step {step}: gpu converged {conv} oracle {so['converged']} ({so['nCorr']}); relL2 D {eD:.2e} sigma {eS:.2e} rho {eR:.2e} points {ePts:.2e}")
        ok = ok and conv and so["converged"] and eD < 1e-6 and eS < 1e-6 and eR < 1e-8 and ePts < 1e-8
    print(f"moving mesh on {world} GPUs ({mode}, {model}, {shape}): " + "; ".join(msg) + f"; max |point motion| {np.abs(o.case.mesh.points - p0).max():.3f}", flush=True)
    return bool(ok)


def main():
    import torch
    import torch.distributed as dist
    from solids4foam_b200 import case as K
    from solids4foam_b200 import cases
    from solids4foam_b200 import foam_io as IO
    from solids4foam_b200 import run_case
    from solids4foam_b200.solid_model import SolidModel, nccl_unique_id
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    case_dir = sys.argv[1]
    mode = sys.argv[2] if len(sys.argv) > 2 else "checker"
    model = sys.argv[3] if len(sys.argv) > 3 else "pointCells"
    shape = sys.argv[4] if len(sys.argv) > 4 else "box"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny, nz = 12, 6, 4
    kw = dict(L=2.0, fieldRelaxD=0.9, nCorrectors=6000, solutionTolerance=1e-11, alternativeTolerance=1e-11, tolerance=1e-13,
              preconditioner=K.PRECOND_GAMG, general=True)
    moving = model in ("ul", "unsul")
    if model == "uns":
        kw["solidModel"] = K.MODEL_UNS_LIN_GEOM
    elif model == "unsul":
        kw["solidModel"] = K.MODEL_UNS_NONLIN_UL
    elif model == "ul":
        kw.update(solidModel=K.MODEL_NONLIN_UL, gradScheme=K.GRAD_POINT_CELLS_LEAST_SQUARES)
    else:
        kw["gradScheme"] = K.GRAD_POINT_CELLS_LEAST_SQUARES
    if moving:
        kw.update(solutionTolerance=1e-9, alternativeTolerance=1e-9, tolerance=1e-12, relTol=0.01, nCorrectors=8000)
        kw.pop("fieldRelaxD", None)
    if rank == 0:
        serial = (cases.neo_hookean_cantilever(nx, ny, nz, traction=(0.0, -1e4, 0.0), **kw) if moving
                  else cases.cantilever(nx, ny, nz, **kw))
        if shape == "warped":
            from solids4foam_b200 import mesh as M

            def warp(p):
                q = p.copy()
                q[:, 0] += 0.04 * np.sin(2.1 * p[:, 1] + 1.3 * p[:, 2])
                q[:, 1] += 0.03 * np.sin(1.7 * p[:, 0]) * (1 + 0.5 * p[:, 2])
                q[:, 2] += 0.03 * np.cos(1.1 * p[:, 0] + 0.9 * p[:, 1])
                return q
            serial.mesh = M.hex_box_general(nx, ny, nz, 2.0, 1.0, 1.0, point_map=warp, names=("fixed", "loaded", "yMin", "yMax", "zMin", "zMax"))
        # a non-uniform prescribed displacement on the clamped end: the boundary points there take boundary-face values, some
        # of them faces of the other rank
        pf = serial.mesh.patch("fixed")
        F = serial.mesh.nInternalFaces
        if not moving:
            serial.bcs["fixed"].value = 1e-2 * field(serial.mesh.Cf[F + pf.start:F + pf.start + pf.size])
        IO.write_case(case_dir, serial, end_time=1.0)
        C = serial.mesh.C
        if mode == "checker":
            # blocks from the cell numbering (x fastest), not from the (possibly warped) centres
            idx = np.arange(serial.mesh.nCells)
            bx = ((idx % nx) >= nx // 2).astype(np.int64)
            by = (((idx // nx) % ny) >= ny // 2).astype(np.int64)
            cell_rank = ((bx + by) % 2) if world == 2 else (bx + 2 * by) % world
        else:
            cell_rank = None
        IO.decompose_case(case_dir, world, cell_rank=cell_rank)
    dist.barrier()
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    case = IO.read_decomposed_case(case_dir, rank, world, run_case.dist_exchange(rank), preconditioner=K.PRECOND_GAMG)
    solid = SolidModel(case, device=local, comm=(world, rank, bytes(uid.cpu().tolist())))
    m = case.mesh
    if moving:
        ok = moving_steps(solid, case, case_dir, rank, world, IO, K, dist, model, mode, shape)
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.broadcast(flag, 0)
        dist.destroy_process_group()
        sys.exit(0 if int(flag.item()) else 1)
    # operators on an analytic field, fresh model
    solid.set("D", field(m.C))
    solid.op_grad()
    g = solid.get("gradD")
    pP = solid.interpolate_to_points("D", with_gradient=False)
    pG = solid.interpolate_to_points("D", with_gradient=True)
    # the case itself
    solid.set("D", np.zeros((m.nCells, 3)))
    solid.op_grad()
    solid.new_timestep(1.0)
    st = solid.evolve()
    Dsol, Ssol = solid.get("D"), solid.get("sigma")
    out = [None] * world
    dist.all_gather_object(out, (m.cellGlobal, Dsol, Ssol, g, m.points, pP, pG))
    ok = True
    if rank == 0:
        from oracle.binding import OracleSolid
        serial = IO.read_case(case_dir, preconditioner=K.PRECOND_DIC)
        o = OracleSolid(serial)
        N = serial.mesh.nCells
        o.set("D", field(serial.mesh.C))
        o.op_grad()
        Go = o.get("gradD")
        oP = o.interpolate_to_points("D", with_gradient=False)
        oG = o.interpolate_to_points("D", with_gradient=True)
        o.set("D", np.zeros((N, 3)))
        o.op_grad()
        o.new_timestep(1.0)
        so = o.evolve()
        D = np.zeros((N, 3)); S = np.zeros((N, 6)); G = np.zeros((N, 9))
        for cg, d, s_, gg, *_ in out:
            D[cg] = d; S[cg] = s_; G[cg] = gg
        eD = np.linalg.norm(D - o.get("D")) / np.linalg.norm(o.get("D"))
        eS = np.linalg.norm(S - o.get("sigma")) / np.linalg.norm(o.get("sigma"))
        eG = np.abs(G - Go).max() / np.abs(Go).max()
        key = {tuple(x): i for i, x in enumerate(np.round(serial.mesh.points, 12).tolist())}
        eP = eGp = 0.0
        shared = {}
        for r, (_, _, _, _, pts, pP_r, pG_r) in enumerate(out):
            ids = np.array([key[tuple(x)] for x in np.round(pts, 12).tolist()])
            eP = max(eP, np.abs(pP_r - oP[ids]).max() / np.abs(oP).max())
            eGp = max(eGp, np.abs(pG_r - oG[ids]).max() / np.abs(oG).max())
            for i, v in zip(ids.tolist(), pP_r.tolist()):
                shared.setdefault(i, []).append(v)
        nShared = sum(1 for v in shared.values() if len(v) > 1)
        # a point held by several ranks gets the same value on each of them (no partial sums to synchronise)
        spread = max((np.ptp(np.array(v), axis=0).max() for v in shared.values() if len(v) > 1), default=0.0) / np.abs(oP).max()
        print(f"point stencils on {world} GPUs ({mode}, {model}, {shape}): grad(D) max rel err {eG:.2e}; vol->point patch {eP:.2e} grad-extrapolated {eGp:.2e}; "
              f"{nShared} shared points, spread between ranks {spread:.2e}; converged gpu {st['converged']} ({st['nCorr']}) "
              f"oracle {so['converged']} ({so['nCorr']}); relL2 D {eD:.2e} sigma {eS:.2e}", flush=True)
        ok = bool(st["converged"] and so["converged"] and eD < 1e-6 and eS < 1e-6 and eG < 1e-11 and eP < 1e-12 and eGp < 1e-11
                  and spread < 1e-13 and nShared > 0)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
