"""N>1 host logic on CPU: two `gloo` ranks (127.0.0.1) run the decomposed-mesh algorithm the CUDA path uses --
processor-patch halo exchange in patch-face order + all-reduced dot products -- with numpy standing in
for the kernels, and must reproduce the single-domain solve.  This pins what the multi-GPU path relies on
from the host side: the slab decomposition, the processor-patch face order / faceCells (the send list of
s4f_build_rows), the coupled-face coefficients, and the rank plumbing used by bench.py."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from solids4foam_b200 import mesh as M

NX, NY, NZ = 12, 4, 3
BOX = (8.0, 1.0, 1.0)
IMPK = 2.7e11


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _coeffs(m):
    """laplacian coefficients of every face: impK * magSf * nonOrthDeltaCoeffs ([OF-ext] gaussLaplacianScheme)."""
    return IMPK * m.magSf * m.nonOrthDeltaCoeffs


def _local_system(m):
    F = m.nInternalFaces
    a = _coeffs(m)
    N = m.nCells
    diag = np.zeros(N)
    np.add.at(diag, m.owner, a[:F])
    np.add.at(diag, m.neighbour, a[:F])
    halo = []
    for p in m.patches:
        sl = slice(p.start, p.start + p.size)
        if p.kind == M.PROCESSOR:
            np.add.at(diag, m.faceCells[sl], a[F:][sl])
            halo.append((p.nbr_rank, m.faceCells[sl].copy(), a[F:][sl].copy()))
        elif p.name == "xMin":              # fixed-value patch: internalCoeffs = impK magSf delta
            np.add.at(diag, m.faceCells[sl], a[F:][sl])
    A = sp.coo_matrix((np.concatenate([-a[:F], -a[:F], diag]),
                       (np.concatenate([m.owner, m.neighbour, np.arange(N)]),
                        np.concatenate([m.neighbour, m.owner, np.arange(N)]))), shape=(N, N)).tocsr()
    return A, diag, halo


def _rhs(cell_global):
    rng = np.random.default_rng(99)
    full = rng.standard_normal(NX * NY * NZ)
    return full[cell_global]


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names = ("xMin", "xMax", "yMin", "yMax", "zMin", "zMax")
    m = M.hex_box_decomposed(NX, NY, NZ, *BOX, rank, world, names=names)
    A, diag, halo = _local_system(m)
    b = _rhs(m.cellGlobal)

    def amul(x):                     # local rows + processor-patch contribution of the neighbour's cells
        y = A @ x
        reqs, recv = [], []
        for nbr, cells, coef in halo:
            send = torch.from_numpy(np.ascontiguousarray(x[cells]))
            buf = torch.empty(len(cells), dtype=torch.float64)
            reqs.append(dist.isend(send, nbr)); reqs.append(dist.irecv(buf, nbr))
            recv.append((cells, coef, buf))
        for r in reqs:
            r.wait()
        for cells, coef, buf in recv:
            np.add.at(y, cells, -coef * buf.numpy())
        return y

    def gsum(v):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    x = np.zeros(m.nCells)
    r = b - amul(x)
    rD = 1.0 / diag
    p = np.zeros_like(x)
    rho_old = 1.0
    for it in range(500):
        z = rD * r
        rho = gsum(z @ r)
        p = z if it == 0 else z + (rho / rho_old) * p
        w = amul(p)
        alpha = rho / gsum(w @ p)
        x += alpha * p
        r -= alpha * w
        rho_old = rho
        if gsum(np.abs(r).sum()) < 1e-9 * gsum(np.abs(b).sum()):
            break
    g = [None] * world
    dist.all_gather_object(g, (m.cellGlobal, x, it))
    if rank == 0:
        full = np.zeros(NX * NY * NZ)
        for cg, xx, _ in g:
            full[cg] = xx
        np.save(out, full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_halo_pcg_matches_single_domain(tmp_path, world):
    import torch.multiprocessing as mp
    out = str(tmp_path / "x.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    x_par = np.load(out)
    names = ("xMin", "xMax", "yMin", "yMax", "zMin", "zMax")
    whole = M.hex_box(NX, NY, NZ, *BOX, names=names)
    A, diag, halo = _local_system(whole)
    assert not halo
    x_ref = spla.spsolve(A.tocsc(), _rhs(np.arange(whole.nCells)))
    assert np.linalg.norm(x_par - x_ref) / np.linalg.norm(x_ref) < 1e-7


def test_bench_reference_arm_runs_on_rank0_only():
    """`bench.py --impl reference` under a 2-rank launch: rank 1 exits 0 without output."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0", "--cells", "8,2,2"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
