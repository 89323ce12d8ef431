"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): launches tests/dist_gpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("precond,dims", [("DIAGONAL", "16 6 6"), ("GAMG", "16 6 6"), ("GAMG", "36 12 12")])
def test_two_gpu_slab_decomposition_matches_oracle(precond, dims):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(HERE, "dist_gpu_check.py")] + dims.split() + [precond]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_decomposed_case_directory_runs_on_two_gpus(tmp_path):
    """processorN/ directories (foam_io.decompose_case = decomposePar simple (2 1 1)) read rank by rank and run by the
    standalone driver, against the single-domain oracle run of the serial case."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29518", os.path.join(HERE, "dist_case_check.py"), str(tmp_path / "case")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("mode,model,shape", [("checker", "pointCells", "box"), ("slab", "pointCells", "warped"),
                                              ("checker", "uns", "box"), ("checker", "uns", "warped"),
                                              ("checker", "ul", "box"), ("slab", "unsul", "box")])
def test_point_stencils_on_a_decomposed_mesh(tmp_path, mode, model, shape):
    """pointCellsLeastSquares gradient, vol->point interpolation and the unsLinearGeometry model across processor patches
    (point-neighbour ghosts, s4f_build_point_ghosts): operators on an analytic field and the converged case against the
    single-domain oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", os.path.join(HERE, "dist_point_check.py"), str(tmp_path / "case"), mode, model, shape]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("model,shape", [("pointCells", "warped"), ("ul", "box")])
def test_point_stencils_on_a_two_by_two_decomposition(tmp_path, model, shape):
    """4 GPUs, blocks dealt out 2 x 2: the diagonal pairs of ranks touch only along the central edge -- no processor patch
    between them -- yet their cells share points; the point-neighbour ghosts find them through the point coordinates."""
    import torch
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "4", "--master-addr", "127.0.0.1",
           "--master-port", "29520", os.path.join(HERE, "dist_point_check.py"), str(tmp_path / "case"), "checker", model, shape]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_gamg_coefficient_refresh_on_two_gpus():
    """Same graph, new coefficients on a decomposed mesh: the hierarchy is refreshed on the devices (interface couplings to the
    other rank's aggregates, gathered rows exchanged) and preconditions like one rebuilt from scratch."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29521", os.path.join(HERE, "dist_refresh_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
