"""ctypes loader of libs4fgpu.so (the CUDA hot path behind include/s4fgpu.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present the
product path raises.  The library is built in-tree by ``solids4foam_b200/build.py`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

from . import case as K

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libs4fgpu.so")
_LIB = None

# every symbol include/s4fgpu.h declares
EXPORTS = ["s4fgpu_create", "s4fgpu_destroy", "s4fgpu_last_error", "s4fgpu_version", "s4fgpu_get_unique_id",
           "s4fgpu_comm_init", "s4fgpu_set_mesh", "s4fgpu_set_geometry", "s4fgpu_set_law", "s4fgpu_set_controls",
           "s4fgpu_set_bc", "s4fgpu_upload", "s4fgpu_download", "s4fgpu_initialise", "s4fgpu_new_timestep",
           "s4fgpu_outer_iteration", "s4fgpu_evolve", "s4fgpu_update_total_fields", "s4fgpu_op_grad",
           "s4fgpu_op_correct", "s4fgpu_op_assemble", "s4fgpu_op_amul", "s4fgpu_op_solve", "s4fgpu_time_kernel",
           "s4fgpu_launch_count", "s4fgpu_gamg_info", "s4fgpu_timer_start", "s4fgpu_timer_stop", "s4fgpu_synchronize",
           "s4fgpu_set_points", "s4fgpu_interpolate_to_points", "s4fgpu_gamg_distributed_levels", "s4fgpu_move_points"]


def _preload_nccl():
    """Bind to the NCCL torch ships when torch is importable (same soname libnccl.so.2), so that one
    NCCL lives in the process; otherwise the system libnccl.so.2 is used."""
    try:
        import torch  # noqa: F401  (loads its bundled libnccl.so.2 when CUDA is available)
    except Exception:
        pass
    for cand in ("libnccl.so.2",):
        try:
            C.CDLL(cand, mode=C.RTLD_GLOBAL)
            return
        except OSError:
            pass
    try:
        import nvidia.nccl  # type: ignore
        p = os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2")
        C.CDLL(p, mode=C.RTLD_GLOBAL)
    except Exception:
        pass


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is missing: build it with `python -m solids4foam_b200.build` "
                           "(there is no CPU fallback for the solid-solver hot path)")
    _preload_nccl()
    L = C.CDLL(SO_PATH)
    H = C.c_void_p
    L.s4fgpu_create.argtypes = [C.POINTER(H), C.c_int]
    L.s4fgpu_create.restype = C.c_int
    L.s4fgpu_destroy.argtypes = [H]
    L.s4fgpu_last_error.argtypes = [H]
    L.s4fgpu_last_error.restype = C.c_char_p
    L.s4fgpu_version.restype = C.c_int
    L.s4fgpu_get_unique_id.argtypes = [C.c_char_p]
    L.s4fgpu_comm_init.argtypes = [H, C.c_int, C.c_int, C.c_char_p]
    L.s4fgpu_time_kernel.argtypes = [H, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.s4fgpu_timer_start.argtypes = [H]
    L.s4fgpu_timer_stop.argtypes = [H, C.POINTER(C.c_double)]
    L.s4fgpu_synchronize.argtypes = [H]
    L.s4fgpu_launch_count.argtypes = [H]
    L.s4fgpu_launch_count.restype = C.c_longlong
    L.s4fgpu_gamg_info.argtypes = [H, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.s4fgpu_gamg_distributed_levels.argtypes = [H]
    L.s4fgpu_move_points.argtypes = [H, C.POINTER(C.c_double)]
    K.declare_api(L, "s4fgpu_", H)
    _LIB = L
    return L
