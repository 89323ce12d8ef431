"""Build libs4fgpu.so (CUDA kernels + C-ABI) in-tree for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libs4fgpu.so")
SOURCES = ["s4f_api.cu", "s4f_comm.cu", "s4f_geom.cu", "s4f_amg_setup.cu", "s4f_setup.cu", "s4f_pcg.cu", "s4f_fv.cu", "s4f_law.cu", "s4f_amg.cu", "s4f_pressure.cu", "s4f_uns.cu", "s4f_dic.cu"]
HEADERS = ["s4f_ctx.h", "s4f_comm.h", "s4f_amg_setup.h", "s4f_dev.cuh", os.path.join("..", "..", "include", "s4fgpu.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=true",
         "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or is_stale()):
        return SO
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    fail = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {s} ---\n{out}\n")
        fail |= p.returncode != 0
    if fail:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-shared", "-Wno-deprecated-gpu-targets", "-o", SO] + objs + ["-lnccl", "-lcudart", "-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
