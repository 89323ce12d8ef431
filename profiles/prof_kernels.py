#!/usr/bin/env python
"""Short driver for ncu: builds a case and launches every hot kernel a few times through the C-ABI timing entry
(s4fgpu_time_kernel).

    S4F_NO_GRAPH=1 S4F_TIME_WARMUP=0 ncu --set full --clock-control none -k regex:'k_amg|k_kc|k_amul|k_source|k_grad|k_pcg|k_law|k_tl' \
        -c 48 -o /tmp/prof python profiles/prof_kernels.py 800,100,100 cantilever GAMG
    (S4F_TIME_WARMUP=0: one launch per kernel; the K-cycle comes last, its first ~35 launches are the two finest levels)

workload: cantilever (linearElastic, orthogonal) | notched_bar (neoHookeanElasticMisesPlastic, non-orthogonal, total Lagrangian) |
neo_hookean.  Numbers printed under ncu are NOT bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from solids4foam_b200 import case as K  # noqa: E402
from solids4foam_b200 import cases  # noqa: E402
from solids4foam_b200.solid_model import SolidModel  # noqa: E402


def main():
    dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "800,100,100").split(","))
    workload = sys.argv[2] if len(sys.argv) > 2 else "cantilever"
    pre = getattr(K, "PRECOND_" + (sys.argv[3] if len(sys.argv) > 3 else "GAMG"))
    fn = dict(cantilever=cases.cantilever, notched_bar=cases.notched_bar, neo_hookean=cases.neo_hookean_cantilever)[workload]
    g = SolidModel(fn(*dims, preconditioner=pre))
    names = ["grad", "rhs", "law"]
    if workload == "cantilever":
        names = ["spmv3", "spmv1", "pcg_p", "pcg_xr"] + names
        if pre == K.PRECOND_GAMG:
            names += ["gamg_vcycle"]
    for name in names:
        ms, by = g.time_kernel(name, reps=1, flush_l2=False)
        print(f"{name:9s} {ms:8.4f} ms  {by / ms / 1e6:8.1f} GB/s (under a profiler: not a bench value)")


if __name__ == "__main__":
    main()
