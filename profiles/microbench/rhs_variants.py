"""k_source_g (right-hand side on non-orthogonal meshes) launch variants at 8 M cells: python profiles/microbench/rhs_variants.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from solids4foam_b200 import case as K
from solids4foam_b200 import cases
from solids4foam_b200.solid_model import SolidModel

dims = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (800, 100, 100)
g = SolidModel(cases.notched_bar(*dims, preconditioner=K.PRECOND_GAMG))
for _ in range(3):
    g.outer_iteration()
for minb in ("2", "3", "6"):
    os.environ["S4F_SRCG_MINB"] = minb
    ms, by = g.time_kernel("rhs", reps=20, flush_l2=False)
    print(f"k_source_g minb {minb}: {ms:.4f} ms  {by / ms / 1e6:.0f} GB/s for {by / 1e9:.3f} GB")
