#!/usr/bin/env python
"""GAMG set-up time on the 8M-cell cantilever, twice in one process (cold, then warm allocator / module state):
S4F_AMG_TIMING=1 python profiles/microbench/gamg_setup_time.py [nx,ny,nz]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from solids4foam_b200 import case as K  # noqa: E402
from solids4foam_b200 import cases  # noqa: E402
from solids4foam_b200.solid_model import SolidModel  # noqa: E402

dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "800,100,100").split(","))
case = cases.cantilever(*dims, preconditioner=K.PRECOND_GAMG)
g = SolidModel(case)
for rep in range(3):
    ctl = K.default_controls(preconditioner=K.PRECOND_GAMG, gamgOverCorrection=2.2 + 0.0001 * rep)     # a changed GAMG parameter: full set-up again
    g.set_controls(ctl)
    g.initialise()
    t0 = time.perf_counter()
    g.outer_iteration(); g.synchronize()
    print(f"rep {rep}: first outer iteration incl. set-up {time.perf_counter() - t0:.3f} s; gamg_info {g.gamg_info()}", flush=True)
