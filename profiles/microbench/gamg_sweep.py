#!/usr/bin/env python
"""GAMG parameter sweep on the 8M-cell cantilever: ms per outer iteration and PCG iterations per component for
smoother degree / over-correction / cycle / precision.  Not a bench value (no clock sampling); a tuning aid."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from solids4foam_b200 import case as K  # noqa: E402
from solids4foam_b200 import cases  # noqa: E402
from solids4foam_b200.solid_model import SolidModel  # noqa: E402


def main():
    dims = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "800,100,100").split(","))
    case = cases.cantilever(*dims, preconditioner=K.PRECOND_GAMG)
    g = SolidModel(case)
    zero = np.zeros((case.mesh.nCells, 3))
    configs = []
    for deg in (1, 2, 3):
        for om in (1.0, 1.4, 1.8, 2.2):
            configs.append(dict(gamgSmootherDegree=deg, gamgOverCorrection=om, gamgCycle=0, gamgSinglePrecision=0))
    configs += [dict(gamgSmootherDegree=2, gamgOverCorrection=1.8, gamgCycle=1, gamgSinglePrecision=0),
                dict(gamgSmootherDegree=1, gamgOverCorrection=1.8, gamgCycle=1, gamgSinglePrecision=0),
                dict(gamgSmootherDegree=2, gamgOverCorrection=1.8, gamgCycle=0, gamgSinglePrecision=1)]
    if len(sys.argv) > 2:
        configs = [eval("dict(" + a + ")") for a in sys.argv[2:]]
    for cfg in configs:
        cfg = dict(cfg)
        if "omegaK" in cfg:          # K-cycle tuning aid: scaling of the fine-level correction (read by the library at set-up)
            os.environ["S4F_GAMG_OMEGA_K"] = str(cfg.pop("omegaK"))
        ctl = K.default_controls(preconditioner=K.PRECOND_GAMG, **cfg)
        g.set_controls(ctl)
        g.set("D", zero); g.set("sigma", np.zeros((case.mesh.nCells, 6)))
        g.initialise()
        for _ in range(5):          # the bench's window: outer iterations 6-25 from D = 0
            g.outer_iteration()
        g.synchronize()
        t0 = time.perf_counter()
        its = []
        for _ in range(20):
            st = g.outer_iteration()
            its.append(st["nIterations"])
        g.synchronize()
        ms = (time.perf_counter() - t0) / 20 * 1e3
        print(cfg, f"{ms:7.2f} ms/outer", "iters", np.mean(its, axis=0).round(1).tolist(), "max", np.max(np.array(its), axis=1).tolist(), flush=True)


if __name__ == "__main__":
    main()
