import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from solids4foam_b200 import case as K, cases
from solids4foam_b200.solid_model import SolidModel
case = cases.cantilever(800, 100, 100, preconditioner=K.PRECOND_GAMG)
mesh = case.mesh
g = SolidModel(case)
for _ in range(5): g.outer_iteration()
N = mesh.nCells
hD = torch.zeros((N, 3), dtype=torch.float64).pin_memory().numpy()
hDold = torch.zeros((N, 3), dtype=torch.float64).pin_memory().numpy()
hOut = [torch.zeros((N, nc), dtype=torch.float64).pin_memory().numpy() for nc in (3, 9, 6)]
hD[:] = g.get("D")
p = [q for q in mesh.patches if q.name == "loaded"][0]
trac = torch.zeros((p.size, 3), dtype=torch.float64).pin_memory().numpy(); trac[:, 1] = -1e6
for rep in range(3):
    g.synchronize(); t0 = time.perf_counter()
    g.set("D", hD); t1 = time.perf_counter()
    g.set("D_old", hDold); t2 = time.perf_counter()
    its = []
    for _ in range(10):
        ta = time.perf_counter(); g.setTraction("loaded", trac); tb = time.perf_counter()
        st = g.outer_iteration(); tc = time.perf_counter()
        its.append((round((tb - ta) * 1e3, 2), round((tc - tb) * 1e3, 2), sum(st["nIterations"])))
    t3 = time.perf_counter()
    td = []
    for name, buf in zip(("D", "gradD", "sigma"), hOut):
        x = time.perf_counter(); g.get(name, out=buf); td.append(round((time.perf_counter() - x) * 1e3, 1))
    g.synchronize(); t4 = time.perf_counter()
    print(f"rep {rep}: setD {1e3*(t1-t0):.1f} setDold {1e3*(t2-t1):.1f} loop {1e3*(t3-t2):.1f} downloads {td} total {1e3*(t4-t0):.1f} ms")
    print("   steps (setTraction ms, outer ms, inner its):", its)
