import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import hydrob200, oracle
from cases import CASES
name = "C1_sod_fe_donor"
cfg, n = CASES[name]
def runR(k=n):
    R = hydrob200.FiniteVolumeSolver(dict(cfg, backend=oracle.OracleBackend))
    for _ in range(k): R.update()
    return R.getState()
def runG(k=n):
    G = hydrob200.FiniteVolumeSolver(dict(cfg, strict_fp=True))
    for _ in range(k): G.update()
    return G.getState()
def d(a, b): return [(int((a[..., q] != b[..., q]).sum())) for q in range(6)]
r0 = runR()
r0b = runR()
print("R twice before CUDA:", d(r0, r0b))
g = runG()
r1 = runR()
print("R before vs after CUDA init:", d(r0, r1))
print("G vs R-before:", d(g, r0), " G vs R-after:", d(g, r1))
# interleaved
R = hydrob200.FiniteVolumeSolver(dict(cfg, backend=oracle.OracleBackend))
G = hydrob200.FiniteVolumeSolver(dict(cfg, strict_fp=True))
for i in range(n):
    G.update(); R.update()
    if G.dt != R.dt:
        print("dt differs at step", i, G.dt, R.dt); break
print("interleaved:", d(G.getState(), R.getState()), "vs r0:", d(R.getState(), r0))
for k in (1, 2, 3, 5, 10, 20):
    print(k, d(runG(k), runR(k)))
