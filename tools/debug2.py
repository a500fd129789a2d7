import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import hydrob200, oracle
from cases import CASES
name = "C1_sod_fe_donor"
cfg, n = CASES[name]
def runR():
    R = hydrob200.FiniteVolumeSolver(dict(cfg, backend=oracle.OracleBackend))
    for _ in range(n): R.update()
    return R.getState()
def runG(sync_each, **kw):
    G = hydrob200.FiniteVolumeSolver(dict(cfg, strict_fp=True, **kw))
    for _ in range(n):
        G.update()
        if sync_each: G.getState()
    return G.getState()
r = runR()
for se in (True, False, True, False):
    g = runG(se)
    print("getState each step:", se, "mismatches per var", [(int((g[..., q] != r[..., q]).sum())) for q in range(6)])
g1 = runG(False); g2 = runG(False)
print("two GPU runs equal:", np.array_equal(g1, g2))
G = hydrob200.FiniteVolumeSolver(dict(cfg, strict_fp=True))
G.update(n)
g = G.getState()
print("update(n) at once:", [(int((g[..., q] != r[..., q]).sum())) for q in range(6)])
