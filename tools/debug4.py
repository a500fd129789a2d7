import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import hydrob200, oracle
from cases import CASES
name = "C1_sod_fe_donor"
cfg, n = CASES[name]
R = hydrob200.FiniteVolumeSolver(dict(cfg, backend=oracle.OracleBackend))
G = hydrob200.FiniteVolumeSolver(dict(cfg, strict_fp=True))
a0 = G.getState(); b0 = R.getState()
print("init equal", np.array_equal(a0, b0))
G.update(); R.update()
a = G.getState(); b = R.getState()
bad = np.argwhere(a != b)
print(len(bad), bad[:3].tolist(), bad[-3:].tolist())
for bb in list(bad[:3]) + list(bad[-3:]):
    print(tuple(bb), "got %r ref %r init %r" % (a[tuple(bb)], b[tuple(bb)], a0[tuple(bb)]))
print("t", G.t, R.t, "dt", G.dt, R.dt)
i = 150
print("row got", a[0,0,148:156].tolist())
print("row ref", b[0,0,148:156].tolist())
