import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import hydrob200, oracle
from cases import CASES

def cmp(tag, a, b):
    nv = a.shape[-1]
    print(tag, "mismatch per var:", [(int((a[..., q] != b[..., q]).sum())) for q in range(nv)],
          "max|diff|", [float(np.abs(a[..., q] - b[..., q]).max()) for q in range(nv)])

for name in sys.argv[1:] or ["C1_sod_fe_donor", "C4_sphere_rk4"]:
    cfg, n = CASES[name]
    R = hydrob200.FiniteVolumeSolver(dict(cfg, backend=oracle.OracleBackend))
    G = hydrob200.FiniteVolumeSolver(dict(cfg, strict_fp=True, use_graph=False))
    print("==", name, G.backend.describe().splitlines()[0])
    for i in range(n):
        a0, b0 = G.getState(), R.getState()
        G.update(); R.update()
        a, b = G.getState(), R.getState()
        if (a != b).any() or G.t != R.t:
            cmp("FIRST MISMATCH after update %d (t %r vs %r)" % (i + 1, G.t, R.t), a, b)
            bad = np.argwhere(a != b)
            print(" where", bad[:8].tolist())
            for bb in bad[:4]:
                k, j, ii, q = bb
                print("  got %r ref %r   prev(all vars) %s" % (a[tuple(bb)], b[tuple(bb)], b0[k, j, ii].tolist()))
            # does the previous state agree? then a single update from identical states differs
            print(" prev states equal:", np.array_equal(a0, b0))
            break
    else:
        print("all", n, "updates identical")
