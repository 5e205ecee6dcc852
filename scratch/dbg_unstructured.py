import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from test_unstructured import _case, _oracle
from helpers import rel_l2, tight
from rheotool_b200 import abi, cases
from rheotool_b200.stress import GpuStressModel
m, models, U, Ub, phi, theta0, thetaB, dt = _case()
fpc = np.bincount(m.owner, minlength=m.n_cells) + np.bincount(m.neighbour, minlength=m.n_cells)
nint = m.n_internal
# pairs of cells sharing more than one face
pairs = {}
for f in range(nint):
    pairs.setdefault((int(m.owner[f]), int(m.neighbour[f])), []).append(f)
multi = {k: v for k, v in pairs.items() if len(v) > 1}
print("cell pairs sharing >1 face:", len(multi))
bcell = np.zeros(m.n_cells, bool); bcell[m.owner[nint:]] = True
for lim in ("upwind", "cubista", "none"):
    sc = tight(cases.scheme_ctl(lim, "PBiCGStab", 1e-10))
    oc, vals, vecs = _oracle(m, models, sc, U, Ub, phi, theta0, thetaB)
    g = GpuStressModel(m, models, sc)
    g.upload_state(0, theta0, np.zeros_like(theta0), vals, vecs, theta_b=thetaB)
    g.upload_velocity(U, Ub, phi)
    oc.store_old_time(); oc.step(dt); g.store_old_time(); st = g.correct(dt, want_stats=True)
    d = np.abs(g.theta() - oc.get(0, 0, abi.FIELD_THETA)).max(1)
    bad = d > 1e-9
    print(lim, "relL2", rel_l2(g.theta(), oc.get(0,0,abi.FIELD_THETA)), "iters gpu", g.last_iterations(), "oracle", oc.last_iterations() if hasattr(oc,'last_iterations') else None,
          "bad cells", int(bad.sum()), "of which boundary", int((bad & bcell).sum()), "faces/cell of bad", np.bincount(fpc[bad]).tolist(), "max", d.max())
    if bad.any():
        w = np.argsort(-d)[:5]
        print("  worst", w.tolist(), d[w].tolist(), "fpc", fpc[w].tolist(), "boundary", bcell[w].tolist(), "in multi pair", [any(c in k for k in multi) for c in w.tolist()])
        print("  conv", list(st[0].converged), list(st[0].n_iterations), [f"{x:.1e}" for x in st[0].final_residual])
