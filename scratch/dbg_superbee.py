import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from helpers import Setup, rel_l2, tight
from rheotool_b200 import abi, cases
spec = cases.by_name("C3", 2/19)
sc = tight(spec.schemes); sc.limiter = abi.LIMITER["superbee"]
s = Setup(spec)
oc, g = s.oracle(sc), s.gpu(sc)
oc.store_old_time(); oc.step(s.dt); g.store_old_time(); g.correct(s.dt)
d = np.abs(g.theta() - oc.get(0,0,0))
print("n cells", len(d), "cells with diff>1e-12:", (d.max(1)>1e-12).sum(), "max", d.max())
idx = np.argsort(-d.max(1))[:10]
print(idx, d[idx].max(1))
print("C of worst", s.mesh.C[idx])
# which components, and recompute phitc on host for the faces of the worst cell
c = idx[0]
print("diff comps", d[c])
m = s.mesh
nint = m.n_internal
faces = np.nonzero((m.owner[:nint]==c)|(m.neighbour==c))[0]
print("faces", faces, "phi", s.phi[faces])
