"""correct(alpha, gradU) with a caller-supplied velocity gradient (constitutiveEq.H:346-350; utils/boilerLog.H:1
`L(gradU == nullptr ? fvc::grad(U)() : *gradU)`; called that way by filmModel.C:408).  `alpha` is part of the signature only:
no *Log model reads it inside correct()."""
import numpy as np
import pytest

from helpers import Setup, rel_l2, tight
from rheotool_b200 import abi, cases


def _own_gradient(s, sc):
    """fvc::grad(U) as the oracle evaluates it, in OpenFOAM's tensor order [3i+j] = d_i U_j"""
    oc = s.oracle(sc)
    import ctypes as C
    from oracle import oracle as orc
    n = s.mesh.n_cells
    out = np.zeros((n, 9))
    fb = s.Ub.copy()
    orc.lib().orc_gauss_grad(oc._h, 0, 3, orc._p(s.U), orc._p(fb), orc._p(out))   # [3k+d] = d_d U_k
    return out.reshape(n, 3, 3).transpose(0, 2, 1).reshape(n, 9).copy()


def test_supplying_the_models_own_gradient_changes_nothing():
    spec = cases.by_name("C3", 3 / 19)
    s = Setup(spec)
    sc = tight(spec.schemes)
    a, b = s.oracle(sc), s.oracle(sc)
    b.set_grad_u(0, _own_gradient(s, sc))
    for oc in (a, b):
        oc.store_old_time(); oc.step(s.dt)
    assert rel_l2(b.get(0, 0, abi.FIELD_THETA), a.get(0, 0, abi.FIELD_THETA)) <= 1e-14
    g = _own_gradient(s, sc) * 1.3
    b.set_grad_u(0, g)
    b.store_old_time(); b.step(s.dt)
    a.store_old_time(); a.step(s.dt)
    assert rel_l2(b.get(0, 0, abi.FIELD_THETA), a.get(0, 0, abi.FIELD_THETA)) > 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name,scale", [("C3", 4 / 19), ("C2", 1 / 9)])
def test_gpu_caller_supplied_gradient_matches_oracle(name, scale):
    spec = cases.by_name(name, scale)
    s = Setup(spec)
    sc = tight(spec.schemes)
    rng = np.random.default_rng(5)
    gU = _own_gradient(s, sc) * (1.0 + 0.2 * rng.standard_normal((s.mesh.n_cells, 9)))   # not the gradient of U any more
    oc, g = s.oracle(sc), s.gpu(sc)
    oc.set_grad_u(0, gU); g.upload_grad_u(gU)
    for _ in range(2):
        oc.store_old_time(); oc.step(s.dt)
        g.store_old_time(); g.correct(s.dt)
    assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-10
    assert rel_l2(g.tau(0), oc.get(0, 0, abi.FIELD_TAU)) <= 1e-10
    with pytest.raises(RuntimeError):
        g.div_tau(abi.STAB_COUPLING)          # divTau's own fvc::grad(U) is not what the device holds now
    g.div_tau(abi.STAB_NONE)
    # back to the model's own gradient
    oc.set_grad_u(0, None); g.upload_grad_u(None)
    oc.store_old_time(); oc.step(s.dt)
    g.store_old_time(); g.correct(s.dt)
    assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-10
    assert rel_l2(g.div_tau(abi.STAB_COUPLING), oc.div_tau(0, abi.STAB_COUPLING)) <= 1e-10
