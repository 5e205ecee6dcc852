"""SURVEY.md §8 row f1 (first half): the explicit part of constitutiveEq::divTau(U) — constitutiveEq.C:72-132, summed over the
modes as multiMode.C:143-157 does — evaluated next to the stress so that tau need not leave the device for the momentum
predictor.  `Gauss linear` for div(tau) and div(grad(U)) (every tutorial's fvSchemes).  CPU tests hold the oracle's
restatement (oracle.cpp: div_tau_explicit) to analytic identities and to partition invariance; GPU tests hold
rheo_gpu_div_tau (csrc/gpu/momentum.cuh) to the oracle."""
import numpy as np
import pytest

from helpers import Setup, rel_l2, tight
from oracle import oracle as orc
from rheotool_b200 import abi, cases


def _sym6(T):
    """(n, 3, 3) symmetric -> (n, 6) xx xy xz yy yz zz"""
    return np.stack([T[:, 0, 0], T[:, 0, 1], T[:, 0, 2], T[:, 1, 1], T[:, 1, 2], T[:, 2, 2]], axis=1)


def _face_centres(s):
    m = s.mesh
    return m.Cf[m.n_internal:]


def _oracle_with(s, spec, tau, tau_b, U=None, Ub=None):
    oc = orc.OracleCase([s.mesh.desc], spec.models, tight(spec.schemes))
    for mi in range(len(spec.models)):
        oc.set_state(0, mi, s.theta0, tau, s.eigvals, s.eigvecs, tau_b=tau_b)
    oc.set_velocity(0, s.U if U is None else U, s.Ub if Ub is None else Ub, s.phi)
    return oc


def _interior(s, layers=2):
    """cells at least `layers` cells away from every boundary of the (uniform) cube"""
    C = s.mesh.C
    h = np.cbrt(s.mesh.V.mean())
    lo, hi = C.min(axis=0), C.max(axis=0)
    return np.all((C > lo + (layers - 0.5) * h) & (C < hi - (layers - 0.5) * h), axis=1)


def test_divergence_of_a_uniform_stress_vanishes_and_of_a_linear_one_is_exact():
    spec = cases.by_name("C5", 12 / 400)   # uniform cube
    s = Setup(spec)
    n, nb = s.mesh.n_cells, s.mesh.n_boundary
    rho = spec.models[0].rho
    T0 = np.array([[1.0, 0.2, -0.3], [0.2, 0.5, 0.1], [-0.3, 0.1, -0.7]])
    oc = _oracle_with(s, spec, np.tile(_sym6(T0[None]), (n, 1)), np.tile(_sym6(T0[None]), (nb, 1)))
    d = oc.div_tau(0, abi.STAB_NONE)
    assert np.abs(d).max() <= 1e-11 * np.abs(T0).max() / np.cbrt(s.mesh.V.min())
    # tau_ij = T0_ij + B_ijk x_k (B symmetric in ij): (div tau)_j = d_i tau_ij = B_iji
    rng = np.random.default_rng(3)
    B = rng.standard_normal((3, 3, 3))
    B = 0.5 * (B + B.transpose(1, 0, 2))
    def field(X):
        return _sym6(T0[None] + np.einsum("ijk,nk->nij", B, X))
    oc = _oracle_with(s, spec, field(s.mesh.C), field(_face_centres(s)))
    d = oc.div_tau(0, abi.STAB_NONE)
    exact = np.einsum("iji->j", B) / rho
    assert np.abs(d - exact).max() <= 1e-9 * np.abs(exact).max()


def test_coupling_term_vanishes_for_a_linear_velocity_and_is_the_laplacian_of_a_quadratic_one():
    spec = cases.by_name("C5", 12 / 400)
    s = Setup(spec)
    m0 = spec.models[0]
    n, nb = s.mesh.n_cells, s.mesh.n_boundary
    tau = np.zeros((n, 6)); tau_b = np.zeros((nb, 6))
    A = np.array([[0.3, -0.2, 0.1], [0.5, 0.1, -0.4], [0.2, 0.6, -0.4]])
    lin = lambda X: X @ A.T
    oc = _oracle_with(s, spec, tau, tau_b, lin(s.mesh.C), lin(_face_centres(s)))
    d = oc.div_tau(0, abi.STAB_COUPLING)
    assert np.abs(d).max() <= 1e-9 * np.abs(A).max() * m0.etaP / m0.rho / np.cbrt(s.mesh.V.min())
    # U_j = c_j |x|^2: laplacian = 6 c_j; Gauss linear is exact for it away from the boundary (central differences)
    c = np.array([0.7, -0.3, 0.2])
    quad = lambda X: np.outer((X ** 2).sum(axis=1), c)
    oc = _oracle_with(s, spec, tau, tau_b, quad(s.mesh.C), quad(_face_centres(s)))
    d = oc.div_tau(0, abi.STAB_COUPLING)
    inner = _interior(s)
    assert inner.sum() > 100
    exact = -(m0.etaP / m0.rho) * 6.0 * c
    assert np.abs(d[inner] - exact).max() <= 1e-8 * np.abs(exact).max()
    assert np.abs(oc.div_tau(0, abi.STAB_BSD)).max() == 0.0   # BSD: only div(tau) is evaluated here (tau = 0)


@pytest.mark.parametrize("name,scale,n", [("C3", 3 / 19, (2, 2, 1)), ("C4", 10 / 252, (2, 1, 2))])
def test_partition_invariance_of_div_tau(name, scale, n):
    """processor faces: interpolation with the neighbour cell's stress / velocity gradient (emulated ranks vs one rank)"""
    spec = cases.by_name(name, scale)
    s = Setup(spec)
    sc = tight(spec.schemes)
    one = s.oracle(sc)
    nr = n[0] * n[1] * n[2]
    c2r = s.mesh.simple_decomp(*n)
    subs = [s.mesh.decompose(c2r, nr, r) for r in range(nr)]
    many = orc.OracleCase([x.desc for x in subs], spec.models, sc)
    addr = []
    for r, sub in enumerate(subs):
        ca, fa = sub.proc_addressing()
        addr.append(ca)
        for mi in range(len(spec.models)):
            many.set_state(r, mi, s.theta_mode(mi)[ca], s.tau0[ca], s.eigvals_mode(mi)[ca], s.eigvecs_mode(mi)[ca])
        gf = np.abs(fa) - 1
        ph = np.where(fa > 0, s.phi[gf], -s.phi[gf])
        gb = gf[sub.n_internal:] - s.mesh.n_internal
        Ub = np.zeros((sub.n_boundary, 3))
        Ub[gb >= 0] = s.Ub[gb[gb >= 0]]
        many.set_velocity(r, s.U[ca], Ub, ph)
    one.store_old_time(); one.step(s.dt)
    many.store_old_time(); many.step(s.dt)
    for stab in (abi.STAB_NONE, abi.STAB_COUPLING):
        ref = one.div_tau(0, stab)
        got = np.empty_like(ref)
        for r in range(nr):
            got[addr[r]] = many.div_tau(r, stab)
        assert np.abs(ref).max() > 0
        assert rel_l2(got, ref) < 1e-11, (name, stab)


@pytest.mark.gpu
@pytest.mark.parametrize("name,scale", [("C2", 1 / 9), ("C3", 4 / 19), ("C4", 12 / 252), ("C5", 20 / 400)])
def test_gpu_div_tau_matches_oracle(name, scale):
    """after two steps: every stabilization option, all modes summed (C4: 4 modes), 2-D (C2) and 3-D"""
    spec = cases.by_name(name, scale)
    s = Setup(spec)
    sc = tight(spec.schemes)
    oc, g = s.oracle(sc), s.gpu(sc)
    for _ in range(2):
        oc.store_old_time(); oc.step(s.dt)
        g.store_old_time(); g.correct(s.dt)
    for stab in (abi.STAB_NONE, abi.STAB_BSD, abi.STAB_COUPLING):
        ref = oc.div_tau(0, stab)
        assert np.abs(ref).max() > 0
        assert rel_l2(g.div_tau(stab), ref) <= 1e-10, (name, stab)
    if name == "C3":   # the coupling term must be visible next to div(tau) (C4's synthetic start has |tau| ~ 1e13: it is not)
        assert rel_l2(oc.div_tau(0, abi.STAB_COUPLING), oc.div_tau(0, abi.STAB_NONE)) > 1e-3


@pytest.mark.gpu
def test_gpu_div_tau_refuses_what_it_does_not_evaluate():
    spec = cases.by_name("C3", 3 / 19)
    s = Setup(spec)
    g = s.gpu(tight(spec.schemes))
    g.store_old_time(); g.correct(s.dt)
    with pytest.raises(RuntimeError):
        g.div_tau(7)
    n = s.mesh.n_cells
    g.upload_thermo(0, np.full(n, 0.1), np.full(n, 0.5))
    with pytest.raises(RuntimeError):
        g.div_tau(abi.STAB_COUPLING)
    g.div_tau(abi.STAB_NONE)
