"""SURVEY.md §8f rank 3, the variants round 1 left out: the steadyState time scheme, `bounded GaussDefCmpw` convection
(EXT-OF9 boundedConvectionScheme; rheoFilmFoam/UCM/system/fvSchemes:35 uses both) and temperature-dependent lambda / etaP
(Oldroyd_BLog.C:133-135 `thermoLambdaPtr_->createField(lambda_)`, of90/src/libs/thermo/thermoFunctions/*).
CPU tests hold the oracle to identities; GPU tests hold the device to the oracle."""
import numpy as np
import pytest

from helpers import Setup, rel_l2, tight
from rheotool_b200 import abi, cases
from rheotool_b200.stress import thermo_factor


def _ctl(spec, **kw):
    sc = tight(spec.schemes)
    for k, v in kw.items():
        setattr(sc, k, v)
    return sc


def _slow_relaxation(name, scale):
    """steadyState has no V/dt on the diagonal: with the synthetic no-slip flows the near-wall cells (tiny fluxes) would take
    O(H / (lambda U)) steps in theta per iteration.  A long relaxation time keeps the pseudo-time iterations of these tests
    well conditioned (the tutorials rely on relaxation factors of 0.4 and on a converging outer loop instead)."""
    spec = cases.by_name(name, scale)
    spec.models = [cases.model_desc("Oldroyd-BLog", rho=1.0, etaS=0.01, etaP=0.99, lambda_=200.0)]
    return spec


def _through_flow(s, phi, u0=0.3):
    """steadyState leaves only the convective out-flux on the diagonal, which is zero in the dead-water corners of the
    synthetic flows (a singular row in OpenFOAM as well): superpose a uniform stream (discretely solenoidal: sum S_f = 0)."""
    return phi + u0 * s.mesh.Sf[:, 0]


def _nonsolenoidal(s, amp=0.05, seed=7):
    """the synthetic fluxes are discretely divergence-free; perturb them so that `bounded` matters"""
    rng = np.random.default_rng(seed)
    phi = s.phi.copy()
    phi[: s.mesh.n_internal] *= 1.0 + amp * rng.standard_normal(s.mesh.n_internal)
    return phi


# ---- thermo functions ------------------------------------------------------------------------------------------------
def test_thermo_functions_follow_the_reference_formulas():
    T = np.linspace(300.0, 420.0, 25)
    assert np.allclose(thermo_factor("Constant", [], T), 1.0)
    assert np.allclose(thermo_factor("Arrhenius", [1720.0, 373.15], T), np.exp(1720.0 * (1 / T - 1 / 373.15)), rtol=1e-15)       # Arrhenius.C:67
    assert np.allclose(thermo_factor("ArrheniusModified", [0.02, 373.15], T), np.exp(-0.02 * (T - 373.15)), rtol=1e-15)            # ArrheniusModified.C:67
    assert np.allclose(thermo_factor("WLF", [4.54, 150.4, 373.15], T), 10 ** (-4.54 * (T - 373.15) / (150.4 + (T - 373.15))), rtol=1e-14)   # WLF.C:67
    assert np.allclose(thermo_factor("VFT", [500.0, -2.0, 200.0], T), 10 ** (-2.0 + 500.0 / (T - 200.0)), rtol=1e-14)              # VFT.C:68
    assert thermo_factor("Arrhenius", [1720.0, 373.15], np.array([373.15]))[0] == 1.0


# ---- oracle identities (CPU) -----------------------------------------------------------------------------------------
def test_bounded_is_the_identity_for_solenoidal_fluxes_and_acts_otherwise():
    spec = cases.by_name("C3", 3 / 19)
    s = Setup(spec)
    a, b = s.oracle(_ctl(spec)), s.oracle(_ctl(spec, bounded=1))
    for oc in (a, b):
        oc.store_old_time(); oc.step(s.dt)
    assert rel_l2(b.get(0, 0, abi.FIELD_THETA), a.get(0, 0, abi.FIELD_THETA)) <= 1e-12   # div(phi) = 0 cell by cell
    phi = _nonsolenoidal(s)
    a, b = s.oracle(_ctl(spec)), s.oracle(_ctl(spec, bounded=1))
    for oc in (a, b):
        oc.set_velocity(0, s.U, s.Ub, phi)
        oc.store_old_time(); oc.step(s.dt)
    d = rel_l2(b.get(0, 0, abi.FIELD_THETA), a.get(0, 0, abi.FIELD_THETA))
    assert 1e-6 < d < 1e-1
    # the bounded form keeps a uniform field uniform whatever div(phi) is: row sum of (convection - Sp(div phi)) vanishes
    n = s.mesh.n_cells
    oc = s.oracle(_ctl(spec, bounded=1, limiter=abi.LIMITER["upwind"]))
    th = np.tile([0.3, 0.1, -0.2, 0.4, 0.05, -0.1], (n, 1))
    thb = np.tile([0.3, 0.1, -0.2, 0.4, 0.05, -0.1], (s.mesh.n_boundary, 1))
    oc.set_state(0, 0, th, s.tau0, s.eigvals, s.eigvecs, theta_b=thb)
    oc.set_velocity(0, np.zeros_like(s.U), np.zeros_like(s.Ub), phi)   # no deformation: the model source only relaxes theta
    oc.store_old_time(); oc.step(1e-9)                                  # ... and a tiny step leaves it where it was
    assert np.abs(oc.get(0, 0, abi.FIELD_THETA) - th).max() <= 1e-7


def test_steady_state_is_the_limit_of_euler_for_a_huge_time_step():
    spec = _slow_relaxation("C3", 3 / 19)
    s = Setup(spec)
    phi = _through_flow(s, s.phi)
    a = s.oracle(_ctl(spec, ddt=abi.DDT_STEADY_STATE, relax=0.4))
    b = s.oracle(_ctl(spec, relax=0.4))
    for oc in (a, b):
        oc.set_velocity(0, s.U, s.Ub, phi)
    a.store_old_time(); a.step(s.dt)
    b.store_old_time(); b.step(1e14)
    ta, tb = a.get(0, 0, abi.FIELD_THETA), b.get(0, 0, abi.FIELD_THETA)
    assert np.isfinite(ta).all() and np.abs(ta).max() < 50
    assert rel_l2(ta, tb) <= 1e-9
    assert rel_l2(ta, s.theta0) > 1e-3


def test_uniform_temperature_equals_scaled_scalar_parameters():
    spec = cases.by_name("C5", 12 / 400)
    s = Setup(spec)
    n = s.mesh.n_cells
    aT = float(thermo_factor("Arrhenius", [1720.0, 373.15], np.array([350.0]))[0])
    m0 = spec.models[0]
    a = s.oracle(tight(spec.schemes))
    a.set_thermo(0, 0, np.full(n, m0.lambda_ * aT), np.full(n, m0.etaP * aT))
    scaled = cases.model_desc("FENE-PLog", rho=m0.rho, etaS=m0.etaS, etaP=m0.etaP * aT, lambda_=m0.lambda_ * aT, L2=m0.L2)
    spec2 = cases.by_name("C5", 12 / 400)
    spec2.models = [scaled]
    b = Setup(spec2).oracle(tight(spec.schemes))
    for oc in (a, b):
        oc.store_old_time(); oc.step(s.dt)
    assert rel_l2(a.get(0, 0, abi.FIELD_THETA), b.get(0, 0, abi.FIELD_THETA)) <= 1e-14
    assert rel_l2(a.get(0, 0, abi.FIELD_TAU), b.get(0, 0, abi.FIELD_TAU)) <= 1e-14


# ---- the device against the oracle (GPU) -----------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,scale", [("C3", 3 / 19), ("C5", 16 / 400)])
def test_gpu_steady_state_bounded_matches_oracle(name, scale):
    """rheoFilmFoam/UCM's combination: ddtSchemes steadyState, div(phi,theta) bounded GaussDefCmpw cubista, theta relaxed —
    with fluxes that are NOT solenoidal, so that the Sp(div phi) term is exercised; three pseudo-time iterations."""
    spec = _slow_relaxation(name, scale)
    s = Setup(spec)
    phi = _through_flow(s, _nonsolenoidal(s))
    sc = _ctl(spec, ddt=abi.DDT_STEADY_STATE, bounded=1, relax=0.4)
    oc, g = s.oracle(sc), s.gpu(sc)
    oc.set_velocity(0, s.U, s.Ub, phi); g.upload_velocity(s.U, s.Ub, phi)
    for _ in range(3):
        oc.store_old_time(); oc.step(s.dt)
        g.store_old_time(); g.correct(s.dt)
    to = oc.get(0, 0, abi.FIELD_THETA)
    assert np.isfinite(to).all() and np.abs(to).max() < 50
    assert rel_l2(g.theta(), to) <= 1e-9
    assert rel_l2(g.tau(0), oc.get(0, 0, abi.FIELD_TAU)) <= 1e-9


@pytest.mark.gpu
def test_gpu_bounded_euler_matches_oracle():
    spec = cases.by_name("C5", 14 / 400)
    s = Setup(spec)
    phi = _nonsolenoidal(s)
    sc = _ctl(spec, bounded=1)
    oc, g = s.oracle(sc), s.gpu(sc)
    oc.set_velocity(0, s.U, s.Ub, phi); g.upload_velocity(s.U, s.Ub, phi)
    oc.store_old_time(); oc.step(s.dt)
    g.store_old_time(); g.correct(s.dt)
    assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-10
    g2 = s.gpu(_ctl(spec))
    g2.upload_velocity(s.U, s.Ub, phi); g2.store_old_time(); g2.correct(s.dt)
    assert rel_l2(g2.theta(), g.theta()) > 1e-7, "bounded must change the answer for non-solenoidal fluxes"


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["Oldroyd-BLog", "GiesekusLog", "PTTLog", "FENE-PLog"])
def test_gpu_thermo_dependent_parameters_match_oracle(model):
    """lambda(T) = lambda a_T(T), etaP(T) = etaP a_T(T) per cell with an Arrhenius / WLF shift over a temperature field that
    varies by 60 K across the domain; two steps; then back to the scalars."""
    spec = cases.by_name("C3", 3 / 19)
    kw = dict(rho=1.0, etaS=0.01, etaP=0.99, lambda_=0.1)
    if model == "GiesekusLog":
        kw["alpha"] = 0.2
    if model == "PTTLog":
        kw.update(epsilon=0.25, zeta=0.1)
    if model == "FENE-PLog":
        kw["L2"] = 100.0
    spec.models = [cases.model_desc(model, **kw)]
    s = Setup(spec)
    T = 340.0 + 60.0 * (s.mesh.C[:, 0] - s.mesh.C[:, 0].min()) / np.ptp(s.mesh.C[:, 0]) + 5.0 * np.sin(3 * s.mesh.C[:, 1])
    lam = kw["lambda_"] * thermo_factor("Arrhenius", [1720.0, 373.15], T)
    eta = kw["etaP"] * thermo_factor("WLF", [4.54, 150.4, 373.15], T)
    sc = tight(spec.schemes)
    oc, g = s.oracle(sc), s.gpu(sc)
    oc.set_thermo(0, 0, lam, eta); g.upload_thermo(0, lam, eta)
    for _ in range(2):
        oc.store_old_time(); oc.step(s.dt)
        g.store_old_time(); g.correct(s.dt)
    assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-10
    assert rel_l2(g.tau(0), oc.get(0, 0, abi.FIELD_TAU)) <= 1e-10
    plain = s.oracle(sc)
    for _ in range(2):
        plain.store_old_time(); plain.step(s.dt)
    assert rel_l2(oc.get(0, 0, abi.FIELD_TAU), plain.get(0, 0, abi.FIELD_TAU)) > 1e-3, "the temperature field must matter"
    oc.set_thermo(0, 0, None, None); g.upload_thermo(0, None, None)
    oc.store_old_time(); oc.step(s.dt)
    g.store_old_time(); g.correct(s.dt)
    assert rel_l2(g.tau(0), oc.get(0, 0, abi.FIELD_TAU)) <= 1e-10
