"""ddtSchemes of the theta equation: Euler (EXT-OF9 EulerDdtScheme) and backward (EXT-OF9
backwardDdtScheme: Euler while the field has fewer than two old times, then the variable-step three-level formula).

CPU part: the oracle against (1) a hand evaluation of the three-level coefficients for unequal time steps and (2) the
exact solution of a homogeneous relaxation, where backward + inner iterations must converge with second order and Euler
with first order.  GPU part: the device against the oracle over steps of varying size."""
import numpy as np
import pytest

from helpers import Setup, rel_l2, tight
from oracle import oracle as orc
from rheotool_b200 import abi, cases, mesh
from rheotool_b200.mesh import GridSpec, PatchSpec


def _box():
    xs = np.linspace(0.0, 1.0, 3)
    patches = [PatchSpec("walls", abi.PATCH_WALL, abi.BC_ZERO_GRADIENT, abi.BC_ZERO_GRADIENT)]
    return mesh.tensor_grid(GridSpec(xs, xs.copy(), xs.copy(), [(0, 2, 0, 2, 0, 2)], patches, [], 0, False))


def _relaxation_case(ddt, th0):
    m = _box()
    model = cases.model_desc("Oldroyd-BLog", etaS=0.1, etaP=0.9, lambda_=0.5)
    oc = orc.OracleCase([m.desc], [model], cases.scheme_ctl("none", "PBiCGStab", 1e-15, ddt=ddt))
    th = np.tile(th0, (m.n_cells, 1))
    vals, vecs = orc.calc_eig(th)
    oc.set_state(0, 0, th, np.zeros_like(th), vals, vecs)
    oc.set_velocity(0, np.zeros((m.n_cells, 3)), np.zeros((m.n_boundary, 3)), np.zeros(m.n_faces))
    return oc, model, m


def test_backward_coefficients_for_unequal_steps_match_a_hand_evaluation():
    th0 = np.array([0.3, 0.1, -0.05, -0.2, 0.07, 0.15])
    oc, model, m = _relaxation_case("backward", th0)
    dts = [0.01, 0.02, 0.005]
    hist = [np.array(oc.get(0, 0, abi.FIELD_THETA)[0])]
    for n, dt in enumerate(dts):
        oc.store_old_time()
        # right-hand side the step will use: eigen-pairs of the current theta, L = 0
        vals = oc.get(0, 0, abi.FIELD_EIGVALS)[:1]; vecs = oc.get(0, 0, abi.FIELD_EIGVECS)[:1]
        rhs, _ = orc.model_rhs(model, np.zeros((1, 9)), hist[-1][None, :], vecs, vals)
        oc.step(dt)
        new = np.array(oc.get(0, 0, abi.FIELD_THETA)[0])
        if n == 0:     # fewer than two old times: Euler
            expect = hist[-1] + dt * rhs[0]
        else:          # EXT-OF9 backwardDdtScheme::fvmDdt
            dt0 = dts[n - 1]
            coefft = 1 + dt / (dt + dt0); c00 = dt * dt / (dt0 * (dt + dt0)); c0 = coefft + c00
            expect = ((c0 * hist[-1] - c00 * hist[-2]) / dt + rhs[0]) / (coefft / dt)
        assert np.abs(new - expect).max() < 1e-13, (n, new - expect)
        hist.append(new)


@pytest.mark.parametrize("ddt,order", [("Euler", 1), ("backward", 2)])
def test_temporal_order_on_homogeneous_relaxation(ddt, order):
    """U = 0: dA/dt = -(A - I)/lambda, A = exp(theta) = I + (A0 - I) exp(-t/lambda) exactly.  The model term is explicit in the
    eigen-pairs of the previous iterate, so the time level is converged with inner iterations (no store_old_time in between),
    as rheoFoam does with nInIter > 1."""
    lam, T = 0.5, 0.4
    d0 = np.array([0.6, -0.3, 0.1])          # theta0 diagonal (A0 = exp(theta0))
    th0 = np.array([d0[0], 0, 0, d0[1], 0, d0[2]])
    exact = np.log(1 + (np.exp(d0) - 1) * np.exp(-T / lam))
    errs = []
    for nsteps in (20, 40):
        oc, _, _ = _relaxation_case(ddt, th0)
        dt = T / nsteps
        for _ in range(nsteps):
            oc.store_old_time()
            for _ in range(25):
                oc.step(dt)
        th = oc.get(0, 0, abi.FIELD_THETA)[0]
        errs.append(np.abs(th[[0, 3, 5]] - exact).max())
    rate = np.log2(errs[0] / errs[1])
    assert abs(rate - order) < 0.25, (errs, rate)


@pytest.mark.gpu
def test_backward_scheme_gpu_matches_oracle_over_varying_steps():
    spec = cases.by_name("C3", 3 / 19)
    s = Setup(spec)
    sc = tight(spec.schemes)
    sc.ddt = abi.DDT_BACKWARD
    oc, g = s.oracle(sc), s.gpu(sc)
    for n, f in enumerate([1.0, 1.0, 0.5, 1.5, 0.8]):
        oc.store_old_time(); oc.step(f * s.dt)
        g.store_old_time(); g.correct(f * s.dt)
        if n in (1, 2):   # inner iteration of the same time level
            oc.step(f * s.dt); g.correct(f * s.dt)
        assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-10 * (n + 1), n
    assert rel_l2(g.tau(0), oc.get(0, 0, abi.FIELD_TAU)) <= 1e-9
    # and it is not Euler in disguise
    oe = s.oracle(tight(spec.schemes))
    for f in [1.0, 1.0, 0.5, 1.5, 0.8]:
        oe.store_old_time(); oe.step(f * s.dt)
    assert rel_l2(oe.get(0, 0, abi.FIELD_THETA), oc.get(0, 0, abi.FIELD_THETA)) > 1e-6


# ---- CrankNicolson (EXT-OF9 CrankNicolsonDdtScheme; tutorial Cavity/Oldroyd-BLog/system/fvSchemes: `CrankNicolson 1`) ----
def test_crank_nicolson_steps_match_a_hand_evaluation():
    """Fresh start: the first step is Euler (the ddt0 field is created during it), the second blends with the Euler rate
    of the first, from the third on ddt0_n = (1+psi)/dt0 (theta_n - theta_{n-1}) - psi ddt0_{n-1}."""
    th0 = np.array([0.3, 0.1, -0.05, -0.2, 0.07, 0.15])
    for psi in (1.0, 0.6):
        m = _box()
        model = cases.model_desc("Oldroyd-BLog", etaS=0.1, etaP=0.9, lambda_=0.5)
        oc = orc.OracleCase([m.desc], [model], cases.scheme_ctl("none", "PBiCGStab", 1e-15, ddt="CrankNicolson", cn_psi=psi))
        th = np.tile(th0, (m.n_cells, 1))
        vals, vecs = orc.calc_eig(th)
        oc.set_state(0, 0, th, np.zeros_like(th), vals, vecs)
        oc.set_velocity(0, np.zeros((m.n_cells, 3)), np.zeros((m.n_boundary, 3)), np.zeros(m.n_faces))
        dts = [0.01, 0.02, 0.005, 0.01]
        hist = [th0.copy()]
        ddt0 = np.zeros(6)
        for n, dt in enumerate(dts):
            oc.store_old_time()
            vals = oc.get(0, 0, abi.FIELD_EIGVALS)[:1]; vecs = oc.get(0, 0, abi.FIELD_EIGVECS)[:1]
            rhs, _ = orc.model_rhs(model, np.zeros((1, 9)), hist[-1][None, :], vecs, vals)
            oc.step(dt)
            new = np.array(oc.get(0, 0, abi.FIELD_THETA)[0])
            k = n + 1
            if k > 1:
                ddt0 = ((1 + psi) if k > 2 else 1.0) / dts[n - 1] * (hist[-1] - hist[-2]) - psi * ddt0
            coef = (1 + psi) if k > 1 else 1.0
            expect = hist[-1] + (dt / coef) * (rhs[0] + psi * ddt0)
            assert np.abs(new - expect).max() < 1e-13, (psi, n, new - expect)
            hist.append(new)


def test_crank_nicolson_is_second_order_on_homogeneous_relaxation():
    lam, T = 0.5, 0.4
    d0 = np.array([0.6, -0.3, 0.1])
    th0 = np.array([d0[0], 0, 0, d0[1], 0, d0[2]])
    exact = np.log(1 + (np.exp(d0) - 1) * np.exp(-T / lam))
    errs = []
    for nsteps in (20, 40):
        oc, _, _ = _relaxation_case("CrankNicolson", th0)
        dt = T / nsteps
        for _ in range(nsteps):
            oc.store_old_time()
            for _ in range(25):
                oc.step(dt)
        th = oc.get(0, 0, abi.FIELD_THETA)[0]
        errs.append(np.abs(th[[0, 3, 5]] - exact).max())
    rate = np.log2(errs[0] / errs[1])
    assert abs(rate - 2) < 0.3, (errs, rate)
