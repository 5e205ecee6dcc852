"""Pin the oracle against analytic material functions (the rheoTestFoam known-answer design,
of90/src/solvers/rheoTestFoam/rheoTestFoam.C:48-90 and createFields.H:61-146: homogeneous flow, phi = 0).

The reference ships no stored expected output for the log models, so the pins are:
  (1) Oldroyd-B steady simple shear:  tau_xy = etaP*gdot,  N1 = tau_xx - tau_yy = 2*etaP*lambda*gdot^2
  (2) for every model, the steady conformation tensor A = R Lambda R^T reached by the log-conformation
      oracle must satisfy the model's ORIGINAL (non-log) constitutive equation
          kappa.A + A.kappa^T - H(A)/lambda = 0 ,   kappa = grad(U)^T,
      with H from the model definitions (user guide / Oldroyd_B.C, Giesekus.C, PTT.C, FENE_P.C).
"""
import numpy as np
import pytest

from oracle import oracle as orc
from rheotool_b200 import abi, cases, mesh
from rheotool_b200.mesh import GridSpec, PatchSpec


def homogeneous_case(model, kappa, n_steps=4000, dt_over_lambda=0.02):
    """2x2x2 uniform cube, all boundaries zeroGradient for theta, U = kappa.x (so fvc::grad(U) = kappa^T exactly)."""
    xs = np.linspace(0.0, 1.0, 3)
    patches = [PatchSpec("walls", abi.PATCH_WALL, abi.BC_ZERO_GRADIENT, abi.BC_ZERO_GRADIENT)]
    m = mesh.tensor_grid(GridSpec(xs, xs.copy(), xs.copy(), [(0, 2, 0, 2, 0, 2)], patches, [], 0, False))
    U = m.C @ kappa.T
    Ub = m.Cf[m.n_internal:] @ kappa.T
    phi = np.zeros(m.n_faces)   # createFields.H:146  phi *= 0
    oc = orc.OracleCase([m.desc], [model], cases.scheme_ctl("cubista", "PBiCGStab", 1e-14))
    rng = np.random.default_rng(7)
    th0 = np.tile(1e-3 * rng.standard_normal(6), (m.n_cells, 1))   # tiny anisotropic seed (theta = 0 is a fixed point of pure shear)
    vals, vecs = orc.calc_eig(th0)
    oc.set_state(0, 0, th0, np.zeros_like(th0), vals, vecs)
    oc.set_velocity(0, U, Ub, phi)
    dt = dt_over_lambda * model.lambda_
    for _ in range(n_steps):
        oc.store_old_time()
        oc.step(dt)
    return oc, m


def conformation(oc):
    vals = oc.get(0, 0, abi.FIELD_EIGVALS)[0].reshape(3, 3)
    R = oc.get(0, 0, abi.FIELD_EIGVECS)[0].reshape(3, 3)
    return R @ vals @ R.T


def sym6(t):
    return np.array([[t[0], t[1], t[2]], [t[1], t[3], t[4]], [t[2], t[4], t[5]]])


def test_oldroyd_b_simple_shear_material_functions():
    etaP, lam, gd = 0.41, 0.7, 1.3
    model = cases.model_desc("Oldroyd-BLog", etaS=0.59, etaP=etaP, lambda_=lam)
    kappa = np.zeros((3, 3)); kappa[0, 1] = gd
    oc, _ = homogeneous_case(model, kappa)
    tau = sym6(oc.get(0, 0, abi.FIELD_TAU)[0])
    assert tau[0, 1] == pytest.approx(etaP * gd, rel=1e-9)
    assert tau[0, 0] - tau[1, 1] == pytest.approx(2 * etaP * lam * gd ** 2, rel=1e-9)
    assert abs(tau[1, 1]) < 1e-9 and abs(tau[2, 2]) < 1e-9
    # all 8 cells identical (homogeneous)
    t_all = oc.get(0, 0, abi.FIELD_TAU)
    assert np.abs(t_all - t_all[0]).max() < 1e-12


def H_of(model, A):
    I = np.eye(3)
    if model.model == abi.MODEL_OLDROYD_B_LOG:
        return A - I
    if model.model == abi.MODEL_GIESEKUS_LOG:
        return (A - I) + model.alpha * (A - I) @ (A - I)
    if model.model == abi.MODEL_PTT_LOG:
        z = model.epsilon / (1 - model.zeta) * (np.trace(A) - 3)
        Y = 1 + z if model.ptt_function == abi.PTT_LINEAR else np.exp(z)
        return Y * (A - I)
    f = model.L2 / (model.L2 - np.trace(A))
    if model.model == abi.MODEL_FENE_CR_LOG:   # FENE_CR.C: f (A - I)
        return f * (A - I)
    a = model.L2 / (model.L2 - 3)
    return f * A - a * I


FLOWS = {
    "shear": np.array([[0, 1.1, 0], [0, 0, 0], [0, 0, 0.0]]),
    "planar_extension": np.diag([0.35, -0.35, 0.0]),
    "mixed3d": np.array([[0.2, 0.7, -0.1], [0.3, -0.05, 0.4], [0.1, -0.2, -0.15]]),
}
MODELS = {
    "Oldroyd-BLog": dict(etaS=0.1, etaP=0.9, lambda_=0.6),
    "GiesekusLog": dict(etaS=0.01, etaP=0.99, lambda_=0.5, alpha=0.2),
    "PTTLog-linear": dict(etaS=0.11, etaP=0.89, lambda_=0.6, epsilon=0.25, zeta=0.0, ptt_function="linear"),
    "PTTLog-exponential": dict(etaS=0.11, etaP=0.89, lambda_=0.6, epsilon=0.1, zeta=0.1, ptt_function="exponential"),
    "FENE-PLog": dict(etaS=0.01, etaP=0.99, lambda_=0.4, L2=50.0),
    "FENE-CRLog": dict(etaS=0.01, etaP=0.99, lambda_=0.4, L2=50.0),
}


@pytest.mark.parametrize("flow", sorted(FLOWS))
@pytest.mark.parametrize("mname", sorted(MODELS))
def test_steady_state_satisfies_original_constitutive_equation(mname, flow):
    model = cases.model_desc(mname.split("-l")[0].split("-e")[0] if mname.startswith("PTT") else mname, **MODELS[mname])
    kappa = FLOWS[flow]
    oc, _ = homogeneous_case(model, kappa, n_steps=3000)
    A = conformation(oc)
    zeta = model.zeta if model.model == abi.MODEL_PTT_LOG else 0.0
    D = 0.5 * (kappa + kappa.T)
    k = kappa - zeta * D                      # Gordon-Schowalter derivative of PTT (boilerLog.H:26-29)
    res = k @ A + A @ k.T - H_of(model, A) / model.lambda_
    assert np.abs(res).max() < 1e-8 * max(1.0, np.abs(A).max()), res
    # and tau is the model's map of A  (Oldroyd_BLog.C:175, GiesekusLog.C:172, PTTLog.C:264, FENE_PLog.C:178)
    tau = sym6(oc.get(0, 0, abi.FIELD_TAU)[0])
    coef = model.etaP / model.lambda_
    if model.model == abi.MODEL_FENE_P_LOG:
        f = model.L2 / (model.L2 - np.trace(A)); a = model.L2 / (model.L2 - 3)
        expect = coef * (f * A - a * np.eye(3))
    elif model.model == abi.MODEL_FENE_CR_LOG:
        expect = coef * model.L2 / (model.L2 - np.trace(A)) * (A - np.eye(3))
    elif model.model == abi.MODEL_PTT_LOG:
        expect = coef / (1 - model.zeta) * (A - np.eye(3))
    else:
        expect = coef * (A - np.eye(3))
    assert np.abs(tau - expect).max() < 1e-8 * max(1.0, np.abs(expect).max())


def test_fene_cr_simple_shear_material_functions():
    """FENE-CR in steady simple shear (closed form): constant shear viscosity  tau_xy = etaP*gdot  and
    N1 = 2 etaP lambda gdot^2 / f  with f the positive root of (L2-3) f^2 - L2 f - 2 (lambda gdot)^2 = 0."""
    etaP, lam, gd, L2 = 0.8, 0.5, 1.7, 30.0
    model = cases.model_desc("FENE-CRLog", etaS=0.2, etaP=etaP, lambda_=lam, L2=L2)
    kappa = np.zeros((3, 3)); kappa[0, 1] = gd
    oc, _ = homogeneous_case(model, kappa)
    tau = sym6(oc.get(0, 0, abi.FIELD_TAU)[0])
    f = (L2 + np.sqrt(L2 ** 2 + 8 * (L2 - 3) * (lam * gd) ** 2)) / (2 * (L2 - 3))
    assert tau[0, 1] == pytest.approx(etaP * gd, rel=1e-8)
    assert tau[0, 0] - tau[1, 1] == pytest.approx(2 * etaP * lam * gd ** 2 / f, rel=1e-8)
    assert abs(tau[1, 1]) < 1e-8 and abs(tau[2, 2]) < 1e-8


def test_white_metzner_cy_simple_shear_material_functions():
    """White-Metzner with Carreau-Yasuda functions (Log version: m = n, L = K, b = a) in steady simple shear behaves like
    Oldroyd-B with the rate-dependent eta(gdot), lambda(gdot):  tau_xy = eta gdot,  N1 = 2 eta lambda gdot^2."""
    etaP, lam, gd, K, n, a = 0.9, 0.4, 2.3, 1.5, 0.6, 1.8
    model = cases.model_desc("WhiteMetznerCYLog", etaS=0.1, etaP=etaP, lambda_=lam, wm_K=K, wm_n=n, wm_a=a)
    kappa = np.zeros((3, 3)); kappa[0, 1] = gd
    oc, _ = homogeneous_case(model, kappa)
    tau = sym6(oc.get(0, 0, abi.FIELD_TAU)[0])
    cy = (1 + (K * gd) ** a) ** ((n - 1) / a)
    assert tau[0, 1] == pytest.approx(etaP * cy * gd, rel=1e-8)
    assert tau[0, 0] - tau[1, 1] == pytest.approx(2 * (etaP * cy) * (lam * cy) * gd ** 2, rel=1e-8)
    assert abs(tau[1, 1]) < 1e-8 and abs(tau[2, 2]) < 1e-8


def _chi_factor(trA, chi):
    c2 = chi * chi
    return ((3 - (trA / 3) / c2) * (1 - 1 / c2)) / ((1 - (trA / 3) / c2) * (3 - 1 / c2))


@pytest.mark.parametrize("flow", sorted(FLOWS))
@pytest.mark.parametrize("chi", [0.0, 4.0])
def test_rolie_poly_log_steady_state_satisfies_the_conformation_form(flow, chi):
    """Cross-file pin: the steady A reached by the LOG oracle (RoliePolyLog.C) must satisfy the NON-log conformation
    equation of RoliePoly.C:139-150:  A.L + L^T.A - (A - I)/lambdaD - M1 (A + beta (trA/3)^delta (A - I)) = 0."""
    lamD, lamR, beta, delta, etaP = 0.8, 0.2, 0.3, -0.5, 0.9
    model = cases.model_desc("Rolie-PolyLog", etaS=0.1, etaP=etaP, lambda_=lamD, rp_lambdaR=lamR, rp_beta=beta, rp_delta=delta, rp_chiMax=chi)
    kappa = FLOWS[flow]
    oc, _ = homogeneous_case(model, kappa, n_steps=3000)
    A = conformation(oc)
    trA = np.trace(A)
    M1 = 2 * (1 - np.sqrt(3 / trA)) / lamR
    if chi > 1:
        M1 *= _chi_factor(trA, chi)
    I = np.eye(3)
    res = kappa @ A + A @ kappa.T - (A - I) / lamD - M1 * (A + beta * (trA / 3) ** delta * (A - I))
    assert np.abs(res).max() < 1e-8 * max(1.0, np.abs(A).max()), res
    tau = sym6(oc.get(0, 0, abi.FIELD_TAU)[0])
    expect = etaP / lamD * (A - I) * (_chi_factor(trA, chi) if chi > 1 else 1.0)   # RoliePoly.C:155-166
    assert np.abs(tau - expect).max() < 1e-8 * max(1.0, np.abs(expect).max())


@pytest.mark.parametrize("flow", sorted(FLOWS))
@pytest.mark.parametrize("n", [0.0, 1.0])
def test_xpompom_log_steady_state_satisfies_the_stress_form(flow, n):
    """Cross-file pin: tau reached by the LOG oracle (XPomPomLog.C) must satisfy the steady STRESS equation of
    XPomPom.C:108-141:  tau.L + L^T.tau + G 2D - (f/lambdaB) tau - (alpha/etaP) tau.tau - (G/lambdaB)(f - 1) I = 0,
    G = etaP/lambdaB, with lambda and f evaluated from tau as that file does."""
    lamB, lamS, alpha, q, etaP = 0.6, 0.3, 0.15, 3.0, 0.85
    model = cases.model_desc("XPomPomLog", etaS=0.15, etaP=etaP, lambda_=lamB, alpha=alpha, xpp_lambdaS=lamS, xpp_q=q, xpp_n=n)
    kappa = FLOWS[flow]
    oc, _ = homogeneous_case(model, kappa, n_steps=3000)
    tau = sym6(oc.get(0, 0, abi.FIELD_TAU)[0])
    G = etaP / lamB
    lam = np.sqrt(1 + np.trace(tau) / (3 * G))
    stretch = (1 - 1 / lam) if n == 0 else (1 - 1 / lam ** (n + 1))
    f = 2 * (lamB / lamS) * np.exp((2 / q) * (lam - 1)) * stretch + (1 / lam ** 2) * (1 - (alpha / 3) * np.trace(tau @ tau) / G ** 2)
    L = kappa.T                                   # L_ij = d_i U_j
    res = tau @ L + L.T @ tau + G * (L + L.T) - (f / lamB) * tau - (alpha / etaP) * (tau @ tau) - (G / lamB) * (f - 1) * np.eye(3)
    assert np.abs(res).max() < 1e-8 * max(1.0, np.abs(tau).max() / lamB), res
    A = conformation(oc)
    assert np.abs(tau - G * (A - np.eye(3))).max() < 1e-8 * max(1.0, np.abs(tau).max())


def test_ptt_generalized_reduces_to_exponential_for_alpha_beta_one():
    """Mittag-Leffler E_{1,1}(z) = exp(z) (PTTLog.C:202-236 with alpha = beta = 1)."""
    rng = np.random.default_rng(3)
    n = 50
    th = 0.4 * rng.standard_normal((n, 6))
    vals, vecs = orc.calc_eig(th)
    L = 0.5 * rng.standard_normal((n, 9))
    m_exp = cases.model_desc("PTTLog", etaP=0.9, lambda_=0.6, epsilon=0.1, zeta=0.05, ptt_function="exponential")
    m_gen = cases.model_desc("PTTLog", etaP=0.9, lambda_=0.6, epsilon=0.1, zeta=0.05, ptt_function="generalized", ml_alpha=1.0, ml_beta=1.0)
    r1, _ = orc.model_rhs(m_exp, L, th, vecs, vals)
    r2, _ = orc.model_rhs(m_gen, L, th, vecs, vals)
    assert np.abs(r1 - r2).max() < 1e-10 * np.abs(r1).max()


@pytest.mark.parametrize("n,k,tau0", [(1.0, None, 0.3), (0.75, 1.5, 0.4)])
def test_saramito_log_steady_state_satisfies_the_stress_form(n, k, tau0):
    """Cross-file pin for SaramitoLog (otherModels/Saramito/SaramitoLog/SaramitoLog.C:143-245): above the yield stress the
    steady tau of the LOG oracle must satisfy the reference's NON-log stress equation (otherModels/Saramito/Saramito.C:
    174-192, zeta = 0, no PTT function):  tau.L + L^T.tau + (etaP/lambda) 2D - fac (etaP/lambda) tau = 0,
    fac = max(0, (|tau_d| - tau0)/(k |tau_d|^n))^(1/n), |tau_d| = mag(dev tau)/sqrt(2)  (Saramito.C:143,160-171)."""
    etaP, lam, gd = 1.0, 0.5, 2.0
    model = cases.model_desc("SaramitoLog", etaS=0.1, etaP=etaP, lambda_=lam, sar_tau0=tau0, sar_n=n, sar_k=k, sar_dims=(1, 1, 1))
    kappa = np.zeros((3, 3)); kappa[0, 1] = gd
    oc, _ = homogeneous_case(model, kappa, n_steps=6000)
    tau = sym6(oc.get(0, 0, abi.FIELD_TAU)[0])
    L = kappa.T                                   # L_ij = d_i U_j
    dev = tau - np.eye(3) * np.trace(tau) / 3.0
    tauDMag = np.sqrt((dev * dev).sum()) / np.sqrt(2.0)
    assert tauDMag > tau0                         # yielded: the relaxation term is active
    kk = model.sar_k
    fac = max(0.0, (tauDMag - tau0) / (kk * tauDMag ** n)) ** (1.0 / n)
    res = tau @ L + L.T @ tau + (etaP / lam) * (L + L.T) - fac * (etaP / lam) * tau
    assert np.abs(res).max() <= 1e-7 * np.abs(tau).max() * gd


def test_saramito_log_below_yield_is_elastic():
    """Below the yield stress fac = 0 (SaramitoLog.C:160-171): no relaxation, the conformation tensor follows the
    upper-convected kinematics exactly: simple shear from rest gives A_xy = gd t, A_xx = 1 + (gd t)^2 (first order in time
    for the Euler scheme: the bar is the time-step error)."""
    etaP, lam, gd = 1.0, 0.5, 0.2
    model = cases.model_desc("SaramitoLog", etaS=0.1, etaP=etaP, lambda_=lam, sar_tau0=50.0, sar_n=1.0, sar_dims=(1, 1, 1))
    kappa = np.zeros((3, 3)); kappa[0, 1] = gd
    steps, dtl = 400, 0.005
    oc, _ = homogeneous_case(model, kappa, n_steps=steps, dt_over_lambda=dtl)
    t = steps * dtl * lam
    A = conformation(oc)
    assert A[0, 1] == pytest.approx(gd * t, rel=2e-2)
    assert A[0, 0] - 1.0 == pytest.approx((gd * t) ** 2, rel=5e-2)
    assert A[1, 1] == pytest.approx(1.0, abs=1e-3)
