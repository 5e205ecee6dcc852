"""The CUDA path, through the C-ABI, against the REFERENCE'S OWN NUMBERS (tests/golden/reference_*.npz: outputs of the
reference text compiled into oracle/_ref, see tests/test_reference_pin.py), incl. the SaramitoLog functor, every limiter row
and the polyhedral fixture; and the CrankNicolson ddt plumbing against the oracle.  First run on B200 in round 1's driver
pass (32 XPASS) — since round 2 these are ordinary tests: a failure is a failure."""
from pathlib import Path

import numpy as np
import pytest

from helpers import Setup, rel_l2, tight
from oracle import mesh_ref
from oracle import oracle as orc
from reference_cases import N_STEPS, REFERENCE_CASES, STORED_STEPS, make_setup
from rheotool_b200 import abi, cases
from test_unstructured import REF_GOLD, _case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]   # pytest-timeout: a kernel that never returns must not hold the box

GOLD = Path(__file__).resolve().parent / "golden"
TOL_GPU_1 = 1e-10
TOL_GPU_N = 1e-8     # N_STEPS chained steps (BASELINE: 1e-6 after 100)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD / "reference_correct.npz")


@pytest.fixture(scope="module")
def cell():
    return np.load(GOLD / "reference_cell.npz")


# ---- the CUDA path against the reference's numbers --------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(REFERENCE_CASES))
def test_gpu_correct_golden(gold, name):
    spec, s = make_setup(name)
    g = s.gpu(spec.schemes)
    for k in range(N_STEPS):
        g.store_old_time(); g.correct(s.dt)
        if k + 1 not in STORED_STEPS:
            continue
        tol = TOL_GPU_1 if k == 0 else TOL_GPU_N
        for fld, key in ((abi.FIELD_THETA, "theta"), (abi.FIELD_TAU, "tau"), (abi.FIELD_THETA_B, "theta_b"), (abi.FIELD_TAU_B, "tau_b")):
            err = rel_l2(g.download(fld, 0), gold[f"{name}/step{k + 1}/{key}"])
            assert err <= tol, f"{name} step {k + 1} {key}: {err:.2e}"


def test_gpu_eig_golden(cell):
    """k_eig_tau against utils/jacobi.H: same eigenvalues (the device sorts ascending like Eigen; jacobi.H does not),
    same conformation tensor R exp(D) R^T."""
    from rheotool_b200.stress import eig_exp
    gv, gV = eig_exp(cell["theta"])
    d = np.sort(cell["expD"], axis=1)
    assert np.abs(np.stack([gv[:, 0], gv[:, 4], gv[:, 8]], 1) - d).max() <= 1e-12 * np.abs(d).max()
    R = gV.reshape(-1, 3, 3)
    A = R @ gv.reshape(-1, 3, 3) @ np.transpose(R, (0, 2, 1))
    Vr = cell["V"]
    Ar = Vr @ (cell["expD"][:, :, None] * np.transpose(Vr, (0, 2, 1)))
    assert rel_l2(A, Ar) <= 1e-13


@pytest.mark.parametrize("limiter", ["cubista", "upwind"])
def test_gpu_matches_the_reference_text_on_the_unstructured_mesh(limiter):
    from rheotool_b200.stress import GpuStressModel
    gold = np.load(REF_GOLD)
    m, models, U, Ub, phi, theta0, thetaB, dt = _case()
    sc = tight(cases.scheme_ctl(limiter, "PBiCGStab", 1e-10))
    vals, vecs = orc.calc_eig(theta0)
    g = GpuStressModel(m, models, sc)
    g.upload_state(0, theta0, np.zeros_like(theta0), vals, vecs, theta_b=thetaB)
    g.upload_velocity(U, Ub, phi)
    g.store_old_time(); g.correct(dt)
    assert rel_l2(g.theta(), gold[f"{limiter}/step1/theta"]) <= 1e-10
    assert rel_l2(g.tau(0), gold[f"{limiter}/step1/tau"]) <= 1e-10
    assert rel_l2(g.download(abi.FIELD_TAU_B, 0), gold[f"{limiter}/step1/tau_b"]) <= 1e-10


def test_gpu_tau_assignment_option_reproduces_the_reference_boundary_stress():
    """rheo_gpu_set_tau_assignment(on): with the alternative reading of `tau_ = ...` the wall faces next to the later-listed
    zeroGradient patch agree with the reference harness too (tests/test_unstructured.py has the CPU side of this)."""
    from rheotool_b200.stress import GpuStressModel
    gold = np.load(REF_GOLD)
    m, models, U, Ub, phi, theta0, thetaB, dt = _case()
    sc = tight(cases.scheme_ctl("cubista", "PBiCGStab", 1e-10))
    vals, vecs = orc.calc_eig(theta0)
    g = GpuStressModel(m, models, sc)
    g.set_tau_assignment(True)
    g.upload_state(0, theta0, np.zeros_like(theta0), vals, vecs, theta_b=thetaB)
    g.upload_velocity(U, Ub, phi)
    for _ in range(2):
        g.store_old_time(); g.correct(dt)
    assert rel_l2(g.download(abi.FIELD_TAU_B, 0), gold["cubista/step2/tau_b"]) <= 1e-9
    assert rel_l2(g.theta(), gold["cubista/step2/theta"]) <= 1e-9


# ---- CrankNicolson ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("psi", [1.0, 0.9])
def test_crank_nicolson_gpu_matches_oracle_over_varying_steps(psi):
    spec = cases.by_name("C3", 3 / 19)
    s = Setup(spec)
    sc = tight(spec.schemes)
    sc.ddt = abi.DDT_CRANK_NICOLSON
    sc.cn_psi = psi
    oc, g = s.oracle(sc), s.gpu(sc)
    for n, f in enumerate([1.0, 1.0, 0.5, 1.5, 0.8]):
        oc.store_old_time(); oc.step(f * s.dt)
        g.store_old_time(); g.correct(f * s.dt)
        if n in (1, 2):   # inner iteration of the same time level: ddt0 must not be evaluated twice
            oc.step(f * s.dt); g.correct(f * s.dt)
        assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-10 * (n + 1), n
    assert rel_l2(g.tau(0), oc.get(0, 0, abi.FIELD_TAU)) <= 1e-9
