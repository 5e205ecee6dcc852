"""GPU tests of everything written after this round's GPU budget was spent — compiled for sm_100a, NOT YET RUN ON HARDWARE:

* the CUDA path against the reference's own numbers (tests/golden/reference_*.npz, see tests/test_reference_pin.py), incl.
  the SaramitoLog functor and the polyhedral fixture;
* PBiCG + DILU on the device (csrc/gpu/pbicg.cuh) — the solver every Log tutorial's fvSolution selects — against the
  oracle's PBiCG, against PBiCGStab on the device and against the reference fixture;
* the CrankNicolson ddt plumbing.

Every case is xfail(strict=False) until its first run: a pass shows as XPASS, a failure cannot hide the rest of the suite
behind `-x`.  The file name sorts last on purpose: should a new kernel fault, the CUDA context of the pytest process is
lost only after every test that HAS run on hardware before is through.  tools/gpu_unverified_first.sh runs this file."""
from pathlib import Path

import numpy as np
import pytest

from helpers import Setup, rel_l2, tight
from oracle import mesh_ref
from oracle import oracle as orc
from reference_cases import N_STEPS, REFERENCE_CASES, STORED_STEPS, make_setup
from rheotool_b200 import abi, cases
from test_unstructured import REF_GOLD, _case

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="written without GPU access: not yet run on hardware"),
              pytest.mark.timeout(900)]   # pytest-timeout: a kernel that never returns must not hold the box

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


# ---- PBiCG on the device ------------------------------------------------------------------------------------------
TOL_1 = 1e-10


@pytest.mark.parametrize("name,scale", [("C1", 0.25), ("C2", 1 / 9), ("C3", 3 / 19), ("C4", 10 / 252), ("C5", 14 / 400)])
def test_pbicg_one_step_parity(name, scale):
    """K = 4 / NR = 4 (2-D), K = 6 / NR = 6 (3-D), several modes batched (C4): theta, tau after one correct()."""
    spec = cases.by_name(name, scale)
    s = Setup(spec)
    sc = tight(spec.schemes, solver="PBiCG")
    oc, g = s.oracle(sc), s.gpu(sc)
    oc.store_old_time(); oc.step(s.dt)
    g.store_old_time(); g.correct(s.dt)
    for mi in range(len(spec.models)):
        assert rel_l2(g.download(abi.FIELD_THETA, mi), oc.get(0, mi, abi.FIELD_THETA)) <= TOL_1
        assert rel_l2(g.download(abi.FIELD_TAU, mi), oc.get(0, mi, abi.FIELD_TAU)) <= TOL_1


def test_pbicg_same_iterations_as_oracle_on_renumbered_mesh():
    """Tutorial tolerance 1e-10, CFL 2: the colour-parallel DILU / DILU^T sweeps are the sequential ones of the oracle on
    the renumbered mesh — same iteration counts and initial residuals per component."""
    spec = cases.by_name("C3", 3 / 19)
    s = Setup(spec, cfl=2.0)
    sc = tight(spec.schemes, tol=1e-10, solver="PBiCG")
    g = s.gpu(sc)
    perm, cstart = g.renumbering()
    rm = mesh_ref.renumbered_mesh(mesh_ref.from_host_mesh(s.mesh), perm)
    desc = mesh_ref.to_desc(rm, abi)
    oc = orc.OracleCase([desc], spec.models, sc)
    oc.set_state(0, 0, s.theta0[perm], s.tau0[perm], s.eigvals[perm], s.eigvecs[perm])
    fa = rm.face_addr
    phi_r = np.where(fa > 0, s.phi[np.abs(fa) - 1], -s.phi[np.abs(fa) - 1])
    oc.set_velocity(0, s.U[perm], s.Ub, phi_r)
    so = (abi.RheoStepStats * 1)()
    oc.store_old_time(); oc.step(s.dt, so)
    g.store_old_time(); sg = g.correct(s.dt, want_stats=True)
    assert list(sg[0].n_iterations) == list(so[0].n_iterations)
    assert max(so[0].n_iterations) >= 2
    np.testing.assert_allclose(list(sg[0].initial_residual), list(so[0].initial_residual), rtol=1e-9)
    th_o = np.empty_like(s.theta0); th_o[perm] = oc.get(0, 0, abi.FIELD_THETA)
    assert rel_l2(g.theta(), th_o) <= 1e-10


def test_pbicg_and_pbicgstab_agree_on_the_device():
    """Both Krylov methods solve the same assembled system: the converged fields agree to the solver tolerance."""
    spec = cases.by_name("C3", 3 / 19)
    s = Setup(spec, cfl=1.0)
    ga, gb = s.gpu(tight(spec.schemes, solver="PBiCG")), s.gpu(tight(spec.schemes))
    for g in (ga, gb):
        g.store_old_time(); g.correct(s.dt)
    assert ga.last_iterations() >= 2
    assert rel_l2(ga.theta(), gb.theta()) <= 1e-11


@pytest.mark.parametrize("name", ["OldroydBLog-2D-cubista", "GiesekusLog-3D-contraction-cubista", "PTTLog-linear-zeta-2D-minmod"])
def test_pbicg_against_the_reference_fixture(name):
    """... and with the reference's own numbers (the solution of the system its text assembles)."""
    gold = np.load(Path(__file__).resolve().parent / "golden" / "reference_correct.npz")
    spec, s = make_setup(name)
    g = s.gpu(tight(spec.schemes, solver="PBiCG"))
    g.store_old_time(); g.correct(s.dt)
    assert 1 in STORED_STEPS
    assert rel_l2(g.download(abi.FIELD_THETA, 0), gold[f"{name}/step1/theta"]) <= TOL_1
    assert rel_l2(g.download(abi.FIELD_TAU, 0), gold[f"{name}/step1/tau"]) <= TOL_1


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
