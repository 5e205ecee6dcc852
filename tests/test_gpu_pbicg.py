"""PBiCG + DILU on the device (csrc/gpu/pbicg.cuh) — the solver every Log tutorial's fvSolution selects — against the
oracle's PBiCG (field parity AND the iteration history on the renumbered mesh), against PBiCGStab on the device and against
the reference fixtures."""
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


