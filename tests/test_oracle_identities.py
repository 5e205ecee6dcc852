"""Pin the oracle through identities the reference algorithm must satisfy (the reference
ships no golden vectors for this path, so the oracle is pinned by algebra, by analytic material functions
(test_oracle_analytic.py) and by partition invariance).  CPU only."""
import numpy as np
import pytest

from helpers import Setup, rel_l2, tight
from oracle import oracle as orc
from rheotool_b200 import abi, cases

rng = np.random.default_rng(11)


def sym6_to_33(t):
    t = np.atleast_2d(t)
    return np.stack([t[:, [0, 1, 2]], t[:, [1, 3, 4]], t[:, [2, 4, 5]]], axis=1)


def random_theta(n, scale=0.6):
    return rng.standard_normal((n, 6)) * scale


def test_jacobi_restatement_against_lapack():
    """CE/utils/jacobi.H:7-158 restated: eigenvalues / eigenvectors agree with LAPACK to round-off."""
    th = random_theta(500)
    th[:20] = 0.0
    th[20:40, [1, 2, 4]] = 0.0
    D, V = orc.jacobi(th)
    A = sym6_to_33(th)
    w = np.linalg.eigvalsh(A)
    assert np.abs(np.sort(D, axis=1) - w).max() < 5e-14
    recon = V @ (D[:, :, None] * np.transpose(V, (0, 2, 1)))
    assert np.abs(recon - A).max() < 5e-14
    assert np.abs(V @ np.transpose(V, (0, 2, 1)) - np.eye(3)).max() < 5e-14


def test_calc_eig_is_exp_of_log_round_trip():
    """constitutiveEq.C:360-416: eigVals holds exp(eig) on the diagonal, eigVecs the eigenvectors in columns,
    so R Lambda R^T = exp(theta)."""
    n = 300
    th = random_theta(n)
    vals, vecs = orc.calc_eig(th)
    R, L = vecs.reshape(-1, 3, 3), vals.reshape(-1, 3, 3)
    A = R @ L @ np.transpose(R, (0, 2, 1))
    w, Q = np.linalg.eigh(sym6_to_33(th))
    Aref = Q @ (np.exp(w)[:, :, None] * np.transpose(Q, (0, 2, 1)))
    assert np.abs(A - Aref).max() < 1e-13 * np.abs(Aref).max()
    assert (np.diff(np.stack([vals[:, 0], vals[:, 4], vals[:, 8]], 1), axis=1) >= 0).all()
    assert np.abs(vals[:, [1, 2, 3, 5, 6, 7]]).max() == 0.0


def test_omega_B_decomposition_identity():
    """constitutiveEq.C:323-358: for distinct eigenvalues  Omega.A - A.Omega + 2 B.A  ==  X.A + A.X^T  with
    X = grad(U)^T, A = R Lambda R^T (the upper-convected terms the log-conformation split replaces)."""
    n = 200
    th = random_theta(n)
    vals, vecs = orc.calc_eig(th)
    L = 0.7 * rng.standard_normal((n, 9))
    om, B = orc.decompose_gradU(L, vecs, vals)
    R, Lam = vecs.reshape(-1, 3, 3), vals.reshape(-1, 3, 3)
    A = R @ Lam @ np.transpose(R, (0, 2, 1))
    X = np.transpose(L.reshape(-1, 3, 3), (0, 2, 1))
    Om, Bm = om.reshape(-1, 3, 3), B.reshape(-1, 3, 3)
    lhs = Om @ A - A @ Om + 2 * Bm @ A
    rhs = X @ A + A @ np.transpose(X, (0, 2, 1))
    gap = np.min(np.abs(np.diff(np.log(np.stack([vals[:, 0], vals[:, 4], vals[:, 8]], 1)), axis=1)), axis=1)
    ok = gap > 1e-3
    assert ok.sum() > n // 2
    assert np.abs(lhs[ok] - rhs[ok]).max() < 1e-10 * np.abs(rhs).max()
    assert np.abs(Om + np.transpose(Om, (0, 2, 1))).max() < 1e-12 * max(1.0, np.abs(Om).max())   # antisymmetric
    assert np.abs(Bm - np.transpose(Bm, (0, 2, 1))).max() < 1e-13 * max(1.0, np.abs(Bm).max())   # symmetric


def test_isotropic_cells_drop_the_commutator_like_the_reference():
    """theta = 0 (every tutorial's t = 0): R = I, Lambda = I, so omega = (..)/(1e-16) * 0-difference terms:
    B = diag(grad(U)^T) and the rotation part is finite*0 (SURVEY.md §7 hard parts)."""
    n = 10
    L = rng.standard_normal((n, 9))
    I9 = np.tile(np.eye(3).reshape(9), (n, 1))
    om, B = orc.decompose_gradU(L, I9, I9)
    assert np.isfinite(om).all() and np.isfinite(B).all()
    Bm = B.reshape(-1, 3, 3)
    Lt = np.transpose(L.reshape(-1, 3, 3), (0, 2, 1))
    assert np.allclose(np.diagonal(Bm, axis1=1, axis2=2), np.diagonal(Lt, axis1=1, axis2=2), rtol=0, atol=1e-15)
    assert np.abs(Bm - np.diagonal(Bm, axis1=1, axis2=2)[:, :, None] * np.eye(3)).max() == 0.0


@pytest.mark.parametrize("name,scale,n", [("C3", 3 / 19, (2, 2, 1)), ("C2", 1 / 9, (3, 1, 1)), ("C5", 12 / 400, (2, 2, 2))])
def test_partition_invariance_of_the_oracle(name, scale, n):
    """decomposePar + mpirun -np N must not change the answer: the oracle on N emulated
    ranks (processor patches, halo copies, rank-ordered reductions) against the oracle on one rank."""
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
    for _ in range(2):
        one.store_old_time(); one.step(s.dt)
        many.store_old_time(); many.step(s.dt)
    for mi in range(len(spec.models)):
        for fld in (abi.FIELD_THETA, abi.FIELD_TAU):
            ref = one.get(0, mi, fld)
            got = np.empty_like(ref)
            for r in range(nr):
                got[addr[r]] = many.get(r, mi, fld)
            assert rel_l2(got, ref) < 1e-11, (name, mi, fld)


def test_pbicg_and_pbicgstab_converge_to_the_same_field():
    """The tutorials select PBiCG (fvSolution:32-45), north_star names PBiCGStab: with a tight tolerance the
    Krylov flavour is immaterial."""
    spec = cases.by_name("C3", 3 / 19)
    s = Setup(spec, cfl=1.5)
    a = s.oracle(tight(spec.schemes, 1e-14, solver="PBiCGStab"))
    b = s.oracle(tight(spec.schemes, 1e-14, solver="PBiCG"))
    sa, sb = (abi.RheoStepStats * 1)(), (abi.RheoStepStats * 1)()
    a.store_old_time(); a.step(s.dt, sa)
    b.store_old_time(); b.step(s.dt, sb)
    assert max(sa[0].n_iterations) >= 2 and max(sb[0].n_iterations) >= 2
    assert rel_l2(a.get(0, 0, abi.FIELD_THETA), b.get(0, 0, abi.FIELD_THETA)) < 1e-11
    np.testing.assert_allclose(list(sa[0].initial_residual), list(sb[0].initial_residual), rtol=1e-12)


def test_limiters_reduce_to_upwind_on_uniform_fields():
    """The high-resolution part is a DEFERRED (explicit) correction evaluated on the current theta
    (gaussDefCmpwConvectionScheme.C:120-167,259-274).  With a spatially uniform theta, theta_N - theta_P = 0
    and grad(theta) = 0, so phi~ = 1 -> (alpha,beta) = (1,0) and the correction vanishes: every limiter then
    gives exactly the implicit-upwind step."""
    spec = cases.by_name("C5", 10 / 400)
    out = {}
    for lim in ("upwind", "cubista", "minmod", "smart", "waceb", "superbee", "none"):
        s = Setup(spec)
        s.theta0[:] = np.array([0.3, 0.05, -0.02, -0.1, 0.04, 0.2])
        s.eigvals, s.eigvecs = orc.calc_eig(s.theta0)
        sc = tight(spec.schemes)
        sc.limiter = abi.LIMITER[lim]
        oc = s.oracle(sc)
        oc.store_old_time(); oc.step(s.dt)
        out[lim] = oc.get(0, 0, abi.FIELD_THETA)
    for lim in ("cubista", "minmod", "smart", "waceb", "superbee"):
        assert rel_l2(out[lim], out["upwind"]) < 1e-13, lim
    assert rel_l2(out["none"], out["upwind"]) > 1e-4   # `none` drops convection altogether


def test_oracle_regression_fixture():
    """Guards the checker itself against accidental edits: one step of a tiny C3 case must reproduce the
    committed fixture tests/golden/oracle_c3_tiny.npz (written by tools/make_golden_oracle.py from THIS
    oracle — a regression pin, not a reference pin; the reference ships no vectors)."""
    from pathlib import Path
    f = Path(__file__).resolve().parent / "golden" / "oracle_c3_tiny.npz"
    g = np.load(f)
    spec = cases.by_name("C3", 1 / 19)
    s = Setup(spec)
    oc = s.oracle(tight(spec.schemes))
    oc.store_old_time(); oc.step(s.dt)
    assert s.mesh.n_cells == int(g["n_cells"])
    assert rel_l2(oc.get(0, 0, abi.FIELD_THETA), g["theta"]) < 1e-12
    assert rel_l2(oc.get(0, 0, abi.FIELD_TAU), g["tau"]) < 1e-12
    assert rel_l2(oc.get(0, 0, abi.FIELD_TAU_B), g["tau_b"]) < 1e-12
