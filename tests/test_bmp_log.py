"""SURVEY.md §8f rank 2, the last Log model: BMPLog (otherModels/BMP/BMPLog/BMPLog.C:142-201) — a thixotropic model whose
fluidity Phi obeys its own transport equation (ddt + GaussDefCmpw convection == -Sp(1/lambda) + Phi0/lambda +
k (PhiInf - Phi) (tau && symm(grad U))), solved before theta in every correct(); theta then relaxes at the rate Phi G0 and
tau = G0 (c - I).  CPU tests hold the oracle to limits with known answers; GPU tests hold the device to the oracle."""
import numpy as np
import pytest

from helpers import Setup, rel_l2, tight
from oracle import oracle as orc
from rheotool_b200 import abi, cases

G0, PHI0 = 0.8, 2.5
BMP_PIN_CASES = [("C3", 3 / 19), ("C2", 1 / 9)]   # live against oracle/_ref (the builder's container)
BMP_FIXTURE_CASES = [("C3", 1 / 19)]                # committed as tests/golden/reference_bmp.npz (small: it travels with the repo)


def _bmp(**kw):
    base = dict(rho=1.0, etaS=0.01, etaP=0.01, lambda_=0.7, bmp_G0=G0, bmp_k=0.0, bmp_Phi0=PHI0, bmp_PhiInf=40.0)
    base.update(kw)
    return cases.model_desc("BMPLog", **base)


def _spec(name, scale, model):
    spec = cases.by_name(name, scale)
    spec.models = [model]
    return spec


def _phi0(s, value=PHI0, vary=0.0):
    Phi = np.full(s.mesh.n_cells, value) * (1.0 + vary * np.sin(5 * s.mesh.C[:, 0]) * np.cos(3 * s.mesh.C[:, 1]))
    own_b = s.mesh.owner[s.mesh.n_internal:]
    return Phi, Phi[own_b].copy()   # boundary values: those of the adjacent cells (fixedValue patches keep them)


def test_without_structure_kinetics_bmp_is_oldroyd_b_with_lambda_from_the_fluidity():
    """k = 0 and Phi = Phi0 everywhere (inlet included): Phi stays Phi0, so theta relaxes at Phi0 G0 = 1/lambda_OB and
    tau = G0 (c - I) = (etaP_OB/lambda_OB)(c - I) with lambda_OB = 1/(Phi0 G0), etaP_OB = 1/Phi0."""
    spec = _spec("C3", 3 / 19, _bmp())
    s = Setup(spec)
    a = s.oracle(tight(spec.schemes))
    Phi, Phi_b = _phi0(s)
    a.set_fluidity(0, 0, Phi, Phi_b)
    spec_ob = _spec("C3", 3 / 19, cases.model_desc("Oldroyd-BLog", rho=1.0, etaS=0.01, etaP=1.0 / PHI0, lambda_=1.0 / (PHI0 * G0)))
    b = Setup(spec_ob).oracle(tight(spec.schemes))
    for oc in (a, b):
        for _ in range(2):
            oc.store_old_time(); oc.step(s.dt)
    assert np.abs(a.get(0, 0, abi.FIELD_FLUIDITY) - PHI0).max() <= 1e-12 * PHI0
    assert rel_l2(a.get(0, 0, abi.FIELD_THETA), b.get(0, 0, abi.FIELD_THETA)) <= 1e-12
    assert rel_l2(a.get(0, 0, abi.FIELD_TAU), b.get(0, 0, abi.FIELD_TAU)) <= 1e-12
    # divTau: the same stress, and a coupling term that scales with each model's own etaP (BMP: 0.01, "only used for stabilization")
    assert rel_l2(a.div_tau(0, abi.STAB_NONE), b.div_tau(0, abi.STAB_NONE)) <= 1e-11
    ca = a.div_tau(0, abi.STAB_COUPLING) - a.div_tau(0, abi.STAB_NONE)
    cb = b.div_tau(0, abi.STAB_COUPLING) - b.div_tau(0, abi.STAB_NONE)
    assert np.abs(cb).max() > 0 and rel_l2(ca * ((1.0 / PHI0) / 0.01), cb) <= 1e-9


def test_fluidity_of_a_fluid_at_rest_relaxes_to_phi0_at_the_rate_one_over_lambda():
    """U = 0: (Phi1 - Phi_old)/dt = (Phi0 - Phi1)/lambda for the Euler scheme, cell by cell"""
    spec = _spec("C5", 10 / 400, _bmp(bmp_k=3.0))
    s = Setup(spec)
    oc = s.oracle(tight(spec.schemes))
    Phi, Phi_b = _phi0(s, value=7.0)
    oc.set_fluidity(0, 0, Phi, Phi_b)
    oc.set_velocity(0, np.zeros_like(s.U), np.zeros_like(s.Ub), np.zeros_like(s.phi))
    dt, lam = 0.05, 0.7
    oc.store_old_time(); oc.step(dt)
    exact = (7.0 / dt + PHI0 / lam) / (1.0 / dt + 1.0 / lam)
    assert np.abs(oc.get(0, 0, abi.FIELD_FLUIDITY) - exact).max() <= 1e-12 * exact


def test_shear_breaks_the_structure_down():
    """k > 0 in a flow with tau && D > 0 somewhere: the fluidity rises above the value the k = 0 model reaches"""
    out = []
    for k in (0.0, 2.0):
        spec = _spec("C3", 3 / 19, _bmp(bmp_k=k))
        s = Setup(spec)
        oc = s.oracle(tight(spec.schemes))
        Phi, Phi_b = _phi0(s, vary=0.05)
        oc.set_fluidity(0, 0, Phi, Phi_b)
        for _ in range(3):
            oc.store_old_time(); oc.step(s.dt)
        out.append((oc.get(0, 0, abi.FIELD_FLUIDITY), oc.get(0, 0, abi.FIELD_THETA)))
    assert np.isfinite(out[1][0]).all() and np.isfinite(out[1][1]).all()
    assert np.abs(out[1][0] - out[0][0]).max() > 1e-6 * PHI0
    assert rel_l2(out[1][1], out[0][1]) > 1e-9


def test_bmp_is_single_mode_only():
    spec = _spec("C3", 3 / 19, _bmp())
    spec.models = [_bmp(), _bmp()]
    with pytest.raises(RuntimeError):
        Setup(spec).oracle(tight(spec.schemes))


@pytest.mark.gpu
@pytest.mark.parametrize("name,scale,ddt", [("C3", 4 / 19, "Euler"), ("C2", 1 / 9, "Euler"), ("C5", 16 / 400, "backward")])
def test_gpu_bmp_log_matches_oracle(name, scale, ddt):
    spec = _spec(name, scale, _bmp(bmp_k=2.0, bmp_relax=0.0))
    sc = tight(spec.schemes)
    sc.ddt = {"Euler": abi.DDT_EULER, "backward": abi.DDT_BACKWARD}[ddt]
    s = Setup(spec)
    oc, g = s.oracle(sc), s.gpu(sc)
    Phi, Phi_b = _phi0(s, vary=0.05)
    oc.set_fluidity(0, 0, Phi, Phi_b); g.upload_fluidity(0, Phi, Phi_b)
    for _ in range(3):
        oc.store_old_time(); oc.step(s.dt)
        g.store_old_time(); g.correct(s.dt)
    assert rel_l2(g.fluidity(0), oc.get(0, 0, abi.FIELD_FLUIDITY)) <= 1e-10
    assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-10
    assert rel_l2(g.tau(0), oc.get(0, 0, abi.FIELD_TAU)) <= 1e-10
    assert rel_l2(g.div_tau(abi.STAB_COUPLING), oc.div_tau(0, abi.STAB_COUPLING)) <= 1e-10


@pytest.mark.gpu
def test_gpu_bmp_log_relaxed_equations_match_oracle():
    """PhiEqn.relax() and thetaEqn.relax() with their own factors (BMPLog.C:165,189)"""
    spec = _spec("C3", 3 / 19, _bmp(bmp_k=1.0, bmp_relax=0.6))
    sc = tight(spec.schemes)
    sc.relax = 0.8
    s = Setup(spec)
    oc, g = s.oracle(sc), s.gpu(sc)
    Phi, Phi_b = _phi0(s, vary=0.05)
    oc.set_fluidity(0, 0, Phi, Phi_b); g.upload_fluidity(0, Phi, Phi_b)
    for _ in range(2):
        oc.store_old_time(); oc.step(s.dt)
        g.store_old_time(); g.correct(s.dt)
    assert rel_l2(g.fluidity(0), oc.get(0, 0, abi.FIELD_FLUIDITY)) <= 1e-10
    assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-10


@pytest.mark.parametrize("n", [(2, 2, 1)])
def test_partition_invariance_of_the_bmp_oracle(n):
    """the fluidity equation across processor patches (emulated ranks vs one rank): its hidden mode takes part in the halo copies,
    the rank-ordered reductions and the per-rank DILU like every other right-hand side"""
    spec = _spec("C3", 3 / 19, _bmp(bmp_k=2.0))
    s = Setup(spec)
    sc = tight(spec.schemes)
    one = s.oracle(sc)
    Phi, Phi_b = _phi0(s, vary=0.05)
    one.set_fluidity(0, 0, Phi, Phi_b)
    nr = n[0] * n[1] * n[2]
    c2r = s.mesh.simple_decomp(*n)
    subs = [s.mesh.decompose(c2r, nr, r) for r in range(nr)]
    many = orc.OracleCase([x.desc for x in subs], spec.models, sc)
    addr = []
    for r, sub in enumerate(subs):
        ca, fa = sub.proc_addressing()
        addr.append(ca)
        many.set_state(r, 0, s.theta_mode(0)[ca], s.tau0[ca], s.eigvals_mode(0)[ca], s.eigvecs_mode(0)[ca])
        gf = np.abs(fa) - 1
        ph = np.where(fa > 0, s.phi[gf], -s.phi[gf])
        gb = gf[sub.n_internal:] - s.mesh.n_internal
        Ub = np.zeros((sub.n_boundary, 3)); Ub[gb >= 0] = s.Ub[gb[gb >= 0]]
        Pb = np.zeros(sub.n_boundary); Pb[gb >= 0] = Phi_b[gb[gb >= 0]]
        many.set_velocity(r, s.U[ca], Ub, ph)
        many.set_fluidity(r, 0, Phi[ca], Pb)
    for _ in range(2):
        one.store_old_time(); one.step(s.dt)
        many.store_old_time(); many.step(s.dt)
    for fld in (abi.FIELD_FLUIDITY, abi.FIELD_THETA, abi.FIELD_TAU):
        ref = one.get(0, 0, fld)
        got = np.empty_like(ref)
        for r in range(nr):
            got[addr[r]] = many.get(r, 0, fld)
        assert rel_l2(got, ref) < 1e-11, fld


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["PBiCG", "RHEO_FLUX=1", "CrankNicolson"])
def test_gpu_bmp_log_on_the_other_code_paths(variant, monkeypatch):
    """the fluidity equation through the round-1 assembly kernels (PBiCG selects them, RHEO_FLUX=1 forces them) and with the ddt0
    field of CrankNicolson"""
    spec = _spec("C3", 3 / 19, _bmp(bmp_k=1.5))
    sc = tight(spec.schemes, solver="PBiCG" if variant == "PBiCG" else None)
    if variant == "CrankNicolson":
        sc.ddt, sc.cn_psi = abi.DDT_CRANK_NICOLSON, 0.9
    if variant == "RHEO_FLUX=1":
        monkeypatch.setenv("RHEO_FLUX", "1")
    s = Setup(spec)
    oc, g = s.oracle(sc), s.gpu(sc)
    Phi, Phi_b = _phi0(s, vary=0.05)
    oc.set_fluidity(0, 0, Phi, Phi_b); g.upload_fluidity(0, Phi, Phi_b)
    for _ in range(3):
        oc.store_old_time(); oc.step(s.dt)
        g.store_old_time(); g.correct(s.dt)
    ref = oc.get(0, 0, abi.FIELD_FLUIDITY)
    assert np.isfinite(ref).all()
    assert rel_l2(g.fluidity(0), ref) <= 1e-9
    assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-9
    assert rel_l2(g.tau(0), oc.get(0, 0, abi.FIELD_TAU)) <= 1e-9


# ---- BMPLog::correct against the reference's own text -----------------------------------------------------------------------------
# oracle/_ref compiles the WHOLE function (BMPLog.C:142-201): boilerLog.H, the fluidity equation (a scalar fvMatrix of the stand-in
# types: fvm::ddt, the reference's GaussDefCmpw scheme instantiated for a scalar, fvm::Sp, the explicit source with tau && symm(L)),
# the theta equation with the new fluidity, calcEig, theta -> tau, the tau boundary conditions.
def bmp_reference_run(name, scale, steps=3):
    """chained reference calls next to the oracle; -> per step (fluidity, theta, tau, tau_b) of the REFERENCE"""
    from oracle import ref
    spec = _spec(name, scale, _bmp(bmp_k=2.0))
    spec.schemes = cases.scheme_ctl("cubista", "PBiCGStab", 1e-15, relax=0.0)
    s = Setup(spec)
    oc = s.oracle(spec.schemes, sort_eig=False)
    Phi, Phi_b = _phi0(s, vary=0.05)
    oc.set_fluidity(0, 0, Phi, Phi_b)
    st = {"theta": s.theta0, "theta_b": oc.get(0, 0, abi.FIELD_THETA_B), "tau": s.tau0, "tau_b": oc.get(0, 0, abi.FIELD_TAU_B),
          "eigvals": s.eigvals, "eigvecs": s.eigvecs, "fluidity": Phi, "fluidity_b": oc.get(0, 0, abi.FIELD_FLUIDITY_B)}
    out = []
    for _ in range(steps):
        oc.store_old_time(); oc.step(s.dt)
        st = ref.correct(s.mesh.desc, spec.models[0], spec.schemes.limiter, s.dt, s.U, s.Ub, s.phi, st["theta"], st["theta_b"], st["tau"],
                         st["tau_b"], st["eigvals"], st["eigvecs"], fluidity=st["fluidity"], fluidity_b=st["fluidity_b"], solve_fluidity=True)
        out.append((st["fluidity"].copy(), st["theta"], st["tau"], st["tau_b"]))
        assert rel_l2(oc.get(0, 0, abi.FIELD_FLUIDITY), st["fluidity"]) <= 1e-12, (name, "fluidity")
        for key, fld in (("theta", abi.FIELD_THETA), ("tau", abi.FIELD_TAU), ("tau_b", abi.FIELD_TAU_B)):
            assert rel_l2(oc.get(0, 0, fld), st[key]) <= 1e-12, (name, key)
    return out


def _ref_available():
    from oracle import ref
    return ref.available()


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref not built and /root/reference not present")
@pytest.mark.parametrize("name,scale", BMP_PIN_CASES)
def test_live_bmp_correct_matches_the_reference_text(name, scale):
    bmp_reference_run(name, scale)


@pytest.mark.parametrize("name,scale", BMP_FIXTURE_CASES)
def test_oracle_bmp_golden(name, scale):
    """tests/golden/reference_bmp.npz (tools/make_golden_reference.py): fluidity, theta, tau and tau_b of the reference's
    BMPLog::correct after one, two and three chained calls"""
    from pathlib import Path
    gold = np.load(Path(__file__).parent / "golden" / "reference_bmp.npz")
    spec = _spec(name, scale, _bmp(bmp_k=2.0))
    spec.schemes = cases.scheme_ctl("cubista", "PBiCGStab", 1e-15, relax=0.0)
    s = Setup(spec)
    oc = s.oracle(spec.schemes, sort_eig=False)
    Phi, Phi_b = _phi0(s, vary=0.05)
    oc.set_fluidity(0, 0, Phi, Phi_b)
    for k in range(3):
        oc.store_old_time(); oc.step(s.dt)
        assert rel_l2(oc.get(0, 0, abi.FIELD_FLUIDITY), gold[f"{name}/step{k + 1}/fluidity"]) <= 1e-12
        for key, fld in (("theta", abi.FIELD_THETA), ("tau", abi.FIELD_TAU), ("tau_b", abi.FIELD_TAU_B)):
            assert rel_l2(oc.get(0, 0, fld), gold[f"{name}/step{k + 1}/{key}"]) <= 1e-12, (name, k, key)


@pytest.mark.gpu
@pytest.mark.parametrize("name,scale", BMP_FIXTURE_CASES)
def test_gpu_bmp_matches_the_reference_fixture(name, scale):
    """the CUDA path against the reference's numbers (not the oracle's) for BMPLog's theta and tau"""
    from pathlib import Path
    gold = np.load(Path(__file__).parent / "golden" / "reference_bmp.npz")
    spec = _spec(name, scale, _bmp(bmp_k=2.0))
    spec.schemes = cases.scheme_ctl("cubista", "PBiCGStab", 1e-15, relax=0.0)
    s = Setup(spec)
    g = s.gpu(spec.schemes)
    Phi, Phi_b = _phi0(s, vary=0.05)
    g.upload_fluidity(0, Phi, Phi_b)
    for k in range(3):
        g.store_old_time(); g.correct(s.dt)
        assert rel_l2(g.fluidity(0), gold[f"{name}/step{k + 1}/fluidity"]) <= 1e-10 * (10 ** k)
        assert rel_l2(g.theta(), gold[f"{name}/step{k + 1}/theta"]) <= 1e-10 * (10 ** k)
        assert rel_l2(g.tau(0), gold[f"{name}/step{k + 1}/tau"]) <= 1e-10 * (10 ** k)
