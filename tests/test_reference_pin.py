"""Pins the oracle on rheoTool's OWN text for the stress step.

oracle/_ref/libref_stress.so is the reference's source compiled from where it lies (utils/jacobi.H,
utils/boilerLog.H, constitutiveEq::decomposeGradU / innerP, the correct() bodies of Oldroyd_BLog / GiesekusLog /
PTTLog / FENE_PLog / FENE_CRLog / WhiteMetznerCYLog / RoliePolyLog / XPomPomLog, gaussDefCmpwConvectionScheme.{H,C} + limiters.H, linearExtrapolationFvPatchField::updateCoeffs)
over a minimal OpenFOAM stand-in (oracle/ref_shim/, recipe oracle/Makefile `ref`).  Its outputs on the cases of
tests/reference_cases.py are committed as tests/golden/reference_*.npz by tools/make_golden_reference.py.

* `test_oracle_*_golden`: the oracle against the committed fixtures — runs everywhere.
* `test_live_*`: the oracle against the library itself on more inputs, and the fixtures against a fresh run — only
  where the library is built or /root/reference is present (skipped on the GPU box's CPU run otherwise).
* the CUDA path, through the C-ABI, against the same fixtures (the reference's numbers, not the oracle's) is in
  tests/test_gpu_reference_golden.py.
Tolerances: 1e-12 relative L2 oracle-vs-reference (measured: <= 4e-15), BASELINE's 1e-10 for the GPU after one step.
"""
from pathlib import Path

import numpy as np
import pytest

from helpers import Setup, rel_l2
from oracle import oracle as orc
from oracle import ref
from reference_cases import MATRIX_CASE, N_STEPS, REFERENCE_CASES, STORED_STEPS, _fixed_theta_walls, digest, make_setup
from rheotool_b200 import abi

GOLD = Path(__file__).resolve().parent / "golden"
TOL_ORACLE = 1e-12
TOL_GPU_1 = 1e-10
TOL_GPU_N = 1e-8     # N_STEPS chained steps (BASELINE: 1e-6 after 100)

live = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference not present")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD / "reference_correct.npz")


@pytest.fixture(scope="module")
def cell():
    return np.load(GOLD / "reference_cell.npz")


def _setup(name):
    spec, s = make_setup(name)
    oc = s.oracle(spec.schemes, sort_eig=False)
    return spec, s, oc


def _check_inputs(gold, name, s, theta_b):
    want = bytes(gold[f"{name}/inputs"]).decode()
    got = digest(s.U, s.Ub, s.phi, s.theta0, theta_b, s.eigvals, s.eigvecs, [s.dt], np.round(s.tau0, 9))
    assert got == want, "the synthetic inputs of this case changed: regenerate with tools/make_golden_reference.py"


# ---- per-cell text: jacobi.H, decomposeGradU, innerP, limiters.H -------------------------------------------------
def test_oracle_jacobi_golden(cell):
    """oracle jacobi == utils/jacobi.H: same rotations, same order (no sort), exp of the eigenvalues."""
    D, V = orc.jacobi(cell["theta"])
    assert np.abs(np.exp(D) - cell["expD"]).max() <= 4e-15 * np.abs(cell["expD"]).max()
    assert np.abs(V - cell["V"]).max() <= 1e-14
    vals, vecs = orc.calc_eig(cell["theta"], sort_eig=False)
    assert np.abs(vals[:, [0, 4, 8]] - cell["expD"]).max() <= 4e-15 * np.abs(cell["expD"]).max()
    assert np.abs(vecs.reshape(-1, 3, 3) - cell["V"]).max() <= 1e-14


def test_oracle_decompose_golden(cell):
    """oracle decomposeGradU == constitutiveEq.C:323-358 (+ innerP :471-518) given the same M = R^T L^T R."""
    R = cell["V"]
    vals = np.zeros((len(R), 9)); vals[:, [0, 4, 8]] = cell["expD"]
    M = cell["M"].reshape(-1, 3, 3)
    # the oracle entry point takes L and forms M itself: L^T = R M R^T
    L = np.transpose(R @ M @ np.transpose(R, (0, 2, 1)), (0, 2, 1))
    om, B = orc.decompose_gradU(L.reshape(-1, 9), R.reshape(-1, 9), vals)
    ok = np.abs(cell["omega"]).max(axis=1) < 1e6          # cells with (nearly) equal eigenvalues amplify round-off by 1/gap
    assert ok.sum() > 400
    assert rel_l2(om[ok], cell["omega"][ok]) <= 1e-11
    assert rel_l2(B, cell["B"]) <= 1e-13
    Rm, Mm = R, M
    assert rel_l2((np.transpose(Rm, (0, 2, 1)) @ Mm @ Rm).reshape(-1, 9), cell["innerP_T"]) <= 1e-14
    assert rel_l2((Rm @ Mm @ np.transpose(Rm, (0, 2, 1))).reshape(-1, 9), cell["innerP"]) <= 1e-14


def test_limiter_rows_golden(cell):
    """limiters.H:48-98 as compiled == the table the product and the oracle use (include/rheo_gpu.h order)."""
    want = {
        abi.LIMITER["upwind"]: ([1.0], [0.0], [1.0]),
        abi.LIMITER["cubista"]: ([7 / 4, 3 / 4, 1 / 4], [0.0, 3 / 8, 3 / 4], [3 / 8, 3 / 4]),
        abi.LIMITER["minmod"]: ([1.5, 0.5, 0.5], [0.0, 0.5, 0.5], [0.5, 1.0]),
        abi.LIMITER["smart"]: ([3.0, 3 / 4, 0.0], [0.0, 3 / 8, 1.0], [1 / 6, 5 / 6]),
        abi.LIMITER["waceb"]: ([2.0, 3 / 4, 0.0], [0.0, 3 / 8, 1.0], [3 / 10, 5 / 6]),
        abi.LIMITER["superbee"]: ([0.5, 1.5, 0.0], [0.5, 0.0, 1.0], [1 / 2, 2 / 3]),
    }
    for l, (a, b, bo) in want.items():
        assert np.array_equal(cell[f"lims/{l}/alpha"], np.array(a))
        assert np.array_equal(cell[f"lims/{l}/beta"], np.array(b))
        assert np.array_equal(cell[f"lims/{l}/bounds"], np.array(bo))


# ---- whole correct() ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(REFERENCE_CASES))
def test_oracle_correct_golden(gold, name):
    """N_STEPS chained XxxLog::correct() calls: theta, tau and their boundary fields of the oracle against the
    reference's (Krylov tolerance 1e-15 on the oracle side; the reference harness solves its own assembled system
    to round-off with a different method, so what is compared is the discrete equation, not an iteration history)."""
    spec, s, oc = _setup(name)
    _check_inputs(gold, name, s, oc.get(0, 0, abi.FIELD_THETA_B))
    for k in range(N_STEPS):
        oc.store_old_time(); oc.step(s.dt)
        if k + 1 not in STORED_STEPS:
            continue
        for fld, key in ((abi.FIELD_THETA, "theta"), (abi.FIELD_TAU, "tau"), (abi.FIELD_THETA_B, "theta_b"), (abi.FIELD_TAU_B, "tau_b")):
            err = rel_l2(oc.get(0, 0, fld), gold[f"{name}/step{k + 1}/{key}"])
            assert err <= TOL_ORACLE, f"{name} step {k + 1} {key}: {err:.2e}"


def test_reference_matrix_structure(gold):
    """The thetaEqn the reference assembled (fvm::ddt + GaussDefCmpw fvmDiv == source): upwind LDU coefficients
    lower = -max(phi,0), upper = min(phi,0) (gaussDefCmpwConvectionScheme.C:92-94), diagonal = V/dt - sum of the
    off-diagonals (negSumDiag), boundary coefficients phi_b x (1,0) on zeroGradient and (0,-value) on fixedValue."""
    name = MATRIX_CASE
    spec, s, oc = _setup(name)
    m = s.mesh
    nif = m.desc.n_internal_faces
    own = np.ctypeslib.as_array(m.desc.owner, (m.desc.n_faces,))
    nei = np.ctypeslib.as_array(m.desc.neighbour, (nif,))
    V = np.ctypeslib.as_array(m.desc.V, (m.n_cells,))
    phi = s.phi[:nif]
    assert np.array_equal(gold[f"{name}/matrix/lower"], -np.where(phi >= 0, 1.0, 0.0) * phi)
    assert np.array_equal(gold[f"{name}/matrix/upper"], (1.0 - np.where(phi >= 0, 1.0, 0.0)) * phi)
    diag = V / s.dt
    np.subtract.at(diag, own[:nif], gold[f"{name}/matrix/lower"])
    np.subtract.at(diag, nei, gold[f"{name}/matrix/upper"])
    assert np.abs(gold[f"{name}/matrix/diag"] - diag).max() <= 1e-13 * np.abs(diag).max()
    iC = gold[f"{name}/matrix/internalCoeffs"]
    for p in range(m.desc.n_patches):
        pd = m.desc.patches[p]
        sl = slice(pd.start - nif, pd.start - nif + pd.size)
        phib = s.phi[pd.start:pd.start + pd.size]
        if pd.theta_bc == abi.BC_ZERO_GRADIENT:
            assert np.array_equal(iC[sl], np.repeat(phib[:, None], 6, 1))
        elif pd.theta_bc == abi.BC_FIXED_VALUE:
            assert not iC[sl].any()


@live
def test_live_fixture_regenerates(gold):
    """The committed fixture is what the library produces today (guards against a stale fixture)."""
    name = "PTTLog-linear-zeta-2D-minmod"
    spec, s, oc = _setup(name)
    st = ref.correct(s.mesh.desc, spec.models[0], spec.schemes.limiter, s.dt, s.U, s.Ub, s.phi, s.theta0,
                     oc.get(0, 0, abi.FIELD_THETA_B), s.tau0, oc.get(0, 0, abi.FIELD_TAU_B), s.eigvals, s.eigvecs)
    for key in ("theta", "tau", "theta_b", "tau_b"):
        assert np.abs(st[key] - gold[f"{name}/step1/{key}"]).max() <= 1e-14 * max(1.0, np.abs(st[key]).max())


@live
def test_live_jacobi_many():
    rng = np.random.default_rng(99)
    th = rng.standard_normal((20000, 6)) * rng.uniform(1e-6, 5.0, (20000, 1))
    D, V, nrot = ref.jacobi(th)
    Do, Vo = orc.jacobi(th)
    assert np.abs(np.exp(Do) - D).max() <= 1e-14 * np.abs(D).max()
    assert np.abs(Vo - V).max() <= 1e-13
    assert nrot.max() <= 50 * 3


@live
@pytest.mark.parametrize("limiter", ["cubista", "minmod", "smart", "waceb", "superbee", "upwind"])
def test_live_every_limiter_larger_mesh(limiter):
    """Every limiter row on a mesh the fixture does not hold (C3 contraction, h = 1/2: 9,216 cells)."""
    from rheotool_b200 import cases
    spec = cases.by_name("C3", 2 / 19)
    if limiter == "superbee":       # see reference_cases._fixed_theta_walls: zeroGradient walls are ill-posed for superbee
        spec = _fixed_theta_walls(spec)
    spec.schemes = cases.scheme_ctl(limiter, "PBiCGStab", 1e-15, relax=0.0)
    s = Setup(spec)
    oc = s.oracle(spec.schemes, sort_eig=False)
    st = ref.correct(s.mesh.desc, spec.models[0], spec.schemes.limiter, s.dt, s.U, s.Ub, s.phi, s.theta0,
                     oc.get(0, 0, abi.FIELD_THETA_B), s.tau0, oc.get(0, 0, abi.FIELD_TAU_B), s.eigvals, s.eigvecs)
    oc.store_old_time(); oc.step(s.dt)
    for fld, key in ((abi.FIELD_THETA, "theta"), (abi.FIELD_TAU, "tau"), (abi.FIELD_TAU_B, "tau_b"), (abi.FIELD_EIGVALS, "eigvals"), (abi.FIELD_EIGVECS, "eigvecs")):
        assert rel_l2(oc.get(0, 0, fld), st[key]) <= TOL_ORACLE, key


@live
def test_live_hundred_chained_calls_stay_on_the_reference():
    """BASELINE's long-run bar (<= 1e-6 after 100 steps) for the oracle against the reference's text itself: 100 chained
    correct() calls with the tutorial Krylov tolerance scaled down to 1e-13 on the oracle side."""
    name = "PTTLog-linear-zeta-2D-minmod"
    spec, s = make_setup(name)
    from rheotool_b200 import cases
    oc = s.oracle(cases.scheme_ctl("minmod", "PBiCG", 1e-13), sort_eig=False)     # the tutorials' solver
    st = {"theta": s.theta0, "theta_b": oc.get(0, 0, abi.FIELD_THETA_B), "tau": s.tau0, "tau_b": oc.get(0, 0, abi.FIELD_TAU_B),
          "eigvals": s.eigvals, "eigvecs": s.eigvecs}
    for _ in range(100):
        st = ref.correct(s.mesh.desc, spec.models[0], spec.schemes.limiter, s.dt, s.U, s.Ub, s.phi, st["theta"], st["theta_b"],
                         st["tau"], st["tau_b"], st["eigvals"], st["eigvecs"])
        oc.store_old_time(); oc.step(s.dt)
    assert rel_l2(oc.get(0, 0, abi.FIELD_THETA), st["theta"]) <= 1e-9
    assert rel_l2(oc.get(0, 0, abi.FIELD_TAU), st["tau"]) <= 1e-9
    assert rel_l2(oc.get(0, 0, abi.FIELD_TAU_B), st["tau_b"]) <= 1e-9


@pytest.mark.parametrize("name,n", [("GiesekusLog-3D-contraction-cubista", (2, 2, 1)), ("PTTLog-linear-zeta-2D-minmod", (3, 1, 1)),
                                    ("FENEPLog-3D-cavity-cubista", (2, 2, 2))])
def test_decomposed_oracle_reproduces_the_single_rank_reference(gold, name, n):
    """decomposePar `simple` + N ranks against the reference's single-rank numbers: what mpirun -np N rheoFoam has to
    reproduce of its own serial run.  Exercises the oracle's restatement of the coupled-patch branches of the scheme
    (gaussDefCmpwConvectionScheme.C:104-110, 158-167, 289-319: plim = upw on processor patches, neighbour gradients, the
    once-per-face source on coupled faces) against numbers that came out of the non-coupled text."""
    spec, s = make_setup(name)
    nr = n[0] * n[1] * n[2]
    c2r = s.mesh.simple_decomp(*n)
    subs = [s.mesh.decompose(c2r, nr, r) for r in range(nr)]
    many = orc.OracleCase([x.desc for x in subs], spec.models, spec.schemes, False)
    addr = []
    for r, sub in enumerate(subs):
        ca, fa = sub.proc_addressing()
        addr.append(ca)
        many.set_state(r, 0, s.theta0[ca], s.tau0[ca], s.eigvals[ca], s.eigvecs[ca])
        gf = np.abs(fa) - 1
        ph = np.where(fa > 0, s.phi[gf], -s.phi[gf])
        gb = gf[sub.n_internal:] - s.mesh.n_internal
        Ub = np.zeros((sub.n_boundary, 3))
        Ub[gb >= 0] = s.Ub[gb[gb >= 0]]
        many.set_velocity(r, s.U[ca], Ub, ph)
    for k in range(N_STEPS):
        many.store_old_time(); many.step(s.dt)
        if k + 1 not in STORED_STEPS:
            continue
        for fld, key in ((abi.FIELD_THETA, "theta"), (abi.FIELD_TAU, "tau")):
            got = np.empty_like(gold[f"{name}/step{k + 1}/{key}"])
            for r in range(nr):
                got[addr[r]] = many.get(r, 0, fld)
            assert rel_l2(got, gold[f"{name}/step{k + 1}/{key}"]) <= 1e-11, (name, k, key)


@live
@pytest.mark.parametrize("name,n", [("GiesekusLog-3D-contraction-cubista", (2, 2, 1)), ("PTTLog-linear-zeta-2D-minmod", (3, 1, 1))])
def test_live_reference_text_on_several_ranks(gold, name, n):
    """The reference's text itself on a decomposed mesh (one thread per rank, processor patches: the coupled branches of
    gaussDefCmpwConvectionScheme.C:104-110, 158-167, 289-319, gradients with neighbour values, interfaces in the solve)
    against (1) its own single-rank numbers and (2) the oracle on the same emulated ranks."""
    spec, s = make_setup(name)
    nr = n[0] * n[1] * n[2]
    c2r = s.mesh.simple_decomp(*n)
    subs = [s.mesh.decompose(c2r, nr, r) for r in range(nr)]
    many = orc.OracleCase([x.desc for x in subs], spec.models, spec.schemes, False)
    states, addr = [], []
    for r, sub in enumerate(subs):
        ca, fa = sub.proc_addressing()
        addr.append(ca)
        gf = np.abs(fa) - 1
        ph = np.where(fa > 0, s.phi[gf], -s.phi[gf])
        gb = gf[sub.n_internal:] - s.mesh.n_internal
        Ub = np.zeros((sub.n_boundary, 3))
        Ub[gb >= 0] = s.Ub[gb[gb >= 0]]
        many.set_state(r, 0, s.theta0[ca], s.tau0[ca], s.eigvals[ca], s.eigvecs[ca])
        many.set_velocity(r, s.U[ca], Ub, ph)
        states.append({"U": s.U[ca], "Ub": Ub, "phi": ph, "theta": s.theta0[ca], "theta_b": many.get(r, 0, abi.FIELD_THETA_B),
                       "tau": s.tau0[ca], "tau_b": many.get(r, 0, abi.FIELD_TAU_B), "eigvals": s.eigvals[ca], "eigvecs": s.eigvecs[ca]})
    out = ref.correct_multi([x.desc for x in subs], spec.models[0], spec.schemes.limiter, s.dt, states)
    many.store_old_time(); many.step(s.dt)
    for key, fld in (("theta", abi.FIELD_THETA), ("tau", abi.FIELD_TAU)):
        whole = np.empty_like(gold[f"{name}/step1/{key}"])
        for r in range(nr):
            whole[addr[r]] = out[r][key]
            assert rel_l2(many.get(r, 0, fld), out[r][key]) <= TOL_ORACLE, (key, r)          # (2)
        assert rel_l2(whole, gold[f"{name}/step1/{key}"]) <= 1e-12, key                       # (1)
    for r in range(nr):   # boundary stresses incl. the walls next to processor faces
        tb_o, tb_r = many.get(r, 0, abi.FIELD_TAU_B), out[r]["tau_b"]
        phys = np.ones(len(tb_o), bool)
        d = subs[r].desc
        for p in range(d.n_patches):
            if d.patches[p].type == abi.PATCH_PROCESSOR:
                phys[d.patches[p].start - d.n_internal_faces: d.patches[p].start - d.n_internal_faces + d.patches[p].size] = False
        assert rel_l2(tb_o[phys], tb_r[phys]) <= TOL_ORACLE, r
