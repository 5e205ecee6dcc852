"""The stress step on an UNSTRUCTURED polyhedral mesh: a 3,000-cell piece of the polyMesh the reference ships
(tests/golden/aneurysm_patch, cut by tools/make_fixture_aneurysm_patch.py around a snappyHexMesh refinement transition and
read through rheo_io_read_polymesh).
Tensor grids only ever have 4 or 6 slots per cell and 2 colours; this mesh drives the generic paths: run-time slot
count (KT = 0 kernels), more than two DILU colours, cells with different numbers of faces, zero-size patches."""
from pathlib import Path

import numpy as np
import pytest

from helpers import rel_l2, tight
from oracle import mesh_ref
from oracle import oracle as orc
from rheotool_b200 import abi, cases, foamio

FIXTURE = Path(__file__).parent / "golden" / "aneurysm_patch"


def _case(wall_tau_bc=abi.BC_LINEAR_EXTRAPOLATION):
    m = foamio.read_polymesh(FIXTURE)
    # BCs as in the reference's viscoelastic tutorials: walls zeroGradient theta / linearExtrapolation tau; the cut surface
    # (in- and outflow) holds fixedValue theta
    bc = {"walls": (abi.BC_ZERO_GRADIENT, wall_tau_bc), "cut": (abi.BC_FIXED_VALUE, abi.BC_ZERO_GRADIENT)}
    for name, p in zip(m.patch_names, m.desc.patches[: m.desc.n_patches]):
        p.theta_bc, p.tau_bc = bc.get(name, (abi.BC_ZERO_GRADIENT, abi.BC_ZERO_GRADIENT))
    c0, L = m.C.mean(0), np.ptp(m.C, axis=0).max()

    def vel(x):
        s = (x - c0) / L * 3.0
        return np.stack([0.8 + 0.3 * np.sin(s[:, 1]), 0.4 * np.cos(s[:, 0] + s[:, 2]), 0.5 * np.sin(s[:, 0]) * np.cos(s[:, 1])], axis=1)

    U, Ub = vel(m.C), vel(m.Cf[m.n_internal:])
    phi = (vel(m.Cf) * m.Sf).sum(1)
    s = (m.C - c0) / L * 4.0
    S = np.stack([0.4 * np.sin(s[:, 0]), 0.2 * np.cos(s[:, 1]), 0.1 * np.sin(s[:, 2]), -0.3 * np.cos(s[:, 0] + s[:, 1]), 0.15 * np.sin(s[:, 1] - s[:, 2]),
                  0.25 * np.cos(s[:, 2])], axis=1)
    rng = np.random.default_rng(12345)
    theta0 = S + 0.05 * (rng.random(S.shape) - 0.5)
    sb = (m.Cf[m.n_internal:] - c0) / L * 4.0
    thetaB = np.stack([0.4 * np.sin(sb[:, 0]), 0.2 * np.cos(sb[:, 1]), 0.1 * np.sin(sb[:, 2]), -0.3 * np.cos(sb[:, 0] + sb[:, 1]),
                       0.15 * np.sin(sb[:, 1] - sb[:, 2]), 0.25 * np.cos(sb[:, 2])], axis=1)
    dt = 0.2 / m.max_courant_rate(phi)
    models = [cases.model_desc("GiesekusLog", rho=1.0, etaS=0.01, etaP=0.99, lambda_=200 * dt, alpha=0.2)]
    return m, models, U, Ub, phi, theta0, thetaB, dt


def _oracle(m, models, sc, U, Ub, phi, theta0, thetaB):
    oc = orc.OracleCase([m.desc], models, sc)
    vals, vecs = orc.calc_eig(theta0)
    oc.set_state(0, 0, theta0, np.zeros_like(theta0), vals, vecs, theta_b=thetaB)
    oc.set_velocity(0, U, Ub, phi)
    return oc, vals, vecs


def test_fixture_is_an_unstructured_mesh_and_the_oracle_runs_on_it():
    m, models, U, Ub, phi, theta0, thetaB, dt = _case()
    assert (m.n_cells, m.n_internal) == (3000, 8375)
    assert m.patch_names == ["walls", "out1", "in1", "out2", "cut"] and [p.size for p in m.patches] == [60, 0, 0, 0, 2069]
    acc = np.zeros((m.n_cells, 3))
    np.add.at(acc, m.owner, m.Sf); np.subtract.at(acc, m.neighbour, m.Sf[: m.n_internal])
    assert np.abs(acc).max() < 1e-12 * np.linalg.norm(m.Sf, axis=1).max() and m.V.min() > 0
    faces_per_cell = np.bincount(m.owner, minlength=m.n_cells) + np.bincount(m.neighbour, minlength=m.n_cells)
    assert len(set(faces_per_cell.tolist())) >= 5 and faces_per_cell.max() == 21   # hexahedra and 9..21-faced transition polyhedra
    perm, colour, cstart = mesh_ref.colour_renumber(m.n_cells, m.owner[: m.n_internal], m.neighbour)
    assert len(cstart) - 1 > 2
    sc = tight(cases.scheme_ctl("cubista", "PBiCGStab", 1e-10))
    oc, _, _ = _oracle(m, models, sc, U, Ub, phi, theta0, thetaB)
    for _ in range(3):
        oc.store_old_time(); oc.step(dt)
    th = oc.get(0, 0, abi.FIELD_THETA)
    assert np.isfinite(th).all() and 1e-4 < rel_l2(th, theta0) < 0.5


@pytest.mark.gpu
@pytest.mark.parametrize("limiter", ["cubista", "upwind"])
def test_gpu_matches_oracle_on_the_unstructured_mesh(limiter):
    from rheotool_b200.stress import GpuStressModel
    m, models, U, Ub, phi, theta0, thetaB, dt = _case()
    sc = tight(cases.scheme_ctl(limiter, "PBiCGStab", 1e-10))
    oc, vals, vecs = _oracle(m, models, sc, U, Ub, phi, theta0, thetaB)
    g = GpuStressModel(m, models, sc)
    g.upload_state(0, theta0, np.zeros_like(theta0), vals, vecs, theta_b=thetaB)
    g.upload_velocity(U, Ub, phi)
    _, cstart = g.renumbering()
    assert len(cstart) - 1 > 2
    for n in range(5):
        oc.store_old_time(); oc.step(dt)
        g.store_old_time(); g.correct(dt)
        if n == 0:
            assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-10
            assert rel_l2(g.tau(0), oc.get(0, 0, abi.FIELD_TAU)) <= 1e-10
    assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-8
    assert rel_l2(g.tau(0), oc.get(0, 0, abi.FIELD_TAU)) <= 1e-8
    assert rel_l2(g.download(abi.FIELD_TAU_B), oc.get(0, 0, abi.FIELD_TAU_B)) <= 1e-8


@pytest.mark.gpu
def test_gpu_regression_walls_match_oracle_on_the_unstructured_mesh():
    """linearExtrapolation with `useRegression true` (linearExtrapolationFvPatchField.C:152-219) on polyhedral wall cells (9 to 21
    faces: the run-time slot loop of k_tau_bc_regress); the oracle's branch is pinned on the reference's text
    (tests/golden/reference_correct.npz: *-regressionWalls)."""
    from rheotool_b200.stress import GpuStressModel
    m, models, U, Ub, phi, theta0, thetaB, dt = _case(abi.BC_LINEAR_EXTRAPOLATION_REG)
    sc = tight(cases.scheme_ctl("cubista", "PBiCGStab", 1e-10))
    oc, vals, vecs = _oracle(m, models, sc, U, Ub, phi, theta0, thetaB)
    g = GpuStressModel(m, models, sc)
    g.upload_state(0, theta0, np.zeros_like(theta0), vals, vecs, theta_b=thetaB)
    g.upload_velocity(U, Ub, phi)
    for n in range(3):
        oc.store_old_time(); oc.step(dt)
        g.store_old_time(); g.correct(dt)
    assert rel_l2(g.tau(0), oc.get(0, 0, abi.FIELD_TAU)) <= 1e-9
    ref_b = oc.get(0, 0, abi.FIELD_TAU_B)
    assert rel_l2(g.download(abi.FIELD_TAU_B), ref_b) <= 1e-9
    plain = _oracle(*( _case()[:2] + (sc,) + _case()[2:7]))[0]
    for n in range(3):
        plain.store_old_time(); plain.step(dt)
    assert rel_l2(plain.get(0, 0, abi.FIELD_TAU_B), ref_b) > 1e-4, "the regression branch must differ from the gradient branch"


@pytest.mark.parametrize("n", [(2, 1, 1), (2, 2, 1)])
def test_partition_invariance_on_the_unstructured_mesh(n):
    """decomposePar `simple` + N ranks must not change the answer on a polyhedral mesh either (processor patches cutting
    through refinement transitions): the oracle on N emulated ranks against the oracle on one rank."""
    m, models, U, Ub, phi, theta0, thetaB, dt = _case()
    sc = tight(cases.scheme_ctl("cubista", "PBiCGStab", 1e-10))
    one, vals, vecs = _oracle(m, models, sc, U, Ub, phi, theta0, thetaB)
    nr = n[0] * n[1] * n[2]
    c2r = m.simple_decomp(*n)
    subs = [m.decompose(c2r, nr, r) for r in range(nr)]
    assert all(any(p.type == abi.PATCH_PROCESSOR for p in s.patches) for s in subs)
    many = orc.OracleCase([x.desc for x in subs], models, sc)
    addr = []
    for r, sub in enumerate(subs):
        ca, fa = sub.proc_addressing()
        addr.append(ca)
        gf = np.abs(fa) - 1
        gb = gf[sub.n_internal:] - m.n_internal          # global boundary face of each local boundary face (< 0: processor face)
        tb = np.zeros((sub.n_boundary, 6)); tb[gb >= 0] = thetaB[gb[gb >= 0]]
        ub = np.zeros((sub.n_boundary, 3)); ub[gb >= 0] = Ub[gb[gb >= 0]]
        many.set_state(r, 0, theta0[ca], np.zeros((len(ca), 6)), vals[ca], vecs[ca], theta_b=tb)
        many.set_velocity(r, U[ca], ub, np.where(fa > 0, phi[gf], -phi[gf]))
    for _ in range(3):
        one.store_old_time(); one.step(dt)
        many.store_old_time(); many.step(dt)
    for fld in (abi.FIELD_THETA, abi.FIELD_TAU):
        ref = one.get(0, 0, fld)
        got = np.empty_like(ref)
        for r in range(nr):
            got[addr[r]] = many.get(r, 0, fld)
        assert rel_l2(got, ref) < 1e-11, fld


# ---- the reference's own text on the polyhedral mesh (oracle/_ref; see tests/test_reference_pin.py) -------------------------
REF_GOLD = Path(__file__).parent / "golden" / "reference_unstructured.npz"


def reference_on_the_unstructured_mesh(limiter, steps=2):
    """[state after each correct()] of rheoTool's text (oracle/_ref) on the fixture; used by tools/make_golden_reference.py."""
    from oracle import ref
    m, models, U, Ub, phi, theta0, thetaB, dt = _case()
    sc = tight(cases.scheme_ctl(limiter, "PBiCGStab", 1e-10))
    oc, vals, vecs = _oracle(m, models, sc, U, Ub, phi, theta0, thetaB)
    st = {"theta": theta0, "theta_b": oc.get(0, 0, abi.FIELD_THETA_B), "tau": np.zeros_like(theta0), "tau_b": oc.get(0, 0, abi.FIELD_TAU_B),
          "eigvals": vals, "eigvecs": vecs}
    out = []
    for _ in range(steps):
        st = ref.correct(m.desc, models[0], abi.LIMITER[limiter], dt, U, Ub, phi, st["theta"], st["theta_b"], st["tau"], st["tau_b"],
                         st["eigvals"], st["eigvecs"])
        out.append(st)
    return out


@pytest.mark.parametrize("limiter", ["cubista", "upwind"])
def test_oracle_matches_the_reference_text_on_the_unstructured_mesh(limiter):
    """Non-orthogonal polyhedra (face area vectors not aligned with the centre-to-centre vectors, 4 to 21 faces per cell):
    Gauss gradients, the limiter's d = C_N - C_P geometry and linearExtrapolation with C_f - C_P are exercised where a tensor
    grid cannot.  Oracle against the committed outputs of the reference's text (tolerance 1e-12; measured 1e-15)."""
    gold = np.load(REF_GOLD)
    m, models, U, Ub, phi, theta0, thetaB, dt = _case()
    sc = tight(cases.scheme_ctl(limiter, "PBiCGStab", 1e-10))
    oc, _, _ = _oracle(m, models, sc, U, Ub, phi, theta0, thetaB)
    for k in range(2):
        oc.store_old_time(); oc.step(dt)
        for fld, key in ((abi.FIELD_THETA, "theta"), (abi.FIELD_TAU, "tau"), (abi.FIELD_TAU_B, "tau_b")):
            if key == "tau_b" and k > 0:
                continue   # see below
            assert rel_l2(oc.get(0, 0, fld), gold[f"{limiter}/step{k + 1}/{key}"]) <= 1e-12, (limiter, k, key)
    # KNOWN DIFFERENCE: this mesh lists `walls` (tau linearExtrapolation) BEFORE `cut` (tau zeroGradient).  While
    # `walls` evaluates its Gauss gradient, the reference harness has `cut` holding the boundary value of the expression just
    # assigned to tau_ (GeometricField::operator= assigns non-fixed patches; with eigVals_/eigVecs_ boundary values never
    # updated after construction that value is 0), the oracle and the device have it holding the previously evaluated
    # zeroGradient value.  Only wall faces whose cell also touches `cut` see it, from the second call on, and only in tau_b:
    tb_o, tb_r = oc.get(0, 0, abi.FIELD_TAU_B), gold[f"{limiter}/step2/tau_b"]
    walls = m.desc.patches[0]
    sl = slice(walls.start - m.n_internal, walls.start - m.n_internal + walls.size)
    rest = np.ones(len(tb_o), bool); rest[sl] = False
    assert rel_l2(tb_o[rest], tb_r[rest]) <= 1e-12                       # every other patch agrees
    touching = set(m.owner[m.desc.patches[4].start: m.desc.patches[4].start + m.desc.patches[4].size].tolist())
    wall_cells = m.owner[walls.start: walls.start + walls.size]
    differs = np.abs(tb_o[sl] - tb_r[sl]).max(axis=1) > 1e-9 * np.abs(tb_r).max()
    assert all(int(c) in touching for c in wall_cells[differs])          # ... and so does every wall face away from `cut`
    # and that reading of the assignment is the WHOLE difference: switched on in the oracle, tau_b agrees everywhere
    oc2, _, _ = _oracle(m, models, sc, U, Ub, phi, theta0, thetaB)
    oc2.set_tau_assignment(True)
    for k in range(2):
        oc2.store_old_time(); oc2.step(dt)
    assert rel_l2(oc2.get(0, 0, abi.FIELD_TAU_B), tb_r) <= 1e-12
    assert rel_l2(oc2.get(0, 0, abi.FIELD_THETA), gold[f"{limiter}/step2/theta"]) <= 1e-12
