"""BASELINE.json config 1: "rheoFoam tutorial: 2D Oldroyd-BLog flow past a confined cylinder (stock tutorial mesh)".
rheotool_b200/blockmesh.py restates blockMesh + mirrorMesh (OpenFOAM-9 utilities, not in /root/reference) for the tutorial's
system/blockMeshDict (8 blocks, arc edges, simpleGrading) and system/mirrorMeshDict; the mesh is body-fitted (not a lattice),
so on the device it takes the cell-colouring order with the K = 4 assembly kernels.  CPU tests hold the generator to the
dictionary and to geometric identities; the GPU test holds the stress step on that mesh to the oracle."""
from pathlib import Path

import numpy as np
import pytest

from helpers import rel_l2, tight
from oracle import oracle as orc
from rheotool_b200 import abi, blockmesh, cases

REF_CASE = Path("/root/reference/of90/tutorials/rheoFoam/Cylinder/Oldroyd-BLog")


@pytest.fixture(scope="module")
def cyl(tmp_path_factory):
    return cases.stock_mesh(cases.cylinder_stock(), tmp_path_factory.mktemp("cylinder"))


def test_expansion_and_arc_follow_blockmesh():
    lam = blockmesh._divisions(43, 30.0)
    d = np.diff(lam)
    assert lam[0] == 0 and lam[-1] == 1 and np.all(d > 0)
    assert d[-1] / d[0] == pytest.approx(30.0, rel=1e-12)                      # simpleGrading: last / first cell size
    assert np.allclose(d[1:] / d[:-1], 30.0 ** (1 / 42), rtol=1e-12)           # geometric progression
    assert np.allclose(blockmesh._divisions(5, 1.0), np.arange(6) / 5)
    a = blockmesh._Arc((0, 1, 0), (0.195090322, 0.9807852804, 0), (0.3826834324, 0.9238795325, 0))   # blockMeshDict:65
    p = a.position(np.linspace(0, 1, 9))
    assert np.allclose(np.hypot(p[:, 0], p[:, 1]), 1.0, atol=1e-9) and np.allclose(a.c, 0, atol=1e-9)
    assert np.allclose(np.degrees(np.arctan2(p[:, 0], p[:, 1])), np.linspace(0, 22.5, 9), atol=1e-6)   # uniform in angle


def test_cylinder_stock_mesh_counts_patches_and_geometry(cyl):
    m = cyl
    # blockMeshDict:60-67: (33x40 + 40x43 + 3 x 23x43 + 20x43 + 60x43 + 50x60) cells, doubled by mirrorMesh
    assert m.n_cells == 2 * (33 * 40 + 40 * 43 + 3 * 23 * 43 + 20 * 43 + 60 * 43 + 50 * 60) == 24894
    assert m.patch_names == ["inlet", "walls", "cylinder", "outlet", "frontAndBack"]                     # no defaultFaces left on y = 0
    assert [p.size for p in m.patches] == [2 * 40, 2 * (33 + 40 + 23 + 23 + 23 + 20 + 50) - 80, 2 * (40 + 3 * 23 + 20 + 60), 2 * 60, 2 * m.n_cells]
    assert [p.type for p in m.patches] == [abi.PATCH_PATCH, abi.PATCH_WALL, abi.PATCH_WALL, abi.PATCH_PATCH, abi.PATCH_EMPTY]
    acc = np.zeros((m.n_cells, 3))
    np.add.at(acc, m.owner, m.Sf); np.subtract.at(acc, m.neighbour, m.Sf[: m.n_internal])
    assert np.abs(acc).max() < 1e-13 and m.V.min() > 0
    assert m.V.sum() == pytest.approx(80 * 4 - np.pi, rel=2e-6)               # channel minus the (polygonal) cylinder
    own, nei = m.owner[: m.n_internal], m.neighbour
    assert np.all(own < nei) and np.all(np.diff(own) >= 0)                    # upper-triangular order
    p = m.patches[m.patch_names.index("cylinder")]
    r = np.hypot(*m.Cf[p.start: p.start + p.size, :2].T)
    assert r.max() < 1.0 and r.min() > 0.9999                                 # chord mid-points of a 378-gon on the unit circle
    # mirrorMesh: the second half of the cells are the mirror images of the first, in the same order
    h = m.n_cells // 2
    assert np.allclose(m.C[h:, 0], m.C[:h, 0], atol=1e-12) and np.allclose(m.C[h:, 1], -m.C[:h, 1], atol=1e-12)
    assert np.allclose(m.V[h:], m.V[:h], rtol=1e-12)
    d = m.C[nei] - m.C[own]
    cosang = (d * m.Sf[: m.n_internal]).sum(1) / np.linalg.norm(d, axis=1) / np.linalg.norm(m.Sf[: m.n_internal], axis=1)
    assert 30 < np.degrees(np.arccos(cosang.min())) < 60                      # a body-fitted O-grid, not a lattice


@pytest.mark.skipif(not REF_CASE.exists(), reason="needs /root/reference (the builder's container)")
def test_parametrised_dictionary_equals_the_tutorials_own(tmp_path):
    """cylinder_tutorial_dicts() against system/blockMeshDict and system/mirrorMeshDict of the reference: same mesh, bit for bit"""
    a = blockmesh.generate(REF_CASE / "system" / "blockMeshDict", (0, 0, 2.5), (0, -1, 0), 1e-7)
    text, bp, nv, tol = blockmesh.cylinder_tutorial_dicts()
    (tmp_path / "blockMeshDict").write_text(text)
    b = blockmesh.generate(tmp_path / "blockMeshDict", bp, nv, tol)
    assert all(np.array_equal(x, y) for x, y in zip(a[:4], b[:4])) and a[4] == b[4]
    md = (REF_CASE / "system" / "mirrorMeshDict").read_text()
    assert "basePoint       (0 0 2.5)" in md and "normalVector    (0 -1 0)" in md and "planeTolerance      1e-7" in md


def _case(m):
    """the tutorial's model and BC kinds with a smooth synthetic velocity that vanishes on the cylinder (cases.cylinder_stock)"""
    spec = cases.cylinder_stock()
    U, Ub, phi, theta0, thetaB = cases.stock_fields(spec, m)
    dt = spec.cfl / m.max_courant_rate(phi)
    return spec.models, U, Ub, phi, theta0, thetaB, dt


def test_oracle_runs_on_the_stock_cylinder_mesh(cyl):
    m = cyl
    models, U, Ub, phi, theta0, thetaB, dt = _case(m)
    assert list(m.desc.solved_components) == [1, 1, 0, 1, 0, 1]              # 2-D: xx xy yy zz (empty frontAndBack)
    sc = tight(cases.scheme_ctl("cubista", "PBiCG", 1e-10))                  # the tutorial's solver (fvSolution:32-45)
    oc = orc.OracleCase([m.desc], models, sc)
    vals, vecs = orc.calc_eig(theta0)
    oc.set_state(0, 0, theta0, np.zeros_like(theta0), vals, vecs, theta_b=thetaB)
    oc.set_velocity(0, U, Ub, phi)
    for _ in range(2):
        oc.store_old_time(); oc.step(dt)
    th = oc.get(0, 0, abi.FIELD_THETA)
    assert np.isfinite(th).all() and 1e-4 < rel_l2(th, theta0) < 0.5
    assert np.abs(th[:, [2, 4]]).max() == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["PBiCGStab", "PBiCG"])
def test_gpu_matches_oracle_on_the_stock_cylinder_mesh(cyl, solver):
    """config 1 on the device: non-lattice quad mesh -> cell colouring, K = 4 rows; PBiCGStab (k_flux3 / k_source_init path) and the
    tutorial's PBiCG (round-1 assembly kernels + pbicg.cuh)"""
    from rheotool_b200.stress import GpuStressModel
    m = cyl
    models, U, Ub, phi, theta0, thetaB, dt = _case(m)
    sc = tight(cases.scheme_ctl("cubista", solver, 1e-10))
    oc = orc.OracleCase([m.desc], models, sc)
    vals, vecs = orc.calc_eig(theta0)
    oc.set_state(0, 0, theta0, np.zeros_like(theta0), vals, vecs, theta_b=thetaB)
    oc.set_velocity(0, U, Ub, phi)
    g = GpuStressModel(m, models, sc)
    g.upload_state(0, theta0, np.zeros_like(theta0), vals, vecs, theta_b=thetaB)
    g.upload_velocity(U, Ub, phi)
    assert "colouring" in g.ordering()
    for n in range(3):
        oc.store_old_time(); oc.step(dt)
        g.store_old_time(); g.correct(dt)
        if n == 0:
            assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-10
    assert rel_l2(g.theta(), oc.get(0, 0, abi.FIELD_THETA)) <= 1e-9
    assert rel_l2(g.tau(0), oc.get(0, 0, abi.FIELD_TAU)) <= 1e-9
    assert rel_l2(g.download(abi.FIELD_TAU_B), oc.get(0, 0, abi.FIELD_TAU_B)) <= 1e-9
    assert rel_l2(g.div_tau(abi.STAB_COUPLING), oc.div_tau(0, abi.STAB_COUPLING)) <= 1e-9


REF_CONTRACTION = Path("/root/reference/of90/tutorials/rheoFoam/Contraction41/Oldroyd-BLog/system/blockMeshDict")


@pytest.mark.skipif(not REF_CONTRACTION.exists(), reason="needs /root/reference (the builder's container)")
def test_contraction41_tutorial_dictionary_gives_the_mesh_config_2_refines(tmp_path):
    """two independent generators on BASELINE config 2's base mesh: the tutorial's own blockMeshDict (24 graded blocks) through
    blockmesh.py, and cases.contraction_2d(1, 1) (the tensor-grid generator behind C2, which refines every block x9): the same 11,991
    cells — centres and volumes to rounding — and the same boundary (the tutorial splits the walls into five patches)"""
    from rheotool_b200 import foamio, mesh
    P, faces, owner, nei, patches = blockmesh.generate(REF_CONTRACTION)
    blockmesh.write_polymesh(tmp_path / "constant" / "polyMesh", P, faces, owner, nei, patches)
    a = foamio.read_polymesh(tmp_path / "constant" / "polyMesh")
    b = mesh.tensor_grid(cases.contraction_2d(1, 1).grid)
    assert a.n_cells == b.n_cells == 11991 and a.n_internal == b.n_internal
    ka = np.lexsort((np.round(a.C[:, 1], 9), np.round(a.C[:, 0], 9)))
    kb = np.lexsort((np.round(b.C[:, 1], 9), np.round(b.C[:, 0], 9)))
    assert np.abs(a.C[ka] - b.C[kb]).max() <= 1e-11 and np.abs(a.V[ka] - b.V[kb]).max() <= 1e-12
    size = {n: p.size for n, p in zip(a.patch_names, a.patches)}
    sizeb = {n: p.size for n, p in zip(b.patch_names, b.patches)}
    assert size["inlet"] == sizeb["inlet"] and size["outlet"] == sizeb["outlet"] and size["frontAndBack"] == sizeb["frontAndBack"]
    assert sum(v for n, v in size.items() if n.startswith("wall")) == sizeb["walls"]
