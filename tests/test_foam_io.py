"""OpenFOAM on-disk formats (include/rheo_io.h).  Pins: the polyMesh the reference ships
(tutorials/rheoFoam/Aneurysm/.../polyMesh.org: 4 files gzip + boundary) and every tau*/theta* field file of its tutorials —
read in this container only (/root/reference does not exist on the GPU box: those tests skip there) — plus write/read
round trips of generated meshes and fields, which run everywhere."""
import json
from pathlib import Path

import numpy as np
import pytest

from rheotool_b200 import abi, cases, foamio, mesh

REF = Path("/root/reference/of90/tutorials")
ANEURYSM = REF / "rheoFoam/Aneurysm/HerschelBulkley/constant/polyMesh.org"
GOLD = Path(__file__).parent / "golden" / "aneurysm_geometry.json"
needs_ref = pytest.mark.skipif(not REF.exists(), reason="reference tree not present (GPU box)")


def _geometry_summary(m):
    nint = m.n_internal
    # closedness: sum of outward face area vectors per cell
    acc = np.zeros((m.n_cells, 3))
    np.add.at(acc, m.owner, m.Sf)
    np.subtract.at(acc, m.neighbour, m.Sf[:nint])
    bnd = m.Sf[nint:]
    return {
        "n_cells": int(m.n_cells), "n_faces": int(m.n_faces), "n_internal_faces": int(nint),
        "patches": [[n, int(p.type), int(p.start), int(p.size)] for n, p in zip(m.patch_names, m.patches)],
        "volume": float(m.V.sum()),
        "volume_by_divergence": float((m.Cf[nint:] * bnd).sum() / 3.0),
        "min_volume": float(m.V.min()),
        "max_open": float(np.abs(acc).max()),
        "patch_area": [float(np.linalg.norm(m.Sf[p.start:p.start + p.size], axis=1).sum()) for p in m.patches],
        "min_weight": float(m.weights[:nint].min()), "max_weight": float(m.weights[:nint].max()),
    }


@needs_ref
def test_reference_aneurysm_polymesh_reads_and_is_a_valid_finite_volume_mesh():
    m = foamio.read_polymesh(ANEURYSM)
    s = _geometry_summary(m)
    # the counts OpenFOAM itself recorded in the owner file header (note "nPoints:... nCells:...")
    import gzip, re
    head = gzip.open(ANEURYSM / "owner.gz", "rt").read(2000)
    note = {k: int(v) for k, v in re.findall(r"(nPoints|nCells|nFaces|nInternalFaces):\s*(\d+)", head)}
    assert (s["n_cells"], s["n_faces"], s["n_internal_faces"]) == (note["nCells"], note["nFaces"], note["nInternalFaces"])
    assert foamio.mesh_counts(m)[0] == note["nPoints"]
    assert s["patches"][0][0] == "walls" and [p[0] for p in s["patches"]] == ["walls", "out1", "in1", "out2"]
    assert s["min_volume"] > 0
    assert s["max_open"] < 1e-12 * s["patch_area"][0]                    # every cell is closed
    assert s["volume"] == pytest.approx(s["volume_by_divergence"], rel=1e-10)   # Gauss: V = 1/3 sum Cf.Sf over the boundary
    assert 0 < s["min_weight"] and s["max_weight"] < 1
    gold = json.loads(GOLD.read_text())
    for k in ("n_cells", "n_faces", "n_internal_faces", "patches"):
        assert s[k] == gold[k]
    for k in ("volume", "min_volume"):
        assert s[k] == pytest.approx(gold[k], rel=1e-12)
    assert s["patch_area"] == pytest.approx(gold["patch_area"], rel=1e-12)


@needs_ref
def test_every_tau_and_theta_file_of_the_reference_tutorials_parses():
    files = sorted(p for p in REF.rglob("*") if p.is_file() and p.parent.name in ("0", "fluid") and (p.name.startswith("tau") or p.name.startswith("theta")))
    assert len(files) >= 60
    known = {"fixedValue", "zeroGradient", "linearExtrapolation", "empty", "symmetryPlane", "wedge", "symmetry", "cyclic", "calculated"}
    n_regex = 0
    for p in files:
        f = foamio.FoamField(p)
        assert f.cls in ("volSymmTensorField", "volTensorField", "volScalarField"), (p, f.cls)
        v = f.internal(3)
        assert v.shape == (3, f.n_comp) and np.isfinite(v).all()
        txt = p.read_text()
        # every boundaryField keyword resolves, regular-expression keys included
        import re
        body = txt[txt.index("boundaryField"):]
        for key in re.findall(r'^\s*("[^"]+"|[A-Za-z_][\w.]*)\s*\n?\s*\{', body, flags=re.M):
            if key == "boundaryField":
                continue
            if key.startswith('"'):
                n_regex += 1
                inner = key.strip('"').strip("()")
                name = inner.split("|")[0].replace(".*", "x")
            else:
                name = key
            ty, _ = f.patch(name, 2)
            assert ty in known, (p, key, ty)
    assert n_regex > 0


def test_polymesh_write_read_round_trip_of_a_generated_mesh(tmp_path):
    spec = cases.by_name("C3", 2 / 19)
    m = mesh.tensor_grid(spec.grid)
    foamio.write_polymesh(m, tmp_path / "polyMesh", gz=True)
    assert (tmp_path / "polyMesh" / "points.gz").exists() and (tmp_path / "polyMesh" / "boundary").exists()
    r = foamio.read_polymesh(tmp_path / "polyMesh")
    assert (r.n_cells, r.n_faces, r.n_internal) == (m.n_cells, m.n_faces, m.n_internal)
    assert np.array_equal(r.owner, m.owner) and np.array_equal(r.neighbour, m.neighbour)
    assert r.patch_names == m.patch_names[: len(m.patches)]
    assert [(p.type, p.start, p.size) for p in r.patches] == [(p.type, p.start, p.size) for p in m.patches]
    # geometry recomputed from the points agrees with the generator's analytic geometry
    for a, b in ((r.Sf, m.Sf), (r.Cf, m.Cf), (r.C, m.C)):
        assert np.abs(a - b).max() < 1e-12
    assert np.abs(r.V - m.V).max() < 1e-15 and np.abs(r.weights - m.weights).max() < 1e-12
    # and a second write of the mesh that was READ reproduces the files bit for bit
    foamio.write_polymesh(r, tmp_path / "again", gz=False)
    foamio.write_polymesh(m, tmp_path / "first", gz=False)
    for name in ("points", "faces", "owner", "neighbour", "boundary"):
        assert (tmp_path / "again" / name).read_bytes() == (tmp_path / "first" / name).read_bytes(), name


def test_field_write_read_round_trip_is_bit_exact_and_bcs_reach_the_mesh(tmp_path):
    spec = cases.by_name("C2", 1 / 9)
    m = mesh.tensor_grid(spec.grid)
    rng = np.random.default_rng(5)
    tau = rng.standard_normal((m.n_cells, 6)) * 10.0 ** rng.integers(-12, 6, (m.n_cells, 1))
    patches = []
    for name, p in zip(m.patch_names, m.patches):
        if p.type == abi.PATCH_EMPTY:
            patches.append((name, "empty", None))
        elif p.type == abi.PATCH_WALL:
            patches.append((name, "linearExtrapolation", rng.standard_normal((p.size, 6))))
        elif name == m.patch_names[0]:
            patches.append((name, "fixedValue", rng.standard_normal((p.size, 6))))
        else:
            patches.append((name, "zeroGradient", None))
    for gz in (False, True):
        path = tmp_path / ("gz" if gz else "plain") / "tau"
        foamio.write_field(path, "tau", tau, patches, "[1 -1 -2 0 0 0 0]", gz=gz)
        f = foamio.FoamField(path)
        assert (f.cls, f.object, f.n_comp, f.internal_uniform, f.n_internal) == ("volSymmTensorField", "tau", 6, False, m.n_cells)
        assert np.array_equal(f.internal(m.n_cells), tau)
        for name, ty, vals in patches:
            size = 0 if vals is None else len(vals)
            t2, v2 = f.patch(name, size)
            assert t2 == ty and (vals is None) == (v2 is None)
            if vals is not None:
                assert np.array_equal(v2, vals)
        with pytest.raises(KeyError):
            f.patch("noSuchPatch")
        f.apply_bcs(m, "tau")
        for (name, ty, _), p in zip(patches, m.patches):
            assert p.tau_bc == {"empty": abi.BC_EMPTY, "linearExtrapolation": abi.BC_LINEAR_EXTRAPOLATION, "fixedValue": abi.BC_FIXED_VALUE,
                                "zeroGradient": abi.BC_ZERO_GRADIENT}[ty]


def test_regular_expression_keys_follow_openfoam_lookup_rules(tmp_path):
    """exact keyword first; otherwise the LAST matching pattern (of90/tutorials/rheoFoam/Cylinder/Oldroyd-BLog/0/theta:33 style)."""
    (tmp_path / "theta").write_text("""
FoamFile { version 2.0; format ascii; class volSymmTensorField; object theta; }
dimensions [0 0 0 0 0 0 0];
internalField uniform (1 2 3 4 5 6);   // comment
boundaryField
{
    ".*"                { type zeroGradient; }
    "(walls|cylinder)"  { type fixedValue; value uniform (0 0 0 0 0 0); }
    inlet               { type fixedValue; value nonuniform List<symmTensor> 2 ((1 0 0 1 0 1) (2 0 0 2 0 2)); }
    /* block comment */
    frontAndBack        { type empty; }
}
""")
    f = foamio.FoamField(tmp_path / "theta")
    assert f.internal_uniform and np.array_equal(f.internal(2), np.tile([1, 2, 3, 4, 5, 6.0], (2, 1)))
    assert f.patch("outlet")[0] == "zeroGradient"
    ty, v = f.patch("cylinder", 3)
    assert ty == "fixedValue" and np.array_equal(v, np.zeros((3, 6)))
    ty, v = f.patch("inlet", 2)
    assert ty == "fixedValue" and np.array_equal(v, [[1, 0, 0, 1, 0, 1], [2, 0, 0, 2, 0, 2.0]])
    assert f.patch("frontAndBack")[0] == "empty"
    with pytest.raises(foamio.FoamError):
        f.patch("inlet", 3)            # value list does not match the patch size


def test_errors_name_the_file_and_the_problem(tmp_path):
    with pytest.raises(foamio.FoamError, match="cannot read"):
        foamio.read_polymesh(tmp_path / "nothing")
    (tmp_path / "bad").write_text("FoamFile { class volSymmTensorField; object x; } internalField uniform (1 2 3); boundaryField { }")
    with pytest.raises(foamio.FoamError, match="components"):
        foamio.FoamField(tmp_path / "bad")


def test_case_round_trip_preserves_the_stress_step(tmp_path):
    """constant/polyMesh + 0/{U,theta,tau,eigVals,eigVecs} written from a generated case, read back, and stepped by the oracle:
    the same answer as stepping the original arrays (field values travel bit-exactly; the geometry is recomputed from the points)."""
    from helpers import Setup, rel_l2, tight
    from oracle import oracle as orc
    spec = cases.by_name("C3", 2 / 19)
    s = Setup(spec)
    m = s.mesh
    phi = foamio.surface_flux(m, s.U, s.Ub)
    foamio.write_case(tmp_path / "case", m, "0", s.theta0, s.tau0, s.U, s.Ub, eigvals=s.eigvals, eigvecs=s.eigvecs, gz=True)
    m2, f = foamio.read_case(tmp_path / "case", "0")
    assert [(p.type, p.theta_bc, p.tau_bc) for p in m2.patches] == [(p.type, p.theta_bc, p.tau_bc) for p in m.patches]
    for k, ref in (("U", s.U), ("theta", s.theta0), ("tau", s.tau0), ("eigvals", s.eigvals), ("eigvecs", s.eigvecs)):
        assert np.array_equal(f[k], ref), k
    assert np.abs(f["phi"] - phi).max() < 1e-14 * np.abs(phi).max()
    sc = tight(spec.schemes)

    def run(mesh_, U, Ub, ph, th, vals, vecs):
        oc = orc.OracleCase([mesh_.desc], spec.models, sc)
        oc.set_state(0, 0, th, np.zeros_like(th), vals, vecs)
        oc.set_velocity(0, U, Ub, ph)
        for _ in range(3):
            oc.store_old_time(); oc.step(s.dt)
        return oc.get(0, 0, abi.FIELD_THETA), oc.get(0, 0, abi.FIELD_TAU)

    th1, ta1 = run(m, s.U, s.Ub, phi, s.theta0, s.eigvals, s.eigvecs)
    th2, ta2 = run(m2, f["U"], f["U_b"], f["phi"], f["theta"], f["eigvals"], f["eigvecs"])
    assert rel_l2(th2, th1) < 1e-12 and rel_l2(ta2, ta1) < 1e-12
    # a cold start: no eigVals / eigVecs files -> READ_IF_PRESENT gives None
    foamio.write_case(tmp_path / "cold", m, "0", s.theta0, s.tau0, s.U, s.Ub)
    _, g = foamio.read_case(tmp_path / "cold", "0")
    assert g["eigvals"] is None and g["eigvecs"] is None


def test_decomposed_case_round_trip_runs_like_the_undecomposed_one(tmp_path):
    """processorN/ directories (decomposePar layout, with cell/face/boundaryProcAddressing) written from the unstructured
    fixture, read back rank by rank, processor patches completed from the neighbours' directories — the oracle on those N
    ranks gives the answer of the oracle on the undecomposed mesh."""
    from helpers import rel_l2, tight
    from oracle import oracle as orc
    from test_unstructured import _case
    m, models, U, Ub, phi_unused, theta0, thetaB, dt = _case()
    phi = foamio.surface_flux(m, U, Ub)
    dt = 0.2 / m.max_courant_rate(phi)
    vals, vecs = orc.calc_eig(theta0)
    c2r = m.simple_decomp(3, 1, 1)
    foamio.write_decomposed_case(tmp_path / "case", m, c2r, "0", {"theta": theta0, "tau": np.zeros_like(theta0), "U": U, "U_b": Ub, "theta_b": thetaB,
                                                                   "eigvals": vals, "eigvecs": vecs}, gz=True)
    for r in range(3):
        pm = tmp_path / "case" / f"processor{r}" / "constant" / "polyMesh"
        assert (pm / "cellProcAddressing.gz").exists() and (pm / "faceProcAddressing.gz").exists() and (pm / "boundaryProcAddressing.gz").exists()
        assert "myProcNo" in (pm / "boundary").read_text()
    ranks = foamio.read_decomposed_case(tmp_path / "case", "0")
    assert len(ranks) == 3 and sum(mm.n_cells for mm, _ in ranks) == m.n_cells
    sc = tight(cases.scheme_ctl("cubista", "PBiCGStab", 1e-10))
    one = orc.OracleCase([m.desc], models, sc)
    one.set_state(0, 0, theta0, np.zeros_like(theta0), vals, vecs, theta_b=thetaB)
    one.set_velocity(0, U, Ub, phi)
    many = orc.OracleCase([mm.desc for mm, _ in ranks], models, sc)
    for r, (mm, f) in enumerate(ranks):
        many.set_state(r, 0, f["theta"], f["tau"], f["eigvals"], f["eigvecs"], theta_b=f["theta_b"])
        many.set_velocity(r, f["U"], f["U_b"], f["phi"])
    for _ in range(3):
        one.store_old_time(); one.step(dt)
        many.store_old_time(); many.step(dt)
    ref = one.get(0, 0, abi.FIELD_TAU)
    got = np.empty_like(ref)
    for r, (mm, _) in enumerate(ranks):
        ca, _fa = mm.proc_addressing()
        got[ca] = many.get(r, 0, abi.FIELD_TAU)
    assert rel_l2(got, ref) < 1e-10


def test_dictionary_variables_directives_and_code_blocks(tmp_path):
    """$variables (innermost scope first), #include-style directives (skipped) and #{ ... #} code blocks, as they occur in the
    reference's 0/U files (e.g. `value uniform ($uStart 0 0);`, `value $internalField;`, `#includeEtc "caseDicts/setConstraintTypes"`)."""
    (tmp_path / "U").write_text("""
FoamFile { version 2.0; format ascii; class volVectorField; object U; }
uStart 1.5;
dimensions [0 1 -1 0 0 0 0];
internalField uniform ($uStart 0 0);
boundaryField
{
    #includeEtc "caseDicts/setConstraintTypes"
    inlet   { type fixedValue; value uniform (${uStart} 0 0); }
    outlet  { type fixedValue; value $internalField; }
    coded   { type codedFixedValue; value uniform (0 0 0); name ramp; code #{ operator==(vector(1, 0, 0)); // not parsed ; { ( #}; }
}
""")
    f = foamio.FoamField(tmp_path / "U")
    assert np.array_equal(f.internal(2), [[1.5, 0, 0], [1.5, 0, 0]])
    assert np.array_equal(f.patch("inlet", 1)[1], [[1.5, 0, 0]])
    assert np.array_equal(f.patch("outlet", 2)[1], [[1.5, 0, 0], [1.5, 0, 0]])
    assert f.patch("coded", 1)[0] == "codedFixedValue"


@needs_ref
def test_every_velocity_file_of_the_reference_tutorials_parses():
    files = sorted(p for p in REF.rglob("U") if p.is_file() and p.parent.name in ("0", "fluid", "0.orig"))
    assert len(files) >= 60
    for p in files:
        f = foamio.FoamField(p)
        assert f.cls == "volVectorField" and f.internal(2).shape == (2, 3), p


@needs_ref
def test_case_dictionaries_of_the_reference_tutorials_give_models_and_schemes():
    """constant/constitutiveProperties + system/fvSchemes + system/fvSolution of every reference tutorial that selects one of the
    log-conformation models of this library (or a multiMode of them): models and schemes come out as the shim would read them."""
    import re
    seen, n_multi, n_schemes, n_refused, n_cn = {}, 0, 0, 0, 0
    for cp in sorted(REF.rglob("constitutiveProperties")):
        txt = cp.read_text()
        types = re.findall(r"^\s*type\s+([\w-]+)\s*;", txt, flags=re.M)
        if not types:
            continue
        try:   # single-phase layout only: a top-level `parameters` dictionary (two-phase cases nest it per phase: out of scope)
            foamio.FoamDict(cp).keys("parameters")
        except foamio.FoamError:
            continue
        first = types[0]
        wanted = first in abi.MODEL_NAMES or (first == "multiMode" and all(t in abi.MODEL_NAMES for t in types[1:] if t.endswith("Log")) and
                                                 all(t.endswith("Log") for t in types[1:len(types)]))
        if not wanted:
            with pytest.raises(foamio.FoamError):
                foamio.read_models(cp)
            continue
        models = foamio.read_models(cp)
        if first == "multiMode":
            n_multi += 1
            assert len(models) >= 2
        for md in models:
            assert md.etaP > 0 and md.lambda_ > 0 and md.rho > 0
        seen[first] = seen.get(first, 0) + 1
        case = cp.parent.parent
        if (case / "system" / "fvSchemes").exists() and "theta" in (case / "system" / "fvSchemes").read_text():
            try:
                ctl, solver = foamio.read_schemes(case, "theta" + foamio.mode_names(cp)[0])
            except foamio.FoamError as e:   # refused loudly: a time scheme (steadyState) or a gradient scheme the stress step does not have
                txt = (case / "system" / "fvSchemes").read_text()
                assert ("ddtSchemes" in str(e) and re.search(r"steadyState", txt)) or ("gradSchemes" in str(e) and not re.search(r"default\s+Gauss linear;", txt)), str(e)
                n_refused += 1
                continue
            n_schemes += 1
            assert solver in ("PBiCG", "PBiCGStab") and ctl.tolerance > 0 and ctl.ddt in (abi.DDT_EULER, abi.DDT_BACKWARD, abi.DDT_CRANK_NICOLSON)
            n_cn += ctl.ddt == abi.DDT_CRANK_NICOLSON
    assert seen.get("Oldroyd-BLog", 0) >= 5 and n_multi >= 1 and len(seen) >= 4, seen
    assert n_schemes >= 10 and n_cn >= 1, (n_schemes, n_refused, n_cn)
    # the SaramitoLog tutorial (SaramitoLog.C:108-165: k is read because n != 1, no PTT function then, dims (1 1 0))
    (m,) = foamio.read_models(REF / "rheoFoam/OtherTests/Channel2D_VE/otherModels/SaramitoLog/constant/constitutiveProperties")
    assert (m.model, m.sar_tau0, m.sar_k, m.sar_n, list(m.sar_dims), m.sar_ptt, m.zeta) == (abi.MODEL_SARAMITO_LOG, 2.5, 1.5, 0.75, [1.0, 1.0, 0.0], 0, 0.0)
    # the Cavity tutorial: `CrankNicolson 1`
    ctl, _ = foamio.read_schemes(REF / "rheoFoam/Cavity/Oldroyd-BLog")
    assert ctl.ddt == abi.DDT_CRANK_NICOLSON and ctl.cn_psi == 1.0
    # a spot value: the Cylinder tutorial
    (m,) = foamio.read_models(REF / "rheoFoam/Cylinder/Oldroyd-BLog/constant/constitutiveProperties")
    assert (m.model, m.rho, m.etaS, m.etaP, m.lambda_) == (abi.MODEL_OLDROYD_B_LOG, 1.0, 0.59, 0.41, 0.7)
    ctl, solver = foamio.read_schemes(REF / "rheoFoam/Cylinder/Oldroyd-BLog")
    assert ctl.limiter == abi.LIMITER["cubista"] and solver == "PBiCG" and ctl.solver == abi.SOLVER["PBiCG"] and ctl.ddt == abi.DDT_EULER


# ---- configurations the stress step does not implement are refused loudly, not run differently ------------------------
_FOAM_HEAD = 'FoamFile\n{\n    version 2.0;\n    format ascii;\n    class %s;\n    object %s;\n}\n'


def _write_case_dicts(case, grad_default="Gauss linear", solver="PBiCG", extra_theta=None):
    (case / "system").mkdir(parents=True, exist_ok=True)
    names = ["theta"] + (extra_theta or [])
    div = "".join(f"    div(phi,{n}) GaussDefCmpw cubista;\n" for n in names)
    (case / "system" / "fvSchemes").write_text(
        _FOAM_HEAD % ("dictionary", "fvSchemes") + "ddtSchemes { default Euler; }\n" + f"gradSchemes {{ default {grad_default}; }}\n" + "divSchemes\n{\n" + div + "}\n")
    sol = "".join(f"    {n} {{ solver {solver if n == 'theta' else 'PBiCGStab'}; preconditioner DILU; tolerance 1e-10; relTol 0; }}\n" for n in names)
    (case / "system" / "fvSolution").write_text(_FOAM_HEAD % ("dictionary", "fvSolution") + "solvers\n{\n" + sol + "}\n")


def test_schemes_the_device_does_not_implement_are_refused(tmp_path):
    """gradSchemes other than Gauss linear, per-mode controls that differ, unknown solvers — an error naming the
    entry, never a silent substitution; the solver named in fvSolution (PBiCG in every Log tutorial) is the solver that runs."""
    _write_case_dicts(tmp_path / "ok")
    ctl, solver = foamio.read_schemes(tmp_path / "ok")
    assert solver == "PBiCG" and ctl.solver == abi.SOLVER["PBiCG"]
    _write_case_dicts(tmp_path / "lsq", grad_default="leastSquares")
    with pytest.raises(foamio.FoamError, match="gradSchemes/default is `leastSquares`"):
        foamio.read_schemes(tmp_path / "lsq")
    _write_case_dicts(tmp_path / "gamg", solver="GAMG")
    with pytest.raises(foamio.FoamError, match="solver GAMG"):
        foamio.read_schemes(tmp_path / "gamg")
    _write_case_dicts(tmp_path / "mm", extra_theta=["thetaM2"])
    with pytest.raises(foamio.FoamError, match="differ from those of theta"):
        foamio.read_schemes_modes(tmp_path / "mm", ["theta", "thetaM2"])
    _write_case_dicts(tmp_path / "mm2", solver="PBiCGStab", extra_theta=["thetaM2"])
    ctl, solver = foamio.read_schemes_modes(tmp_path / "mm2", ["theta", "thetaM2"])
    assert solver == "PBiCGStab"


def test_linear_extrapolation_with_regression_is_read_and_written(tmp_path):
    """linearExtrapolationFvPatchField.C:72,118,152-219: `useRegression true` selects the least-squares branch — its own BC code,
    kept through a write / read round trip."""
    spec = cases.by_name("C3", 2 / 19)
    m = mesh.tensor_grid(spec.grid)
    n = m.n_cells
    zeros6 = np.zeros((n, 6))
    U, Ub, phi, th = m.synth_fields(spec.synth)
    foamio.write_case(tmp_path, m, "0", th, zeros6, U, Ub)
    f = foamio.FoamField(tmp_path / "0" / "tau")
    f.apply_bcs(m, "tau")   # as written: plain linearExtrapolation on the walls
    walls = [i for i, p in enumerate(m.patches) if p.tau_bc == abi.BC_LINEAR_EXTRAPOLATION]
    assert walls
    txt = (tmp_path / "0" / "tau").read_text()
    assert "linearExtrapolation" in txt
    (tmp_path / "0" / "tau").write_text(txt.replace("linearExtrapolation;", "linearExtrapolation;\n        useRegression   true;"))
    f2 = foamio.FoamField(tmp_path / "0" / "tau")
    f2.apply_bcs(m, "tau")
    assert all(m.patches[i].tau_bc == abi.BC_LINEAR_EXTRAPOLATION_REG for i in walls)
    foamio.write_case(tmp_path / "again", m, "0", th, zeros6, U, Ub)
    assert "useRegression   true" in (tmp_path / "again" / "0" / "tau").read_text()
    m2 = mesh.tensor_grid(spec.grid)
    foamio.FoamField(tmp_path / "again" / "0" / "tau").apply_bcs(m2, "tau")
    assert [p.tau_bc for p in m2.patches] == [p.tau_bc for p in m.patches]
