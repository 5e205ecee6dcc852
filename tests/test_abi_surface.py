"""The drop-in boundary (CPU only, no compute calls): librheo_b200.so loads, exports every symbol the
headers under include/ declare, its struct layouts match the ctypes mirror, and — on a machine without a
CUDA device — the compute entry points fail loudly instead of falling back to a CPU path."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from rheotool_b200 import abi, cases, mesh

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    names = set()
    for h in sorted((ROOT / "include").glob("*.h")):
        txt = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names |= set(re.findall(r"\b(rheo_[a-z0-9_]+)\s*\(", txt))
    return names


def test_every_declared_symbol_is_exported_and_bound():
    decl = declared_functions()
    assert len(decl) >= 35
    exported = {}
    for path in (abi.lib_path(), abi.host_lib_path()):
        out = subprocess.run(["nm", "-D", "--defined-only", str(path)], capture_output=True, text=True, check=True).stdout
        exported[path.name] = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    gpu_decl = {n for n in decl if n.startswith("rheo_gpu_")}
    assert gpu_decl <= exported["librheo_b200.so"], f"declared but not exported: {sorted(gpu_decl - exported['librheo_b200.so'])}"
    assert decl - gpu_decl <= exported["librheo_host.so"], f"declared but not exported: {sorted(decl - gpu_decl - exported['librheo_host.so'])}"
    assert not any(n.startswith("rheo_gpu_") for n in exported["librheo_host.so"]), "the host library holds no compute entry point"
    bound = set(abi.MESH_SYMBOLS) | set(abi.GPU_SYMBOLS) | set(abi.IO_SYMBOLS)
    assert decl == bound, f"header / ctypes mismatch: {sorted(decl ^ bound)}"
    lib = abi.lib()
    for n in decl:
        assert getattr(lib, n) is not None


def test_no_oracle_or_torch_in_the_product_library():
    """The product never links or loads the oracle (test infrastructure) nor torch."""
    out = subprocess.run(["ldd", str(abi.lib_path())], capture_output=True, text=True).stdout
    assert "liboracle" not in out and "torch" not in out and "libc10" not in out
    host = subprocess.run(["ldd", str(abi.host_lib_path())], capture_output=True, text=True).stdout
    assert "liboracle" not in host and "cudart" not in host and "libcuda" not in host, "librheo_host.so must not depend on CUDA"
    for src in (ROOT / "rheotool_b200").rglob("*"):
        if src.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp", ".inl") and src.is_file():
            txt = src.read_text()
            assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, src


def test_struct_layouts_match_the_ctypes_mirror():
    sizes = np.zeros(8, dtype=np.int32)
    assert abi.lib().rheo_gpu_abi_sizes(sizes.ctypes.data_as(C.c_void_p)) == 0
    mirror = [abi.RheoPatchDesc, abi.RheoMeshDesc, abi.RheoModelDesc, abi.RheoSchemeCtl, abi.RheoStepStats, abi.RheoSynthSpec,
              abi.RheoPatchRule, abi.RheoPatchSpec]
    assert [C.sizeof(t) for t in mirror] == sizes.tolist()


def test_constants_match_the_headers():
    txt = (ROOT / "include" / "rheo_gpu.h").read_text() + (ROOT / "include" / "rheo_mesh.h").read_text()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(RHEO_[A-Z0-9_]+)\s+(\d+)", txt)}
    assert defs["RHEO_MODEL_OLDROYD_B_LOG"] == abi.MODEL_NAMES["Oldroyd-BLog"]
    assert defs["RHEO_MODEL_GIESEKUS_LOG"] == abi.MODEL_NAMES["GiesekusLog"]
    assert defs["RHEO_MODEL_PTT_LOG"] == abi.MODEL_NAMES["PTTLog"]
    assert defs["RHEO_MODEL_FENE_P_LOG"] == abi.MODEL_NAMES["FENE-PLog"]
    for macro, name in (("FENE_CR_LOG", "FENE-CRLog"), ("WM_CY_LOG", "WhiteMetznerCYLog"), ("ROLIE_POLY_LOG", "Rolie-PolyLog"),
                        ("XPOMPOM_LOG", "XPomPomLog"), ("SARAMITO_LOG", "SaramitoLog")):
        assert defs["RHEO_MODEL_" + macro] == abi.MODEL_NAMES[name]
    assert (defs["RHEO_DDT_EULER"], defs["RHEO_DDT_BACKWARD"], defs["RHEO_DDT_CRANK_NICOLSON"]) == (abi.DDT_EULER, abi.DDT_BACKWARD, abi.DDT_CRANK_NICOLSON)
    for name, val in abi.LIMITER.items():
        assert defs["RHEO_LIMITER_" + name.upper()] == val
    assert defs["RHEO_SOLVER_PBICGSTAB"] == abi.SOLVER["PBiCGStab"] and defs["RHEO_SOLVER_PBICG"] == abi.SOLVER["PBiCG"]
    assert (defs["RHEO_FIELD_THETA"], defs["RHEO_FIELD_TAU"], defs["RHEO_FIELD_EIGVALS"], defs["RHEO_FIELD_EIGVECS"], defs["RHEO_FIELD_THETA_B"],
            defs["RHEO_FIELD_TAU_B"], defs["RHEO_FIELD_TAU_TOTAL"], defs["RHEO_FIELD_THETA_OLD"], defs["RHEO_FIELD_TAU_B_TOTAL"]) == tuple(range(9))
    assert (defs["RHEO_PATCH_PATCH"], defs["RHEO_PATCH_WALL"], defs["RHEO_PATCH_EMPTY"], defs["RHEO_PATCH_PROCESSOR"]) == (0, 1, 2, 3)
    assert (defs["RHEO_BC_FIXED_VALUE"], defs["RHEO_BC_ZERO_GRADIENT"], defs["RHEO_BC_LINEAR_EXTRAPOLATION"], defs["RHEO_BC_EMPTY"],
            defs["RHEO_BC_PROCESSOR"]) == (0, 1, 2, 3, 4)
    # round 2: BMPLog, the fluidity fields, the regression flavour of linearExtrapolation, the stabilization options of divTau
    assert defs["RHEO_MODEL_BMP_LOG"] == abi.MODEL_NAMES["BMPLog"] and defs["RHEO_MODEL_BMP_FLUIDITY"] not in abi.MODEL_NAMES.values()
    assert (defs["RHEO_FIELD_FLUIDITY"], defs["RHEO_FIELD_FLUIDITY_B"]) == (abi.FIELD_FLUIDITY, abi.FIELD_FLUIDITY_B)
    assert defs["RHEO_BC_LINEAR_EXTRAPOLATION_REG"] == abi.BC_LINEAR_EXTRAPOLATION_REG
    assert (defs["RHEO_STAB_NONE"], defs["RHEO_STAB_BSD"], defs["RHEO_STAB_COUPLING"]) == (abi.STAB_NONE, abi.STAB_BSD, abi.STAB_COUPLING)
    assert defs["RHEO_DDT_STEADY_STATE"] == abi.DDT_STEADY_STATE


def test_compute_calls_fail_loudly_without_a_gpu():
    if abi.lib().rheo_gpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from rheotool_b200.stress import GpuStressModel, RheoError, eig_exp
    spec = cases.by_name("C3", 2 / 19)
    m = mesh.tensor_grid(spec.grid)
    with pytest.raises(RheoError, match="no CPU fallback"):
        GpuStressModel(m, spec.models, spec.schemes)
    with pytest.raises(RheoError, match="no CPU fallback"):
        eig_exp(np.zeros((4, 6)))


def test_run_time_selection_mirror_rejects_unknown_types_like_the_reference():
    """constitutiveEq::New (newConstitutiveEq.C:40-60): unknown `type` is a fatal error listing the valid
    types; missing `type` is reported as such.  multiMode expands its `models` list (multiMode.C:73-92)."""
    from rheotool_b200.stress import RheoError, models_from_dict
    with pytest.raises(RheoError, match="Unknown constitutiveEq type Oldroyd-B"):
        models_from_dict({"type": "Oldroyd-B", "etaP": 1.0, "lambda": 1.0})
    with pytest.raises(RheoError, match="type"):
        models_from_dict({"etaP": 1.0})
    mm = models_from_dict({"type": "multiMode", "models": [
        ("M1", {"type": "GiesekusLog", "rho": 1, "etaS": 0.0, "etaP": 0.3, "lambda": 0.1, "alpha": 0.2}),
        ("M2", {"type": "PTTLog", "rho": 1, "etaS": 0.0, "etaP": 0.2, "lambda": 0.5, "epsilon": 0.1, "zeta": 0.0,
                "destructionFunctionType": "exponential"})]})
    assert [m.model for m in mm] == [abi.MODEL_GIESEKUS_LOG, abi.MODEL_PTT_LOG]
    assert mm[1].ptt_function == abi.PTT_EXPONENTIAL and mm[0].alpha == 0.2
