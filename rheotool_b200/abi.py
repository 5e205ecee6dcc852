"""ctypes mirror of include/rheo_mesh.h and include/rheo_gpu.h, and the loader of librheo_b200.so.

The product path fails loudly: there is no CPU fallback — if the shared library (or, for compute
calls, a CUDA device) is missing, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os as _os
# RHEO_LIB_PATH: load another build of the same library (kernel tuning experiments with RHEO_NVCC_EXTRA variants)
_LIB_PATH = Path(_os.environ.get("RHEO_LIB_PATH") or (Path(__file__).resolve().parent / "librheo_b200.so"))
_HOST_LIB_PATH = Path(__file__).resolve().parent / "librheo_host.so"

# ---- constants (keep in sync with include/*.h) ---------------------------------------------------
PATCH_PATCH, PATCH_WALL, PATCH_EMPTY, PATCH_PROCESSOR = 0, 1, 2, 3
BC_FIXED_VALUE, BC_ZERO_GRADIENT, BC_LINEAR_EXTRAPOLATION, BC_EMPTY, BC_PROCESSOR, BC_LINEAR_EXTRAPOLATION_REG = 0, 1, 2, 3, 4, 5
MODEL_OLDROYD_B_LOG, MODEL_GIESEKUS_LOG, MODEL_PTT_LOG, MODEL_FENE_P_LOG, MODEL_FENE_CR_LOG, MODEL_WM_CY_LOG, MODEL_ROLIE_POLY_LOG, MODEL_XPOMPOM_LOG, MODEL_SARAMITO_LOG, MODEL_BMP_LOG = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9
PTT_LINEAR, PTT_EXPONENTIAL, PTT_GENERALIZED = 0, 1, 2
LIMITER = {"upwind": 0, "cubista": 1, "minmod": 2, "smart": 3, "waceb": 4, "superbee": 5, "none": 6}
DDT_EULER, DDT_BACKWARD, DDT_CRANK_NICOLSON, DDT_STEADY_STATE = 0, 1, 2, 3
THERMO = {"Constant": 0, "Arrhenius": 1, "ArrheniusModified": 2, "WLF": 3, "VFT": 4}   # thermoFunctions/* type names
SOLVER = {"PBiCGStab": 0, "PBiCG": 1}
STAB_NONE, STAB_BSD, STAB_COUPLING = range(3)   # constitutiveProperties `stabilization`
FIELD_THETA, FIELD_TAU, FIELD_EIGVALS, FIELD_EIGVECS, FIELD_THETA_B, FIELD_TAU_B, FIELD_TAU_TOTAL, FIELD_THETA_OLD, FIELD_TAU_B_TOTAL, FIELD_FLUIDITY, FIELD_FLUIDITY_B = range(11)
FLOW_CONTRACTION_2D, FLOW_VORTEX, FLOW_CONTRACTION_3D = 0, 1, 2

MODEL_NAMES = {
    "Oldroyd-BLog": MODEL_OLDROYD_B_LOG,
    "GiesekusLog": MODEL_GIESEKUS_LOG,
    "PTTLog": MODEL_PTT_LOG,
    "FENE-PLog": MODEL_FENE_P_LOG,
    "FENE-CRLog": MODEL_FENE_CR_LOG,
    "WhiteMetznerCYLog": MODEL_WM_CY_LOG,
    "Rolie-PolyLog": MODEL_ROLIE_POLY_LOG,
    "XPomPomLog": MODEL_XPOMPOM_LOG,
    "SaramitoLog": MODEL_SARAMITO_LOG,
    "BMPLog": MODEL_BMP_LOG,
}


class RheoPatchDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("start", C.c_int32), ("size", C.c_int32), ("nbr_rank", C.c_int32),
                ("theta_bc", C.c_int32), ("tau_bc", C.c_int32)]


class RheoMeshDesc(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_faces", C.c_int32), ("n_internal_faces", C.c_int32), ("n_patches", C.c_int32),
                ("owner", C.POINTER(C.c_int32)), ("neighbour", C.POINTER(C.c_int32)),
                ("Sf", C.POINTER(C.c_double)), ("Cf", C.POINTER(C.c_double)), ("C", C.POINTER(C.c_double)),
                ("V", C.POINTER(C.c_double)), ("weights", C.POINTER(C.c_double)), ("nbr_C", C.POINTER(C.c_double)),
                ("patches", C.POINTER(RheoPatchDesc)), ("solved_components", C.c_int32 * 6)]


class RheoPatchRule(C.Structure):
    _fields_ = [("patch", C.c_int32), ("lo", C.c_double * 3), ("hi", C.c_double * 3)]


class RheoPatchSpec(C.Structure):
    _fields_ = [("type", C.c_int32), ("theta_bc", C.c_int32), ("tau_bc", C.c_int32)]


class RheoSynthSpec(C.Structure):
    _fields_ = [("flow", C.c_int32), ("amplitude", C.c_double), ("h_up", C.c_double), ("h_down", C.c_double),
                ("x_ramp", C.c_double), ("theta_amp", C.c_double), ("noise", C.c_double), ("seed", C.c_uint64)]


class RheoModelDesc(C.Structure):
    _fields_ = [("model", C.c_int32), ("rho", C.c_double), ("etaS", C.c_double), ("etaP", C.c_double),
                ("lambda_", C.c_double), ("alpha", C.c_double), ("epsilon", C.c_double), ("zeta", C.c_double),
                ("ptt_function", C.c_int32), ("ml_alpha", C.c_double), ("ml_beta", C.c_double), ("ml_rtol", C.c_double),
                ("ml_max_iter", C.c_int32), ("L2", C.c_double),
                ("wm_K", C.c_double), ("wm_n", C.c_double), ("wm_a", C.c_double),
                ("rp_lambdaR", C.c_double), ("rp_beta", C.c_double), ("rp_delta", C.c_double), ("rp_chiMax", C.c_double),
                ("xpp_lambdaS", C.c_double), ("xpp_q", C.c_double), ("xpp_n", C.c_double),
                ("sar_tau0", C.c_double), ("sar_k", C.c_double), ("sar_n", C.c_double), ("sar_dims", C.c_double * 3),
                ("sar_ptt", C.c_int32),
                ("bmp_G0", C.c_double), ("bmp_k", C.c_double), ("bmp_Phi0", C.c_double), ("bmp_PhiInf", C.c_double), ("bmp_relax", C.c_double)]


class RheoSchemeCtl(C.Structure):
    _fields_ = [("limiter", C.c_int32), ("ddt", C.c_int32), ("solver", C.c_int32), ("tolerance", C.c_double),
                ("rel_tol", C.c_double), ("min_iter", C.c_int32), ("max_iter", C.c_int32), ("relax", C.c_double),
                ("cn_psi", C.c_double), ("bounded", C.c_int32), ("pad_", C.c_int32)]


class RheoStepStats(C.Structure):
    _fields_ = [("initial_residual", C.c_double * 6), ("final_residual", C.c_double * 6),
                ("n_iterations", C.c_int32 * 6), ("converged", C.c_int32 * 6)]


# every symbol include/*.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_I = C.c_int32
_D = C.c_double
MESH_SYMBOLS = {
    "rheo_mesh_tensor_grid": (_P, [_I, _I, _I, _P, _P, _P, _I, _P, _I, _P, _I, _P, _I, _D, _I]),
    "rheo_mesh_tensor_grid_part": (_P, [_I, _I, _I, _P, _P, _P, _I, _P, _I, _P, _I, _P, _I, _D, _I, _I, _I, _I, _I]),
    "rheo_mesh_from_desc": (_P, [_P]),
    "rheo_mesh_free": (None, [_P]),
    "rheo_mesh_desc": (C.c_int, [_P, _P]),
    "rheo_mesh_n_boundary_faces": (C.c_int, [_P]),
    "rheo_mesh_simple_decomp": (C.c_int, [_P, _I, _I, _I, _D, _P]),
    "rheo_mesh_decompose": (_P, [_P, _P, _I, _I]),
    "rheo_mesh_proc_addressing": (C.c_int, [_P, _P, _P]),
    "rheo_mesh_colour_renumber": (C.c_int, [_P, _P, _P, _P]),
    "rheo_mesh_block_renumber": (C.c_int, [_P, _P, _P, _P]),
    "rheo_synth_fields": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "rheo_thermo_factor": (C.c_int, [_I, _P, C.c_int64, _P, _P]),
    "rheo_mesh_max_courant_rate": (_D, [_P, _P]),
    "rheo_mesh_last_error": (C.c_char_p, []),
}
GPU_SYMBOLS = {
    "rheo_gpu_device_count": (C.c_int, []),
    "rheo_gpu_create": (C.c_int, [_P, _P, _I, _P, _I, _P]),
    "rheo_gpu_destroy": (None, [_P]),
    "rheo_gpu_nccl_unique_id": (C.c_int, [_P]),
    "rheo_gpu_comm_init": (C.c_int, [_P, _I, _I, _P]),
    "rheo_gpu_upload_state": (C.c_int, [_P, _I, _P, _P, _P, _P, _P, _P]),
    "rheo_gpu_upload_velocity": (C.c_int, [_P, _P, _P, _P]),
    "rheo_gpu_upload_thermo": (C.c_int, [_P, _I, _P, _P]),
    "rheo_gpu_store_old_time": (C.c_int, [_P]),
    "rheo_gpu_set_tau_assignment": (C.c_int, [_P, C.c_int32]),
    "rheo_gpu_step": (C.c_int, [_P, _D, _P]),
    "rheo_gpu_download": (C.c_int, [_P, _I, _I, _P]),
    "rheo_gpu_correct": (C.c_int, [_P, _P, _P, _P, _D, _I, _P, _P, _P]),
    "rheo_gpu_div_tau": (C.c_int, [_P, _I, _P]),
    "rheo_gpu_upload_fluidity": (C.c_int, [_P, _I, _P, _P]),
    "rheo_gpu_upload_grad_u": (C.c_int, [_P, _P]),
    "rheo_gpu_get_renumbering": (C.c_int, [_P, _P, _P, _P]),
    "rheo_gpu_get_ell": (C.c_int, [_P, _P, _P, _P]),
    "rheo_gpu_get_ordering": (C.c_int, [_P, _P, _I]),
    "rheo_gpu_get_levels": (C.c_int, [_P, _P, _P]),
    "rheo_gpu_launch_count": (C.c_int64, [_P]),
    "rheo_gpu_last_iterations": (C.c_int, [_P]),
    "rheo_gpu_transfer_bytes": (C.c_int, [_P, _P, _P]),
    "rheo_gpu_comm_stats": (C.c_int, [_P, _P, _P, _P]),
    "rheo_gpu_set_phase_timing": (C.c_int, [_P, _I]),
    "rheo_gpu_get_phase_times": (C.c_int, [_P, _P]),
    "rheo_gpu_set_kernel_timing": (C.c_int, [_P, _I]),
    "rheo_gpu_get_kernel_times": (C.c_int, [_P, _P, _I]),
    "rheo_gpu_stream": (C.c_int, [_P, _P]),
    "rheo_gpu_synchronize": (C.c_int, [_P]),
    "rheo_gpu_eig_exp": (C.c_int, [_I, _I, _P, _P, _P]),
    "rheo_gpu_last_error": (C.c_char_p, []),
    "rheo_gpu_abi_sizes": (C.c_int, [_P]),
}

IO_SYMBOLS = {
    "rheo_io_read_polymesh": (_P, [C.c_char_p]),
    "rheo_io_write_polymesh": (C.c_int, [_P, C.c_char_p, _I]),
    "rheo_io_mesh_counts": (C.c_int, [_P, _P, _P]),
    "rheo_io_set_nbr_centres": (C.c_int, [_P, _I, _P]),
    "rheo_io_patch_name": (C.c_int, [_P, _I, _P, _I]),
    "rheo_io_set_patch_name": (C.c_int, [_P, _I, C.c_char_p]),
    "rheo_io_read_field": (_P, [C.c_char_p]),
    "rheo_io_field_free": (None, [_P]),
    "rheo_io_field_info": (C.c_int, [_P, _P, _I, _P, _I, _P, _P, _P]),
    "rheo_io_field_internal": (C.c_int, [_P, C.c_int64, _P]),
    "rheo_io_field_patch": (C.c_int, [_P, C.c_char_p, _I, _P, _I, _P, _P]),
    "rheo_io_apply_field_bcs": (C.c_int, [_P, _P, _I]),
    "rheo_io_dict_open": (_P, [C.c_char_p]),
    "rheo_io_dict_free": (None, [_P]),
    "rheo_io_dict_lookup": (C.c_int, [_P, C.c_char_p, _P, _I]),
    "rheo_io_write_field": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, _I, C.c_int64, _P, _I, _P, _P, _P, _P, _I]),
}

_gpu = None
_host = None


def _bind(L, table):
    for name, (res, args) in table.items():
        fn = getattr(L, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return L


def host_lib() -> C.CDLL:
    """librheo_host.so: mesh generator, decomposition, synthetic fields, OpenFOAM file formats.  No CUDA."""
    global _host
    if _host is None:
        if not _HOST_LIB_PATH.exists():
            raise RuntimeError(f"{_HOST_LIB_PATH} is missing: run `python -m rheotool_b200.build` (or __graft_entry__.build()).")
        _host = _bind(C.CDLL(str(_HOST_LIB_PATH)), {**MESH_SYMBOLS, **IO_SYMBOLS})
    return _host


def gpu_lib() -> C.CDLL:
    """librheo_b200.so: the CUDA stress step (include/rheo_gpu.h)."""
    global _gpu
    if _gpu is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} is missing: run `python -m rheotool_b200.build` (or __graft_entry__.build()). "
                "There is no CPU fallback for the stress step.")
        _gpu = _bind(C.CDLL(str(_LIB_PATH)), GPU_SYMBOLS)
    return _gpu


class _Libs:
    """One namespace over both libraries; the CUDA library is loaded on the first use of a rheo_gpu_* entry point, so a
    process that only builds meshes or reads files (bench.py --impl reference, the oracle-only tests) never maps it."""

    def __getattr__(self, name):
        if name in GPU_SYMBOLS:
            return getattr(gpu_lib(), name)
        return getattr(host_lib(), name)


_libs = _Libs()


def lib() -> _Libs:
    return _libs


def lib_path() -> Path:
    return _LIB_PATH


def host_lib_path() -> Path:
    return _HOST_LIB_PATH
