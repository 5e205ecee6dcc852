"""Python mirror of the reference's plugin surface for the stress step, on top of the C-ABI.

`GpuStressModel` plays the role of one `Foam::constitutiveEq` object (constitutiveEq.H:62-351): it is
constructed from (mesh, U/phi providers, dictionary-like model description), owns theta/tau/eigVals/
eigVecs on the device, and exposes `correct()` and `tau()`.  `constitutive_model()` mirrors
`constitutiveModel` + `constitutiveEq::New` (constitutiveModel.C:47-58, newConstitutiveEq.C:32-61):
it selects by the `type` keyword, including `multiMode` (multiMode.C:73-92).

Everything here calls librheo_b200.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .cases import model_desc, scheme_ctl
from .mesh import HostMesh


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class RheoError(RuntimeError):
    """The C++ shim turns a non-zero status into FatalError; here it is an exception."""


def _check(rc):
    if rc:
        raise RheoError(abi.lib().rheo_gpu_last_error().decode())


class GpuStressModel:
    def __init__(self, mesh: HostMesh, models, schemes: abi.RheoSchemeCtl, device: int = 0):
        self.mesh = mesh
        self.n_modes = len(models)
        self._models = (abi.RheoModelDesc * len(models))(*models)
        self._schemes = schemes
        self._h = C.c_void_p()
        _check(abi.lib().rheo_gpu_create(mesh.desc_ptr(), C.cast(self._models, C.c_void_p), len(models), C.byref(schemes), device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h:
            abi.lib().rheo_gpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU ----
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(abi.lib().rheo_gpu_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, n_ranks: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        _check(abi.lib().rheo_gpu_comm_init(self._h, rank, n_ranks, buf))

    # ---- state ----
    def upload_state(self, mode=0, theta=None, tau=None, eigvals=None, eigvecs=None, theta_b=None, tau_b=None):
        arrs = [_c(a) for a in (theta, tau, eigvals, eigvecs, theta_b, tau_b)]
        _check(abi.lib().rheo_gpu_upload_state(self._h, mode, *[_p(a) for a in arrs]))

    def upload_grad_u(self, gradU=None):
        """correct(alpha, gradU) with a caller-supplied velocity gradient (boilerLog.H:1; filmModel.C:408): [n_cells, 9] in OpenFOAM's
        tensor order, used instead of fvc::grad(U) from now on; None returns to the device's own gradient."""
        a = _c(gradU)
        _check(abi.lib().rheo_gpu_upload_grad_u(self._h, _p(a)))

    def upload_fluidity(self, mode=0, Phi=None, Phi_b=None):
        """BMPLog: the fluidity field Phi (MUST_READ, BMPLog.C:112-122), cell values and (fixedValue patches) boundary values."""
        a, b = _c(Phi), _c(Phi_b)
        _check(abi.lib().rheo_gpu_upload_fluidity(self._h, mode, _p(a), _p(b)))

    def fluidity(self, mode=0) -> np.ndarray:
        out = np.zeros(self.mesh.n_cells)
        _check(abi.lib().rheo_gpu_download(self._h, mode, abi.FIELD_FLUIDITY, _p(out)))
        return out

    def fluidity_b(self, mode=0) -> np.ndarray:
        out = np.zeros(self.mesh.n_boundary)
        _check(abi.lib().rheo_gpu_download(self._h, mode, abi.FIELD_FLUIDITY_B, _p(out)))
        return out

    def upload_velocity(self, U, U_b, phi):
        U, U_b, phi = _c(U), _c(U_b), _c(phi)
        _check(abi.lib().rheo_gpu_upload_velocity(self._h, _p(U), _p(U_b), _p(phi)))
        self.synchronize()

    def upload_velocity_ptrs(self, U_ptr, Ub_ptr, phi_ptr):
        """Raw host pointers (e.g. pinned torch tensors); asynchronous on the model's stream."""
        _check(abi.lib().rheo_gpu_upload_velocity(self._h, U_ptr, Ub_ptr, phi_ptr))

    def upload_thermo(self, mode=0, lambda_cell=None, etaP_cell=None):
        """Thermo-dependent lambda / etaP per cell (Oldroyd_BLog.C:133-135: thermoLambdaPtr_->createField(lambda_)); None, None
        restores the scalars.  `thermo_factor` evaluates the reference's thermoFunctions."""
        a, b = _c(lambda_cell), _c(etaP_cell)
        _check(abi.lib().rheo_gpu_upload_thermo(self._h, mode, _p(a), _p(b)))

    def store_old_time(self):
        _check(abi.lib().rheo_gpu_store_old_time(self._h))

    def set_tau_assignment(self, on: bool):
        """Alternative reading of `tau_ = ...` before tau_.correctBoundaryConditions() (include/rheo_gpu.h; off by default)."""
        _check(abi.lib().rheo_gpu_set_tau_assignment(self._h, 1 if on else 0))

    def correct(self, dt: float, want_stats: bool = False):
        """constitutiveEq::correct() with U/phi already resident on the device."""
        stats = (abi.RheoStepStats * self.n_modes)() if want_stats else None
        _check(abi.lib().rheo_gpu_step(self._h, float(dt), None if stats is None else C.cast(stats, C.c_void_p)))
        return stats

    def correct_host(self, U_ptr, Ub_ptr, phi_ptr, dt, new_time_step, tau_out_ptr, tau_b_out_ptr=None):
        """upload + correct + tau download through the one-call C entry point (host buffers)."""
        _check(abi.lib().rheo_gpu_correct(self._h, U_ptr, Ub_ptr, phi_ptr, float(dt), 1 if new_time_step else 0, tau_out_ptr, tau_b_out_ptr, None))

    def div_tau(self, stabilization: int = abi.STAB_COUPLING, out: np.ndarray | None = None) -> np.ndarray:
        """Explicit part of constitutiveEq::divTau(U) (constitutiveEq.C:72-132; multiMode.C:143-157), evaluated on the device:
        sum over modes of fvc::div(tau/rho) [- fvc::div((etaP/rho) grad(U)) with stabilization coupling], 3 per cell."""
        if out is None:
            out = np.zeros((self.mesh.n_cells, 3))
        _check(abi.lib().rheo_gpu_div_tau(self._h, int(stabilization), _p(out)))
        return out

    def div_tau_host(self, stabilization: int, out_ptr):
        """div_tau into a caller-owned (pinned) host buffer."""
        _check(abi.lib().rheo_gpu_div_tau(self._h, int(stabilization), out_ptr))

    def download(self, field: int, mode: int = 0) -> np.ndarray:
        n = self.mesh.n_boundary if field in (abi.FIELD_THETA_B, abi.FIELD_TAU_B, abi.FIELD_TAU_B_TOTAL) else self.mesh.n_cells
        w = 9 if field in (abi.FIELD_EIGVALS, abi.FIELD_EIGVECS) else 6
        out = np.zeros((n, w))
        _check(abi.lib().rheo_gpu_download(self._h, mode, field, _p(out)))
        return out

    def theta(self, mode=0):
        return self.download(abi.FIELD_THETA, mode)

    def tau(self, mode=None):
        """tau() of the reference: the mode's tau, or the sum over modes (multiMode::tau)."""
        return self.download(abi.FIELD_TAU_TOTAL, 0) if mode is None else self.download(abi.FIELD_TAU, mode)

    # ---- introspection ----
    def renumbering(self):
        perm = np.zeros(self.mesh.n_cells, dtype=np.int32)
        nc = C.c_int32()
        cs = np.zeros(65, dtype=np.int32)
        _check(abi.lib().rheo_gpu_get_renumbering(self._h, _p(perm), C.byref(nc), _p(cs)))
        return perm, cs[: nc.value + 1].copy()

    def ell(self):
        K = C.c_int32()
        _check(abi.lib().rheo_gpu_get_ell(self._h, C.byref(K), None, None))
        nbr = np.zeros((K.value, self.mesh.n_cells), dtype=np.int32)
        face = np.zeros((K.value, self.mesh.n_cells), dtype=np.int32)
        _check(abi.lib().rheo_gpu_get_ell(self._h, C.byref(K), _p(nbr), _p(face)))
        return nbr, face

    def ordering(self) -> str:
        """The cell ordering the DILU substitutions run in."""
        buf = C.create_string_buffer(256)
        _check(abi.lib().rheo_gpu_get_ordering(self._h, buf, len(buf)))
        return buf.value.decode()

    def levels(self):
        """(fwd, bwd) in-chunk levels per cell in device numbering (block ordering only)."""
        fwd = np.zeros(self.mesh.n_cells, dtype=np.int32)
        bwd = np.zeros(self.mesh.n_cells, dtype=np.int32)
        _check(abi.lib().rheo_gpu_get_levels(self._h, _p(fwd), _p(bwd)))
        return fwd, bwd

    def launch_count(self) -> int:
        return int(abi.lib().rheo_gpu_launch_count(self._h))

    def comm_stats(self) -> dict:
        """Multi-GPU path in use and, on the peer-memory path, the time spent waiting for the other ranks."""
        mode = C.c_int32(0)
        wait = (C.c_double * 2)()
        n = (C.c_int64 * 2)()
        _check(abi.lib().rheo_gpu_comm_stats(self._h, C.byref(mode), wait, n))
        return {"mode": {0: "single", 1: "nccl", 2: "nvlink-peer-memory"}[mode.value], "halo_wait_ms": wait[0], "reduce_wait_ms": wait[1],
                "halo_waits": int(n[0]), "reduce_waits": int(n[1])}

    def transfer_bytes(self) -> tuple[int, int]:
        """(host->device, device->host) bytes copied by this handle so far."""
        a, b = C.c_int64(0), C.c_int64(0)
        abi.lib().rheo_gpu_transfer_bytes(self._h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def last_iterations(self) -> int:
        return int(abi.lib().rheo_gpu_last_iterations(self._h))

    def set_phase_timing(self, on: bool):
        _check(abi.lib().rheo_gpu_set_phase_timing(self._h, 1 if on else 0))

    def phase_times(self):
        out = np.zeros(7)
        _check(abi.lib().rheo_gpu_get_phase_times(self._h, _p(out)))
        return dict(zip(["halo_bc", "flux_matrix", "assemble", "solve", "eig_tau", "tau_bc", "total"], out.tolist()))

    def set_kernel_timing(self, on: bool):
        _check(abi.lib().rheo_gpu_set_kernel_timing(self._h, 1 if on else 0))

    def kernel_times(self) -> dict:
        """{kernel: (launches, total_ms)} accumulated since set_kernel_timing(True)."""
        buf = C.create_string_buffer(1 << 16)
        _check(abi.lib().rheo_gpu_get_kernel_times(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(" ", 2)
            out[name.strip("()")] = (int(cnt), float(ms))
        return out

    def stream_ptr(self) -> int:
        s = C.c_void_p()
        _check(abi.lib().rheo_gpu_stream(self._h, C.byref(s)))
        return int(s.value or 0)

    def synchronize(self):
        _check(abi.lib().rheo_gpu_synchronize(self._h))


def thermo_factor(kind: str, params, T) -> np.ndarray:
    """a_T(T) of the reference's thermoFunctions (of90/src/libs/thermo/thermoFunctions): kind in Constant | Arrhenius (alpha, T0) |
    ArrheniusModified (alpha, T0) | WLF (c1, c2, T0) | VFT (A, B, T0); lambda(T) = lambda a_T, etaP(T) = etaP a_T."""
    T = np.ascontiguousarray(T, dtype=np.float64)
    p = np.zeros(3)
    p[: len(params)] = params
    out = np.zeros_like(T)
    if abi.lib().rheo_thermo_factor(abi.THERMO[kind], _p(p), T.size, _p(T), _p(out)):
        raise RheoError(f"unknown thermoFunction {kind}")
    return out


def eig_exp(theta6: np.ndarray, device: int = 0):
    """calcEig on the GPU for an AoS symmTensor array (constitutiveEq.C:360-416)."""
    t = np.ascontiguousarray(theta6, dtype=np.float64).reshape(-1, 6)
    vals = np.zeros((len(t), 9))
    vecs = np.zeros((len(t), 9))
    _check(abi.lib().rheo_gpu_eig_exp(device, len(t), _p(t), _p(vals), _p(vecs)))
    return vals, vecs


# ---- run-time selection mirror -----------------------------------------------------------------------
def models_from_dict(params: dict) -> list:
    """`parameters` dictionary of constant/constitutiveProperties -> list of RheoModelDesc.
    `type multiMode` expands `models` ( M1 {...} M2 {...} ) like multiMode.C:73-92."""
    t = params.get("type")
    if t is None:
        raise RheoError("keyword type is undefined in dictionary parameters")
    if t == "multiMode":
        out = []
        for _name, sub in params["models"]:
            out.extend(models_from_dict(sub))
        return out
    if t not in abi.MODEL_NAMES:
        raise RheoError(f"Unknown constitutiveEq type {t}\n\nValid constitutiveEq types on the GPU path are :\n{sorted(abi.MODEL_NAMES) + ['multiMode']}")
    kw = {}
    for src, dst in (("rho", "rho"), ("etaS", "etaS"), ("etaP", "etaP"), ("lambda", "lambda_"), ("alpha", "alpha"),
                     ("epsilon", "epsilon"), ("zeta", "zeta"), ("L2", "L2")):
        if src in params:
            kw[dst] = float(params[src])
    if t == "PTTLog":
        kw["ptt_function"] = params.get("destructionFunctionType", "linear")
        if kw["ptt_function"] == "generalized":
            kw["ml_alpha"], kw["ml_beta"] = float(params["alpha"]), float(params["beta"])
            kw.pop("alpha", None)
            kw["ml_rtol"] = float(params.get("rTolMittagLeffler", 1e-12))
            kw["ml_max_iter"] = int(params.get("maxIterMittagLeffler", 200))
    if t == "BMPLog":                 # BMPLog.C:129-136
        for src in ("G0", "k", "Phi0", "PhiInf"):
            kw["bmp_" + src] = float(params[src])
    if t == "SaramitoLog":            # SaramitoLog.C:108-165
        kw["sar_tau0"], kw["sar_n"] = float(params["tau0"]), float(params["n"])
        kw["sar_k"] = None if kw["sar_n"] == 1.0 else float(params["k"])
        kw["sar_dims"] = tuple(float(x) for x in params["dims"])
        if kw["sar_n"] == 1.0:
            fn = params.get("PTTfunction")
            if fn not in ("none", "linear", "exponential"):
                raise RheoError("The PTT function specified does not exist.\n\nAvailable PTT functions are:\n. none\n. linear\n. exponential")
            kw["sar_ptt"] = fn
    return [model_desc(t, **kw)]


def constitutive_model(mesh: HostMesh, constitutive_properties: dict, fv_schemes: dict | None = None,
                       fv_solution: dict | None = None, device: int = 0) -> GpuStressModel:
    """constitutiveModel constEq(U, phi): select the model(s) from `parameters.type`, the limiter from
    divSchemes `div(phi,theta)` ("GaussDefCmpw cubista"), the solver controls from fvSolution."""
    models = models_from_dict(constitutive_properties["parameters"])
    fv_schemes = fv_schemes or {}
    fv_solution = fv_solution or {}
    div = fv_schemes.get("div(phi,theta)", "GaussDefCmpw cubista").split()
    bounded = div[0] == "bounded"   # EXT-OF9 boundedConvectionScheme (rheoFilmFoam/UCM/system/fvSchemes:35)
    if bounded:
        div = div[1:]
    if div[0] != "GaussDefCmpw":
        raise RheoError(f"div(phi,theta) must use GaussDefCmpw, got {div[0]}")
    ddt_tok = str(fv_schemes.get("ddt", "Euler")).split()
    if ddt_tok[0] not in ("Euler", "backward", "CrankNicolson", "steadyState"):
        raise RheoError(f"ddtSchemes Euler, backward, CrankNicolson and steadyState are implemented on the GPU path, not {ddt_tok[0]}")
    cn_psi = float(ddt_tok[1]) if (ddt_tok[0] == "CrankNicolson" and len(ddt_tok) > 1) else 1.0
    sol = fv_solution.get("theta", {})
    ctl = scheme_ctl(limiter=div[1], solver=sol.get("solver", "PBiCGStab"), tolerance=float(sol.get("tolerance", 1e-10)),
                     rel_tol=float(sol.get("relTol", 0.0)), min_iter=int(sol.get("minIter", 0)), max_iter=int(sol.get("maxIter", 1000)),
                     relax=float(fv_solution.get("relaxationFactors", {}).get("theta", 0.0)), ddt=ddt_tok[0], cn_psi=cn_psi, bounded=bounded)
    return GpuStressModel(mesh, models, ctl, device)
