"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never by the product package.  Mesh/model/scheme descriptors are the public C structs of
include/*.h (passed by pointer), so a test feeds the same descriptors to both sides.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "_build" / "liboracle.so"
_lib = None


def build() -> Path:
    subprocess.run(["make", "-s", "-C", str(_HERE)], check=True)
    return _LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB.exists():
            build()
        L = C.CDLL(str(_LIB))
        P, I, D = C.c_void_p, C.c_int, C.c_double
        sig = {
            "orc_create": (P, [I]), "orc_destroy": (None, [P]), "orc_set_mesh": (I, [P, I, P]),
            "orc_add_mode": (I, [P, P]), "orc_set_schemes": (I, [P, P]), "orc_set_sort_eig": (I, [P, I]), "orc_set_tau_assignment": (I, [P, I]),
            "orc_set_state": (I, [P, I, I, P, P, P, P, P, P]), "orc_set_velocity": (I, [P, I, P, P, P]),
            "orc_store_old_time": (I, [P]), "orc_step": (I, [P, D, P]), "orc_last_iterations": (I, [P]),
            "orc_get": (I, [P, I, I, I, P]), "orc_jacobi": (None, [I, P, P, P]),
            "orc_calc_eig": (None, [I, P, P, P, I]), "orc_decompose_gradU": (None, [I, P, P, P, D, I, P, P]),
            "orc_model_rhs": (None, [P, I, P, P, P, P, P, P]), "orc_model_rhs_tau": (None, [P, I, P, P, P, P, P, P, P]), "orc_tau": (None, [P, I, P, P, P, P]),
            "orc_gauss_grad": (I, [P, I, I, P, P, P]), "orc_div_tau": (I, [P, I, I, P]), "orc_set_fluidity": (I, [P, I, I, P, P]), "orc_set_grad_u": (I, [P, I, P]), "orc_last_error": (C.c_char_p, []), "orc_set_num_threads": (I, [I]), "orc_set_thermo": (I, [P, I, I, P, P]),
        }
        for n, (r, a) in sig.items():
            f = getattr(L, n)
            f.restype, f.argtypes = r, a
        _lib = L
    return _lib


def set_num_threads(n: int) -> int:
    """Threads of the oracle's rank loop (one sub-domain per thread); returns how many a parallel region really gets."""
    return int(lib().orc_set_num_threads(int(n)))


def _p(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.c_void_p)


FIELD_WIDTH = {0: 6, 1: 6, 2: 9, 3: 9, 4: 6, 5: 6, 6: 6, 7: 6, 9: 1, 10: 1}


class OracleCase:
    """R sub-domain meshes + modes; step() = one constitutiveEq::correct() of every mode."""

    def __init__(self, mesh_descs, models, schemes, sort_eig=True):
        L = lib()
        self._descs = list(mesh_descs)   # keep alive
        self.n_ranks = len(self._descs)
        self._h = L.orc_create(self.n_ranks)
        self.sizes = []
        for r, d in enumerate(self._descs):
            if L.orc_set_mesh(self._h, r, C.byref(d)):
                raise RuntimeError(L.orc_last_error().decode())
            self.sizes.append((d.n_cells, d.n_faces - d.n_internal_faces))
        for m in models:
            if L.orc_add_mode(self._h, C.byref(m)):
                raise RuntimeError(L.orc_last_error().decode())
        self.n_modes = len(models)
        L.orc_set_schemes(self._h, C.byref(schemes))
        L.orc_set_sort_eig(self._h, 1 if sort_eig else 0)

    def __del__(self):
        try:
            lib().orc_destroy(self._h)
        except Exception:
            pass

    def set_state(self, rank, mode, theta=None, tau=None, eigvals=None, eigvecs=None, theta_b=None, tau_b=None):
        keep = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (theta, tau, eigvals, eigvecs, theta_b, tau_b)]
        lib().orc_set_state(self._h, rank, mode, *[_p(a) for a in keep])

    def set_grad_u(self, rank, gradU=None):
        """correct(alpha, gradU): a caller-supplied velocity gradient instead of fvc::grad(U) (boilerLog.H:1); None resets"""
        a = None if gradU is None else np.ascontiguousarray(gradU, dtype=np.float64)
        lib().orc_set_grad_u(self._h, rank, _p(a))

    def set_fluidity(self, rank, mode, Phi, Phi_b=None):
        """BMPLog: the fluidity field (BMPLog.C:112-122)"""
        a = np.ascontiguousarray(Phi, dtype=np.float64)
        b = None if Phi_b is None else np.ascontiguousarray(Phi_b, dtype=np.float64)
        if lib().orc_set_fluidity(self._h, rank, mode, _p(a), _p(b)):
            raise RuntimeError(lib().orc_last_error().decode())

    def set_velocity(self, rank, U, Ub, phi):
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (U, Ub, phi)]
        lib().orc_set_velocity(self._h, rank, *[_p(a) for a in keep])

    def set_thermo(self, rank, mode, lambda_cell=None, etaP_cell=None):
        """Per-cell lambda / etaP (thermo-dependent parameters, Oldroyd_BLog.C:133-135); None restores the scalars."""
        keep = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (lambda_cell, etaP_cell)]
        if lib().orc_set_thermo(self._h, rank, mode, _p(keep[0]), _p(keep[1])):
            raise RuntimeError(lib().orc_last_error().decode())

    def set_tau_assignment(self, on: bool):
        """Alternative reading of `tau_ = ...` before tau_.correctBoundaryConditions() (oracle.cpp: Case::tauAssign)."""
        lib().orc_set_tau_assignment(self._h, 1 if on else 0)

    def store_old_time(self):
        lib().orc_store_old_time(self._h)

    def step(self, dt, stats=None):
        rc = lib().orc_step(self._h, float(dt), None if stats is None else C.cast(stats, C.c_void_p))
        if rc:
            raise RuntimeError(lib().orc_last_error().decode())

    def last_iterations(self):
        return lib().orc_last_iterations(self._h)

    def get(self, rank, mode, field):
        n, nb = self.sizes[rank]
        rows = nb if field in (4, 5, 10) else n
        out = np.zeros((rows, FIELD_WIDTH[field]))
        if lib().orc_get(self._h, rank, mode, field, _p(out)):
            raise RuntimeError(lib().orc_last_error().decode())
        return out[:, 0] if FIELD_WIDTH[field] == 1 else out

    def div_tau(self, rank, stabilization):
        """explicit part of constitutiveEq::divTau (constitutiveEq.C:72-132; multiMode.C:143-157), 3 per cell of `rank`"""
        n, _ = self.sizes[rank]
        out = np.zeros((n, 3))
        if lib().orc_div_tau(self._h, rank, int(stabilization), _p(out)):
            raise RuntimeError(lib().orc_last_error().decode())
        return out


# ---- stand-alone per-cell pieces --------------------------------------------------------------------
def jacobi(theta6):
    t = np.ascontiguousarray(theta6, dtype=np.float64).reshape(-1, 6)
    D = np.zeros((len(t), 3)); V = np.zeros((len(t), 9))
    lib().orc_jacobi(len(t), _p(t), _p(D), _p(V))
    return D, V.reshape(-1, 3, 3)


def calc_eig(theta6, sort_eig=True):
    t = np.ascontiguousarray(theta6, dtype=np.float64).reshape(-1, 6)
    vals = np.zeros((len(t), 9)); vecs = np.zeros((len(t), 9))
    lib().orc_calc_eig(len(t), _p(t), _p(vals), _p(vecs), 1 if sort_eig else 0)
    return vals, vecs


def decompose_gradU(L9, R9, Lam9, zeta=0.0, ptt=False):
    L9 = np.ascontiguousarray(L9, dtype=np.float64).reshape(-1, 9)
    R9 = np.ascontiguousarray(R9, dtype=np.float64).reshape(-1, 9)
    Lam9 = np.ascontiguousarray(Lam9, dtype=np.float64).reshape(-1, 9)
    om = np.zeros_like(L9); B = np.zeros_like(L9)
    lib().orc_decompose_gradU(len(L9), _p(L9), _p(R9), _p(Lam9), float(zeta), 1 if ptt else 0, _p(om), _p(B))
    return om, B


def model_rhs(model, L9, theta6, R9, Lam9, tau6=None):
    """tau6: the model's current stress per cell (SaramitoLog's yield criterion); None = zero."""
    L9 = np.ascontiguousarray(L9, dtype=np.float64).reshape(-1, 9)
    th = np.ascontiguousarray(theta6, dtype=np.float64).reshape(-1, 6)
    R9 = np.ascontiguousarray(R9, dtype=np.float64).reshape(-1, 9)
    Lam9 = np.ascontiguousarray(Lam9, dtype=np.float64).reshape(-1, 9)
    rhs = np.zeros_like(th); f = np.zeros(len(th))
    if tau6 is None:
        lib().orc_model_rhs(C.byref(model), len(th), _p(L9), _p(th), _p(R9), _p(Lam9), _p(rhs), _p(f))
    else:
        t6 = np.ascontiguousarray(tau6, dtype=np.float64).reshape(-1, 6)
        lib().orc_model_rhs_tau(C.byref(model), len(th), _p(L9), _p(th), _p(R9), _p(Lam9), _p(t6), _p(rhs), _p(f))
    return rhs, f


def tau_from_eig(model, R9, Lam9, f=None):
    R9 = np.ascontiguousarray(R9, dtype=np.float64).reshape(-1, 9)
    Lam9 = np.ascontiguousarray(Lam9, dtype=np.float64).reshape(-1, 9)
    tau = np.zeros((len(R9), 6))
    ff = None if f is None else np.ascontiguousarray(f, dtype=np.float64)
    lib().orc_tau(C.byref(model), len(R9), _p(R9), _p(Lam9), _p(ff), _p(tau))
    return tau
