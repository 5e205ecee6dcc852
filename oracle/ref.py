"""ctypes front-end of oracle/_ref/libref_stress.so — rheoTool's OWN stress-step text compiled from
/root/reference over a minimal OpenFOAM stand-in (oracle/ref_shim/).  TEST INFRASTRUCTURE ONLY.

Used (a) by tools/make_golden_reference.py, which runs it in the container that has /root/reference and commits
its outputs as fixtures under tests/golden/, and (b) directly by tests when the built library is present.
Never imported by the product package.  /root/reference is only needed to BUILD the library (`make -C oracle ref`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "_ref" / "libref_stress.so"
REFERENCE_ROOT = Path(os.environ.get("RHEO_REFERENCE_ROOT", "/root/reference"))
_lib = None


def can_build() -> bool:
    return (REFERENCE_ROOT / "of90" / "src" / "libs" / "gaussDefCmpwConvectionScheme" / "limiters.H").exists()


def build() -> Path | None:
    """Compile the reference text (only possible where /root/reference exists); returns the library path or None."""
    if can_build():
        subprocess.run(["make", "-s", "-C", str(_HERE), "ref", f"REF={REFERENCE_ROOT}/of90/src/libs"], check=True)
    return _LIB if _LIB.exists() else None


def available() -> bool:
    return _LIB.exists() or can_build()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB.exists() and build() is None:
            raise RuntimeError("oracle/_ref/libref_stress.so is not built and /root/reference is not present")
        L = C.CDLL(str(_LIB))
        P, I, D = C.c_void_p, C.c_int, C.c_double
        L.ref_jacobi.restype, L.ref_jacobi.argtypes = None, [I, P, P, P, P]
        L.ref_lims.restype, L.ref_lims.argtypes = I, [I, P, P, P]
        L.ref_decompose_gradU.restype, L.ref_decompose_gradU.argtypes = None, [I, P, P, P, P, P]
        L.ref_innerP.restype, L.ref_innerP.argtypes = None, [I, P, P, I, P]
        L.ref_correct.restype, L.ref_correct.argtypes = I, [P, P, I, D, I] + [P] * 15
        L.ref_correct_multi.restype, L.ref_correct_multi.argtypes = I, [I, P, P, I, D, I] + [P] * 9
        L.ref_set_fluidity.restype, L.ref_set_fluidity.argtypes = None, [P, P, I]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def jacobi(theta6):
    """utils/jacobi.H: (exp(eigenvalues) [n,3] in jacobi's own order, eigenvector tensors [n,3,3], rotations [n])."""
    t = _f(theta6, (-1, 6))
    D = np.zeros((len(t), 3)); V = np.zeros((len(t), 9)); nrot = np.zeros(len(t), dtype=np.int32)
    lib().ref_jacobi(len(t), _p(t), _p(D), _p(V), _p(nrot))
    return D, V.reshape(-1, 3, 3), nrot


def lims(limiter: int):
    """limiters.H rows of one limiter: (alpha, beta, bounds) as the reference appends them."""
    a = np.full(3, np.nan); b = np.full(3, np.nan); bo = np.full(2, np.nan)
    n = lib().ref_lims(int(limiter), _p(a), _p(b), _p(bo))
    return a[:n], b[:n], bo[:max(0, n - 1)] if n > 1 else bo[:1]


def decompose_gradU(M9, eigvals9, eigvecs9):
    M = _f(M9, (-1, 9)); va = _f(eigvals9, (-1, 9)); ve = _f(eigvecs9, (-1, 9))
    om = np.zeros_like(M); B = np.zeros_like(M)
    lib().ref_decompose_gradU(len(M), _p(M), _p(va), _p(ve), _p(om), _p(B))
    return om, B


def innerP(t1, t2, is_first_T: bool):
    a = _f(t1, (-1, 9)); b = _f(t2, (-1, 9))
    out = np.zeros_like(a)
    lib().ref_innerP(len(a), _p(a), _p(b), 1 if is_first_T else 0, _p(out))
    return out


def correct(mesh_desc, model_desc, limiter: int, dt: float, U, Ub, phi, theta, theta_b, tau, tau_b, eigvals, eigvecs,
            use_regression=False, want_matrix=False, fluidity=None, fluidity_b=None, solve_fluidity=False):
    """One XxxLog::correct() of the reference.  Returns a dict of the state after the call (+ the assembled thetaEqn).
    BMPLog, solve_fluidity False: `fluidity` = Phi per cell AFTER PhiEqn.solve() — BMPLog::correct without its fluidity equation
    (BMPLog.C:168-199: the theta equation and theta -> tau).  solve_fluidity True: `fluidity`, `fluidity_b` = Phi BEFORE the call;
    the whole BMPLog::correct runs (BMPLog.C:142-201, PhiEqn as a scalar fvMatrix of the stand-in types) and the state returned
    holds the new "fluidity" / "fluidity_b"."""
    n, nb = mesh_desc.n_cells, mesh_desc.n_faces - mesh_desc.n_internal_faces
    st = {
        "theta": _f(theta, (n, 6)).copy(), "theta_b": _f(theta_b, (nb, 6)).copy(),
        "tau": _f(tau, (n, 6)).copy(), "tau_b": _f(tau_b, (nb, 6)).copy(),
        "eigvals": _f(eigvals, (n, 9)).copy(), "eigvecs": _f(eigvecs, (n, 9)).copy(),
    }
    mats = {}
    if want_matrix:
        nif = mesh_desc.n_internal_faces
        mats = {"lower": np.zeros(nif), "upper": np.zeros(nif), "diag": np.zeros(n), "source": np.zeros((n, 6)),
                "internalCoeffs": np.zeros((nb, 6)), "boundaryCoeffs": np.zeros((nb, 6))}
    U, Ub, phi = _f(U), _f(Ub), _f(phi)
    fl = None if fluidity is None else _f(fluidity).copy()
    flb = None if fluidity_b is None else _f(fluidity_b).copy()
    if solve_fluidity and flb is None:
        flb = np.zeros(nb)
    lib().ref_set_fluidity(_p(fl), _p(flb), 1 if solve_fluidity else 0)
    rc = lib().ref_correct(C.cast(C.byref(mesh_desc), C.c_void_p), C.cast(C.byref(model_desc), C.c_void_p), int(limiter),
                           float(dt), 1 if use_regression else 0, _p(U), _p(Ub), _p(phi),
                           _p(st["theta"]), _p(st["theta_b"]), _p(st["tau"]), _p(st["tau_b"]), _p(st["eigvals"]), _p(st["eigvecs"]),
                           _p(mats.get("lower")), _p(mats.get("upper")), _p(mats.get("diag")), _p(mats.get("source")),
                           _p(mats.get("internalCoeffs")), _p(mats.get("boundaryCoeffs")))
    lib().ref_set_fluidity(None, None, 0)
    if fl is not None:
        st["fluidity"], st["fluidity_b"] = fl, flb
    if rc:
        raise RuntimeError("the reference harness holds no correct() text for this model")
    st.update(mats)
    return st


def correct_multi(mesh_descs, model_desc, limiter: int, dt: float, states, use_regression=False):
    """One XxxLog::correct() of the reference on R sub-domain meshes with processor patches (one thread per rank).
    `states`: per rank a dict with U, Ub, phi, theta, theta_b, tau, tau_b, eigvals, eigvecs; returns the per-rank states after."""
    R = len(mesh_descs)
    keys_in = ("U", "Ub", "phi")
    keys_io = ("theta", "theta_b", "tau", "tau_b", "eigvals", "eigvecs")
    keep = [{k: _f(st[k]).copy() for k in keys_in + keys_io} for st in states]
    ptrs = {k: (C.c_void_p * R)(*[a[k].ctypes.data for a in keep]) for k in keys_in + keys_io}
    descs = (C.c_void_p * R)(*[C.cast(C.byref(d), C.c_void_p).value for d in mesh_descs])
    rc = lib().ref_correct_multi(R, C.cast(descs, C.c_void_p), C.cast(C.byref(model_desc), C.c_void_p), int(limiter), float(dt),
                                 1 if use_regression else 0, *[C.cast(ptrs[k], C.c_void_p) for k in keys_in + keys_io])
    if rc:
        raise RuntimeError("the reference harness holds no correct() text for this model")
    out = []
    for r, d in enumerate(mesh_descs):
        n, nb = d.n_cells, d.n_faces - d.n_internal_faces
        shapes = {"theta": (n, 6), "theta_b": (nb, 6), "tau": (n, 6), "tau_b": (nb, 6), "eigvals": (n, 9), "eigvecs": (n, 9)}
        out.append({k: keep[r][k].reshape(shapes[k]) for k in keys_io})
    return out
