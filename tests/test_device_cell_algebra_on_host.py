"""The DEVICE per-cell source compiled for the host (tests/host/cell_algebra_host.cpp includes
rheotool_b200/csrc/gpu/cell_algebra.cuh unchanged, with only `__device__` / `__forceinline__` defined away) and checked
against the oracle — which tests/test_reference_pin.py pins on the reference's own text.  This is the very text the kernels
k_cell_source2 (model_rhs<MODEL>) and k_eig_tau (jacobi_eig, tau_from_eig<MODEL>) execute per cell, so a functor that has not
run on a GPU yet (SaramitoLog) is at least known to compute the right numbers where the same source runs on a CPU.
Not a GPU test and no substitute for one: memory layout, launch and fusion are not exercised here."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import rel_l2
from oracle import oracle as orc
from rheotool_b200 import abi, cases

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "host" / "cell_algebra_host.cpp"
HDR = ROOT / "rheotool_b200" / "csrc" / "gpu" / "cell_algebra.cuh"
LIB = ROOT / "build" / "host" / "libcell_algebra_host.so"


@pytest.fixture(scope="module")
def hca():
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, HDR.stat().st_mtime, (ROOT / "include" / "rheo_gpu.h").stat().st_mtime):
        LIB.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run(["/usr/bin/g++", "-O2", "-fPIC", "-shared", "-std=c++17", f"-I{ROOT / 'include'}", "-o", str(LIB), str(SRC)], check=True)
    L = C.CDLL(str(LIB))
    P, I = C.c_void_p, C.c_int
    L.hca_model_rhs.restype, L.hca_model_rhs.argtypes = None, [P, I, P, P, P, P, P, P, P]
    L.hca_tau.restype, L.hca_tau.argtypes = None, [P, I, P, P, P, P]
    L.hca_eig.restype, L.hca_eig.argtypes = None, [I, P, P, P]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


MODELS = {
    "Oldroyd-BLog": dict(),
    "GiesekusLog": dict(alpha=0.2),
    "PTTLog-linear": dict(epsilon=0.25, zeta=0.1),
    "PTTLog-exponential": dict(epsilon=0.25, zeta=0.05, ptt_function="exponential"),
    "PTTLog-generalized": dict(epsilon=0.25, zeta=0.0, ptt_function="generalized", ml_alpha=0.8, ml_beta=1.2),
    "FENE-PLog": dict(L2=100.0),
    "FENE-CRLog": dict(L2=50.0),
    "WhiteMetznerCYLog": dict(wm_K=0.5, wm_n=0.6, wm_a=1.7),
    "Rolie-PolyLog": dict(rp_lambdaR=0.05, rp_beta=0.5, rp_delta=-0.5, rp_chiMax=0.0),
    "Rolie-PolyLog-chiMax": dict(rp_lambdaR=0.2, rp_beta=0.5, rp_delta=-0.5, rp_chiMax=10.0),
    "XPomPomLog-n0": dict(alpha=0.15, xpp_lambdaS=0.04, xpp_q=3.0, xpp_n=0.0),
    "XPomPomLog-n1": dict(alpha=0.1, xpp_lambdaS=0.3, xpp_q=2.0, xpp_n=1.0),
    "SaramitoLog-n075": dict(sar_tau0=2.5, sar_k=1.5, sar_n=0.75, sar_dims=(1, 1, 0)),
    "SaramitoLog-n1-linear": dict(epsilon=0.1, zeta=0.1, sar_tau0=1.0, sar_n=1.0, sar_ptt="linear"),
    "SaramitoLog-n1-exponential": dict(epsilon=0.1, zeta=0.05, sar_tau0=1.0, sar_n=1.0, sar_ptt="exponential"),
    "SaramitoLog-n1-none": dict(sar_tau0=1.0, sar_n=1.0, sar_ptt="none"),
}


def _inputs(n=4000, seed=5):
    rng = np.random.default_rng(seed)
    th = rng.standard_normal((n, 6)) * np.array([0.6, 0.3, 0.2, 0.6, 0.25, 0.6])
    th[:50] = 0.0
    th[50:100, [2, 4]] = 0.0                       # 2-D tensors
    L = rng.standard_normal((n, 9))
    tau = rng.standard_normal((n, 6)) * 3.0
    tau[:200] *= 0.05                              # below the yield stress
    return th, L, tau


def test_device_jacobi_eig_source_matches_the_oracle(hca):
    th, _, _ = _inputs()
    d = np.zeros((len(th), 3)); V = np.zeros((len(th), 9))
    hca.hca_eig(len(th), _p(th), _p(d), _p(V))
    vals, vecs = orc.calc_eig(th)                   # sorted ascending like the device
    assert np.abs(np.exp(d) - vals[:, [0, 4, 8]]).max() <= 1e-13 * np.abs(vals).max()
    R = V.reshape(-1, 3, 3)
    A = R @ (np.exp(d)[:, :, None] * np.transpose(R, (0, 2, 1)))
    Ro = vecs.reshape(-1, 3, 3)
    Ao = Ro @ vals.reshape(-1, 3, 3) @ np.transpose(Ro, (0, 2, 1))
    assert rel_l2(A, Ao) <= 1e-13
    assert np.abs(R @ np.transpose(R, (0, 2, 1)) - np.eye(3)).max() <= 1e-13


@pytest.mark.parametrize("name", sorted(MODELS))
def test_device_model_source_and_stress_map_match_the_oracle(hca, name):
    kw = MODELS[name]
    md = cases.model_desc(name.split("-n")[0] if name.startswith(("XPomPom", "Saramito")) else name.replace("-linear", "").replace("-exponential", "").replace("-generalized", "").replace("-chiMax", ""),
                          rho=1.0, etaS=0.05, etaP=0.95, lambda_=0.3, **kw)
    th, L, tau = _inputs()
    vals, vecs = orc.calc_eig(th)
    lam3 = np.ascontiguousarray(vals[:, [0, 4, 8]])
    rhs = np.zeros_like(th); f = np.zeros(len(th))
    hca.hca_model_rhs(C.byref(md), len(th), _p(L), _p(th), _p(vecs), _p(lam3), _p(tau), _p(rhs), _p(f))
    rhs_o, f_o = orc.model_rhs(md, L, th, vecs, vals, tau6=tau)
    # cells with (nearly) equal eigenvalues amplify round-off by 1/gap in Omega (constitutiveEq.C:339-341: /(gap + 1e-16))
    gap = np.minimum(np.abs(np.diff(lam3, axis=1)).min(axis=1), np.abs(lam3[:, 2] - lam3[:, 0]))
    ok = gap > 1e-6
    assert ok.sum() > 3500
    assert rel_l2(rhs[ok], rhs_o[ok]) <= 1e-11, name
    assert np.abs(f - f_o).max() <= 1e-12 * max(1.0, np.abs(f_o).max())
    t6 = np.zeros_like(th)
    hca.hca_tau(C.byref(md), len(th), _p(vecs), _p(lam3), _p(f_o), _p(t6))
    assert rel_l2(t6, orc.tau_from_eig(md, vecs, vals, f_o)) <= 1e-13, name
    if name.startswith("SaramitoLog"):              # the yield switch is exercised on both sides
        r0, _ = orc.model_rhs(md, L, th, vecs, vals, tau6=np.zeros_like(tau))
        assert rel_l2(rhs_o[ok], r0[ok]) > 1e-3
