"""The cases on which the oracle (and the GPU path) are pinned against rheoTool's OWN stress-step text
(oracle/_ref, built by `make -C oracle ref` where /root/reference exists).  Shared by
tools/make_golden_reference.py (which writes tests/golden/reference_correct.npz) and tests/test_reference_pin.py.
Small meshes on purpose: the fixture travels with the repo."""
from __future__ import annotations

import hashlib

import numpy as np

from rheotool_b200 import abi, cases


def _with(spec, models, limiter):
    spec.models = models
    spec.schemes = cases.scheme_ctl(limiter, "PBiCGStab", 1e-15, relax=0.0)
    return spec


def _fixed_theta_walls(spec):
    """theta fixedValue (0) on every wall instead of zeroGradient.  superbee's first row (alpha, beta) = (1/2, 1/2) does
    not vanish at phi~ = 0 (limiters.H:77-82), and a wall-adjacent cell with zeroGradient theta and flow away from the wall
    has phi~ = 1 - (tN - tP)/(2 grad.d + 1e-18) = 0 up to the rounding of its Gauss gradient: the reference algorithm itself
    picks `upwind` or `row 0` by the last bit there (measured: ~40 faces of a 729-cell cavity differ between two builds of
    the same text), so that combination is not a parity case for any implementation."""
    import copy
    spec.grid.patches = [copy.copy(p) for p in spec.grid.patches]   # cases.py shares its inlet / outlet PatchSpec objects
    for p in spec.grid.patches:
        if p.theta_bc == abi.BC_ZERO_GRADIENT:
            p.theta_bc = abi.BC_FIXED_VALUE
    return spec


def _regression_walls(spec):
    """linearExtrapolation walls with `useRegression true` (linearExtrapolationFvPatchField.C:72,152-219)"""
    import copy
    spec.grid.patches = [copy.copy(p) for p in spec.grid.patches]
    for p in spec.grid.patches:
        if p.tau_bc == abi.BC_LINEAR_EXTRAPOLATION:
            p.tau_bc = abi.BC_LINEAR_EXTRAPOLATION_REG
    return spec


# name -> (spec factory, limiter).  Every model whose correct() the reference harness compiles, every limiter row,
# 2-D (empty patches, fixedValue inlet, zeroGradient outlet, linearExtrapolation walls) and 3-D meshes.
REFERENCE_CASES = {
    "OldroydBLog-2D-cubista": lambda: _with(cases.channel_2d(30, 12), [cases.model_desc("Oldroyd-BLog", 1.0, 0.59, 0.41, 0.7)], "cubista"),
    "PTTLog-linear-zeta-2D-minmod": lambda: _with(cases.channel_2d(30, 12), [cases.model_desc("PTTLog", 1.0, 0.1, 0.9, 0.6, epsilon=0.25, zeta=0.1)], "minmod"),
    "PTTLog-exponential-2D-smart": lambda: _with(cases.channel_2d(24, 10), [cases.model_desc("PTTLog", 1.0, 0.1, 0.9, 0.6, epsilon=0.25, zeta=0.05, ptt_function="exponential")], "smart"),
    "PTTLog-generalized-2D-waceb": lambda: _with(cases.channel_2d(24, 10), [cases.model_desc("PTTLog", 1.0, 0.1, 0.9, 0.6, epsilon=0.25, zeta=0.0, ptt_function="generalized", ml_alpha=0.8, ml_beta=1.2)], "waceb"),
    "GiesekusLog-3D-contraction-cubista": lambda: _with(cases.by_name("C3", 1 / 19), [cases.model_desc("GiesekusLog", 1.0, 0.01, 0.99, 0.1, alpha=0.2)], "cubista"),
    "FENEPLog-3D-cavity-cubista": lambda: _with(cases.cube(9, "cavity", 1, "FENE-PLog"), [cases.model_desc("FENE-PLog", 1.0, 0.01, 0.99, 0.1, L2=100.0)], "cubista"),
    "FENEPLog-3D-box-superbee": lambda: _with(_fixed_theta_walls(cases.cube(8, "box", 1, "FENE-PLog")), [cases.model_desc("FENE-PLog", 1.0, 0.01, 0.99, 0.1, L2=100.0)], "superbee"),
    "FENECRLog-3D-cavity-cubista": lambda: _with(cases.cube(7, "cavity", 1, "FENE-CRLog"), [cases.model_desc("FENE-CRLog", 1.0, 0.01, 0.99, 0.1, L2=50.0)], "cubista"),
    "WhiteMetznerCYLog-2D-cubista": lambda: _with(cases.channel_2d(24, 10), [cases.model_desc("WhiteMetznerCYLog", 1.0, 0.01, 0.99, 0.1, wm_K=0.5, wm_n=0.6, wm_a=1.7)], "cubista"),
    "RoliePolyLog-3D-cavity-minmod": lambda: _with(cases.cube(7, "cavity", 1, "Rolie-PolyLog"), [cases.model_desc("Rolie-PolyLog", 1.0, 0.01, 0.99, 0.1, rp_lambdaR=0.05, rp_beta=0.5, rp_delta=-0.5, rp_chiMax=0.0)], "minmod"),
    "RoliePolyLog-chiMax-2D-cubista": lambda: _with(cases.channel_2d(24, 10), [cases.model_desc("Rolie-PolyLog", 1.0, 0.01, 0.99, 0.1, rp_lambdaR=0.2, rp_beta=0.5, rp_delta=-0.5, rp_chiMax=10.0)], "cubista"),
    "XPomPomLog-n0-3D-cavity-cubista": lambda: _with(cases.cube(7, "cavity", 1, "XPomPomLog"), [cases.model_desc("XPomPomLog", 1.0, 0.01, 0.99, 0.1, alpha=0.15, xpp_lambdaS=0.04, xpp_q=3.0, xpp_n=0.0)], "cubista"),
    "XPomPomLog-n1-2D-smart": lambda: _with(cases.channel_2d(24, 10), [cases.model_desc("XPomPomLog", 1.0, 0.01, 0.99, 0.5, alpha=0.1, xpp_lambdaS=0.3, xpp_q=2.0, xpp_n=1.0)], "smart"),
    "SaramitoLog-n075-2D-cubista": lambda: _with(cases.channel_2d(24, 10), [cases.model_desc("SaramitoLog", 1.0, 0.01, 0.99, 0.1, sar_tau0=2.5, sar_k=1.5, sar_n=0.75, sar_dims=(1, 1, 0))], "cubista"),
    "SaramitoLog-n1-linearPTT-3D-minmod": lambda: _with(cases.cube(7, "cavity", 1, "Oldroyd-BLog"), [cases.model_desc("SaramitoLog", 1.0, 0.01, 0.99, 0.1, epsilon=0.1, zeta=0.1, sar_tau0=1.0, sar_n=1.0, sar_ptt="linear")], "minmod"),
    "SaramitoLog-n1-expPTT-3D-cubista": lambda: _with(cases.cube(7, "cavity", 1, "Oldroyd-BLog"), [cases.model_desc("SaramitoLog", 1.0, 0.01, 0.99, 0.1, epsilon=0.1, zeta=0.05, sar_tau0=1.0, sar_n=1.0, sar_ptt="exponential")], "cubista"),
    "OldroydBLog-2D-cubista-regressionWalls": lambda: _with(_regression_walls(cases.channel_2d(30, 12)), [cases.model_desc("Oldroyd-BLog", 1.0, 0.59, 0.41, 0.7)], "cubista"),
    "GiesekusLog-3D-contraction-cubista-regressionWalls": lambda: _with(_regression_walls(cases.by_name("C3", 1 / 19)), [cases.model_desc("GiesekusLog", 1.0, 0.01, 0.99, 0.1, alpha=0.2)], "cubista"),
    "OldroydBLog-3D-cavity-upwind": lambda: _with(cases.cube(8, "cavity", 1, "Oldroyd-BLog"), [cases.model_desc("Oldroyd-BLog", 1.0, 0.01, 0.99, 0.1)], "upwind"),
}

def make_setup(name):
    """(spec, Setup) of a case.  SaramitoLog reads its CURRENT tau (yield criterion), so those cases start from the stress
    of the initial conformation tensor instead of tau = 0 (which would switch the relaxation term off in the first call)."""
    from helpers import Setup
    from oracle import oracle as orc
    spec = REFERENCE_CASES[name]()
    s = Setup(spec)
    if spec.models[0].model == abi.MODEL_SARAMITO_LOG:
        s.tau0 = orc.tau_from_eig(spec.models[0], s.eigvecs, s.eigvals)
    return spec, s


N_STEPS = 3   # correct() calls chained in the fixture (each is the first correct() of a new time step)
STORED_STEPS = (1, 3)          # steps whose fields the fixture holds
MATRIX_CASE = "GiesekusLog-3D-contraction-cubista"   # the case whose assembled thetaEqn the fixture holds


def digest(*arrays) -> str:
    """sha256 of the inputs of a fixture case.  The SaramitoLog cases start from a stress the ORACLE computes (make_setup): its
    last bit follows the compiler's FMA contraction choices, so callers pass that one array rounded to 9 decimals."""
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()[:16]
