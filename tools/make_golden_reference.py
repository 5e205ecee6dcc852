"""Writes tests/golden/reference_correct.npz, reference_unstructured.npz and reference_cell.npz: outputs of rheoTool's OWN text
for the stress step (oracle/_ref/libref_stress.so, compiled from /root/reference by `make -C oracle ref`).
Run in the container that has /root/reference:  python tools/make_golden_reference.py
The GPU box has no /root/reference; tests there (and the -m gpu parity tests) read these fixtures."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

import helpers  # noqa: E402
from oracle import ref  # noqa: E402
from reference_cases import MATRIX_CASE, N_STEPS, REFERENCE_CASES, STORED_STEPS, digest, make_setup  # noqa: E402
from rheotool_b200 import abi  # noqa: E402


def cell_vectors():
    rng = np.random.default_rng(2024)
    th = rng.standard_normal((512, 6)) * np.array([1.0, 0.5, 0.3, 1.0, 0.4, 1.0])
    th[:16] = 0.0
    th[16:32, [1, 2, 4]] = 0.0
    th[32:48] = np.array([0.3, 0, 0, 0.3, 0, -0.2])
    th[48:64, [2, 4]] = 0.0
    th[64:80] *= 1e-9
    th[80:96] *= 8.0
    return th


def main():
    assert ref.build() is not None, "needs /root/reference"
    out = {}
    for name, make in REFERENCE_CASES.items():
        spec, s = make_setup(name)
        oc = s.oracle(spec.schemes, sort_eig=False)     # only used to evaluate the initial boundary values
        st = {"theta": s.theta0, "theta_b": oc.get(0, 0, abi.FIELD_THETA_B), "tau": s.tau0, "tau_b": oc.get(0, 0, abi.FIELD_TAU_B),
              "eigvals": s.eigvals, "eigvecs": s.eigvecs}
        out[f"{name}/inputs"] = np.frombuffer(digest(s.U, s.Ub, s.phi, s.theta0, st["theta_b"], s.eigvals, s.eigvecs, [s.dt], np.round(s.tau0, 9)).encode(), dtype=np.uint8)
        for k in range(N_STEPS):
            st = ref.correct(s.mesh.desc, spec.models[0], spec.schemes.limiter, s.dt, s.U, s.Ub, s.phi, st["theta"], st["theta_b"],
                             st["tau"], st["tau_b"], st["eigvals"], st["eigvecs"], want_matrix=(k == 0 and name == MATRIX_CASE))
            if k == 0 and name == MATRIX_CASE:
                for f in ("lower", "upper", "diag", "source", "internalCoeffs", "boundaryCoeffs"):
                    out[f"{name}/matrix/{f}"] = st[f]
            for f in ("theta", "tau", "theta_b", "tau_b") if k + 1 in STORED_STEPS else ():
                out[f"{name}/step{k + 1}/{f}"] = st[f]
        print(name, s.mesh.n_cells, "cells")
    np.savez_compressed(ROOT / "tests" / "golden" / "reference_correct.npz", **out)

    import test_unstructured
    un = {}
    for limiter in ("cubista", "upwind"):
        for k, st in enumerate(test_unstructured.reference_on_the_unstructured_mesh(limiter)):
            for f in ("theta", "tau", "tau_b"):
                un[f"{limiter}/step{k + 1}/{f}"] = st[f]
    np.savez_compressed(ROOT / "tests" / "golden" / "reference_unstructured.npz", **un)

    th = cell_vectors()
    D, V, nrot = ref.jacobi(th)
    rng = np.random.default_rng(7)
    M = rng.standard_normal((512, 9)); vals = np.zeros((512, 9)); vals[:, [0, 4, 8]] = D
    om, B = ref.decompose_gradU(M, vals, V.reshape(-1, 9))
    lims = {f"lims/{l}/{k}": v for l in range(6) for k, v in zip(("alpha", "beta", "bounds"), ref.lims(l))}
    np.savez_compressed(ROOT / "tests" / "golden" / "reference_cell.npz", theta=th, expD=D, V=V, nrot=nrot, M=M, omega=om, B=B,
                        innerP_T=ref.innerP(V.reshape(-1, 9), M, True), innerP=ref.innerP(V.reshape(-1, 9), M, False), **lims)
    # BMPLog: the theta equation and theta -> tau of the reference text, fed with the fluidity of the oracle's PhiEqn
    import test_bmp_log
    bmp = {}
    for name, scale in test_bmp_log.BMP_FIXTURE_CASES:
        for k, (fl, th, ta, tab) in enumerate(test_bmp_log.bmp_reference_run(name, scale)):
            bmp[f"{name}/step{k + 1}/fluidity"], bmp[f"{name}/step{k + 1}/theta"], bmp[f"{name}/step{k + 1}/tau"], bmp[f"{name}/step{k + 1}/tau_b"] = fl, th, ta, tab
    np.savez_compressed(ROOT / "tests" / "golden" / "reference_bmp.npz", **bmp)
    print("written")


if __name__ == "__main__":
    main()
