#!/usr/bin/env python
"""Advance the stress of an OpenFOAM case on the GPU with the velocity field frozen (what rheoFoam does with the momentum
balance switched off): read constant/polyMesh and <time>/{U, theta, tau[, eigVals, eigVecs]}, run n steps of
constitutiveEq::correct() on cuda:0, write the new time directory (theta, tau, eigVals, eigVecs) for restart / ParaView.

    python tools/run_case.py CASE --time 0 --dt 0.02 --steps 50

The model(s) come from constant/constitutiveProperties (one log-conformation model or a multiMode of them), the convection
limiter, time scheme and solver controls from system/fvSchemes and system/fvSolution — like rheoFoam reads them; --model and
the flags after it replace the dictionaries.  The flux is createPhi's (linear interpolation of U dotted with Sf).  Needs a CUDA device: there is no CPU path."""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rheotool_b200 import abi, cases, foamio  # noqa: E402
from rheotool_b200.stress import GpuStressModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("case")
    ap.add_argument("--time", default="0")
    ap.add_argument("--name", default="", help="field name suffix (multi-region / multi-phase cases)")
    ap.add_argument("--model", default=None, choices=sorted(abi.MODEL_NAMES), help="override constant/constitutiveProperties")
    ap.add_argument("--etaS", type=float, default=0.0)
    ap.add_argument("--etaP", type=float, default=1.0)
    ap.add_argument("--lambda", dest="lambda_", type=float, default=1.0)
    ap.add_argument("--param", action="append", default=[], metavar="KEY=VALUE", help="further model_desc parameters (alpha, epsilon, zeta, L2, ...)")
    ap.add_argument("--limiter", default="cubista", choices=sorted(abi.LIMITER))
    ap.add_argument("--ddt", default="Euler", choices=["Euler", "backward"])
    ap.add_argument("--tolerance", type=float, default=1e-10)
    ap.add_argument("--dt", type=float, required=True)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--gz", action="store_true")
    a = ap.parse_args()

    case = Path(a.case)
    if a.model:
        extra = {k: (v if k == "ptt_function" else float(v)) for k, v in (kv.split("=", 1) for kv in a.param)}
        models, names = [cases.model_desc(a.model, etaS=a.etaS, etaP=a.etaP, lambda_=a.lambda_, **extra)], [a.name]
        schemes = cases.scheme_ctl(a.limiter, "PBiCGStab", a.tolerance, ddt=a.ddt)
    else:
        cp = case / "constant" / "constitutiveProperties"
        models, names = foamio.read_models(cp), [a.name + n for n in foamio.mode_names(cp)]
        schemes, solver = foamio.read_schemes_modes(case, ["theta" + n[len(a.name):] if a.name else "theta" + n for n in names])
        print(f"fvSolution selects {solver} + DILU for theta: that is the solver the device runs")
    m, f = foamio.read_case(case, a.time, names[0])
    per_mode = [f] + [foamio.read_case(case, a.time, n)[1] for n in names[1:]]
    g = GpuStressModel(m, models, schemes, 0)
    for mi, fm in enumerate(per_mode):
        g.upload_state(mi, fm["theta"], fm["tau"], fm["eigvals"], fm["eigvecs"], theta_b=fm["theta_b"])
        if models[mi].model == abi.MODEL_BMP_LOG:   # BMPLog.C:112-122: Phi is MUST_READ
            if fm["Phi"] is None:
                raise SystemExit(f"{case / a.time / ('Phi' + names[mi])}: MUST_READ field is missing (BMPLog)")
            g.upload_fluidity(mi, fm["Phi"], fm["Phi_b"])
    g.upload_velocity(f["U"], f["U_b"], f["phi"])
    for n in range(a.steps):
        g.store_old_time()
        g.correct(a.dt)
        print(f"step {n + 1}: Krylov iterations {g.last_iterations()}")
    t_new = f"{float(a.time) + a.steps * a.dt:g}"
    for mi, n in enumerate(names):
        foamio.write_case(case, m, t_new, g.theta(mi), g.tau(mi), f["U"], f["U_b"], theta_b=g.download(abi.FIELD_THETA_B, mi), tau_b=g.download(abi.FIELD_TAU_B, mi),
                          eigvals=g.download(abi.FIELD_EIGVALS, mi), eigvecs=g.download(abi.FIELD_EIGVECS, mi), name=n, gz=a.gz,
                          Phi=g.fluidity(mi) if models[mi].model == abi.MODEL_BMP_LOG else None,
                          Phi_b=g.fluidity_b(mi) if models[mi].model == abi.MODEL_BMP_LOG else None)
    print(f"wrote {Path(a.case) / t_new}")


if __name__ == "__main__":
    main()
