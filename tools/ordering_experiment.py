#!/usr/bin/env python
"""Krylov iteration counts of the CPU oracle under different cell orderings.

DILU is (D+L) D^-1 (D+U) with L/U split by the cell numbering, so the ordering decides how much of the (upwind) matrix
the forward substitution solves exactly.  This script renumbers a tensor-grid case in several ways, runs the oracle
(test infrastructure, CPU) on each renumbered mesh and prints the iteration counts per step.

  natural     the mesh generator's order (= the reference's, blockMesh i-fastest)
  redblack    greedy colouring, colour-major (what the device used in round 1)
  tileRB      bx x by x bz tiles, natural order inside a tile, tiles red-black, tile-major
  lineRB      whole x-lines, natural order inside a line, lines red-black by (j+k)
usage: ordering_experiment.py CASE SCALE [steps]
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import mesh_ref                      # noqa: E402
from oracle import oracle as orc                 # noqa: E402
from rheotool_b200 import abi, cases, mesh       # noqa: E402


def ijk_of(m):
    out = []
    for d in range(3):
        u, inv = np.unique(np.round(m.C[:, d], 9), return_inverse=True)
        out.append(inv.astype(np.int64))
    return out


TILES = [(4, 4, 2), (8, 4, 1), (8, 2, 2)]


def perms(m):
    n = m.n_cells
    i, j, k = ijk_of(m)
    nat = np.arange(n, dtype=np.int32)
    out = {"natural": nat}
    out["redblack"] = np.lexsort((nat, (i + j + k) & 1)).astype(np.int32)
    for (bx, by, bz) in TILES:
        if k.max() == 0:
            bz = 1
        I, J, K = i // bx, j // by, k // bz
        tid = (K * (J.max() + 1) + J) * (I.max() + 1) + I
        col = (I + J + K) & 1
        out[f"tileRB{bx}x{by}x{bz}"] = np.lexsort((nat, tid, col)).astype(np.int32)
    col = (j + k) & 1
    out["lineRB"] = np.lexsort((nat, col)).astype(np.int32)
    br = m.block_renumber()   # what the device uses (csrc/host/ordering.hpp)
    if br is not None:
        out["device-blocks"] = br[0]
    return out


def run(spec, m, perm, steps, solver="PBiCGStab"):
    U, Ub, phi, theta0 = m.synth_fields(spec.synth)
    dt = spec.cfl / m.max_courant_rate(phi)
    rm = mesh_ref.renumbered_mesh(mesh_ref.from_host_mesh(m), perm)
    desc = mesh_ref.to_desc(rm, abi)
    sc = spec.schemes
    sc.solver = abi.SOLVER[solver]
    oc = orc.OracleCase([desc], spec.models, sc)
    vals, vecs = orc.calc_eig(theta0)
    for mi in range(len(spec.models)):
        oc.set_state(0, mi, theta0[perm], np.zeros_like(theta0), vals[perm], vecs[perm])
    fa = rm.face_addr
    phi_r = np.where(fa > 0, phi[np.abs(fa) - 1], -phi[np.abs(fa) - 1])
    oc.set_velocity(0, U[perm], Ub, phi_r)
    its = []
    so = (abi.RheoStepStats * len(spec.models))()
    for _ in range(steps):
        oc.store_old_time(); oc.step(dt, so)
        its.append((max(so[0].n_iterations), max(so[0].final_residual)))
    return its


if __name__ == "__main__":
    name, scale = sys.argv[1], float(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    only = sys.argv[4].split(",") if len(sys.argv) > 4 else None
    spec = cases.by_name(name, scale)
    m = mesh.tensor_grid(spec.grid)
    print(name, scale, m.n_cells, "cells", flush=True)
    for label, p in perms(m).items():
        if only and label not in only:
            continue
        t = time.time()
        its = run(spec, m, p, steps)
        print(f"{label:16s} iterations {[a for a, _ in its]}  final residuals {[f'{b:.1e}' for _, b in its]}  ({time.time() - t:.0f} s)", flush=True)
