"""Shared set-up for the parity tests: one synthetic case fed identically to the oracle and the GPU."""
from __future__ import annotations

import numpy as np

from oracle import mesh_ref
from oracle import oracle as orc
from rheotool_b200 import abi, cases, mesh


class Setup:
    """mesh + fields of one configuration (optionally decomposed for the oracle / the GPU ranks)."""

    def __init__(self, spec: cases.CaseSpec, cold_start: bool = False, cfl: float | None = None):
        self.spec = spec
        self.mesh = mesh.tensor_grid(spec.grid)
        self.U, self.Ub, self.phi, self.theta0 = self.mesh.synth_fields(spec.synth)
        rate = self.mesh.max_courant_rate(self.phi)
        self.dt = (cfl if cfl is not None else spec.cfl) / rate
        if cold_start:   # no eigVals/eigVecs files: READ_IF_PRESENT defaults (identity)
            n = self.mesh.n_cells
            self.eigvals = np.tile(np.eye(3).reshape(9), (n, 1))
            self.eigvecs = self.eigvals.copy()
        else:            # restart: eigen-pairs of theta0 as the previous correct() left them
            self.eigvals, self.eigvecs = orc.calc_eig(self.theta0)
        self.tau0 = np.zeros_like(self.theta0)

    def oracle(self, schemes=None, sort_eig=True) -> orc.OracleCase:
        oc = orc.OracleCase([self.mesh.desc], self.spec.models, schemes or self.spec.schemes, sort_eig)
        for mi in range(len(self.spec.models)):
            oc.set_state(0, mi, self.theta0 * (1.0 + 0.1 * mi), self.tau0, self.eigvals_mode(mi), self.eigvecs_mode(mi))
        oc.set_velocity(0, self.U, self.Ub, self.phi)
        return oc

    # modes get slightly different initial theta so that a mode mix-up cannot go unnoticed
    def theta_mode(self, mi):
        return self.theta0 * (1.0 + 0.1 * mi)

    def eigvals_mode(self, mi):
        if mi == 0 or np.allclose(self.eigvals[:, [1, 2, 3, 5, 6, 7]], 0) and np.allclose(self.eigvecs, np.tile(np.eye(3).reshape(9), (len(self.eigvecs), 1))):
            return self.eigvals if mi == 0 else self.eigvals
        return orc.calc_eig(self.theta_mode(mi))[0]

    def eigvecs_mode(self, mi):
        if mi == 0 or np.allclose(self.eigvecs, np.tile(np.eye(3).reshape(9), (len(self.eigvecs), 1))):
            return self.eigvecs
        return orc.calc_eig(self.theta_mode(mi))[1]

    def gpu(self, schemes=None, device=0):
        from rheotool_b200.stress import GpuStressModel
        g = GpuStressModel(self.mesh, self.spec.models, schemes or self.spec.schemes, device)
        for mi in range(len(self.spec.models)):
            g.upload_state(mi, self.theta_mode(mi), self.tau0, self.eigvals_mode(mi), self.eigvecs_mode(mi))
        g.upload_velocity(self.U, self.Ub, self.phi)
        return g


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


def tight(schemes: abi.RheoSchemeCtl, tol=1e-15, solver=None) -> abi.RheoSchemeCtl:
    """Same schemes with a Krylov tolerance far below the parity tolerance (SURVEY.md §7 hard parts)."""
    c = abi.RheoSchemeCtl()
    for f, _ in abi.RheoSchemeCtl._fields_:
        setattr(c, f, getattr(schemes, f))
    c.tolerance = tol
    if solver is not None:
        c.solver = abi.SOLVER[solver]
    return c
