"""CPU check of the FORMULATION pbicg.cuh implements (the kernels themselves need a GPU: tests/test_gpu_pbicg.py).

numpy restatement of exactly what the device does —
  * A^T without a transposed mesh structure: slot coefficient of A is min(F, 0), of A^T min(-F, 0), F = signed outflow
    of the row's cell through the slot's face (written by k_flux_assemble next to A);
  * rD = 1/diag shared by A and A^T;
  * DILU / DILU^T substitutions as one fully parallel pass per colour (forward colour 0..nc-1, backward nc-2..0) on the
    colour-sorted numbering;
  * the per-RHS scalar control of krylov.cuh (ctl_pb_beta_one / ctl_pb_alpha_one / ctl_pb_end_one)
— run on the system the oracle assembles, against the oracle's own PBiCG (EXT-OF9 PBiCG.C restated with sequential
face-order sweeps) on the renumbered mesh: same iteration counts per component, same solution."""
import numpy as np

from helpers import Setup, rel_l2, tight
from oracle import mesh_ref
from oracle import oracle as orc
from rheotool_b200 import abi, cases


def _device_style_pbicg(n, own, nei, phi_int, D, b, x0, cstart, tol, max_iter=1000):
    """One component.  Slots: every cell's list of (neighbour, F_out) like the ELL rows of the device."""
    nbr = [[] for _ in range(n)]
    for f, (o, q) in enumerate(zip(own, nei)):
        nbr[o].append((q, phi_int[f]))      # owner: outflow = phi
        nbr[q].append((o, -phi_int[f]))     # neighbour: outflow = -phi
    A = [[(q, min(F, 0.0)) for q, F in row] for row in nbr]
    AT = [[(q, min(-F, 0.0)) for q, F in row] for row in nbr]
    rD = 1.0 / D

    def spmv(M, p):
        return np.array([D[c] * p[c] + sum(a * p[q] for q, a in M[c]) for c in range(n)])

    def precond(M, r):
        w = np.zeros(n)
        for k in range(len(cstart) - 1):                       # forward, colour by colour (cells of a colour independent)
            cells = range(cstart[k], cstart[k + 1])
            new = {c: rD[c] * (r[c] - sum(a * w[q] for q, a in M[c] if q < c)) for c in cells}
            for c, v in new.items():
                w[c] = v
        for k in range(len(cstart) - 3, -1, -1):               # backward: nc-2 .. 0
            cells = range(cstart[k], cstart[k + 1])
            new = {c: w[c] - rD[c] * sum(a * w[q] for q, a in M[c] if q > c) for c in cells}
            for c, v in new.items():
                w[c] = v
        return w

    x = x0.copy()
    wA = spmv(A, x)
    rowsum = np.array([D[c] + sum(a for _, a in A[c]) for c in range(n)])
    xref = x.mean()
    norm = np.abs(wA - xref * rowsum).sum() + np.abs(b - xref * rowsum).sum() + 1e-20   # lduMatrix::solver::normFactor
    rA = b - wA
    rT = rA.copy()
    init = fin = np.abs(rA).sum() / norm
    iters, rho = 0, 0.0
    pA = pT = None
    if fin < tol:
        return x, 0, init
    while True:
        wA, wT = precond(A, rA), precond(AT, rT)
        rho_old, rho = rho, float(wA @ rT)
        if iters == 0:
            pA, pT = wA.copy(), wT.copy()
        else:
            beta = rho / rho_old
            pA, pT = wA + beta * pA, wT + beta * pT
        wA, wT = spmv(A, pA), spmv(AT, pT)
        wApT = float(wA @ pT)
        if not abs(wApT) / norm > 1e-300:
            break
        alpha = rho / wApT
        x += alpha * pA; rA -= alpha * wA; rT -= alpha * wT
        fin = np.abs(rA).sum() / norm
        iters += 1
        if not (iters < max_iter and not fin < tol):
            break
    return x, iters, init


def test_device_pbicg_formulation_matches_the_oracle_pbicg_on_the_renumbered_mesh():
    spec = cases.cube(6, "cavity", 1, "Oldroyd-BLog")
    spec.schemes = cases.scheme_ctl("upwind", "PBiCG", 1e-10)
    s = Setup(spec, cfl=1.0)                   # several Krylov iterations
    perm, colour, cstart = s.mesh.colour_renumber()
    rm = mesh_ref.renumbered_mesh(mesh_ref.from_host_mesh(s.mesh), perm)
    desc = mesh_ref.to_desc(rm, abi)
    fa = rm.face_addr
    phi_r = np.where(fa > 0, s.phi[np.abs(fa) - 1], -s.phi[np.abs(fa) - 1])

    def oracle(solver, tol):
        oc = orc.OracleCase([desc], spec.models, tight(spec.schemes, tol=tol, solver=solver))
        oc.set_state(0, 0, s.theta0[perm], s.tau0[perm], s.eigvals[perm], s.eigvecs[perm])
        oc.set_velocity(0, s.U[perm], s.Ub, phi_r)
        st = (abi.RheoStepStats * 1)()
        oc.store_old_time(); oc.step(s.dt, st)
        return oc.get(0, 0, abi.FIELD_THETA), st[0]

    exact, _ = oracle("PBiCGStab", 1e-15)
    got_o, st = oracle("PBiCG", 1e-10)
    assert max(st.n_iterations) >= 3

    # the assembled system (upwind + Euler, all patches zeroGradient walls): gaussDefCmpwConvectionScheme.C:92-110
    n, nint = rm.n_cells, rm.n_internal
    own, nei = rm.owner[:nint], rm.neighbour
    D = rm.V / s.dt
    np.add.at(D, own, np.maximum(phi_r[:nint], 0.0))
    np.add.at(D, nei, np.maximum(-phi_r[:nint], 0.0))
    np.add.at(D, rm.owner[nint:], phi_r[nint:])                # internalCoeffs = phi_b * 1 on zeroGradient patches
    th0 = s.theta0[perm]
    got = []
    for cmp in range(6):
        # right-hand side: b = A x* with the tightly converged solution x*
        xs = exact[:, cmp]
        b = D * xs
        lo, up = -np.maximum(phi_r[:nint], 0.0), np.minimum(phi_r[:nint], 0.0)
        np.add.at(b, nei, lo * xs[own]); np.add.at(b, own, up * xs[nei])
        x, iters, init = _device_style_pbicg(n, own, nei, phi_r[:nint], D, b, th0[:, cmp].copy(), cstart, 1e-10)
        got.append(iters)
        assert abs(init - st.initial_residual[cmp]) <= 1e-6 * st.initial_residual[cmp]
        assert rel_l2(x, got_o[:, cmp]) <= 1e-9
    # b is rebuilt from the converged solution, so it differs from the oracle's right-hand side in the last bits: allow one
    # iteration of difference on a component that ends within rounding of the tolerance, none on most
    assert all(abs(a - b) <= 1 for a, b in zip(got, st.n_iterations)), (got, list(st.n_iterations))
    assert sum(a == b for a, b in zip(got, st.n_iterations)) >= 4, (got, list(st.n_iterations))
