"""Rank programs for the multi-process tests (importable so torch.multiprocessing can spawn them).

  gloo_rank   CPU, backend gloo, world_size >= 2: host-side logic of a decomposed run
  gpu_rank    one GPU per rank, backend nccl: the stress step on a decomposed mesh against the oracle
"""
from __future__ import annotations

import os
import sys
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _init(rank, world, port, backend):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import datetime
    dist.init_process_group(backend=backend, rank=rank, world_size=world, timeout=datetime.timedelta(seconds=90))
    return dist


def gloo_rank(rank, world, port, case, scale, decomp, out_dir):
    try:
        dist = _init(rank, world, port, "gloo")
        from rheotool_b200 import cases, distributed, mesh
        spec = cases.by_name(case, scale)
        part = mesh.tensor_grid_part(spec.grid, *decomp, rank)
        U, Ub, phi, theta0 = part.synth_fields(spec.synth)
        info = distributed.rank_info()
        assert (info.rank, info.world) == (rank, world)
        n_total = distributed.global_cell_count(part)
        dt = distributed.global_time_step(part, phi, spec.cfl)
        distributed.check_processor_patches(part, phi)
        uid = distributed.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
        assert uid == bytes(range(128))
        # patchNeighbourField of theta: what the GPU halo exchange must deliver
        own_b = part.owner[part.n_internal:]
        nbr_theta = distributed.exchange_processor_patches(part, theta0[own_b])
        np.savez(Path(out_dir) / f"rank{rank}.npz", n_total=n_total, dt=dt, cells=part.global_cells(), theta0=theta0, U=U,
                 nbr_theta=nbr_theta, patch_type=np.array([p.type for p in part.patches]), patch_start=np.array([p.start for p in part.patches]),
                 patch_size=np.array([p.size for p in part.patches]), patch_nbr=np.array([p.nbr_rank for p in part.patches]),
                 n_internal=part.n_internal, owner=part.owner)
        # a deliberately broken flux must be caught by the start-up check on every rank that owns a cut face
        bad = phi.copy()
        has_proc = False
        for p in part.patches:
            if p.nbr_rank >= 0 and p.size:
                bad[p.start] += 1.0
                has_proc = True
        caught = False
        try:
            distributed.check_processor_patches(part, bad)
        except RuntimeError:
            caught = True
        assert caught == has_proc
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        traceback.print_exc()
        raise


def bmp_model():
    """the BMPLog model of the multi-GPU fluidity case (shared with the parent test)"""
    from rheotool_b200 import cases
    return cases.model_desc("BMPLog", rho=1.0, etaS=0.01, etaP=0.01, lambda_=0.7, bmp_G0=0.8, bmp_k=2.0, bmp_Phi0=2.5, bmp_PhiInf=40.0)


def bmp_fluidity(C, owner_b):
    """smooth initial fluidity from the cell centres; boundary values = those of the adjacent cells"""
    Phi = 2.5 * (1.0 + 0.05 * np.sin(5 * C[:, 0]) * np.cos(3 * C[:, 1]))
    return Phi, Phi[owner_b].copy()


def gpu_rank(rank, world, port, case, scale, decomp, steps, out_dir, tol, solver="PBiCGStab", bmp=False):
    """Each rank: its part of the mesh on cuda:<rank>; all ranks: the oracle on the same decomposition
    (in-process emulation) is run by rank 0 only and compared by the parent test."""
    try:
        import torch
        torch.cuda.set_device(rank)
        dist = _init(rank, world, port, "nccl")
        from helpers import tight
        from rheotool_b200 import abi, cases, distributed, mesh
        from rheotool_b200.stress import GpuStressModel, eig_exp
        spec = cases.by_name(case, scale)
        if bmp:
            spec.models = [bmp_model()]
        part = mesh.tensor_grid_part(spec.grid, *decomp, rank)
        U, Ub, phi, theta0 = part.synth_fields(spec.synth)
        info = distributed.rank_info()
        dt = distributed.global_time_step(part, phi, spec.cfl)
        distributed.check_processor_patches(part, phi)
        sc = tight(spec.schemes, tol, solver=solver)
        g = GpuStressModel(part, spec.models, sc, rank)
        distributed.connect(g, info)
        for mi in range(len(spec.models)):
            th = theta0 * (1.0 + 0.1 * mi)
            vals, vecs = eig_exp(th, rank)
            g.upload_state(mi, th, np.zeros_like(th), vals, vecs)
        g.upload_velocity(U, Ub, phi)
        if bmp:
            g.upload_fluidity(0, *bmp_fluidity(part.C, part.owner[part.n_internal:]))
        iters = []
        for _ in range(steps):
            g.store_old_time()
            g.correct(dt)
            iters.append(g.last_iterations())
        div_tau = g.div_tau(abi.STAB_COUPLING)   # collective: swaps the velocity gradient of the ghost cells
        np.savez(Path(out_dir) / f"gpu_rank{rank}.npz", cells=part.global_cells(), dt=dt, iters=np.array(iters), div_tau=div_tau,
                 fluidity=g.fluidity(0) if bmp else np.zeros(0),
                 **{f"theta{mi}": g.theta(mi) for mi in range(len(spec.models))},
                 **{f"tau{mi}": g.tau(mi) for mi in range(len(spec.models))})
        g.close()
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        traceback.print_exc()
        raise
