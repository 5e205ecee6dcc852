"""One process per GPU: the host-side plumbing of a decomposed run (what `mpirun -np N <solver> -parallel`
gives the reference through OpenFOAM's Pstream).

torch.distributed is only the rendezvous/plumbing layer here (backend "nccl" on GPUs, "gloo" in the CPU
tests): it distributes the NCCL unique id the C-ABI needs (rheo_gpu_nccl_unique_id / rheo_gpu_comm_init),
agrees on global scalars (cell count, time step) and cross-checks the processor patches of the rank meshes
before the first step.  The halo swaps and Krylov reductions of the stress step itself run inside
librheo_b200.so on NCCL (csrc/gpu/engine.cu), not through this module.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from . import abi
from .mesh import HostMesh


@dataclass
class RankInfo:
    rank: int
    world: int
    local_rank: int


def rank_info() -> RankInfo:
    """RANK / WORLD_SIZE / LOCAL_RANK as torchrun (or the test launcher) exports them."""
    return RankInfo(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def _dist():
    import torch.distributed as dist
    return dist


def _device():
    import torch
    dist = _dist()
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_reduce_scalar(x: float, op: str = "sum") -> float:
    """gSum / gMax of one double over the ranks (identity when not initialised)."""
    import torch
    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op])
    return float(t.item())


def global_cell_count(mesh: HostMesh) -> int:
    return int(round(all_reduce_scalar(mesh.n_cells, "sum")))


def global_time_step(mesh: HostMesh, phi: np.ndarray, cfl: float) -> float:
    """dt such that the face Courant number max_cells(dt * sum(outflow)/V) equals `cfl` on the GLOBAL mesh."""
    return cfl / all_reduce_scalar(mesh.max_courant_rate(phi), "max")


def broadcast_bytes(payload: bytes | None, n: int, src: int = 0) -> bytes:
    """Broadcast an n-byte blob (the 128-byte NCCL unique id) from rank `src`."""
    import torch
    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bytes(payload)
    t = torch.zeros(n, dtype=torch.uint8, device=_device())
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def exchange_processor_patches(mesh: HostMesh, values: np.ndarray) -> np.ndarray:
    """patchNeighbourField() on the host: `values` [n_boundary, w] holds this rank's patch-internal values;
    returns the neighbour ranks' values on the processor faces (other boundary faces: zeros).
    All sends/receives of a rank are posted as ONE batch (ncclGroupStart/End under the nccl backend)."""
    import torch
    dist = _dist()
    vals = np.ascontiguousarray(values, dtype=np.float64)
    vals = vals.reshape(len(vals), -1)
    out = np.zeros_like(vals)
    ops, recvs = [], []
    dev = _device()
    for p in mesh.patches:
        if p.type != abi.PATCH_PROCESSOR or p.size == 0:
            continue
        b0 = p.start - mesh.n_internal
        send = torch.from_numpy(vals[b0:b0 + p.size].copy()).to(dev)
        recv = torch.zeros_like(send)
        ops.append(dist.P2POp(dist.isend, send, p.nbr_rank))
        ops.append(dist.P2POp(dist.irecv, recv, p.nbr_rank))
        recvs.append((b0, p.size, recv))
    if ops:   # one group: NCCL would deadlock on a send/send ordering of separate calls
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    for b0, n, recv in recvs:
        out[b0:b0 + n] = recv.cpu().numpy()
    return out.reshape(values.shape)


def check_processor_patches(mesh: HostMesh, phi: np.ndarray | None = None, rtol: float = 1e-10) -> None:
    """Start-up cross-check of a decomposition (what checkMesh -parallel guards in an OpenFOAM run): the k-th
    face of my patch towards rank r must be the k-th face of r's patch towards me — same neighbour cell
    centre as recorded in nbr_C, opposite area vector and, when given, opposite flux."""
    nb = mesh.n_boundary
    if nb == 0:
        return
    own_b = mesh.owner[mesh.n_internal:]
    pack = np.concatenate([mesh.C[own_b], mesh.Sf[mesh.n_internal:], (phi[mesh.n_internal:, None] if phi is not None else np.zeros((nb, 1)))], axis=1)
    got = exchange_processor_patches(mesh, pack)
    for p in mesh.patches:
        if p.type != abi.PATCH_PROCESSOR or p.size == 0:
            continue
        b0 = p.start - mesh.n_internal
        sl = slice(b0, b0 + p.size)
        scale = max(1.0, float(np.abs(mesh.C).max()))
        if not np.allclose(got[sl, 0:3], mesh.nbr_C[sl], rtol=rtol, atol=rtol * scale):
            raise RuntimeError(f"processor patch to rank {p.nbr_rank}: neighbour cell centres do not match nbr_C")
        if not np.allclose(got[sl, 3:6], -pack[sl, 3:6], rtol=rtol, atol=rtol * float(np.abs(pack[:, 3:6]).max())):
            raise RuntimeError(f"processor patch to rank {p.nbr_rank}: face area vectors are not opposite")
        if phi is not None and not np.allclose(got[sl, 6], -pack[sl, 6], rtol=rtol, atol=rtol * float(np.abs(phi).max() + 1e-300)):
            raise RuntimeError(f"processor patch to rank {p.nbr_rank}: face fluxes are not opposite")


def connect(model, info: RankInfo) -> None:
    """Join the ranks' GpuStressModel objects into one NCCL communicator (id from rank 0)."""
    if info.world <= 1:
        return
    from .stress import GpuStressModel
    uid = GpuStressModel.nccl_unique_id() if info.rank == 0 else None
    uid = broadcast_bytes(uid, 128, src=0)
    model.comm_init(info.rank, info.world, uid)
