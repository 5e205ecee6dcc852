"""N > 1: one rank per GPU (processor patches + NCCL halo swaps + all-reduced Krylov scalars) against the
oracle on ONE rank, i.e. decomposePar + mpirun must not change the answer.
Needs >= 2 CUDA devices (gpurun --gpus 2); skipped otherwise."""
import socket
import tempfile
from pathlib import Path

import numpy as np
import pytest

from helpers import Setup, rel_l2, tight
from rheotool_b200 import abi, cases

pytestmark = pytest.mark.gpu


def _n_gpus():
    return abi.lib().rheo_gpu_device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case,scale,decomp,steps", [
    ("C3", 4 / 19, (2, 1, 1), 2),
    ("C2", 1 / 9, (2, 1, 1), 2),
    ("C4", 14 / 252, (2, 1, 1), 1),
    ("C5", 20 / 400, (2, 2, 1), 2),
    ("C5", 20 / 400, (2, 2, 2), 1),
])
def test_decomposed_gpu_run_matches_single_rank_oracle(case, scale, decomp, steps):
    _decomposed_run(case, scale, decomp, steps, "PBiCGStab")


@pytest.mark.gpu
@pytest.mark.parametrize("case,scale,decomp,steps", [("C3", 4 / 19, (2, 1, 1), 2), ("C2", 1 / 9, (2, 1, 1), 2)])
def test_decomposed_pbicg_run_matches_single_rank_oracle(case, scale, decomp, steps):
    """the tutorials' solver (fvSolution: PBiCG + DILU) on several ranks: Amul and Tmul with processor interfaces, DILU / DILU^T
    rank-local, every dot all-reduced (solve.inl: solve_batch_pbicg)"""
    _decomposed_run(case, scale, decomp, steps, "PBiCG")


@pytest.mark.gpu
def test_decomposed_bmp_log_run_matches_single_rank_oracle():
    """BMPLog on two ranks: the hidden fluidity mode takes part in the halo swaps and all-reduces like every other right-hand side"""
    _decomposed_run("C3", 4 / 19, (2, 1, 1), 2, "PBiCGStab", bmp=True)


def _decomposed_run(case, scale, decomp, steps, solver, bmp=False):
    world = decomp[0] * decomp[1] * decomp[2]
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs, have {_n_gpus()}")
    import torch.multiprocessing as mp
    from mp_worker import gpu_rank
    tol = 1e-15
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(gpu_rank, args=(world, _free_port(), case, scale, decomp, steps, td, tol, solver, bmp), nprocs=world, join=True)
        ranks = [dict(np.load(Path(td) / f"gpu_rank{r}.npz")) for r in range(world)]
    spec = cases.by_name(case, scale)
    if bmp:
        from mp_worker import bmp_fluidity, bmp_model
        spec.models = [bmp_model()]
    s = Setup(spec)
    assert float(ranks[0]["dt"]) == pytest.approx(s.dt, rel=1e-12)
    oc = s.oracle(tight(spec.schemes, tol, solver=solver))
    if bmp:
        oc.set_fluidity(0, 0, *bmp_fluidity(s.mesh.C, s.mesh.owner[s.mesh.n_internal:]))
    for _ in range(steps):
        oc.store_old_time(); oc.step(s.dt)
    for mi in range(len(spec.models)):
        for name, fld in (("theta", abi.FIELD_THETA), ("tau", abi.FIELD_TAU)):
            ref = oc.get(0, mi, fld)
            got = np.full_like(ref, np.nan)
            for d in ranks:
                got[d["cells"]] = d[f"{name}{mi}"]
            err = rel_l2(got, ref)
            assert err <= 1e-10, f"{case} {decomp} mode {mi} {name}: rel L2 {err:.3e}"
    if bmp:
        ref = oc.get(0, 0, abi.FIELD_FLUIDITY)
        got = np.full_like(ref, np.nan)
        for d in ranks:
            got[d["cells"]] = d["fluidity"]
        assert rel_l2(got, ref) <= 1e-10, f"BMPLog fluidity on {decomp}: rel L2 {rel_l2(got, ref):.3e}"
    # explicit part of constitutiveEq::divTau across the processor patches (constitutiveEq.C:72-132, stabilization coupling)
    ref = oc.div_tau(0, abi.STAB_COUPLING)
    got = np.full_like(ref, np.nan)
    for d in ranks:
        got[d["cells"]] = d["div_tau"]
    err = rel_l2(got, ref)
    assert err <= 1e-10, f"{case} {decomp} divTau: rel L2 {err:.3e}"
