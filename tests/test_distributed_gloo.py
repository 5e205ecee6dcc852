"""world_size-2/4 runs of the host-side logic of a decomposed stress step on CPU (backend gloo): direct
generation of each rank's processor mesh, global time step / cell count reductions, NCCL-id style broadcast,
processor-patch cross-check and patchNeighbourField exchange — compared with the single-process answer.
The data path itself (halo swaps + Krylov reductions on NCCL) is covered on GPUs by test_multi_gpu.py."""
import socket
import tempfile
from pathlib import Path

import numpy as np
import pytest

from rheotool_b200 import abi, cases, mesh


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case,scale,decomp", [("C3", 3 / 19, (2, 1, 1)), ("C5", 12 / 400, (2, 2, 1))])
def test_decomposed_host_logic_gloo(case, scale, decomp):
    import torch.multiprocessing as mp
    from mp_worker import gloo_rank
    world = decomp[0] * decomp[1] * decomp[2]
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(gloo_rank, args=(world, _free_port(), case, scale, decomp, td), nprocs=world, join=True)
        ranks = [np.load(Path(td) / f"rank{r}.npz") for r in range(world)]
    spec = cases.by_name(case, scale)
    m = mesh.tensor_grid(spec.grid)
    U, Ub, phi, theta0 = m.synth_fields(spec.synth)
    dt = spec.cfl / m.max_courant_rate(phi)
    seen = np.zeros(m.n_cells, dtype=int)
    for r, d in enumerate(ranks):
        assert int(d["n_total"]) == m.n_cells
        assert float(d["dt"]) == pytest.approx(dt, rel=1e-12)
        cells = d["cells"]
        seen[cells] += 1
        # fields keyed by GLOBAL cell id: a rank's synthetic fields are the global ones restricted to its cells
        assert np.allclose(d["theta0"], theta0[cells], rtol=1e-11, atol=1e-13)   # cell centres agree to the last bits only
        assert np.allclose(d["U"], U[cells], rtol=1e-12, atol=1e-14)
    assert (seen == 1).all()
    # patchNeighbourField: values received on a processor patch are the neighbour rank's owner-cell values
    for r, d in enumerate(ranks):
        nint = int(d["n_internal"])
        for t, st, sz, nb in zip(d["patch_type"], d["patch_start"], d["patch_size"], d["patch_nbr"]):
            if t != abi.PATCH_PROCESSOR or sz == 0:
                continue
            o = ranks[nb]
            k = next(i for i in range(len(o["patch_type"])) if o["patch_type"][i] == abi.PATCH_PROCESSOR and o["patch_nbr"][i] == r)
            ost, osz, onint = int(o["patch_start"][k]), int(o["patch_size"][k]), int(o["n_internal"])
            assert osz == sz
            expect = o["theta0"][o["owner"][ost:ost + osz]]
            assert np.array_equal(d["nbr_theta"][st - nint: st - nint + sz], expect)
