#!/usr/bin/env python
"""Generate tests/golden/aneurysm_renumber.json from the reference's shipped polyMesh
(of90/tutorials/rheoFoam/Aneurysm/HerschelBulkley/constant/polyMesh.org/{owner,neighbour}.gz) with the
numpy restatement of the renumbering (oracle/mesh_ref.py).  Run in the build container only
(/root/reference does not exist on the GPU box); the JSON (digests, not the mesh) is committed."""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import mesh_ref  # noqa: E402
from test_mesh_integers import ANEURYSM, _read_label_list  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int32).tobytes()).hexdigest()


own = _read_label_list(ANEURYSM / "owner.gz")
nei = _read_label_list(ANEURYSM / "neighbour.gz")
ncell = int(own.max()) + 1
perm, colour, cstart = mesh_ref.colour_renumber(ncell, own[: len(nei)], nei)
out = {"source": "of90/tutorials/rheoFoam/Aneurysm/HerschelBulkley/constant/polyMesh.org/{owner,neighbour}.gz",
       "generator": "tools/make_golden_aneurysm.py (oracle/mesh_ref.py colour_renumber)",
       "n_cells": ncell, "n_faces": int(len(own)), "n_internal_faces": int(len(nei)),
       "n_colours": int(len(cstart) - 1), "colour_start": [int(x) for x in cstart],
       "perm_sha256": digest(perm), "colour_sha256": digest(colour),
       "owner_sha256": digest(own), "neighbour_sha256": digest(nei)}
(ROOT / "tests" / "golden" / "aneurysm_renumber.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
