#!/usr/bin/env python
"""Generate tests/golden/aneurysm_geometry.json: counts, patches, total volume and patch areas of the polyMesh the reference
ships (of90/tutorials/rheoFoam/Aneurysm/HerschelBulkley/constant/polyMesh.org), as read by rheo_io_read_polymesh and with
the EXT-OF9 primitiveMesh geometry of csrc/host/mesh.cpp.  Run in the build container only (/root/reference does not exist
on the GPU box); the JSON (a summary, not the mesh) is committed."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from rheotool_b200 import foamio  # noqa: E402
from test_foam_io import ANEURYSM, _geometry_summary  # noqa: E402

s = _geometry_summary(foamio.read_polymesh(ANEURYSM))
s["source"] = "of90/tutorials/rheoFoam/Aneurysm/HerschelBulkley/constant/polyMesh.org/{points,faces,owner,neighbour}.gz + boundary"
s["generator"] = "tools/make_golden_aneurysm_geometry.py (rheo_io_read_polymesh)"
(ROOT / "tests" / "golden" / "aneurysm_geometry.json").write_text(json.dumps(s, indent=1) + "\n")
print(json.dumps(s, indent=1))
