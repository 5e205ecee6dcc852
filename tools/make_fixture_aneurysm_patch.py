#!/usr/bin/env python
"""Cut a small connected piece out of the polyMesh the reference ships (tutorials/rheoFoam/Aneurysm/.../polyMesh.org:
unstructured, polyhedral) and commit it as an OpenFOAM polyMesh fixture under tests/golden/aneurysm_patch/: the only
UNSTRUCTURED mesh the GPU tests can use on the GPU box (where /root/reference does not exist).  It exercises the generic
slot-count paths of the kernels (KT = 0) and multi-colour DILU (more than 2 colours).

Cells: breadth-first from a many-faced refinement-transition polyhedron next to the wall, N_CELLS cells, kept in ascending original order
(so the faces stay in upper-triangular order).  Faces towards cells that were not selected become patch `cut`.
Run in the build container only; rheo_io_read_polymesh reads the result."""
import gzip
import re
import sys
from collections import deque
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference/of90/tutorials/rheoFoam/Aneurysm/HerschelBulkley/constant/polyMesh.org")
DST = ROOT / "tests" / "golden" / "aneurysm_patch"
N_CELLS = 3000


def body(path):
    txt = gzip.open(path, "rt").read()
    mm = re.search(r"\n(\d+)\s*\n?\(", txt)
    return int(mm.group(1)), txt[mm.end(): txt.rindex(")")]


n, b = body(SRC / "points.gz")
pts = np.array(re.findall(r"[-+0-9.eE]+", b), dtype=np.float64).reshape(n, 3)
n, b = body(SRC / "faces.gz")
faces = [np.array(m.split(), dtype=np.int64) for m in re.findall(r"\d+\(([^)]*)\)", b)]
assert len(faces) == n
n, b = body(SRC / "owner.gz"); own = np.array(b.split(), dtype=np.int64)
n, b = body(SRC / "neighbour.gz"); nei = np.array(b.split(), dtype=np.int64)
nint, ncell = len(nei), int(own.max()) + 1
btxt = (SRC / "boundary").read_text()
patches = [(m[0], int(m[1]), int(m[2])) for m in re.findall(r"(\w+)\s*\{\s*type\s+patch;\s*nFaces\s+(\d+);\s*startFace\s+(\d+);", btxt)]
assert [p[0] for p in patches] == ["walls", "out1", "in1", "out2"]

# adjacency
adj = [[] for _ in range(ncell)]
for f in range(nint):
    adj[own[f]].append(nei[f]); adj[nei[f]].append(own[f])
# seed: the refinement-transition polyhedron (snappyHexMesh cells with 9-24 faces) nearest to the wall, so that the piece holds
# tetrahedra/prisms/hexahedra AND many-faced polyhedra (slots per cell != 4, 6 -> the kernels' run-time slot-count paths)
fpc = np.bincount(own, minlength=ncell) + np.bincount(nei, minlength=ncell)
wall_cells = set(own[patches[0][2]: patches[0][2] + patches[0][1]].tolist())
# the many-faced cell closest (in graph distance) to the wall: breadth-first from all wall cells
dist = -np.ones(ncell, dtype=np.int64)
dq = deque(wall_cells)
for c in wall_cells:
    dist[c] = 0
seed = None
while dq and seed is None:
    c = dq.popleft()
    if fpc[c] >= 12:
        seed = int(c)
    for nb in adj[c]:
        if dist[nb] < 0:
            dist[nb] = dist[c] + 1; dq.append(nb)
sel, q, seen = [], deque([seed]), {seed}
while q and len(sel) < N_CELLS:
    c = q.popleft(); sel.append(c)
    for nb in adj[c]:
        if nb not in seen:
            seen.add(nb); q.append(nb)
sel = np.array(sorted(sel))
new_of = -np.ones(ncell, dtype=np.int64); new_of[sel] = np.arange(len(sel))

int_faces, bnd = [], {p[0]: [] for p in patches}
bnd["cut"] = []
for f in range(nint):
    o, nb = new_of[own[f]], new_of[nei[f]]
    if o >= 0 and nb >= 0:
        int_faces.append((o, nb, faces[f]))
    elif o >= 0:
        bnd["cut"].append((o, faces[f]))
    elif nb >= 0:
        bnd["cut"].append((nb, faces[f][::-1]))      # the selected cell was the neighbour: flip the face
for pname, nf, sf in patches:
    for f in range(sf, sf + nf):
        if new_of[own[f]] >= 0:
            bnd[pname].append((new_of[own[f]], faces[f]))
int_faces.sort(key=lambda t: (t[0], t[1]))           # upper-triangular order
order = [p[0] for p in patches] + ["cut"]
all_faces = [t[2] for t in int_faces] + [t[1] for p in order for t in bnd[p]]
new_own = [t[0] for t in int_faces] + [t[0] for p in order for t in bnd[p]]
new_nei = [t[1] for t in int_faces]
used = np.unique(np.concatenate(all_faces))
pid = -np.ones(len(pts), dtype=np.int64); pid[used] = np.arange(len(used))

HEAD = "FoamFile\n{{\n    version     2.0;\n    format      ascii;\n    class       {cls};\n    location    \"constant/polyMesh\";\n    object      {obj};\n}}\n\n"
DST.mkdir(parents=True, exist_ok=True)


def put(name, cls, text, gz=True):
    data = HEAD.format(cls=cls, obj=name) + text
    if gz:
        with gzip.GzipFile(DST / (name + ".gz"), "wb", mtime=0) as fh:
            fh.write(data.encode())
    else:
        (DST / name).write_text(data)


put("points", "vectorField", f"{len(used)}\n(\n" + "".join("(%.17g %.17g %.17g)\n" % tuple(pts[p]) for p in used) + ")\n")
put("faces", "faceList", f"{len(all_faces)}\n(\n" + "".join("%d(%s)\n" % (len(f), " ".join(str(pid[v]) for v in f)) for f in all_faces) + ")\n")
put("owner", "labelList", f"{len(new_own)}\n(\n" + "".join(f"{v}\n" for v in new_own) + ")\n")
put("neighbour", "labelList", f"{len(new_nei)}\n(\n" + "".join(f"{v}\n" for v in new_nei) + ")\n")
start, lines = len(int_faces), []
for p in order:
    kind = "wall" if p == "walls" else "patch"
    lines.append(f"    {p}\n    {{\n        type            {kind};\n        nFaces          {len(bnd[p])};\n        startFace       {start};\n    }}\n")
    start += len(bnd[p])
put("boundary", "polyBoundaryMesh", f"{len(order)}\n(\n" + "".join(lines) + ")\n", gz=False)
(DST / "README").write_text("Cut out of of90/tutorials/rheoFoam/Aneurysm/HerschelBulkley/constant/polyMesh.org by tools/make_fixture_aneurysm_patch.py\n"
                            f"({len(sel)} cells breadth-first from a wall-adjacent transition polyhedron; faces towards unselected cells form patch `cut`; `walls` typed wall).\n")
print(len(sel), "cells", len(int_faces), "internal faces", {p: len(bnd[p]) for p in order}, "points", len(used))
