"""blockMesh + mirrorMesh, restated for the stock meshes of the reference's tutorials (host-side tool, numpy).

BASELINE.json config 1 is "the stock tutorial mesh" of rheoFoam/Cylinder/Oldroyd-BLog: `blockMesh` on
system/blockMeshDict (8 hex blocks, arc edges around the cylinder, simpleGrading) followed by `mirrorMesh` about y = 0
(system/mirrorMeshDict; Allrun:9-13).  Neither utility is part of /root/reference (they are OpenFOAM-9's), so this
module restates what they do — EXT-OF9, the builder's reading of
    blockMesh/blockDescriptors/blockDescriptorEdges.C   edge points + weights, curved edges used in either sense
    blockMesh/gradingDescriptor + lineDivide.C          geometric expansion along an edge
    blockMesh/blockEdges/arcEdge/arcEdge.C              circle through three points
    blockMesh/blocks/block/blockCreate.C                weighted blend of the 12 edges + curved-edge correction
    polyMeshFromShapeMesh.C                             faces in upper-triangular order, patches in dictionary order
    mirrorFvMesh.C                                      original cells first, then their mirror images
and writes an ordinary constant/polyMesh that `foamio.read_polymesh` (and OpenFOAM) reads.  Cell count, cell order,
patch names / types and the geometry of the block edges follow the dictionary exactly; interior point positions follow
the published blend and are not pinned against a blockMesh binary (none in this environment).  The order of the faces
INSIDE a boundary patch is by owner cell, not blockMesh's block-face order.
"""
from __future__ import annotations

import re
from pathlib import Path

import numpy as np

# hexModel: the six faces of a hex (outward-pointing for a right-handed cell), and the 12 edges in blockDescriptor order
HEX_FACES = [(0, 4, 7, 3), (1, 2, 6, 5), (0, 1, 5, 4), (3, 7, 6, 2), (0, 3, 2, 1), (4, 5, 6, 7)]
BLOCK_EDGES = [(0, 1), (3, 2), (7, 6), (4, 5), (0, 3), (1, 2), (5, 6), (4, 7), (0, 4), (1, 5), (2, 6), (3, 7)]


# ---------------------------------------------------------------------------------------------- dictionary
def _tokens(text: str):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return re.findall(r"[(){};]|[^\s(){};]+", text)


def _parse_list(tok, i):
    """tok[i] == '(' -> (nested python list, index after the closing parenthesis)"""
    assert tok[i] == "("
    out, i = [], i + 1
    while tok[i] != ")":
        if tok[i] == "(":
            sub, i = _parse_list(tok, i)
            out.append(sub)
        else:
            out.append(tok[i])
            i += 1
    return out, i + 1


def parse_block_mesh_dict(path) -> dict:
    tok = _tokens(Path(path).read_text())
    d = {"scale": 1.0, "vertices": [], "blocks": [], "edges": [], "boundary": []}
    i = 0
    while i < len(tok):
        t = tok[i]
        if t == "FoamFile":
            while tok[i] != "}":
                i += 1
            i += 1
        elif t in ("convertToMeters", "scale"):
            d["scale"] = float(tok[i + 1]); i += 3
        elif t == "vertices":
            lst, i = _parse_list(tok, i + 1)
            d["vertices"] = [[float(x) for x in v] for v in lst]
        elif t == "blocks":
            lst, i = _parse_list(tok, i + 1)
            k = 0
            while k < len(lst):
                assert lst[k] == "hex", "only hex blocks"
                verts = [int(x) for x in lst[k + 1]]
                dens = [int(x) for x in lst[k + 2]]
                assert lst[k + 3] == "simpleGrading", "only simpleGrading"
                grad = [float(x) for x in lst[k + 4]]
                d["blocks"].append((verts, dens, grad))
                k += 5
        elif t == "edges":
            lst, i = _parse_list(tok, i + 1)
            k = 0
            while k < len(lst):
                assert lst[k] == "arc", "only arc edges"
                d["edges"].append((int(lst[k + 1]), int(lst[k + 2]), [float(x) for x in lst[k + 3]]))
                k += 4
        elif t == "boundary":
            assert tok[i + 1] == "("
            i += 2
            while tok[i] != ")":
                name = tok[i]; assert tok[i + 1] == "{"
                i += 2
                ptype, faces = "patch", []
                while tok[i] != "}":
                    if tok[i] == "type":
                        ptype = tok[i + 1]; i += 3
                    elif tok[i] == "faces":
                        lst, i = _parse_list(tok, i + 1)
                        faces = [[int(x) for x in f] for f in lst]
                    else:
                        i += 1
                i += 1
                d["boundary"].append((name, ptype, faces))
            i += 1
        else:
            i += 1
    return d


# ---------------------------------------------------------------------------------------------- edges
def _divisions(n: int, ratio: float) -> np.ndarray:
    """lineDivide.C with one grading section: lambda_i, i = 0..n"""
    lam = np.zeros(n + 1)
    lam[n] = 1.0
    if ratio == 1.0 or n == 1:
        lam[1:n] = np.arange(1, n) / n
    else:
        g = ratio ** (1.0 / (n - 1))
        i = np.arange(1, n)
        lam[1:n] = (1.0 - g ** i) / (1.0 - g ** n)
    return lam


class _Arc:
    """arcEdge.C: the circle through p1, pm, p3; position(lambda) = rotation of p1 about the centre by lambda * angle"""

    def __init__(self, p1, pm, p3):
        p1, pm, p3 = (np.asarray(x, dtype=float) for x in (p1, pm, p3))
        a, b = pm - p1, p3 - p1
        asqr, bsqr, adotb = a @ a, b @ b, a @ b
        denom = asqr * bsqr - adotb * adotb
        fact = 0.5 * (bsqr - adotb) / denom
        self.c = p1 + 0.5 * a + fact * np.cross(np.cross(a, b), a)
        r1, r2, r3 = p1 - self.c, pm - self.c, p3 - self.c
        ang = np.arccos(np.clip((r3 @ r1) / (np.linalg.norm(r3) * np.linalg.norm(r1)), -1.0, 1.0))
        if np.cross(r1, r2) @ np.cross(r1, r3) < 0.0:
            ang = 2 * np.pi - ang
        axis = np.cross(r1, r3) if ang <= np.pi else np.cross(r3, r1)
        self.angle, self.radius = ang, np.linalg.norm(r3)
        self.e1 = r1 / np.linalg.norm(r1)
        e3 = axis / np.linalg.norm(axis)
        self.e2 = np.cross(e3, self.e1)
        self.p1, self.p3 = p1, p3

    def position(self, lam):
        lam = np.asarray(lam)
        out = self.c + self.radius * (np.cos(lam * self.angle)[:, None] * self.e1 + np.sin(lam * self.angle)[:, None] * self.e2)
        out[lam < 1e-15] = self.p1
        out[lam > 1 - 1e-15] = self.p3
        return out


def _edge_points_weights(pa, pb, va, vb, n, ratio, arcs):
    """blockDescriptorEdges.C::setEdge: points and weights of the block edge va -> vb with n divisions"""
    if (va, vb) in arcs:
        lam = _divisions(n, ratio)
        return arcs[(va, vb)].position(lam), lam
    if (vb, va) in arcs:   # curved edge defined in the opposite sense: inverse expansion, reversed
        lam = _divisions(n, 1.0 / ratio)
        return arcs[(vb, va)].position(lam)[::-1].copy(), (1.0 - lam)[::-1].copy()
    lam = _divisions(n, ratio)
    return pa + (pb - pa) * lam[:, None], lam


def _block_points(P8, vlabels, dens, grad, arcs):
    """blockCreate.C::createPoints -> array [nk+1][nj+1][ni+1][3]"""
    ni, nj, nk = dens
    nd = [ni] * 4 + [nj] * 4 + [nk] * 4
    ex = [grad[0]] * 4 + [grad[1]] * 4 + [grad[2]] * 4
    p, w = [], []
    curved = False
    for e, (a, b) in enumerate(BLOCK_EDGES):
        pts, lam = _edge_points_weights(P8[a], P8[b], vlabels[a], vlabels[b], nd[e], ex[e], arcs)
        curved |= (vlabels[a], vlabels[b]) in arcs or (vlabels[b], vlabels[a]) in arcs
        p.append(pts); w.append(lam)
    K, J, I = np.meshgrid(np.arange(nk + 1), np.arange(nj + 1), np.arange(ni + 1), indexing="ij")
    w0, w1, w2, w3 = (w[e][I] for e in range(4))
    w4, w5, w6, w7 = (w[e][J] for e in range(4, 8))
    w8, w9, w10, w11 = (w[e][K] for e in range(8, 12))
    wx = [(1 - w0) * (1 - w4) * (1 - w8) + w0 * (1 - w5) * (1 - w9), (1 - w1) * w4 * (1 - w11) + w1 * w5 * (1 - w10),
          (1 - w2) * w7 * w11 + w2 * w6 * w10, (1 - w3) * (1 - w7) * w8 + w3 * (1 - w6) * w9]
    wy = [(1 - w4) * (1 - w0) * (1 - w8) + w4 * (1 - w1) * (1 - w11), (1 - w5) * w0 * (1 - w9) + w5 * w1 * (1 - w10),
          (1 - w6) * w3 * w9 + w6 * w2 * w10, (1 - w7) * (1 - w3) * w8 + w7 * (1 - w2) * w11]
    wz = [(1 - w8) * (1 - w0) * (1 - w4) + w8 * (1 - w3) * (1 - w7), (1 - w9) * w0 * (1 - w5) + w9 * w3 * (1 - w6),
          (1 - w10) * w1 * w5 + w10 * w2 * w6, (1 - w11) * (1 - w1) * w4 + w11 * (1 - w2) * w7]
    for ws in (wx, wy, wz):
        s = ws[0] + ws[1] + ws[2] + ws[3]
        for q in range(4):
            ws[q] = ws[q] / s
    p000, p100, p110, p010, p001, p101, p111, p011 = P8
    lin = lambda a, b, t: a + (b - a) * t[..., None]
    edge = [lin(p000, p100, w0), lin(p010, p110, w1), lin(p011, p111, w2), lin(p001, p101, w3),
            lin(p000, p010, w4), lin(p100, p110, w5), lin(p101, p111, w6), lin(p001, p011, w7),
            lin(p000, p001, w8), lin(p100, p101, w9), lin(p110, p111, w10), lin(p010, p011, w11)]
    wt = wx + wy + wz
    pts = sum(wt[e][..., None] * edge[e] for e in range(12)) / 3.0
    if curved:
        idx = [I] * 4 + [J] * 4 + [K] * 4
        pts = pts + sum(wt[e][..., None] * (p[e][idx[e]] - edge[e]) for e in range(12))
    # the block vertices themselves
    for (k, j, i), v in zip([(0, 0, 0), (0, 0, ni), (0, nj, ni), (0, nj, 0), (nk, 0, 0), (nk, 0, ni), (nk, nj, ni), (nk, nj, 0)], P8):
        pts[k, j, i] = v
    return pts


# ---------------------------------------------------------------------------------------------- hexes -> polyMesh
def _build_hexes(d):
    from scipy.spatial import cKDTree
    V = np.asarray(d["vertices"], dtype=float) * d["scale"]
    arcs = {(a, b): _Arc(V[a], np.asarray(m) * d["scale"], V[b]) for a, b, m in d["edges"]}
    all_pts, hexes = [], []
    base = 0
    for verts, dens, grad in d["blocks"]:
        ni, nj, nk = dens
        pts = _block_points(V[verts], verts, dens, grad, arcs)
        lab = base + np.arange((ni + 1) * (nj + 1) * (nk + 1)).reshape(nk + 1, nj + 1, ni + 1)
        all_pts.append(pts.reshape(-1, 3))
        k, j, i = np.meshgrid(np.arange(nk), np.arange(nj), np.arange(ni), indexing="ij")
        h = np.stack([lab[k, j, i], lab[k, j, i + 1], lab[k, j + 1, i + 1], lab[k, j + 1, i],
                      lab[k + 1, j, i], lab[k + 1, j, i + 1], lab[k + 1, j + 1, i + 1], lab[k + 1, j + 1, i]], axis=-1)
        hexes.append(h.reshape(-1, 8))   # i fastest, then j, then k: blockMesh's cell order inside a block
        base += lab.size
    P = np.concatenate(all_pts)
    H = np.concatenate(hexes)
    # merge the points of coincident block faces (blockMesh merges them topologically; geometrically they agree to rounding)
    tol = 1e-9 * np.ptp(P, axis=0).max()
    tree = cKDTree(P)
    rep = np.arange(len(P))
    for a, b in tree.query_pairs(tol, output_type="ndarray"):
        ra, rb = rep[a], rep[b]
        while rep[ra] != ra: ra = rep[ra]
        while rep[rb] != rb: rb = rep[rb]
        if ra != rb:
            rep[max(ra, rb)] = min(ra, rb)
    for q in range(len(rep)):
        r = q
        while rep[r] != r: r = rep[r]
        rep[q] = r
    uniq, inv = np.unique(rep, return_inverse=True)
    return P[uniq], inv[H], V, tol


def _mirror(P, H, point, normal, tol):
    """mirrorFvMesh.C: points on the plane are shared, the mirrored cells follow the original ones"""
    n = np.asarray(normal, dtype=float); n /= np.linalg.norm(n)
    dist = (P - np.asarray(point, dtype=float)) @ n
    on = np.abs(dist) <= tol
    new_lab = np.where(on, np.arange(len(P)), -1)
    extra = np.flatnonzero(~on)
    new_lab[extra] = len(P) + np.arange(len(extra))
    P2 = np.concatenate([P, P[extra] - 2.0 * dist[extra, None] * n])
    Hm = new_lab[H][:, [0, 3, 2, 1, 4, 7, 6, 5]]   # reflection flips the handedness: reverse the winding of both quad layers
    return P2, np.concatenate([H, Hm]), new_lab


def _polymesh(P, H, patches, default_name="defaultFaces"):
    """polyMeshFromShapeMesh.C: internal faces by owner, then by neighbour (upper-triangular order), owner's outward
    orientation; boundary faces patch by patch (dictionary order), inside a patch by owner cell and cell-face index"""
    nC = len(H)
    F = H[:, HEX_FACES].reshape(-1, 4)                         # [6 nC][4], cell-face q of cell c at 6 c + q
    key = np.sort(F, axis=1)
    order = np.lexsort((key[:, 3], key[:, 2], key[:, 1], key[:, 0]))
    ks = key[order]
    same = np.all(ks[1:] == ks[:-1], axis=1)
    cell = order // 6
    internal_pairs = np.flatnonzero(same)
    assert not np.any(same[1:] & same[:-1]), "a face shared by more than two cells"
    a, b = order[internal_pairs], order[internal_pairs + 1]
    own_f = np.where(cell[internal_pairs] < cell[internal_pairs + 1], a, b)       # the cell-face of the lower-numbered cell
    nei_c = np.maximum(cell[internal_pairs], cell[internal_pairs + 1])
    own_c = np.minimum(cell[internal_pairs], cell[internal_pairs + 1])
    srt = np.lexsort((nei_c, own_c))
    faces = [F[own_f[srt]]]
    owner = [own_c[srt]]
    neighbour = nei_c[srt]
    paired = np.zeros(len(F), dtype=bool)
    paired[a] = True; paired[b] = True
    bfaces = np.flatnonzero(~paired)
    bkey = {tuple(k): f for k, f in zip(map(tuple, key[bfaces]), bfaces)}
    used = np.zeros(len(F), dtype=bool)
    out_patches = []
    for name, ptype, quads in patches:
        ids = []
        for q in quads:
            f = bkey.get(tuple(sorted(q)))
            if f is None:
                raise ValueError(f"patch {name}: face {q} is not a boundary face of the mesh")
            ids.append(f)
        ids = np.array(sorted(set(ids)), dtype=int)
        used[ids] = True
        out_patches.append((name, ptype, sum(len(x) for x in faces), len(ids)))
        faces.append(F[ids]); owner.append(ids // 6)
    rest = np.array([f for f in bfaces if not used[f]], dtype=int)
    if len(rest):
        out_patches.append((default_name, "empty", sum(len(x) for x in faces), len(rest)))
        faces.append(F[rest]); owner.append(rest // 6)
    return np.concatenate(faces), np.concatenate(owner), neighbour, out_patches


def _expand_patch_faces(d, P, H, V, tol):
    """the dictionary names BLOCK faces (4 block vertices); find the cell faces lying on each of them"""
    from scipy.spatial import cKDTree
    vlab = cKDTree(P).query(V)[1]                                   # block vertex -> mesh point
    F = H[:, HEX_FACES].reshape(-1, 4)
    key = np.sort(F, axis=1)
    _, first, counts = np.unique(key, axis=0, return_index=True, return_counts=True)
    bfaces = first[counts == 1]
    # a block face is one side (i, j or k = 0 / n) of one block: collect the boundary cell-faces block by block
    out = []
    cell0 = 0
    block_side_faces = {}
    for verts, dens, grad in d["blocks"]:
        ni, nj, nk = dens
        c = cell0 + np.arange(ni * nj * nk).reshape(nk, nj, ni)
        sides = {(0, 4, 7, 3): (c[:, :, 0], 0), (1, 2, 6, 5): (c[:, :, -1], 1), (0, 1, 5, 4): (c[:, 0, :], 2),
                 (3, 7, 6, 2): (c[:, -1, :], 3), (0, 3, 2, 1): (c[0, :, :], 4), (4, 5, 6, 7): (c[-1, :, :], 5)}
        for loc, (cells, q) in sides.items():
            block_side_faces[frozenset(verts[x] for x in loc)] = (cells.reshape(-1) * 6 + q)
        cell0 += ni * nj * nk
    for name, ptype, quads in d["boundary"]:
        ids = []
        for q in quads:
            fs = block_side_faces.get(frozenset(q))
            if fs is None:
                raise ValueError(f"patch {name}: {q} is not a face of any block")
            ids.extend(F[fs].tolist())
        out.append((name, ptype, ids))
    return out


def generate(block_mesh_dict, mirror_point=None, mirror_normal=None, plane_tolerance=1e-7):
    """-> (points [nP][3], faces [nF][4], owner [nF], neighbour [nInt], patches [(name, type, start, size)])"""
    d = parse_block_mesh_dict(block_mesh_dict)
    P, H, V, tol = _build_hexes(d)
    patches = _expand_patch_faces(d, P, H, V, tol)
    if mirror_normal is not None:
        P2, H2, new_lab = _mirror(P, H, mirror_point, mirror_normal, plane_tolerance)
        patches = [(name, ptype, quads + [[int(new_lab[v]) for v in (q[0], q[3], q[2], q[1])] for q in quads]) for name, ptype, quads in patches]
        P, H = P2, H2
    faces, owner, neighbour, out_patches = _polymesh(P, H, patches)
    return P, faces, owner, neighbour, out_patches


# ---------------------------------------------------------------------------------------------- constant/polyMesh
def _header(cls, obj, note=""):
    return ("FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       %s;\n" % cls + (f"    note        \"{note}\";\n" if note else "") +
            "    location    \"constant/polyMesh\";\n    object      %s;\n}\n\n" % obj)


def write_polymesh(directory, P, faces, owner, neighbour, patches):
    directory = Path(directory)
    directory.mkdir(parents=True, exist_ok=True)
    nC = int(owner.max()) + 1
    note = f"nPoints:{len(P)}  nCells:{nC}  nFaces:{len(faces)}  nInternalFaces:{len(neighbour)}"
    (directory / "points").write_text(_header("vectorField", "points") + f"{len(P)}\n(\n" + "\n".join("(%.17g %.17g %.17g)" % tuple(p) for p in P) + "\n)\n")
    (directory / "faces").write_text(_header("faceList", "faces") + f"{len(faces)}\n(\n" + "\n".join("4(%d %d %d %d)" % tuple(f) for f in faces) + "\n)\n")
    (directory / "owner").write_text(_header("labelList", "owner", note) + f"{len(owner)}\n(\n" + "\n".join(str(int(o)) for o in owner) + "\n)\n")
    (directory / "neighbour").write_text(_header("labelList", "neighbour", note) + f"{len(neighbour)}\n(\n" + "\n".join(str(int(o)) for o in neighbour) + "\n)\n")
    body = f"{len(patches)}\n(\n"
    for name, ptype, start, size in patches:
        body += f"    {name}\n    {{\n        type            {ptype};\n" + ("        inGroups        1(wall);\n" if ptype == "wall" else "") + \
                f"        nFaces          {size};\n        startFace       {start};\n    }}\n"
    (directory / "boundary").write_text(_header("polyBoundaryMesh", "boundary") + body + ")\n")


def cylinder_tutorial_mesh(case_dir, out_dir):
    """rheoFoam/Cylinder/Oldroyd-BLog: blockMesh, then mirrorMesh -overwrite about (0 0 2.5), normal (0 -1 0) (Allrun:9-13,
    system/mirrorMeshDict).  Writes out_dir/constant/polyMesh and returns the HostMesh read back from it."""
    from . import foamio
    case_dir = Path(case_dir)
    md = (case_dir / "system" / "mirrorMeshDict").read_text()
    bp = [float(x) for x in re.search(r"basePoint\s*\(([^)]*)\)", md).group(1).split()]
    nv = [float(x) for x in re.search(r"normalVector\s*\(([^)]*)\)", md).group(1).split()]
    tol = float(re.search(r"planeTolerance\s+([^;]+);", md).group(1))
    P, faces, owner, neighbour, patches = generate(case_dir / "system" / "blockMeshDict", bp, nv, tol)
    pm = Path(out_dir) / "constant" / "polyMesh"
    write_polymesh(pm, P, faces, owner, neighbour, patches)
    return foamio.read_polymesh(pm)


# ---------------------------------------------------------------------------------------------- BASELINE config 1
def cylinder_tutorial_dicts():
    """The geometry of of90/tutorials/rheoFoam/Cylinder/Oldroyd-BLog (system/blockMeshDict:17-170, system/mirrorMeshDict:17-23)
    re-expressed from its parameters, so that the stock mesh can be generated where /root/reference does not exist (the GPU
    box): a cylinder of radius 1 on the centre line of a channel of half-height 2, inlet at x = -20, outlet at x = 60, one
    cell thick; the upper half is meshed with 8 blocks (O-grid of 6 blocks around the half cylinder out to the 2.83 / 2 box,
    an inflow and an outflow block) and mirrored about y = 0.  tests/test_blockmesh.py checks, where the reference is
    present, that this text and the reference's own dictionary give the same mesh bit for bit.
    -> (blockMeshDict text, mirror base point, mirror normal, plane tolerance)"""
    s2, c8, s8, c16, s16, c316, s316 = 0.7071067812, 0.9238795325, 0.3826834324, 0.9807852804, 0.195090322, 0.8314696123, 0.555570233
    layer = [(-20, 0), (-2.83, 0), (-1, 0), (-s2, s2), (-s8, c8), (0, 1), (s8, c8), (s2, s2), (1, 0), (2.83, 0), (60, 0), (60, 2), (2, 2),
             (0.83, 2), (0, 2), (-0.83, 2), (-2, 2), (-20, 2)]
    verts = [(x, y, z) for z in (0, 1) for x, y in layer]
    blocks = [((0, 1, 16, 17), (33, 40, 1), (0.12, 1, 1)), ((2, 3, 16, 1), (40, 43, 1), (1, 30, 1)), ((3, 4, 15, 16), (23, 43, 1), (0.5, 30, 1)),
              ((4, 5, 14, 15), (23, 43, 1), (1, 30, 1)), ((5, 6, 13, 14), (23, 43, 1), (1, 30, 1)), ((6, 7, 12, 13), (20, 43, 1), (2, 30, 1)),
              ((7, 8, 9, 12), (60, 43, 1), (0.2, 30, 1)), ((9, 10, 11, 12), (50, 60, 1), (20, 5, 1))]
    arcs = [(5, 6, (s16, c16)), (6, 7, (s316, c316)), (7, 8, (c8, s8)), (12, 9, (2.614579077, 1.0829941136)),
            (4, 5, (-s16, c16)), (3, 4, (-s316, c316)), (2, 3, (-c8, s8)), (1, 16, (-2.614579077, 1.0829941136))]
    quad = lambda a, b: f"({a} {b} {b + 18} {a + 18})"
    patch = {"inlet": ("patch", [quad(0, 17)]), "walls": ("wall", [quad(17, 16), quad(16, 15), quad(15, 14), quad(14, 13), quad(13, 12), quad(12, 11)]),
             "cylinder": ("wall", [quad(2, 3), quad(3, 4), quad(4, 5), quad(5, 6), quad(6, 7), quad(7, 8)]), "outlet": ("patch", [quad(10, 11)])}
    fab = [f"({b[0][0]} {b[0][1]} {b[0][2]} {b[0][3]})" for b in blocks] + [f"({b[0][0] + 18} {b[0][1] + 18} {b[0][2] + 18} {b[0][3] + 18})" for b in blocks]
    t = "vertices\n(\n" + "\n".join("  (%r %r %r)" % v for v in verts) + "\n);\n\nblocks\n(\n"
    for (a, b, c, e), dens, gr in blocks:
        t += f"    hex ({a} {b} {c} {e} {a + 18} {b + 18} {c + 18} {e + 18}) ({dens[0]} {dens[1]} {dens[2]}) simpleGrading ({gr[0]!r} {gr[1]!r} {gr[2]!r})\n"
    t += ");\n\nedges\n(\n"
    for z in (0, 1):
        for a, b, (x, y) in arcs:
            t += f"  arc {a + 18 * z} {b + 18 * z} ({x!r} {y!r} {z})\n"
    t += ");\n\nboundary\n(\n"
    for name, (ptype, faces) in patch.items():
        t += f"    {name}\n    {{\n        type {ptype};\n        faces\n        (\n" + "\n".join("            " + f for f in faces) + "\n        );\n    }\n"
    t += "    frontAndBack\n    {\n        type empty;\n        faces\n        (\n" + "\n".join("            " + f for f in fab) + "\n        );\n    }\n);\n"
    return t, (0.0, 0.0, 2.5), (0.0, -1.0, 0.0), 1e-7


def cylinder_stock_mesh(out_dir):
    """BASELINE.json config 1's mesh (24,894 cells) from cylinder_tutorial_dicts(); returns the HostMesh read back from
    out_dir/constant/polyMesh"""
    from . import foamio
    text, bp, nv, tol = cylinder_tutorial_dicts()
    out_dir = Path(out_dir)
    (out_dir / "system").mkdir(parents=True, exist_ok=True)
    (out_dir / "system" / "blockMeshDict").write_text(text)
    P, faces, owner, neighbour, patches = generate(out_dir / "system" / "blockMeshDict", bp, nv, tol)
    pm = out_dir / "constant" / "polyMesh"
    write_polymesh(pm, P, faces, owner, neighbour, patches)
    return foamio.read_polymesh(pm)
