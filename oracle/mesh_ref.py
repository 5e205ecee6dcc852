"""numpy restatement of the INTEGER contracts of the GPU library.  TEST INFRASTRUCTURE ONLY.

  * block_renumber    block ordering of lattice meshes (csrc/host/ordering.hpp): what the device uses for PBiCGStab
  * colour_renumber   greedy multi-colouring in cell order + colour-major stable sort (what
                      rheo_gpu_create applies on the device; DESIGN.md "Renumbering")
  * renumbered_mesh   the mesh after that permutation in OpenFOAM upper-triangular order — what
                      `renumberMesh` would write — so the C++ oracle can run on it
  * ell_tables        slot-major ELL neighbour / face tables in the new numbering
  * simple_decomp     EXT-OF9 simpleGeomDecomp (decomposeParDict method simple)
  * sub_mesh          EXT-OF9 domainDecomposition processor mesh (cells ascending, internal faces in
                      global order, physical patches, processor patches by neighbour rank)
Everything here must agree BIT-EXACTLY with the product's integer arrays (tests/test_mesh_integers.py).
Written independently (numpy, different algorithmic route) from rheotool_b200/csrc/host/*.cpp.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

PATCH_EMPTY, PATCH_PROCESSOR = 2, 3


@dataclass
class RefMesh:
    n_cells: int
    owner: np.ndarray          # [n_faces] int32
    neighbour: np.ndarray      # [n_internal] int32
    Sf: np.ndarray
    Cf: np.ndarray
    C: np.ndarray
    V: np.ndarray
    weights: np.ndarray
    nbr_C: np.ndarray
    patches: list              # [(type, start, size, nbr_rank, theta_bc, tau_bc)]
    solved: list
    cell_addr: np.ndarray | None = None
    face_addr: np.ndarray | None = None
    _keep: list = field(default_factory=list)

    @property
    def n_faces(self):
        return len(self.owner)

    @property
    def n_internal(self):
        return len(self.neighbour)


def from_host_mesh(m) -> RefMesh:
    pat = [(p.type, p.start, p.size, p.nbr_rank, p.theta_bc, p.tau_bc) for p in m.patches]
    return RefMesh(m.n_cells, m.owner.copy(), m.neighbour.copy(), m.Sf.copy(), m.Cf.copy(), m.C.copy(), m.V.copy(),
                   m.weights.copy(), m.nbr_C.copy(), pat, list(m.solved))


def to_desc(rm: RefMesh, abi):
    """Build a RheoMeshDesc (ctypes) over contiguous copies kept alive inside rm."""
    import ctypes as C
    d = abi.RheoMeshDesc()
    own = np.ascontiguousarray(rm.owner, dtype=np.int32)
    nei = np.ascontiguousarray(rm.neighbour if len(rm.neighbour) else np.zeros(1), dtype=np.int32)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (rm.Sf, rm.Cf, rm.C, rm.V, rm.weights,
                                                                rm.nbr_C if len(rm.nbr_C) else np.zeros((1, 3)))]
    pats = (abi.RheoPatchDesc * len(rm.patches))()
    for i, p in enumerate(rm.patches):
        pats[i].type, pats[i].start, pats[i].size, pats[i].nbr_rank, pats[i].theta_bc, pats[i].tau_bc = p
    rm._keep = [own, nei, arrs, pats]
    d.n_cells, d.n_faces, d.n_internal_faces, d.n_patches = rm.n_cells, rm.n_faces, rm.n_internal, len(rm.patches)
    d.owner = own.ctypes.data_as(C.POINTER(C.c_int32))
    d.neighbour = nei.ctypes.data_as(C.POINTER(C.c_int32))
    dp = C.POINTER(C.c_double)
    d.Sf, d.Cf, d.C, d.V, d.weights, d.nbr_C = [a.ctypes.data_as(dp) for a in arrs]
    d.patches = pats
    for q in range(6):
        d.solved_components[q] = rm.solved[q]
    return d


# ------------------------------------------------------------------------------------------ colouring
def colour_renumber(n_cells: int, owner: np.ndarray, neighbour: np.ndarray):
    """Returns (perm[new]=old, colour[old], colour_start)."""
    nint = len(neighbour)
    lower = [[] for _ in range(n_cells)]   # already-coloured neighbours = neighbours with a smaller index
    for f in range(nint):
        o, n = int(owner[f]), int(neighbour[f])
        if o < n:
            lower[n].append(o)
        else:
            lower[o].append(n)
    colour = np.zeros(n_cells, dtype=np.int32)
    for c in range(n_cells):
        used = {int(colour[q]) for q in lower[c]}
        k = 0
        while k in used:
            k += 1
        colour[c] = k
    perm = np.argsort(colour, kind="stable").astype(np.int32)
    ncol = int(colour.max()) + 1
    cstart = np.zeros(ncol + 1, dtype=np.int32)
    cstart[1:] = np.cumsum(np.bincount(colour, minlength=ncol))
    return perm, colour, cstart


CHUNK = 32


def block_renumber(rm: "RefMesh"):
    """Independent numpy restatement of csrc/host/ordering.hpp::block_renumber (the integer contract of the device's block
    ordering).  Returns (perm[new]=old, colour_start, tile) or None when the mesh is not a lattice mesh."""
    N, nint = rm.n_cells, rm.n_internal
    ijk, dims = [], []
    tol = 1e-10 * max(float((rm.C.max(axis=0) - rm.C.min(axis=0)).max()), 1e-300)
    for d in range(3):
        x = rm.C[:, d]
        u = np.unique(x)                      # sorted; merge planes closer than tol
        keep = np.concatenate([[True], np.diff(u) > tol])
        planes = u[keep]
        if len(planes) > 8192:
            return None
        idx = np.searchsorted(planes, x - tol)
        idx = np.clip(idx, 0, len(planes) - 1)
        if not np.all(np.abs(planes[idx] - x) <= tol):
            return None
        ijk.append(idx.astype(np.int64)); dims.append(len(planes))
    own, nei = rm.owner[:nint].astype(np.int64), rm.neighbour.astype(np.int64)
    dist = sum(np.abs(a[own] - a[nei]) for a in ijk)
    if nint and not np.all(dist == 1):
        return None
    thick = [d > 1 for d in dims]
    t = [1, 1, 1]
    if sum(thick) == 3:
        t = [4, 4, 2]
    elif sum(thick) == 2:
        w = [8, 4]
        t = [w.pop(0) if th else 1 for th in thick]
    elif sum(thick) == 1:
        t = [CHUNK if th else 1 for th in thick]
    nT = [(dims[d] + t[d] - 1) // t[d] for d in range(3)]
    i, j, k = ijk
    tid = ((k // t[2]) * nT[1] + (j // t[1])) * nT[0] + (i // t[0])
    local = ((k % t[2]) * t[1] + (j % t[1])) * t[0] + (i % t[0])
    seq = np.lexsort((np.arange(N), local, tid))
    n_chunks = (N + CHUNK - 1) // CHUNK
    chunk_of = np.empty(N, dtype=np.int64)
    chunk_of[seq] = np.arange(N) // CHUNK
    a, b = chunk_of[own], chunk_of[nei]
    cut = a != b
    hi, lo = np.maximum(a, b)[cut], np.minimum(a, b)[cut]
    order = np.argsort(hi, kind="stable")
    hi, lo = hi[order], lo[order]
    starts = np.searchsorted(hi, np.arange(n_chunks + 1))
    colour = np.zeros(n_chunks, dtype=np.int64)
    for q in range(n_chunks):
        used = set(colour[lo[starts[q]:starts[q + 1]]].tolist())
        c = 0
        while c in used:
            c += 1
        if c >= 63:
            return None
        colour[q] = c
    ncol = int(colour.max()) + 1
    if N % CHUNK != 0:
        cp, cl = int(colour[-1]), ncol - 1
        if cp != cl:
            was_cp, was_cl = colour == cp, colour == cl
            colour[was_cp], colour[was_cl] = cl, cp
    sizes = np.minimum(CHUNK, N - np.arange(n_chunks) * CHUNK)
    cstart = np.zeros(ncol + 1, dtype=np.int32)
    cstart[1:] = np.cumsum(np.bincount(colour, weights=sizes, minlength=ncol)).astype(np.int64)
    chunk_order = np.argsort(colour, kind="stable")
    perm = np.concatenate([seq[q * CHUNK:q * CHUNK + sizes[q]] for q in chunk_order]).astype(np.int32)
    return perm, cstart, tuple(t)


def chunk_levels(nbr: np.ndarray, n_cells: int):
    """(fwd, bwd) levels of the in-chunk dependency graphs for the slot-major neighbour table nbr[K, N] in the NEW numbering
    (csrc/host/ordering.hpp::chunk_levels): longest chain of lower- / higher-numbered neighbours inside the cell's chunk."""
    K, N = nbr.shape
    fwd = np.zeros(N, dtype=np.int64); bwd = np.zeros(N, dtype=np.int64)
    for c in range(N):
        base = c - c % CHUNK
        for s in range(K):
            nb = nbr[s, c]
            if base <= nb < c:
                fwd[c] = max(fwd[c], fwd[nb] + 1)
    for c in range(N - 1, -1, -1):
        end = min(n_cells, c - c % CHUNK + CHUNK)
        for s in range(K):
            nb = nbr[s, c]
            if c < nb < end:
                bwd[c] = max(bwd[c], bwd[nb] + 1)
    return fwd, bwd


def face_order(n_cells, owner, neighbour, perm):
    """New internal-face order: signed old face ids (+(f+1), negative = flipped), new owner / neighbour."""
    iperm = np.empty(n_cells, dtype=np.int64)
    iperm[perm] = np.arange(n_cells)
    nint = len(neighbour)
    o = iperm[owner[:nint]]
    n = iperm[neighbour]
    flip = o > n
    lo, hi = np.minimum(o, n), np.maximum(o, n)
    order = np.lexsort((hi, lo))
    signed = np.where(flip[order], -(order + 1), order + 1).astype(np.int32)
    return signed, lo[order].astype(np.int32), hi[order].astype(np.int32), iperm


def renumbered_mesh(rm: RefMesh, perm: np.ndarray) -> RefMesh:
    signed, no, nn, iperm = face_order(rm.n_cells, rm.owner, rm.neighbour, perm)
    nint = rm.n_internal
    old = np.abs(signed) - 1
    sg = np.sign(signed).astype(np.float64)
    owner = np.concatenate([no, iperm[rm.owner[nint:]].astype(np.int32)])
    Sf = np.concatenate([rm.Sf[old] * sg[:, None], rm.Sf[nint:]])
    Cf = np.concatenate([rm.Cf[old], rm.Cf[nint:]])
    w = np.concatenate([np.where(signed > 0, rm.weights[old], 1.0 - rm.weights[old]), rm.weights[nint:]])
    out = RefMesh(rm.n_cells, owner, nn, Sf, Cf, rm.C[perm], rm.V[perm], w, rm.nbr_C.copy(), list(rm.patches), list(rm.solved))
    out.face_addr = np.concatenate([signed, np.arange(nint, rm.n_faces, dtype=np.int32) + 1])
    out.cell_addr = perm.copy()
    return out


def ell_tables(rm: RefMesh, perm: np.ndarray):
    """(nbr[K,N], face[K,N]) as rheo_gpu_get_ell returns them."""
    signed, no, nn, iperm = face_order(rm.n_cells, rm.owner, rm.neighbour, perm)
    N, nint = rm.n_cells, rm.n_internal
    rows = [[] for _ in range(N)]
    for q in range(nint):
        rows[no[q]].append((int(nn[q]), q))
        rows[nn[q]].append((int(no[q]), ~q))
    for c in range(N):
        rows[c].sort(key=lambda t: t[0])
    h = 0
    for (ptype, start, size, _r, _a, _b) in rm.patches:
        for f in range(start, start + size):
            if ptype == PATCH_EMPTY:
                continue
            c = int(iperm[rm.owner[f]])
            b = f - nint
            if ptype == PATCH_PROCESSOR:
                rows[c].append((N + h, f))
                h += 1
            else:
                rows[c].append((-(b + 2), f))
    K = max(len(r) for r in rows)
    nbr = np.full((K, N), -1, dtype=np.int32)
    face = np.zeros((K, N), dtype=np.int32)
    for c, r in enumerate(rows):
        for s, (nb, fq) in enumerate(r):
            nbr[s, c] = nb
            face[s, c] = fq
    return nbr, face


# ------------------------------------------------------------------------------------------ decomposition
def simple_decomp(C: np.ndarray, n, delta=1e-3):
    px, py, pz = n
    d = 1 - 0.5 * delta * delta
    d2, a, a2 = d * d, delta, delta * delta
    R = np.array([[d2, -a * d, a], [a * d - a2 * d, a * a2 + d2, -2 * a * d], [a * d2 + a2, a * d - a2 * d, d2 - a2]])
    rp = C @ R.T
    ncell = len(C)
    out = np.zeros(ncell, dtype=np.int32)
    mult = [1, px, px * py]
    for dirn, npd in enumerate((px, py, pz)):
        order = np.argsort(rp[:, dirn], kind="stable")
        per, rem = divmod(ncell, npd)
        counts = [per + (1 if g < rem else 0) for g in range(npd)]
        group = np.repeat(np.arange(npd), counts)
        out[order] += (mult[dirn] * group).astype(np.int32)
    return out


def sub_mesh(rm: RefMesh, c2r: np.ndarray, n_ranks: int, rank: int) -> RefMesh:
    nint = rm.n_internal
    mine = np.nonzero(c2r == rank)[0]
    g2l = -np.ones(rm.n_cells, dtype=np.int64)
    g2l[mine] = np.arange(len(mine))
    ro, rn = c2r[rm.owner[:nint]], c2r[rm.neighbour]
    fin = np.nonzero((ro == rank) & (rn == rank))[0]
    owner = [g2l[rm.owner[fin]]]
    neigh = g2l[rm.neighbour[fin]]
    faces = [fin + 1]
    w = [rm.weights[fin]]
    nbrC = []
    patches = []
    pos = len(fin)
    for (ptype, start, size, _r, tb, ub) in rm.patches:
        fs = np.arange(start, start + size)
        fs = fs[c2r[rm.owner[fs]] == rank]
        owner.append(g2l[rm.owner[fs]])
        faces.append(fs + 1)
        w.append(rm.weights[fs])
        nbrC.append(np.zeros((len(fs), 3)))
        patches.append((ptype, pos, len(fs), -1, tb, ub))
        pos += len(fs)
    for r in range(n_ranks):
        if r == rank:
            continue
        a = np.nonzero((ro == rank) & (rn == r))[0]     # local cell is the global owner
        b = np.nonzero((ro == r) & (rn == rank))[0]     # local cell is the global neighbour -> flipped
        fs = np.concatenate([a, b])
        if len(fs) == 0:
            continue
        flip = np.concatenate([np.zeros(len(a), bool), np.ones(len(b), bool)])
        srt = np.argsort(fs, kind="stable")
        fs, flip = fs[srt], flip[srt]
        loc = np.where(flip, rm.neighbour[fs], rm.owner[fs])
        oth = np.where(flip, rm.owner[fs], rm.neighbour[fs])
        owner.append(g2l[loc])
        faces.append(np.where(flip, -(fs + 1), fs + 1))
        w.append(np.where(flip, 1.0 - rm.weights[fs], rm.weights[fs]))
        nbrC.append(rm.C[oth])
        patches.append((PATCH_PROCESSOR, pos, len(fs), r, 4, 4))
        pos += len(fs)
    faces = np.concatenate(faces).astype(np.int32)
    of = np.abs(faces) - 1
    sg = np.sign(faces).astype(np.float64)
    out = RefMesh(len(mine), np.concatenate(owner).astype(np.int32), neigh.astype(np.int32), rm.Sf[of] * sg[:, None], rm.Cf[of],
                  rm.C[mine], rm.V[mine], np.concatenate(w), np.concatenate(nbrC) if nbrC else np.zeros((0, 3)), patches, list(rm.solved))
    out.cell_addr = mine.astype(np.int32)
    out.face_addr = faces
    return out
