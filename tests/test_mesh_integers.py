"""Integer contracts of the host mesh services (CPU only): OpenFOAM face ordering, decomposePar `simple`,
processor sub-meshes, colour renumbering and ELL tables must agree BIT-EXACTLY with the independent numpy
restatement in oracle/mesh_ref.py (BASELINE.json: "integer mesh renumbering/addressing is bit-exact").

EXT-OF9 semantics restated: primitiveMesh upper-triangular order, simpleGeomDecomp, domainDecomposition
(SURVEY.md §8e); reference call sites that consume this addressing:
of90/src/libs/gaussDefCmpwConvectionScheme/gaussDefCmpwConvectionScheme.C:88-91,147-167.
"""
import ctypes as C
import gzip
import hashlib
import json
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import mesh_ref
from rheotool_b200 import abi, cases, mesh

GOLDEN = Path(__file__).resolve().parent / "golden"
ANEURYSM = Path("/root/reference/of90/tutorials/rheoFoam/Aneurysm/HerschelBulkley/constant/polyMesh.org")


def small_cases():
    return {
        "C1": cases.by_name("C1", 0.2),
        "C2": cases.by_name("C2", 1 / 9),
        "C3": cases.by_name("C3", 3 / 19),
        "C5": cases.by_name("C5", 12 / 400),
    }


@pytest.mark.parametrize("name", sorted(small_cases()))
def test_upper_triangular_order_and_closed_cells(name):
    m = mesh.tensor_grid(small_cases()[name].grid)
    own, nei = m.owner[: m.n_internal], m.neighbour
    assert (own < nei).all()
    key = own.astype(np.int64) * m.n_cells + nei
    assert (np.diff(key) > 0).all(), "internal faces sorted by (owner, neighbour)"
    # boundary faces grouped by patch, patches contiguous and covering all boundary faces
    pos = m.n_internal
    for p in m.patches:
        assert p.start == pos
        pos += p.size
    assert pos == m.n_faces
    # closed cells: sum of outward face area vectors vanishes; volumes positive and add up
    acc = np.zeros((m.n_cells, 3))
    np.add.at(acc, m.owner, m.Sf)
    np.subtract.at(acc, nei, m.Sf[: m.n_internal])
    scale = np.abs(m.Sf).max()
    assert np.abs(acc).max() < 1e-12 * scale
    assert (m.V > 0).all()
    w = m.weights[: m.n_internal]
    assert ((w > 0) & (w < 1)).all()


def test_geometry_matches_primitive_mesh_formulas_on_uniform_grid():
    spec = cases.by_name("C5", 8 / 400)
    m = mesh.tensor_grid(spec.grid)
    h = 1.0 / 8
    assert np.allclose(m.V, h ** 3, rtol=1e-13)
    assert np.allclose(np.linalg.norm(m.Sf, axis=1), h * h, rtol=1e-13)
    assert np.allclose(m.weights[: m.n_internal], 0.5, rtol=1e-13)
    assert m.V.sum() == pytest.approx(1.0, rel=1e-13)
    d = m.C[m.neighbour] - m.C[m.owner[: m.n_internal]]
    assert np.allclose(np.linalg.norm(d, axis=1), h, rtol=1e-12)
    # Sf points from owner to neighbour
    assert (np.einsum("ij,ij->i", d, m.Sf[: m.n_internal]) > 0).all()


@pytest.mark.parametrize("n", [(2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 1, 2)])
def test_simple_decomposition_bit_exact(n):
    m = mesh.tensor_grid(cases.by_name("C3", 3 / 19).grid)
    got = m.simple_decomp(*n)
    ref = mesh_ref.simple_decomp(m.C, n)
    assert np.array_equal(got, ref)
    counts = np.bincount(got, minlength=n[0] * n[1] * n[2])
    assert counts.max() - counts.min() <= max(n) * 2 + counts.max() // 20   # balanced like simpleGeomDecomp


@pytest.mark.parametrize("name,n", [("C2", (2, 1, 1)), ("C3", (2, 2, 1)), ("C5", (2, 2, 2))])
def test_processor_meshes_bit_exact_and_round_trip(name, n):
    m = mesh.tensor_grid(small_cases()[name].grid)
    rm = mesh_ref.from_host_mesh(m)
    nr = n[0] * n[1] * n[2]
    c2r = m.simple_decomp(*n)
    seen_cells = np.zeros(m.n_cells, dtype=int)
    seen_faces = np.zeros(m.n_faces, dtype=int)
    subs = []
    for r in range(nr):
        sub = m.decompose(c2r, nr, r)
        ref = mesh_ref.sub_mesh(rm, c2r, nr, r)
        ca, fa = sub.proc_addressing()
        assert np.array_equal(ca, ref.cell_addr) and np.array_equal(fa, ref.face_addr)
        assert np.array_equal(sub.owner, ref.owner) and np.array_equal(sub.neighbour, ref.neighbour)
        got_p = [(p.type, p.start, p.size, p.nbr_rank) for p in sub.patches]
        assert got_p == [(p[0], p[1], p[2], p[3]) for p in ref.patches]
        assert np.array_equal(sub.Sf, ref.Sf) and np.array_equal(sub.weights, ref.weights) and np.array_equal(sub.V, ref.V)
        seen_cells[ca] += 1
        np.add.at(seen_faces, np.abs(fa) - 1, 1)
        subs.append((sub, ca, fa))
    assert (seen_cells == 1).all()                                   # cellProcAddressing is a partition
    assert (seen_faces[m.n_internal:] == 1).all()                    # every boundary face on exactly one rank
    assert set(np.unique(seen_faces[: m.n_internal])) <= {1, 2}      # internal: once, or twice when cut
    # the two sides of every processor patch list the same global faces in the same order, flipped
    for r, (sub, ca, fa) in enumerate(subs):
        for p in sub.patches:
            if p.type != abi.PATCH_PROCESSOR:
                continue
            other, oca, ofa = subs[p.nbr_rank]
            q = next(x for x in other.patches if x.type == abi.PATCH_PROCESSOR and x.nbr_rank == r)
            mine, theirs = fa[p.start:p.start + p.size], ofa[q.start:q.start + q.size]
            assert np.array_equal(mine, -theirs)
            # halo geometry: the centre of the cell across the face is the neighbour's own cell centre
            b0 = p.start - sub.n_internal
            assert np.array_equal(sub.nbr_C[b0:b0 + p.size], other.C[other.owner[q.start:q.start + q.size]])


@pytest.mark.parametrize("name,n", [("C2", (3, 1, 1)), ("C3", (2, 2, 1)), ("C5", (2, 2, 2))])
def test_direct_part_generation_equals_decomposition_of_the_global_mesh(name, n):
    """rheo_mesh_tensor_grid_part never builds the global mesh (64 M cells would not fit on every rank)."""
    spec = small_cases()[name]
    m = mesh.tensor_grid(spec.grid)
    nr = n[0] * n[1] * n[2]
    parts = [mesh.tensor_grid_part(spec.grid, *n, r) for r in range(nr)]
    c2r = np.full(m.n_cells, -1, dtype=np.int32)
    for r, pm in enumerate(parts):
        c2r[pm.global_cells()] = r
    assert (c2r >= 0).all()
    for r, pm in enumerate(parts):
        sub = m.decompose(c2r, nr, r)
        assert np.array_equal(pm.owner, sub.owner) and np.array_equal(pm.neighbour, sub.neighbour)
        assert [(p.type, p.start, p.size, p.nbr_rank, p.theta_bc, p.tau_bc) for p in pm.patches] == \
               [(p.type, p.start, p.size, p.nbr_rank, p.theta_bc, p.tau_bc) for p in sub.patches]
        # integers are bit-exact; geometry is recomputed from the same points but sums the faces of a cell in
        # the local face order, so it may differ from the global mesh in the last bits
        for a, b in ((pm.Sf, sub.Sf), (pm.Cf, sub.Cf), (pm.C, sub.C), (pm.V, sub.V), (pm.weights, sub.weights), (pm.nbr_C, sub.nbr_C)):
            assert np.allclose(a, b, rtol=1e-11, atol=1e-13 * max(1.0, np.abs(b).max()))


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_colour_renumbering_bit_exact(name):
    m = mesh.tensor_grid(small_cases()[name].grid)
    perm, colour, cstart = m.colour_renumber()
    rperm, rcolour, rcstart = mesh_ref.colour_renumber(m.n_cells, m.owner[: m.n_internal], m.neighbour)
    assert np.array_equal(perm, rperm) and np.array_equal(colour, rcolour) and np.array_equal(cstart, rcstart)
    assert len(cstart) - 1 == 2, "hex meshes are two-colourable (red-black)"
    # proper colouring, and perm is a permutation sorted by (colour, old index)
    assert (colour[m.owner[: m.n_internal]] != colour[m.neighbour]).all()
    assert np.array_equal(np.sort(perm), np.arange(m.n_cells))
    key = colour[perm].astype(np.int64) * m.n_cells + perm
    assert (np.diff(key) > 0).all()


@pytest.mark.parametrize("name,scale", [("C1", 0.3), ("C2", 2 / 9), ("C3", 5 / 19), ("C4", 30 / 252), ("C5", 32 / 400), ("C5", 0.1)])
def test_block_renumbering_bit_exact(name, scale):
    """The device's block ordering for PBiCGStab on lattice meshes (csrc/host/ordering.hpp) against its numpy restatement:
    permutation and chunk-colour offsets bit-exact; chunks of one colour share no face; natural order inside a chunk's blocks."""
    from rheotool_b200 import cases
    m = mesh.tensor_grid(cases.by_name(name, scale).grid)
    got = m.block_renumber()
    assert got is not None, "tensor grids are lattice meshes"
    perm, cstart, tile = got
    rm = mesh_ref.from_host_mesh(m)
    rperm, rcstart, rtile = mesh_ref.block_renumber(rm)
    assert np.array_equal(perm, rperm) and np.array_equal(cstart, rcstart) and tile == rtile
    assert np.array_equal(np.sort(perm), np.arange(m.n_cells))
    assert all(int(c) % 32 == 0 for c in cstart[:-1]) and cstart[-1] == m.n_cells
    # chunks of one colour are pairwise non-adjacent
    iperm = np.empty(m.n_cells, dtype=np.int64); iperm[perm] = np.arange(m.n_cells)
    o, n = iperm[m.owner[: m.n_internal]], iperm[m.neighbour]
    colour_of = np.searchsorted(cstart, np.arange(m.n_cells), side="right") - 1
    cut = (o >> 5) != (n >> 5)
    assert (colour_of[o[cut]] != colour_of[n[cut]]).all()
    if name == "C5" and scale == 32 / 400:
        assert len(cstart) - 1 == 2 and tile == (4, 4, 2), "a box of whole 4x4x2 blocks is two-colourable"
    # levels of the in-chunk dependency graphs: a chain can only grow by one per neighbour
    nbr, _ = mesh_ref.ell_tables(rm, perm)
    fwd, bwd = mesh_ref.chunk_levels(nbr, m.n_cells)
    assert fwd.max() <= 31 and bwd.max() <= 31
    if tile == (4, 4, 2) and len(cstart) - 1 == 2:
        assert fwd.max() == 3 + 3 + 1 and bwd.max() == 3 + 3 + 1


def test_block_renumbering_declines_unstructured_meshes():
    ncell, own, nei = _prism_like_mesh()
    m, _ = _mesh_from_addressing(ncell, own, nei)
    assert m.block_renumber() is None


def _prism_like_mesh(nx=7, ny=6, nz=3, seed=3):
    """An unstructured, NOT two-colourable addressing: a hex grid with the squares of every z-layer split
    into two triangular prisms (odd cycles), cells randomly renumbered.  Geometry is irrelevant here."""
    rng = np.random.default_rng(seed)
    ncell = 2 * nx * ny * nz
    new_id = rng.permutation(ncell)

    def cid(i, j, k, t):
        return new_id[((k * ny + j) * nx + i) * 2 + t]

    faces = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                a, b = cid(i, j, k, 0), cid(i, j, k, 1)
                faces.append((a, b))                                  # diagonal
                if i + 1 < nx:
                    faces.append((b, cid(i + 1, j, k, 0)))            # +x
                if j + 1 < ny:
                    faces.append((b, cid(i, j + 1, k, 0)))            # +y
                if k + 1 < nz:
                    faces.append((a, cid(i, j, k + 1, 0)))
                    faces.append((b, cid(i, j, k + 1, 1)))
                if i + 1 < nx and j + 1 < ny:
                    faces.append((a, cid(i + 1, j + 1, k, 1)))        # extra diagonal link -> triangles in the graph
    f = np.array([(min(p), max(p)) for p in faces], dtype=np.int64)
    f = f[np.lexsort((f[:, 1], f[:, 0]))]
    return ncell, f[:, 0].astype(np.int32), f[:, 1].astype(np.int32)


def _mesh_from_addressing(ncell, own, nei):
    nint = len(own)
    rm = mesh_ref.RefMesh(ncell, own.copy(), nei.copy(), np.ones((nint, 3)), np.zeros((nint, 3)), np.zeros((ncell, 3)),
                          np.ones(ncell), np.full(nint, 0.5), np.zeros((0, 3)), [], [1] * 6)
    d = mesh_ref.to_desc(rm, abi)
    h = abi.lib().rheo_mesh_from_desc(C.byref(d))
    return mesh.HostMesh(h), rm


def test_colour_renumbering_unstructured_multicolour():
    ncell, own, nei = _prism_like_mesh()
    m, rm = _mesh_from_addressing(ncell, own, nei)
    perm, colour, cstart = m.colour_renumber()
    rperm, rcolour, rcstart = mesh_ref.colour_renumber(ncell, own, nei)
    assert np.array_equal(perm, rperm) and np.array_equal(colour, rcolour) and np.array_equal(cstart, rcstart)
    assert len(cstart) - 1 >= 3
    assert (colour[own] != colour[nei]).all()


def _read_label_list(path: Path) -> np.ndarray:
    txt = gzip.open(path, "rt").read() if path.suffix == ".gz" else path.read_text()
    mm = re.search(r"\n(\d+)\s*\n?\(", txt)
    n = int(mm.group(1))
    body = txt[mm.end(): txt.rindex(")")]
    a = np.array(body.split(), dtype=np.int64)
    assert len(a) == n
    return a.astype(np.int32)


def _digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int32).tobytes()).hexdigest()


def test_aneurysm_polymesh_fixture_against_golden():
    """The only polyMesh the reference ships (279,966 cells, unstructured): its renumbering
    must reproduce the digests committed in tests/golden/aneurysm_renumber.json, which were generated by
    the numpy restatement (tools/make_golden_aneurysm.py).  Skipped where /root/reference is absent."""
    if not ANEURYSM.exists():
        pytest.skip("/root/reference not present on this machine")
    gold = json.loads((GOLDEN / "aneurysm_renumber.json").read_text())
    own = _read_label_list(ANEURYSM / "owner.gz")
    nei = _read_label_list(ANEURYSM / "neighbour.gz")
    nint = len(nei)
    ncell = int(own.max()) + 1
    assert (ncell, len(own), nint) == (gold["n_cells"], gold["n_faces"], gold["n_internal_faces"])
    m, _ = _mesh_from_addressing(ncell, own[:nint], nei)
    perm, colour, cstart = m.colour_renumber()
    assert len(cstart) - 1 == gold["n_colours"]
    assert list(map(int, cstart)) == gold["colour_start"]
    assert _digest(perm) == gold["perm_sha256"] and _digest(colour) == gold["colour_sha256"]
