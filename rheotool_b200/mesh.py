"""Host mesh objects (numpy views over the C++ RheoHostMesh) and the blockMesh-lite helpers."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import abi


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def graded_points(x0: float, x1: float, n: int, ratio: float = 1.0) -> np.ndarray:
    """blockMesh `simpleGrading`: n cells between x0 and x1, last/first cell-size ratio = `ratio`
    (EXT-OF9 lineDivide: geometric progression with factor ratio**(1/(n-1)))."""
    if n == 1 or abs(ratio - 1.0) < 1e-12:
        return np.linspace(x0, x1, n + 1)
    r = ratio ** (1.0 / (n - 1))
    i = np.arange(n + 1, dtype=np.float64)
    s = (1.0 - r ** i) / (1.0 - r ** n)
    return x0 + (x1 - x0) * s


def multi_graded(segments) -> np.ndarray:
    """Concatenate graded segments [(x0, x1, n, ratio), ...] into one coordinate array."""
    out = [np.array([segments[0][0]], dtype=np.float64)]
    for (a, b, n, r) in segments:
        out.append(graded_points(a, b, n, r)[1:])
    return np.concatenate(out)


@dataclass
class PatchSpec:
    name: str
    type: int
    theta_bc: int
    tau_bc: int


@dataclass
class GridSpec:
    """Arguments of rheo_mesh_tensor_grid (a multi-block blockMeshDict without curved edges)."""
    xs: np.ndarray
    ys: np.ndarray
    zs: np.ndarray
    boxes: list            # [(i0,i1,j0,j1,k0,k1)]
    patches: list          # [PatchSpec]
    rules: list            # [(patch index, (lo3), (hi3))]
    default_patch: int
    two_d: bool
    tol: float = 1e-9


class HostMesh:
    """Owns a RheoHostMesh*; exposes its arrays as numpy views (valid while this object lives)."""

    def __init__(self, handle, patch_names=None):
        if not handle:
            raise RuntimeError("mesh construction failed: " + abi.lib().rheo_mesh_last_error().decode())
        self._h = handle
        self.desc = abi.RheoMeshDesc()
        abi.lib().rheo_mesh_desc(self._h, C.byref(self.desc))
        d = self.desc
        self.n_cells, self.n_faces, self.n_internal = d.n_cells, d.n_faces, d.n_internal_faces
        self.n_boundary = self.n_faces - self.n_internal
        as_np = np.ctypeslib.as_array
        self.owner = as_np(d.owner, (self.n_faces,))
        self.neighbour = as_np(d.neighbour, (self.n_internal,)) if self.n_internal else np.zeros(0, dtype=np.int32)
        self.Sf = as_np(d.Sf, (self.n_faces, 3))
        self.Cf = as_np(d.Cf, (self.n_faces, 3))
        self.C = as_np(d.C, (self.n_cells, 3))
        self.V = as_np(d.V, (self.n_cells,))
        self.weights = as_np(d.weights, (self.n_faces,))
        self.nbr_C = as_np(d.nbr_C, (self.n_boundary, 3)) if self.n_boundary else np.zeros((0, 3))
        self.patches = [d.patches[i] for i in range(d.n_patches)]
        self.solved = [int(v) for v in d.solved_components]
        self.patch_names = list(patch_names) if patch_names else []
        while len(self.patch_names) < len(self.patches):
            p = self.patches[len(self.patch_names)]
            self.patch_names.append(f"procBoundaryTo{p.nbr_rank}" if p.type == abi.PATCH_PROCESSOR else f"patch{len(self.patch_names)}")

    def __del__(self):
        try:
            if self._h:
                abi.lib().rheo_mesh_free(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def desc_ptr(self):
        return C.byref(self.desc)

    # ---- decomposition -------------------------------------------------------------------------
    def simple_decomp(self, px: int, py: int, pz: int, delta: float = 1e-3) -> np.ndarray:
        out = np.zeros(self.n_cells, dtype=np.int32)
        rc = abi.lib().rheo_mesh_simple_decomp(self._h, px, py, pz, delta, _ptr(out))
        if rc:
            raise RuntimeError(abi.lib().rheo_mesh_last_error().decode())
        return out

    def decompose(self, cell_to_rank: np.ndarray, n_ranks: int, rank: int) -> "HostMesh":
        c2r = np.ascontiguousarray(cell_to_rank, dtype=np.int32)
        sub = HostMesh(abi.lib().rheo_mesh_decompose(self._h, _ptr(c2r), n_ranks, rank), self.patch_names[: len([p for p in self.patches if p.type != abi.PATCH_PROCESSOR])])
        return sub

    def proc_addressing(self):
        ca = np.zeros(self.n_cells, dtype=np.int32)
        fa = np.zeros(self.n_faces, dtype=np.int32)
        rc = abi.lib().rheo_mesh_proc_addressing(self._h, _ptr(ca), _ptr(fa))
        if rc == 2:
            return ca, None
        return ca, fa

    def global_cells(self) -> np.ndarray:
        ca = np.zeros(self.n_cells, dtype=np.int32)
        abi.lib().rheo_mesh_proc_addressing(self._h, _ptr(ca), None)
        return ca

    def colour_renumber(self):
        perm = np.zeros(self.n_cells, dtype=np.int32)
        colour = np.zeros(self.n_cells, dtype=np.int32)
        cstart = np.zeros(65, dtype=np.int32)
        nc = abi.lib().rheo_mesh_colour_renumber(self._h, _ptr(perm), _ptr(colour), _ptr(cstart))
        if nc < 0:
            raise RuntimeError(abi.lib().rheo_mesh_last_error().decode())
        return perm, colour, cstart[: nc + 1].copy()

    def block_renumber(self):
        """(perm, colour_start, tile) of the block ordering, or None when the mesh is not a lattice mesh."""
        perm = np.zeros(self.n_cells, dtype=np.int32)
        cstart = np.zeros(65, dtype=np.int32)
        tile = np.zeros(3, dtype=np.int32)
        nc = abi.lib().rheo_mesh_block_renumber(self._h, _ptr(perm), _ptr(cstart), _ptr(tile))
        if nc < 0:
            raise RuntimeError(abi.lib().rheo_mesh_last_error().decode())
        if nc == 0:
            return None
        return perm, cstart[: nc + 1].copy(), tuple(int(t) for t in tile)

    # ---- synthetic fields ------------------------------------------------------------------------
    def synth_fields(self, spec: abi.RheoSynthSpec):
        U = np.zeros((self.n_cells, 3))
        Ub = np.zeros((self.n_boundary, 3))
        phi = np.zeros(self.n_faces)
        theta0 = np.zeros((self.n_cells, 6))
        rc = abi.lib().rheo_synth_fields(self._h, C.byref(spec), None, _ptr(U), _ptr(Ub), _ptr(phi), _ptr(theta0))
        if rc:
            raise RuntimeError(abi.lib().rheo_mesh_last_error().decode())
        return U, Ub, phi, theta0

    def max_courant_rate(self, phi: np.ndarray) -> float:
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        return float(abi.lib().rheo_mesh_max_courant_rate(self._h, _ptr(phi)))


def _grid_args(spec: GridSpec):
    xs = np.ascontiguousarray(spec.xs, dtype=np.float64)
    ys = np.ascontiguousarray(spec.ys, dtype=np.float64)
    zs = np.ascontiguousarray(spec.zs, dtype=np.float64)
    boxes = np.ascontiguousarray(np.array(spec.boxes, dtype=np.int32).reshape(-1, 6))
    ps = (abi.RheoPatchSpec * len(spec.patches))()
    for i, p in enumerate(spec.patches):
        ps[i].type, ps[i].theta_bc, ps[i].tau_bc = p.type, p.theta_bc, p.tau_bc
    rs = (abi.RheoPatchRule * max(len(spec.rules), 1))()
    for i, (pi, lo, hi) in enumerate(spec.rules):
        rs[i].patch = pi
        for d in range(3):
            rs[i].lo[d] = lo[d]
            rs[i].hi[d] = hi[d]
    keep = (xs, ys, zs, boxes, ps, rs)
    args = [len(xs) - 1, len(ys) - 1, len(zs) - 1, _ptr(xs), _ptr(ys), _ptr(zs), len(boxes), _ptr(boxes),
            len(spec.patches), C.cast(ps, C.c_void_p), len(spec.rules), C.cast(rs, C.c_void_p), spec.default_patch,
            spec.tol, 1 if spec.two_d else 0]
    return args, keep


def tensor_grid(spec: GridSpec) -> HostMesh:
    args, keep = _grid_args(spec)
    return HostMesh(abi.lib().rheo_mesh_tensor_grid(*args), [p.name for p in spec.patches])


def tensor_grid_part(spec: GridSpec, px: int, py: int, pz: int, rank: int) -> HostMesh:
    args, keep = _grid_args(spec)
    return HostMesh(abi.lib().rheo_mesh_tensor_grid_part(*args, px, py, pz, rank), [p.name for p in spec.patches])
