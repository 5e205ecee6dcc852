"""OpenFOAM on-disk formats either side of the stress step (include/rheo_io.h): polyMesh directories
and vol<Type>Field files, ASCII, plain or gzip.  Thin ctypes mirror of the C++ implementation (csrc/host/foamfile.cpp)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import abi
from .mesh import HostMesh

CLASS_OF_NCOMP = {1: "volScalarField", 3: "volVectorField", 6: "volSymmTensorField", 9: "volTensorField"}


class FoamError(RuntimeError):
    pass


def _err() -> str:
    return abi.lib().rheo_mesh_last_error().decode()


def read_polymesh(directory) -> HostMesh:
    """constant/polyMesh -> HostMesh (geometry computed like EXT-OF9 primitiveMesh); patch names from the boundary file."""
    h = abi.lib().rheo_io_read_polymesh(str(directory).encode())
    if not h:
        raise FoamError(_err())
    m = HostMesh(h)
    buf = C.create_string_buffer(256)
    m.patch_names = []
    for p in range(len(m.patches)):
        abi.lib().rheo_io_patch_name(h, p, buf, len(buf))
        m.patch_names.append(buf.value.decode())
    return m


def write_polymesh(m: HostMesh, directory, gz: bool = False):
    Path(directory).mkdir(parents=True, exist_ok=True)
    for p, name in enumerate(m.patch_names[: len(m.patches)]):
        abi.lib().rheo_io_set_patch_name(m.handle, p, name.encode())
    if abi.lib().rheo_io_write_polymesh(m.handle, str(directory).encode(), 1 if gz else 0):
        raise FoamError(_err())


def mesh_counts(m: HostMesh) -> tuple[int, int]:
    a, b = C.c_int64(0), C.c_int64(0)
    abi.lib().rheo_io_mesh_counts(m.handle, C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


class FoamField:
    """A parsed vol<Type>Field file."""

    def __init__(self, path):
        self._h = abi.lib().rheo_io_read_field(str(path).encode())
        if not self._h:
            raise FoamError(_err())
        cls, obj = C.create_string_buffer(64), C.create_string_buffer(128)
        nc, uni, n = C.c_int32(0), C.c_int32(0), C.c_int64(0)
        abi.lib().rheo_io_field_info(self._h, cls, len(cls), obj, len(obj), C.byref(nc), C.byref(uni), C.byref(n))
        self.cls, self.object, self.n_comp = cls.value.decode(), obj.value.decode(), int(nc.value)
        self.internal_uniform, self.n_internal = bool(uni.value), int(n.value)

    def __del__(self):
        try:
            if self._h:
                abi.lib().rheo_io_field_free(self._h)
                self._h = None
        except Exception:
            pass

    def internal(self, n_cells: int) -> np.ndarray:
        out = np.zeros((n_cells, self.n_comp))
        if abi.lib().rheo_io_field_internal(self._h, n_cells, out.ctypes.data_as(C.c_void_p)):
            raise FoamError(_err())
        return out

    def patch(self, name: str, n_faces: int = 0):
        """(type, values or None) of the boundaryField entry that applies to patch `name` (regex keys honoured)."""
        ty = C.create_string_buffer(128)
        has = C.c_int32(0)
        vals = np.zeros((n_faces, self.n_comp))
        rc = abi.lib().rheo_io_field_patch(self._h, name.encode(), n_faces, ty, len(ty), C.byref(has), vals.ctypes.data_as(C.c_void_p))
        if rc:
            raise (KeyError if rc == 2 else FoamError)(_err())
        return ty.value.decode(), (vals if has.value else None)

    def apply_bcs(self, m: HostMesh, which: str):
        """Set theta_bc / tau_bc of the mesh's patches from this file's boundaryField types."""
        for p, name in enumerate(m.patch_names[: len(m.patches)]):
            abi.lib().rheo_io_set_patch_name(m.handle, p, name.encode())
        if abi.lib().rheo_io_apply_field_bcs(m.handle, self._h, {"theta": 0, "tau": 1}[which]):
            raise FoamError(_err())
        abi.lib().rheo_mesh_desc(m.handle, C.byref(m.desc))
        m.patches = [m.desc.patches[i] for i in range(m.desc.n_patches)]


def write_field(path, obj: str, internal: np.ndarray, patches: list, dimensions: str = "[0 0 0 0 0 0 0]", gz: bool = False):
    """patches: [(name, type, values or None)]; values [n_faces, n_comp].  The class follows from the component count."""
    a = np.ascontiguousarray(internal, dtype=np.float64)
    a = a.reshape(len(a), -1)
    nc = a.shape[1]
    names = (C.c_char_p * len(patches))(*[p[0].encode() for p in patches])
    types = (C.c_char_p * len(patches))(*[p[1].encode() for p in patches])
    keep = [None if p[2] is None else np.ascontiguousarray(p[2], dtype=np.float64).reshape(-1, nc) for p in patches]
    sizes = (C.c_int32 * len(patches))(*[0 if v is None else len(v) for v in keep])
    vals = (C.c_void_p * len(patches))(*[None if v is None else v.ctypes.data for v in keep])
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    if abi.lib().rheo_io_write_field(str(path).encode(), CLASS_OF_NCOMP[nc].encode(), obj.encode(), dimensions.encode(), nc, len(a),
                                     a.ctypes.data_as(C.c_void_p), len(patches), names, types, sizes, vals, 1 if gz else 0):
        raise FoamError(_err())
