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


# ---------------------------------------------------------------- whole cases (constant/polyMesh + a time directory)
BC_WORD = {abi.BC_FIXED_VALUE: "fixedValue", abi.BC_ZERO_GRADIENT: "zeroGradient", abi.BC_LINEAR_EXTRAPOLATION: "linearExtrapolation",
           abi.BC_EMPTY: "empty", abi.BC_PROCESSOR: "processor",
           # the writer prints `type <word>;`: the regression flavour carries its extra keyword along (linearExtrapolationFvPatchField.C:230)
           abi.BC_LINEAR_EXTRAPOLATION_REG: "linearExtrapolation;\n        useRegression   true"}
DIMENSIONS = {"tau": "[1 -1 -2 0 0 0 0]", "U": "[0 1 -1 0 0 0 0]"}


def surface_flux(m: HostMesh, U: np.ndarray, U_b: np.ndarray) -> np.ndarray:
    """phi = linearInterpolate(U) & Sf  (what createPhi.H computes for an incompressible solver: EXT-OF9
    surfaceInterpolationScheme::interpolate, w U_P + (1 - w) U_N on internal faces, patch values on the boundary)."""
    n = m.n_internal
    w = m.weights[:n, None]
    Uf = w * (U[m.owner[:n]] - U[m.neighbour]) + U[m.neighbour]
    return np.concatenate([(Uf * m.Sf[:n]).sum(1), (U_b * m.Sf[n:]).sum(1)])


def write_case(case_dir, m: HostMesh, time: str, theta, tau, U, U_b, theta_b=None, tau_b=None, eigvals=None, eigvecs=None, name: str = "", gz: bool = False,
               Phi=None, Phi_b=None):
    """constant/polyMesh + <time>/{U, theta<name>, tau<name>[, eigVals<name>, eigVecs<name>]} as rheoFoam reads / writes them
    (CE/Oldroyd-B/Oldroyd-BLog/Oldroyd_BLog.C:52-113).  Boundary types come from the mesh's patch kinds and theta/tau BCs."""
    case_dir = Path(case_dir)
    write_polymesh(m, case_dir / "constant" / "polyMesh", gz=gz)
    nint = m.n_internal

    def patches(bc_of, values, fixed_word="fixedValue"):
        out = []
        for pname, p in zip(m.patch_names, m.patches):
            sl = slice(p.start - nint, p.start - nint + p.size)
            if p.type == abi.PATCH_EMPTY:
                out.append((pname, "empty", None))
            elif p.type == abi.PATCH_PROCESSOR:
                out.append((pname, "processor", None if values is None else values[sl]))
            else:
                word = BC_WORD[bc_of(p)] if bc_of else fixed_word
                has_value = word == "fixedValue" or word.startswith("linearExtrapolation")
                out.append((pname, word, values[sl] if (has_value and values is not None) else None))
        return out

    tdir = case_dir / time
    write_field(tdir / "U", "U", U, patches(None, U_b), DIMENSIONS["U"], gz=gz)
    nb = m.n_boundary
    write_field(tdir / f"theta{name}", f"theta{name}", theta, patches(lambda p: p.theta_bc, theta_b if theta_b is not None else np.zeros((nb, 6))), gz=gz)
    write_field(tdir / f"tau{name}", f"tau{name}", tau, patches(lambda p: p.tau_bc, tau_b if tau_b is not None else np.zeros((nb, 6))), DIMENSIONS["tau"], gz=gz)
    calc = [(pname, "empty" if p.type == abi.PATCH_EMPTY else "zeroGradient", None) for pname, p in zip(m.patch_names, m.patches)]
    if eigvals is not None:
        write_field(tdir / f"eigVals{name}", f"eigVals{name}", eigvals, calc, gz=gz)
    if eigvecs is not None:
        write_field(tdir / f"eigVecs{name}", f"eigVecs{name}", eigvecs, calc, gz=gz)
    if Phi is not None:   # BMPLog's fluidity (BMPLog.C:112-122): the BC kinds are theta's
        write_field(tdir / f"Phi{name}", f"Phi{name}", np.asarray(Phi).reshape(-1, 1),
                    patches(lambda p: p.theta_bc, np.asarray(Phi_b).reshape(-1, 1) if Phi_b is not None else np.zeros((nb, 1))), "[-1 1 1 0 0 0 0]", gz=gz)


def read_case(case_dir, time: str, name: str = ""):
    """-> (mesh with theta/tau BCs set from the field files, dict of arrays: U, U_b, phi, theta, theta_b, tau, tau_b, eigvals, eigvecs).
    eigVals/eigVecs are READ_IF_PRESENT (None when absent: the model then starts from the identity, Oldroyd_BLog.C:76-113)."""
    case_dir = Path(case_dir)
    m = read_polymesh(case_dir / "constant" / "polyMesh")
    tdir = case_dir / time
    nint, nb = m.n_internal, m.n_boundary

    def load(fname, ncomp, required=True):
        path = tdir / fname
        if not (path.exists() or Path(str(path) + ".gz").exists()):
            if required:
                raise FoamError(f"{path}: MUST_READ field is missing")
            return None, None, None
        f = FoamField(path)
        if f.n_comp != ncomp:
            raise FoamError(f"{path}: class {f.cls} where {CLASS_OF_NCOMP[ncomp]} is expected")
        internal = f.internal(m.n_cells)
        bvals = np.zeros((nb, ncomp))
        for pname, p in zip(m.patch_names, m.patches):
            if p.type == abi.PATCH_EMPTY or p.size == 0:
                continue
            _, v = f.patch(pname, p.size)
            if v is not None:
                bvals[p.start - nint: p.start - nint + p.size] = v
        return f, internal, bvals

    fU, U, U_b = load("U", 3)
    fth, theta, theta_b = load(f"theta{name}", 6)
    fta, tau, tau_b = load(f"tau{name}", 6)
    fth.apply_bcs(m, "theta")
    fta.apply_bcs(m, "tau")
    _, eigvals, _ = load(f"eigVals{name}", 9, required=False)
    _, eigvecs, _ = load(f"eigVecs{name}", 9, required=False)
    _, Phi, Phi_b = load(f"Phi{name}", 1, required=False)   # BMPLog's fluidity (MUST_READ there: the caller checks)
    # velocity on patches without a value entry (zeroGradient outlets): the internal value
    for pname, p in zip(m.patch_names, m.patches):
        if p.type == abi.PATCH_EMPTY or p.size == 0:
            continue
        ty, v = fU.patch(pname, p.size)
        if v is None:
            U_b[p.start - nint: p.start - nint + p.size] = U[m.owner[p.start: p.start + p.size]]
    return m, {"U": U, "U_b": U_b, "phi": surface_flux(m, U, U_b), "theta": theta, "theta_b": theta_b, "tau": tau, "tau_b": tau_b,
               "eigvals": eigvals, "eigvecs": eigvecs, "Phi": None if Phi is None else Phi[:, 0], "Phi_b": None if Phi_b is None else Phi_b[:, 0]}


# ---------------------------------------------------------------- decomposed cases (processorN/ directories, decomposePar layout)
def write_decomposed_case(case_dir, m: HostMesh, cell_to_rank: np.ndarray, time: str, fields: dict, name: str = "", gz: bool = False):
    """processor<r>/constant/polyMesh (+ cell/face/boundaryProcAddressing) and processor<r>/<time>/ fields for every rank of
    `cell_to_rank` — what `decomposePar` leaves for `mpirun -np N rheoFoam -parallel`.
    fields: theta, tau, U, U_b (+ optional theta_b, tau_b, eigvals, eigvecs) of the undecomposed mesh."""
    case_dir = Path(case_dir)
    n_ranks = int(np.max(cell_to_rank)) + 1
    nint = m.n_internal
    subs = []
    for r in range(n_ranks):
        sub = m.decompose(cell_to_rank, n_ranks, r)
        ca, fa = sub.proc_addressing()
        gb = (np.abs(fa) - 1)[sub.n_internal:] - nint       # undecomposed boundary face of each local boundary face (< 0: processor face)

        def bnd(key, ncomp):
            v = fields.get(key)
            if v is None:
                return None
            out = np.zeros((sub.n_boundary, ncomp))
            out[gb >= 0] = v[gb[gb >= 0]]
            return out

        def cells(key):
            v = fields.get(key)
            return None if v is None else v[ca]

        write_case(case_dir / f"processor{r}", sub, time, cells("theta"), cells("tau"), cells("U"), bnd("U_b", 3), theta_b=bnd("theta_b", 6),
                   tau_b=bnd("tau_b", 6), eigvals=cells("eigvals"), eigvecs=cells("eigvecs"), name=name, gz=gz)
        subs.append(sub)
    return subs


def read_decomposed_case(case_dir, time: str, name: str = ""):
    """-> [(mesh, fields)] per rank, processor patches completed with the cell centres across them (read from the
    neighbour's directory — inside an MPI run each rank would receive them from its neighbour at start-up), and phi taken
    consistently on both sides of every processor face."""
    case_dir = Path(case_dir)
    n_ranks = len([p for p in case_dir.iterdir() if p.is_dir() and p.name.startswith("processor")])
    ranks = [read_case(case_dir / f"processor{r}", time, name) for r in range(n_ranks)]
    for r, (m, f) in enumerate(ranks):
        for pi, p in enumerate(m.patches):
            if p.type != abi.PATCH_PROCESSOR:
                continue
            om, of = ranks[p.nbr_rank]
            q = next(x for x in om.patches if x.type == abi.PATCH_PROCESSOR and x.nbr_rank == r)
            if q.size != p.size:
                raise FoamError(f"processor patches {r}<->{p.nbr_rank} differ in size")
            centres = np.ascontiguousarray(om.C[om.owner[q.start: q.start + q.size]])
            if abi.lib().rheo_io_set_nbr_centres(m.handle, pi, centres.ctypes.data_as(C.c_void_p)):
                raise FoamError(_err())
    # weights changed on the processor patches: flux through them = linear interpolation of the two cell values
    for r, (m, f) in enumerate(ranks):
        nint = m.n_internal
        for p in m.patches:
            if p.type != abi.PATCH_PROCESSOR:
                continue
            om, of = ranks[p.nbr_rank]
            q = next(x for x in om.patches if x.type == abi.PATCH_PROCESSOR and x.nbr_rank == r)
            sl = slice(p.start, p.start + p.size)
            w = m.weights[sl, None]
            Uf = w * f["U"][m.owner[sl]] + (1 - w) * of["U"][om.owner[q.start: q.start + q.size]]
            f["phi"][sl] = (Uf * m.Sf[sl]).sum(1)
            f["U_b"][p.start - nint: p.start - nint + p.size] = Uf
    return ranks


# ---------------------------------------------------------------- case dictionaries: constitutiveProperties, fvSchemes, fvSolution
class FoamDict:
    """A plain OpenFOAM dictionary file; entries are addressed by '/'-separated keyword paths."""

    def __init__(self, path):
        self.path = str(path)
        self._h = abi.lib().rheo_io_dict_open(self.path.encode())
        if not self._h:
            raise FoamError(_err())

    def __del__(self):
        try:
            if self._h:
                abi.lib().rheo_io_dict_free(self._h)
                self._h = None
        except Exception:
            pass

    def _lookup(self, key):
        buf = C.create_string_buffer(1 << 16)
        rc = abi.lib().rheo_io_dict_lookup(self._h, key.encode(), buf, len(buf))
        return rc, buf.value.decode()

    def get(self, key, default=None):
        """Value tokens of a primitive entry joined by blanks; `default` when there is no such entry."""
        rc, s = self._lookup(key)
        if rc == 2:
            return default
        if rc == 3:
            raise FoamError(f"{self.path}: {key} is a dictionary, not a value")
        if rc:
            raise FoamError(_err())
        return s

    def keys(self, key):
        rc, s = self._lookup(key)
        if rc != 3:
            raise FoamError(f"{self.path}: {key} is not a dictionary")
        return s.split()

    def scalar(self, key, default=None) -> float:
        """`name [dims] value` (dimensionedScalar) or a bare number: the last token."""
        s = self.get(key)
        if s is None:
            if default is None:
                raise FoamError(f"{self.path}: keyword {key} is undefined")
            return default
        return float(s.split()[-1])


def _read_one_model(d: FoamDict, at: str):
    from . import cases
    ty = d.get(f"{at}/type")
    if ty is None:
        raise FoamError(f"{d.path}: {at} has no type")
    if ty not in abi.MODEL_NAMES:
        raise FoamError(f"{d.path}: constitutiveEq type {ty} is not a log-conformation model of this library (valid: {sorted(abi.MODEL_NAMES)})")
    sc = lambda k, default=None: d.scalar(f"{at}/{k}", default)   # noqa: E731
    kw = dict(rho=sc("rho"), etaS=sc("etaS"), etaP=sc("etaP"))
    if ty == "Rolie-PolyLog":      # RoliePolyLog.C:114-121
        kw.update(lambda_=sc("lambdaD"), rp_lambdaR=sc("lambdaR"), rp_beta=sc("beta"), rp_delta=sc("delta"), rp_chiMax=sc("chiMax"))
    elif ty == "XPomPomLog":       # XPomPomLog.C:115-122
        kw.update(lambda_=sc("lambdaB"), xpp_lambdaS=sc("lambdaS"), alpha=sc("alpha"), xpp_q=sc("q"), xpp_n=sc("n"))
    else:
        kw.update(lambda_=sc("lambda"))
    if ty == "GiesekusLog":
        kw.update(alpha=sc("alpha"))
    elif ty == "PTTLog":           # PTTLog.C:41-50, 129-170
        fn = d.get(f"{at}/destructionFunctionType")
        if fn not in ("linear", "exponential", "generalized"):
            raise FoamError(f"{d.path}: destructionFunctionType {fn}: valid are linear exponential generalized")
        kw.update(epsilon=sc("epsilon"), zeta=sc("zeta"), ptt_function=fn)
        if fn == "generalized":
            kw.update(ml_alpha=sc("alpha"), ml_beta=sc("beta"))
    elif ty in ("FENE-PLog", "FENE-CRLog"):
        kw.update(L2=sc("L2"))
    elif ty == "WhiteMetznerCYLog":   # WhiteMetznerCYLog.C:132-140
        K, L, n, m_, a, b = (sc(k) for k in ("K", "L", "n", "m", "a", "b"))
        if m_ != n or K != L or a != b:
            raise FoamError("The Log version of the WhiteMetznerCY model can only be used if:  m=n   and   K=L   and   a=b")
        kw.update(wm_K=K, wm_n=n, wm_a=a)
    elif ty == "SaramitoLog":         # SaramitoLog.C:108-165
        n = sc("n")
        dims = d.get(f"{at}/dims")
        if dims is None:
            raise FoamError(f"{d.path}: keyword {at}/dims is undefined")
        kw.update(epsilon=sc("epsilon"), zeta=sc("zeta"), sar_tau0=sc("tau0"), sar_n=n, sar_k=None if n == 1.0 else sc("k"),
                  sar_dims=tuple(float(t) for t in dims.replace("(", " ").replace(")", " ").split()))
        if n == 1.0:
            fn = d.get(f"{at}/PTTfunction")
            if fn not in ("none", "linear", "exponential"):
                raise FoamError(f"{d.path}: The PTT function specified does not exist. Available PTT functions are: none linear exponential")
            kw.update(sar_ptt=fn)
    return cases.model_desc(ty, **kw)


def read_models(constitutive_properties, section: str = "parameters"):
    """constant/constitutiveProperties -> [RheoModelDesc]: one model, or the modes of a multiMode entry (multiMode.C:41-92)."""
    d = FoamDict(constitutive_properties)
    if d.get(f"{section}/type") == "multiMode":
        return [_read_one_model(d, f"{section}/models/{k}") for k in d.keys(f"{section}/models")]
    return [_read_one_model(d, section)]


def mode_names(constitutive_properties, section: str = "parameters"):
    """Field-name suffixes of the modes: [""] for a single model, the keys of `models` for multiMode (theta<name>, tau<name>:
    multiMode.C:73-87)."""
    d = FoamDict(constitutive_properties)
    return d.keys(f"{section}/models") if d.get(f"{section}/type") == "multiMode" else [""]


def read_schemes(case_dir, theta_name: str = "theta"):
    """system/fvSchemes + system/fvSolution -> RheoSchemeCtl, as the OpenFOAM shim reads them (of90_LogConformationGPU.C):
    div(phi,theta) must be `GaussDefCmpw <limiter>`, ddt Euler / backward / CrankNicolson, gradSchemes Gauss linear; the solver
    entry (PBiCG or PBiCGStab) is the solver that runs."""
    from . import cases
    case_dir = Path(case_dir)
    fs, sol = FoamDict(case_dir / "system" / "fvSchemes"), FoamDict(case_dir / "system" / "fvSolution")
    div = fs.get(f"divSchemes/div(phi,{theta_name})") or fs.get("divSchemes/default")
    if div is None:
        raise FoamError(f"{fs.path}: no div(phi,{theta_name}) scheme")
    tok = div.split()
    bounded = bool(tok) and tok[0] == "bounded"   # EXT-OF9 boundedConvectionScheme (rheoFilmFoam/UCM/system/fvSchemes:35)
    if bounded:
        tok = tok[1:]
    if len(tok) != 2 or tok[0] != "GaussDefCmpw" or tok[1] not in abi.LIMITER:
        raise FoamError(f"{fs.path}: div(phi,{theta_name}) is `{div}`; the stress step needs `[bounded] GaussDefCmpw <limiter>` with one of {sorted(abi.LIMITER)}")
    ddt = fs.get(f"ddtSchemes/ddt({theta_name})") or fs.get("ddtSchemes/default")
    cn_psi = 1.0
    if ddt is not None and ddt.split()[0] == "CrankNicolson":      # `CrankNicolson <psi>` (Cavity/Oldroyd-BLog/system/fvSchemes)
        tokd = ddt.split()
        cn_psi = float(tokd[1]) if len(tokd) > 1 else 1.0
        if not 0.0 <= cn_psi <= 1.0:
            raise FoamError(f"{fs.path}: CrankNicolson off-centring coefficient {cn_psi} is outside [0, 1]")
        ddt = "CrankNicolson"
    if ddt not in ("Euler", "backward", "CrankNicolson", "steadyState"):
        raise FoamError(f"{fs.path}: ddtSchemes Euler, backward, CrankNicolson and steadyState are available, not {ddt}")
    # the device hard-codes Gauss linear for grad(U) (boilerLog.H:1), for the per-component grad(theta) of phifDefC
    # (gaussDefCmpwConvectionScheme.C:254) and for `linExtrapGrad` (linearExtrapolationFvPatchField.C:128): anything else in
    # gradSchemes would silently differ from the reference
    for key in ("gradSchemes/default", "gradSchemes/grad(U)", "gradSchemes/linExtrapGrad"):
        g = fs.get(key)
        if g is not None and g.split() != ["Gauss", "linear"]:
            raise FoamError(f"{fs.path}: {key} is `{g}`; the stress step implements `Gauss linear` gradients only")
    if fs.get("gradSchemes/default") is None and (fs.get("gradSchemes/grad(U)") is None or fs.get("gradSchemes/linExtrapGrad") is None):
        raise FoamError(f"{fs.path}: gradSchemes needs `default Gauss linear` (or grad(U) and linExtrapGrad entries)")
    at = f"solvers/{theta_name}"
    solver = sol.get(f"{at}/solver")
    if solver is None:
        raise FoamError(f"{sol.path}: no solver entry for {theta_name}")
    if solver not in abi.SOLVER:
        raise FoamError(f"{sol.path}: solver {solver} for {theta_name}; the stress step implements {sorted(abi.SOLVER)} (+ DILU)")
    relax = sol.scalar(f"relaxationFactors/equations/{theta_name}", 0.0)
    # the solver the file names is the solver that runs: PBiCG (what the Log tutorials select; pbicg.cuh) or PBiCGStab (the tuned
    # path), each on any number of ranks
    ctl = cases.scheme_ctl(tok[1], solver, sol.scalar(f"{at}/tolerance", 1e-6), sol.scalar(f"{at}/relTol", 0.0), int(sol.scalar(f"{at}/minIter", 0)),
                           int(sol.scalar(f"{at}/maxIter", 1000)), relax, ddt=ddt, cn_psi=cn_psi, bounded=bounded)
    return ctl, solver


def read_schemes_modes(case_dir, theta_names):
    """multiMode: the modes are batched on ONE matrix with ONE set of controls, so every mode's schemes and solver entry
    (thetaM1, thetaM2, ...) must be identical — checked here instead of applying the first mode's to all."""
    first = None
    for name in theta_names:
        ctl, solver = read_schemes(case_dir, name)
        key = tuple(getattr(ctl, f) for f, _ in abi.RheoSchemeCtl._fields_)
        if first is None:
            first = (key, ctl, solver, name)
        elif key != first[0]:
            raise FoamError(f"{case_dir}: schemes / solver controls of {name} differ from those of {first[3]}; the batched multiMode solve needs identical controls")
    return first[1], first[2]
