/* rheo_mesh.h — host-side finite-volume mesh services of the B200 stress-step library.
 *
 * C-ABI (plain pointers and sizes).  These entry points stand in for what OpenFOAM-9 gives the
 * reference's hot path for free and which therefore is NOT under /root/reference (EXT-OF9):
 *   - polyMesh addressing in upper-triangular order (owner/neighbour, patches)       [primitiveMesh]
 *   - face/cell geometry (triangle-fan face centres/areas, pyramid cell centres/volumes)
 *   - linear interpolation weights                                   [surfaceInterpolation::makeWeights]
 *   - decomposePar `simple` decomposition + processor sub-mesh construction [domainDecomposition]
 * plus the synthetic solenoidal flux / conformation fields SURVEY.md §8(d) specifies for the
 * benchmark configurations C1..C5.
 *
 * Reference call sites that consume this data on the hot path:
 *   of90/src/libs/gaussDefCmpwConvectionScheme/gaussDefCmpwConvectionScheme.C:88-91 (owner/neighbour)
 *   of90/src/libs/gaussDefCmpwConvectionScheme/gaussDefCmpwConvectionScheme.C:232   (mesh.C())
 *   of90/src/libs/boundaryConditions/linearExtrapolation/linearExtrapolationFvPatchField.C:136-144
 *   of90/tutorials/rheoFoam/Cylinder/Oldroyd-BLog/system/decomposeParDict:16-31 (simple/scotch)
 *
 * All integer labels are int32 (OpenFOAM `label` default).  Vectors are AoS (x,y,z), symmTensors
 * AoS (xx,xy,xz,yy,yz,zz), tensors AoS row-major — the layouts of OpenFOAM `Field<Type>`.
 */
#ifndef RHEO_MESH_H
#define RHEO_MESH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- polyPatch kinds ------------------------------------------------------------------------ */
#define RHEO_PATCH_PATCH      0   /* `patch`  (inlet/outlet)                               */
#define RHEO_PATCH_WALL       1   /* `wall`                                                */
#define RHEO_PATCH_EMPTY      2   /* `empty`  (2-D front/back; carries no field values)    */
#define RHEO_PATCH_PROCESSOR  3   /* `processor` (halo to another rank)                    */

/* ---- fvPatchField kinds for theta / tau ------------------------------------------------------- */
#define RHEO_BC_FIXED_VALUE          0
#define RHEO_BC_ZERO_GRADIENT        1
#define RHEO_BC_LINEAR_EXTRAPOLATION 2   /* rheoTool's wall BC for tau (fixedValue-derived)  */
#define RHEO_BC_EMPTY                3
#define RHEO_BC_PROCESSOR            4
#define RHEO_BC_LINEAR_EXTRAPOLATION_REG 5   /* the same patch type with `useRegression true`: least-squares line through the wall
                                              cell's face and centre values (linearExtrapolationFvPatchField.C:72,152-219)        */

typedef struct RheoPatchDesc {
    int32_t type;      /* RHEO_PATCH_*                                              */
    int32_t start;     /* first face of the patch (index into the face list)        */
    int32_t size;      /* number of faces                                           */
    int32_t nbr_rank;  /* processor patches: the rank on the other side; else -1    */
    int32_t theta_bc;  /* RHEO_BC_* applied to theta on this patch                  */
    int32_t tau_bc;    /* RHEO_BC_* applied to tau on this patch                    */
} RheoPatchDesc;

/* A finite-volume mesh as the stress step sees it (what fvMesh exposes to the reference). */
typedef struct RheoMeshDesc {
    int32_t n_cells;
    int32_t n_faces;            /* internal + boundary                                     */
    int32_t n_internal_faces;
    int32_t n_patches;
    const int32_t* owner;       /* [n_faces]                                               */
    const int32_t* neighbour;   /* [n_internal_faces]                                      */
    const double*  Sf;          /* [3*n_faces]  face area vectors (owner -> neighbour)     */
    const double*  Cf;          /* [3*n_faces]  face centres                               */
    const double*  C;           /* [3*n_cells]  cell centres                               */
    const double*  V;           /* [n_cells]    cell volumes                               */
    const double*  weights;     /* [n_faces]    linear weights (owner side); non-coupled boundary = 1 */
    const double*  nbr_C;       /* [3*(n_faces-n_internal_faces)] centre of the cell across each
                                   processor face (ignored elsewhere); may be NULL if no processor patch */
    const RheoPatchDesc* patches;
    int32_t solved_components[6]; /* 1 = component solved (xx,xy,xz,yy,yz,zz); 2-D: xz,yz = 0
                                     (fvMesh::validComponents<symmTensor>)                 */
} RheoMeshDesc;

typedef struct RheoHostMesh RheoHostMesh;   /* opaque; owns its arrays */

/* Patch rule for the tensor-grid generator: boundary faces whose centre lies inside the closed box
 * [lo,hi] (each extended by `tol`) go to patch `patch`; first matching rule wins; faces matching
 * no rule go to `default_patch`. */
typedef struct RheoPatchRule {
    int32_t patch;
    double lo[3];
    double hi[3];
} RheoPatchRule;

typedef struct RheoPatchSpec {
    int32_t type;      /* RHEO_PATCH_* (not PROCESSOR) */
    int32_t theta_bc;
    int32_t tau_bc;
} RheoPatchSpec;

/* blockMesh-lite: a tensor-product hex grid xs[nx+1] x ys[ny+1] x zs[nz+1] in which only the cells
 * covered by one of the `n_boxes` index boxes (i0,i1,j0,j1,k0,k1; half-open) exist.  Cells are
 * numbered x-fastest over existing cells; faces in OpenFOAM upper-triangular order; boundary faces
 * grouped by patch (ordered by owner cell, then -x,+x,-y,+y,-z,+z).  Covers every multi-block
 * blockMeshDict of the log-conformation tutorials that has no curved edges (Contraction41, Cavity,
 * Channel, CrossSlot) and the synthetic 3-D configurations.
 * `two_d` != 0 marks xz,yz as not solved (empty front/back). Returns NULL on error. */
RheoHostMesh* rheo_mesh_tensor_grid(int32_t nx, int32_t ny, int32_t nz,
                                    const double* xs, const double* ys, const double* zs,
                                    int32_t n_boxes, const int32_t* boxes6,
                                    int32_t n_patches, const RheoPatchSpec* patches,
                                    int32_t n_rules, const RheoPatchRule* rules,
                                    int32_t default_patch, double tol, int32_t two_d);

/* The same grid restricted to the cells of sub-domain `rank` of a `simple`-style index-space
 * decomposition into px*py*pz boxes (rank = ix + px*iy + px*py*iz; equal index splits), with
 * processor patches towards the neighbouring sub-domains.  Produces, without ever building the
 * global mesh, exactly what rheo_mesh_decompose would produce for that assignment.
 * global_cell_ids (may be NULL) receives the global cell index of each local cell. */
RheoHostMesh* rheo_mesh_tensor_grid_part(int32_t nx, int32_t ny, int32_t nz,
                                    const double* xs, const double* ys, const double* zs,
                                    int32_t n_boxes, const int32_t* boxes6,
                                    int32_t n_patches, const RheoPatchSpec* patches,
                                    int32_t n_rules, const RheoPatchRule* rules,
                                    int32_t default_patch, double tol, int32_t two_d,
                                    int32_t px, int32_t py, int32_t pz, int32_t rank);

/* Wrap caller-owned arrays (copied) into a host mesh, e.g. arrays exported from an fvMesh. */
RheoHostMesh* rheo_mesh_from_desc(const RheoMeshDesc* desc);

void rheo_mesh_free(RheoHostMesh* m);
/* Fill `out` with pointers into `m` (valid until rheo_mesh_free). */
int  rheo_mesh_desc(const RheoHostMesh* m, RheoMeshDesc* out);
int  rheo_mesh_n_boundary_faces(const RheoHostMesh* m);

/* decomposePar `simple` (n = px,py,pz; delta = 0.001) : cell -> rank.  EXT-OF9 simpleGeomDecomp. */
int rheo_mesh_simple_decomp(const RheoHostMesh* m, int32_t px, int32_t py, int32_t pz,
                            double delta, int32_t* cell_to_rank);

/* Build the sub-mesh of `rank` for a given cell->rank map (EXT-OF9 domainDecomposition):
 * cells ascending in global id, internal faces in global face order, physical patches (all kept,
 * possibly empty) then one processor patch per neighbouring rank in ascending rank order, faces in
 * global face order.  cell_addr[n_cells_local], face_addr[n_faces_local] (OpenFOAM convention:
 * global face index + 1, negative when the local face is flipped), may be NULL. */
RheoHostMesh* rheo_mesh_decompose(const RheoHostMesh* m, const int32_t* cell_to_rank,
                                  int32_t n_ranks, int32_t rank);
int rheo_mesh_proc_addressing(const RheoHostMesh* sub, int32_t* cell_addr, int32_t* face_addr);

/* ---- GPU renumbering (the integer contract that must be bit-exact) --------------------------- */
/* Greedy multi-colouring in cell order + stable sort by colour:
 *   colour[c]   smallest colour unused by already coloured neighbours (internal faces only)
 *   perm[new]   = old cell;  colour_start[n_colours+1] offsets in the new numbering.
 * Returns the number of colours (>=1), or <0 on error.  perm/colour/colour_start sized by caller
 * (colour_start: at least 65 entries). */
int rheo_mesh_colour_renumber(const RheoHostMesh* m, int32_t* perm, int32_t* colour,
                              int32_t* colour_start);

/* Block ordering of lattice (tensor-product, blockMesh-like) meshes — what the device uses for PBiCGStab's DILU on such
 * meshes (csrc/host/ordering.hpp): cells sorted block by block (8x8x4 lattice cells in 3-D, 16x16 in 2-D; natural order, i
 * fastest, inside a block), the sequence cut into chunks of 256 cells, the chunk graph coloured greedily in chunk order, new
 * numbering = colour by colour, chunk by chunk.  perm[new] = old cell; colour_start[n_colours+1] (caller: >= 65 entries);
 * tile3 (may be NULL) receives the block shape.  Returns the number of chunk colours, 0 if the mesh is not a lattice mesh
 * (the device then falls back to rheo_mesh_colour_renumber's ordering), < 0 on error. */
int rheo_mesh_block_renumber(const RheoHostMesh* m, int32_t* perm, int32_t* colour_start, int32_t* tile3);

/* ---- thermoFunctions (of90/src/libs/thermo/thermoFunctions/): factor a_T(T) that multiplies lambda / etaP --------- */
#define RHEO_THERMO_CONSTANT           0   /* Constant: 1                                                         */
#define RHEO_THERMO_ARRHENIUS          1   /* Arrhenius.C:67:          exp(alpha (1/T - 1/T0));     p = {alpha, T0}   */
#define RHEO_THERMO_ARRHENIUS_MODIFIED 2   /* ArrheniusModified.C:67:  exp(-alpha (T - T0));        p = {alpha, T0}   */
#define RHEO_THERMO_WLF                3   /* WLF.C:67:                10^(-c1 (T - T0) / (c2 + (T - T0))); p = {c1, c2, T0} */
#define RHEO_THERMO_VFT                4   /* VFT.C:68:                10^(B + A / (T - T0));       p = {A, B, T0}    */
/* out[i] = a_T(T[i]) for n cells; returns non-zero for an unknown kind */
int rheo_thermo_factor(int32_t kind, const double* params3, int64_t n, const double* T, double* out);

/* ---- synthetic benchmark fields ------------------------------------------------ */
#define RHEO_FLOW_CONTRACTION_2D 0  /* psi = Q g(y/h(x)), h: H_up -> H_down, no-slip/no-penetration walls */
#define RHEO_FLOW_VORTEX         1  /* Psi_z = A sin(pi x^) sin(pi y^) (1 + 0.3 sin(pi z^)) on the bounding box */
#define RHEO_FLOW_CONTRACTION_3D 2  /* contraction stream function in (x,y), modulated in z */

typedef struct RheoSynthSpec {
    int32_t flow;        /* RHEO_FLOW_*                           */
    double  amplitude;   /* velocity scale                        */
    double  h_up, h_down;/* contraction half-heights              */
    double  x_ramp;      /* contraction: h(x) ramps over [-x_ramp,0] */
    double  theta_amp;   /* amplitude of smooth theta0 field      */
    double  noise;       /* +-noise uniform per component         */
    uint64_t seed;       /* mt19937_64 seed; keyed by GLOBAL cell id so partitions agree */
} RheoSynthSpec;

/* Fill U[3*n_cells], U_b[3*n_bfaces], phi[n_faces] (discretely divergence-free: circulation of the
 * vector potential round each face) and theta0[6*n_cells].  global_ids may be NULL (identity).
 * For the tensor-grid meshes only (needs the point coordinates kept inside the host mesh). */
int rheo_synth_fields(const RheoHostMesh* m, const RheoSynthSpec* spec, const int32_t* global_ids,
                      double* U, double* U_b, double* phi, double* theta0);

/* max over cells of dt*sum(max(outflow,0))/V for dt = 1 (to choose dt for a target CFL). */
double rheo_mesh_max_courant_rate(const RheoHostMesh* m, const double* phi);

const char* rheo_mesh_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* RHEO_MESH_H */
