/* rheo_io.h — OpenFOAM on-disk formats either side of the stress step.
 *
 * C-ABI (plain pointers and sizes).  What this stands for in the reference (EXT-OF9 facilities rheoTool relies on):
 *   - constant/polyMesh/{points,faces,owner,neighbour,boundary}[.gz] as shipped with
 *     of90/tutorials/rheoFoam/Aneurysm/HerschelBulkley/constant/polyMesh.org  (polyMesh read)
 *   - time-directory field files: MUST_READ of tau/theta, READ_IF_PRESENT of eigVals/eigVecs and their AUTO_WRITE
 *     (CE/Oldroyd-B/Oldroyd-BLog/Oldroyd_BLog.C:52-113); boundaryField dictionaries with regular-expression patch keys
 *     (e.g. of90/tutorials/rheoFoam/Cylinder/Oldroyd-BLog/0/theta:33)
 * so that restart files and ParaView keep working when the GPU type is selected, and so that a polyMesh written by
 * blockMesh / snappyHexMesh can be fed to the library outside an OpenFOAM process.
 *
 * ASCII only (`format ascii;`), plain or gzip-compressed.  Errors: NULL / non-zero + rheo_mesh_last_error().
 */
#ifndef RHEO_IO_H
#define RHEO_IO_H

#include <stdint.h>
#include "rheo_mesh.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- polyMesh -------------------------------------------------------------------------------- */
/* Read <dir>/{points,faces,owner,neighbour,boundary}[.gz]; geometry (Sf, Cf, C, V, weights) is computed as
 * EXT-OF9 primitiveMesh / surfaceInterpolation do.  Patch kinds come from `type` (patch, wall, empty, processor with
 * neighbProcNo; symmetryPlane / wedge / cyclic are refused: the stress step has no such patch).  theta/tau BCs default
 * to zeroGradient and are set from field files with rheo_io_apply_field_bcs. */
RheoHostMesh* rheo_io_read_polymesh(const char* dir);
/* Write a mesh that carries its points and faces (read from disk, or a generated tensor grid) as an OpenFOAM polyMesh. */
int rheo_io_write_polymesh(const RheoHostMesh* m, const char* dir, int32_t gz);
/* nPoints / nFaces-with-points of the mesh (0 when the mesh carries no points), patch names */
int rheo_io_mesh_counts(const RheoHostMesh* m, int64_t* n_points, int64_t* n_face_points);
/* processor patches of a mesh read from a processorN directory: the cell centres across the patch (what the neighbour rank
 * would send at start-up) — also recomputes the patch's interpolation weights (EXT-OF9 makeWeights on coupled patches) */
int rheo_io_set_nbr_centres(RheoHostMesh* m, int32_t patch, const double* centres3);
int rheo_io_patch_name(const RheoHostMesh* m, int32_t patch, char* buf, int32_t buflen);
int rheo_io_set_patch_name(RheoHostMesh* m, int32_t patch, const char* name);

/* ---- field files ------------------------------------------------------------------------------ */
typedef struct RheoFoamField RheoFoamField;
RheoFoamField* rheo_io_read_field(const char* path);
void rheo_io_field_free(RheoFoamField* f);
/* class ("volSymmTensorField", ...), object name, components per value (1, 3, 6, 9), whether internalField is uniform,
 * number of values of a nonuniform internalField (0 when uniform) */
int rheo_io_field_info(const RheoFoamField* f, char* cls, int32_t cls_len, char* object, int32_t object_len, int32_t* n_comp,
                       int32_t* internal_uniform, int64_t* n_internal);
/* internalField expanded to n_cells values (AoS, n_comp doubles each) */
int rheo_io_field_internal(const RheoFoamField* f, int64_t n_cells, double* out);
/* boundaryField entry that applies to `patch_name`: exact keyword first, then the regular-expression keys, last one
 * wins (EXT-OF9 dictionary lookup).  type -> buf; has_value = 1 if a `value` entry exists; values (n_faces*n_comp,
 * uniform values expanded) written when `values` != NULL.  Returns 2 when no entry matches. */
int rheo_io_field_patch(const RheoFoamField* f, const char* patch_name, int32_t n_faces, char* type, int32_t type_len, int32_t* has_value,
                        double* values);
/* Set theta_bc / tau_bc of every patch of `m` from the boundaryField types of a field file (which = 0 theta, 1 tau):
 * fixedValue, zeroGradient, linearExtrapolation, empty, processor; anything else is an error naming the patch. */
int rheo_io_apply_field_bcs(RheoHostMesh* m, const RheoFoamField* f, int32_t which);
/* Write a vol<Type>Field: internal [n_cells*n_comp]; per patch a type word and, when patch_values[p] != NULL, a
 * nonuniform `value` list of patch_sizes[p] entries.  dimensions e.g. "[1 -1 -2 0 0 0 0]".  %.17g: reading the file back
 * gives the same bits. */
int rheo_io_write_field(const char* path, const char* cls, const char* object, const char* dimensions, int32_t n_comp, int64_t n_cells,
                        const double* internal, int32_t n_patches, const char* const* patch_names, const char* const* patch_types,
                        const int32_t* patch_sizes, const double* const* patch_values, int32_t gz);

/* ---- plain dictionaries (constant/constitutiveProperties, system/fvSchemes, system/fvSolution) ------------------ */
typedef struct RheoFoamDict RheoFoamDict;
RheoFoamDict* rheo_io_dict_open(const char* path);
void rheo_io_dict_free(RheoFoamDict* d);
/* Entry at a '/'-separated path of keywords (regular-expression keys honoured at every level; a list of named
 * dictionaries such as multiMode's `models ( M1 {..} M2 {..} )` is descended like a dictionary).  Returns 0 and the
 * value's tokens joined by blanks ($variables expanded), 3 and the keys of the sub-dictionary, or 2 when there is no
 * such entry. */
int rheo_io_dict_lookup(const RheoFoamDict* d, const char* path, char* buf, int32_t buflen);

#ifdef __cplusplus
}
#endif
#endif /* RHEO_IO_H */
