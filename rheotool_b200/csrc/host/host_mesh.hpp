// host_mesh.hpp — in-memory finite-volume mesh used on the host side of the B200 stress-step library.
// Restates the parts of OpenFOAM-9's polyMesh/fvMesh (EXT-OF9, not under /root/reference) that the
// reference hot path reads: owner/neighbour addressing, Sf, Cf, C, V, linear weights, patches.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "rheo_mesh.h"

struct RheoHostMesh {
    int32_t n_cells = 0, n_faces = 0, n_internal = 0;
    std::vector<int32_t> owner, neighbour;
    std::vector<double> Sf, Cf, C, V, weights, nbr_C;
    std::vector<RheoPatchDesc> patches;
    int32_t solved[6] = {1, 1, 1, 1, 1, 1};

    // --- tensor-grid provenance (only for generated meshes; needed by the synthetic fields) ---
    bool has_grid = false;
    std::vector<double> xs, ys, zs;
    std::vector<int32_t> cell_ijk;      // 3 per cell: global grid indices
    std::vector<int8_t>  face_dir;      // per face: 0..5 = -x,+x,-y,+y,-z,+z seen from the owner
    std::vector<int32_t> global_cell;   // global cell id of each local cell (identity for whole mesh)

    // --- polyMesh provenance (meshes read from disk; generated on demand for tensor grids): include/rheo_io.h ---
    std::vector<double> points;            // 3 per point
    std::vector<int32_t> face_start;       // n_faces + 1 offsets into face_pts
    std::vector<int32_t> face_pts;
    std::vector<std::string> patch_names;  // may be shorter than patches (unnamed: patch<i>)
    int32_t my_rank = -1;                  // processor sub-meshes: myProcNo

    // --- decomposition provenance (EXT-OF9 cellProcAddressing / faceProcAddressing) ---
    std::vector<int32_t> cell_addr, face_addr;

    int n_boundary_faces() const { return n_faces - n_internal; }
};

namespace rheo {

void set_error(const std::string& msg);

// EXT-OF9 primitiveMesh::makeFaceCentresAndAreas for one polygon given its points (n>=3).
void face_centre_area(const double (*p)[3], int n, double* fC, double* fS);

// EXT-OF9 primitiveMesh::makeCellCentresAndVols from face data + owner/neighbour.
void cell_centres_volumes(int n_cells, int n_faces, int n_internal, const int32_t* own,
                          const int32_t* nei, const double* fC, const double* fS,
                          std::vector<double>& C, std::vector<double>& V);

// EXT-OF9 surfaceInterpolation::makeWeights (internal faces); boundary weights = 1.
void linear_weights(RheoHostMesh& m);

// The 4 corner points of face (cell ijk, dir) of a tensor grid, ordered so that the right-hand
// normal points out of the cell.
void grid_face_points(const RheoHostMesh& m, const int32_t* ijk, int dir, double (*p)[3]);

}  // namespace rheo
