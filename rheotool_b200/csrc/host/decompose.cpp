// decompose.cpp — decomposePar `simple` + processor sub-mesh construction + the GPU colour renumbering.
//
// EXT-OF9 semantics restated (not under /root/reference): simpleGeomDecomp / geomDecomp rotation,
// domainDecomposition ordering (cells ascending, internal faces in global order, physical patches,
// then processor patches by neighbour rank, faces in global order, flipped when the local cell is
// the global neighbour).  Reference use: `decomposePar` + `mpirun -np N rheoFoam -parallel`
// (of90/tutorials/rheoHeatFoam/channel/PTTLog/Allrun:16-20; .../Cylinder/Oldroyd-BLog/system/decomposeParDict:16-31).
#include "ordering.hpp"
#include <algorithm>
#include <cmath>
#include <numeric>

#include "host_mesh.hpp"

namespace {

// EXT-OF9 simpleGeomDecomp::assignToProcessorGroup: equal counts, remainder spread one each
void assign_groups(std::vector<int32_t>& group, int n_groups, size_t n) {
    size_t per = n / n_groups, rem = n % n_groups;
    size_t idx = 0;
    for (int g = 0; g < n_groups; ++g) {
        size_t cnt = per + ((size_t)g < rem ? 1 : 0);
        for (size_t q = 0; q < cnt; ++q) group[idx++] = g;
    }
}

}  // namespace

extern "C" {

int rheo_mesh_simple_decomp(const RheoHostMesh* m, int32_t px, int32_t py, int32_t pz, double delta,
                            int32_t* cell_to_rank) {
    if (!m || px < 1 || py < 1 || pz < 1) { rheo::set_error("rheo_mesh_simple_decomp: bad arguments"); return 1; }
    const size_t n = (size_t)m->n_cells;
    // geomDecomp: small rotation that breaks ties of grid-aligned meshes
    const double d = 1 - 0.5 * delta * delta, d2 = d * d, a = delta, a2 = a * a;
    const double R[9] = {d2, -a * d, a, a * d - a2 * d, a * a2 + d2, -2 * a * d, a * d2 + a2, a * d - a2 * d, d2 - a2};
    std::vector<double> rp(3 * n);
    for (size_t c = 0; c < n; ++c) {
        const double* x = &m->C[3 * c];
        for (int r = 0; r < 3; ++r) rp[3 * c + r] = R[3 * r] * x[0] + R[3 * r + 1] * x[1] + R[3 * r + 2] * x[2];
    }
    std::vector<int32_t> order(n), group(n);
    std::fill(cell_to_rank, cell_to_rank + n, 0);
    const int np[3] = {px, py, pz};
    const int mult[3] = {1, px, px * py};
    for (int dir = 0; dir < 3; ++dir) {
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int32_t u, int32_t v) { return rp[3 * (size_t)u + dir] < rp[3 * (size_t)v + dir]; });
        assign_groups(group, np[dir], n);
        for (size_t q = 0; q < n; ++q) cell_to_rank[order[q]] += mult[dir] * group[q];
    }
    return 0;
}

RheoHostMesh* rheo_mesh_decompose(const RheoHostMesh* m, const int32_t* c2r, int32_t n_ranks, int32_t rank) {
    if (!m || !c2r || rank < 0 || rank >= n_ranks) { rheo::set_error("rheo_mesh_decompose: bad arguments"); return nullptr; }
    auto* s = new RheoHostMesh();
    for (int q = 0; q < 6; ++q) s->solved[q] = m->solved[q];
    s->has_grid = m->has_grid;
    s->xs = m->xs; s->ys = m->ys; s->zs = m->zs;

    std::vector<int32_t> g2l((size_t)m->n_cells, -1);
    for (int32_t c = 0; c < m->n_cells; ++c)
        if (c2r[c] == rank) {
            g2l[c] = s->n_cells++;
            s->cell_addr.push_back(c);
            s->global_cell.push_back(m->global_cell.empty() ? c : m->global_cell[c]);
            if (m->has_grid) for (int q = 0; q < 3; ++q) s->cell_ijk.push_back(m->cell_ijk[3 * (size_t)c + q]);
        }
    if (s->n_cells == 0) { rheo::set_error("rheo_mesh_decompose: rank owns no cells"); delete s; return nullptr; }

    auto push_face = [&](int32_t gf, bool flip, int32_t own_local) {
        s->owner.push_back(own_local);
        const double sg = flip ? -1.0 : 1.0;
        for (int q = 0; q < 3; ++q) { s->Sf.push_back(sg * m->Sf[3 * (size_t)gf + q]); s->Cf.push_back(m->Cf[3 * (size_t)gf + q]); }
        s->face_addr.push_back(flip ? -(gf + 1) : (gf + 1));
        if (m->has_grid) s->face_dir.push_back(flip ? (int8_t)(m->face_dir[gf] ^ 1) : m->face_dir[gf]);
    };
    // internal faces in global order
    for (int32_t f = 0; f < m->n_internal; ++f) {
        int32_t o = g2l[m->owner[f]], n = g2l[m->neighbour[f]];
        if (o >= 0 && n >= 0) { push_face(f, false, o); s->neighbour.push_back(n); s->weights.push_back(m->weights[f]); }
    }
    s->n_internal = (int32_t)s->neighbour.size();
    std::vector<double> nbrC_tmp;  // per boundary face, filled as faces are appended
    // physical patches (all kept)
    for (auto& pd : m->patches) {
        if (pd.type == RHEO_PATCH_PROCESSOR) { rheo::set_error("rheo_mesh_decompose: mesh is already decomposed"); delete s; return nullptr; }
        RheoPatchDesc q = pd;
        q.start = (int32_t)s->owner.size();
        for (int32_t f = pd.start; f < pd.start + pd.size; ++f) {
            int32_t o = g2l[m->owner[f]];
            if (o < 0) continue;
            push_face(f, false, o);
            s->weights.push_back(m->weights[f]);
            for (int e = 0; e < 3; ++e) nbrC_tmp.push_back(0.0);
        }
        q.size = (int32_t)s->owner.size() - q.start;
        s->patches.push_back(q);
    }
    // processor patches by neighbour rank, global face order
    std::vector<std::vector<int32_t>> pf(n_ranks);
    for (int32_t f = 0; f < m->n_internal; ++f) {
        int ro = c2r[m->owner[f]], rn = c2r[m->neighbour[f]];
        if (ro == rn) continue;
        if (ro == rank) pf[rn].push_back(f);
        else if (rn == rank) pf[ro].push_back(f);
    }
    for (int r = 0; r < n_ranks; ++r) {
        if (pf[r].empty()) continue;
        RheoPatchDesc q;
        q.type = RHEO_PATCH_PROCESSOR; q.theta_bc = RHEO_BC_PROCESSOR; q.tau_bc = RHEO_BC_PROCESSOR;
        q.nbr_rank = r; q.start = (int32_t)s->owner.size(); q.size = (int32_t)pf[r].size();
        for (int32_t f : pf[r]) {
            const bool mine_is_owner = (c2r[m->owner[f]] == rank);
            const int32_t gl = mine_is_owner ? m->owner[f] : m->neighbour[f];
            const int32_t go = mine_is_owner ? m->neighbour[f] : m->owner[f];
            push_face(f, !mine_is_owner, g2l[gl]);
            s->weights.push_back(mine_is_owner ? m->weights[f] : 1.0 - m->weights[f]);
            for (int e = 0; e < 3; ++e) nbrC_tmp.push_back(m->C[3 * (size_t)go + e]);
        }
        s->patches.push_back(q);
    }
    s->n_faces = (int32_t)s->owner.size();
    s->nbr_C = nbrC_tmp;
    s->my_rank = rank;
    for (size_t p = 0; p < m->patches.size() && p < m->patch_names.size(); ++p) s->patch_names.push_back(m->patch_names[p]);
    if (!m->points.empty()) {   // polyMesh provenance: the sub-mesh's faces are the parent's (reversed where this side is the neighbour)
        std::vector<int32_t> pid(m->points.size() / 3, -1);
        s->face_start.assign(1, 0);
        for (int32_t fa : s->face_addr) {
            const int32_t gf = (fa > 0 ? fa : -fa) - 1;
            const int32_t b = m->face_start[gf], e = m->face_start[gf + 1];
            for (int32_t q = 0; q < e - b; ++q) {
                // a flipped face keeps its first point and reverses the rest (EXT-OF9 face::reverseFace)
                const int32_t gp = m->face_pts[fa > 0 ? b + q : (q == 0 ? b : e - q)];
                if (pid[gp] < 0) { pid[gp] = (int32_t)(s->points.size() / 3); for (int c = 0; c < 3; ++c) s->points.push_back(m->points[3 * (size_t)gp + c]); }
                s->face_pts.push_back(pid[gp]);
            }
            s->face_start.push_back((int32_t)s->face_pts.size());
        }
    }
    s->C.resize(3 * (size_t)s->n_cells);
    s->V.resize((size_t)s->n_cells);
    for (int32_t c = 0; c < s->n_cells; ++c) {
        for (int e = 0; e < 3; ++e) s->C[3 * (size_t)c + e] = m->C[3 * (size_t)s->cell_addr[c] + e];
        s->V[c] = m->V[s->cell_addr[c]];
    }
    return s;
}

int rheo_mesh_block_renumber(const RheoHostMesh* m, int32_t* perm, int32_t* colour_start, int32_t* tile3) {
    if (!m || !perm || !colour_start) { rheo::set_error("rheo_mesh_block_renumber: null argument"); return -1; }
    rk_host::BlockOrdering bo;
    if (!rk_host::block_renumber(m->n_cells, m->n_internal, m->owner.data(), m->neighbour.data(), m->C.data(), bo)) return 0;
    std::copy(bo.perm.begin(), bo.perm.end(), perm);
    std::copy(bo.colourStart.begin(), bo.colourStart.end(), colour_start);
    if (tile3) { tile3[0] = bo.tile[0]; tile3[1] = bo.tile[1]; tile3[2] = bo.tile[2]; }
    return bo.nColours;
}

int rheo_mesh_colour_renumber(const RheoHostMesh* m, int32_t* perm, int32_t* colour, int32_t* colour_start) {
    if (!m || !perm || !colour || !colour_start) { rheo::set_error("rheo_mesh_colour_renumber: null argument"); return -1; }
    const int32_t n = m->n_cells;
    // cell -> internal-face neighbours (CSR)
    std::vector<int32_t> start((size_t)n + 1, 0);
    for (int32_t f = 0; f < m->n_internal; ++f) { start[m->owner[f] + 1]++; start[m->neighbour[f] + 1]++; }
    for (int32_t c = 0; c < n; ++c) start[c + 1] += start[c];
    std::vector<int32_t> adj((size_t)start[n]), fill(start.begin(), start.end() - 1);
    for (int32_t f = 0; f < m->n_internal; ++f) {
        adj[fill[m->owner[f]]++] = m->neighbour[f];
        adj[fill[m->neighbour[f]]++] = m->owner[f];
    }
    int n_col = 0;
    for (int32_t c = 0; c < n; ++c) {
        uint64_t used = 0;
        for (int32_t q = start[c]; q < start[c + 1]; ++q)
            if (adj[q] < c) used |= (uint64_t)1 << colour[adj[q]];
        int col = 0;
        while (used & ((uint64_t)1 << col)) ++col;
        if (col >= 63) { rheo::set_error("rheo_mesh_colour_renumber: more than 63 colours"); return -2; }
        colour[c] = col;
        n_col = std::max(n_col, col + 1);
    }
    for (int q = 0; q <= 64; ++q) colour_start[q] = 0;
    for (int32_t c = 0; c < n; ++c) colour_start[colour[c] + 1]++;
    for (int q = 0; q < 64; ++q) colour_start[q + 1] += colour_start[q];
    std::vector<int32_t> pos(colour_start, colour_start + 64);
    for (int32_t c = 0; c < n; ++c) perm[pos[colour[c]]++] = c;
    return n_col;
}

}  // extern "C"
