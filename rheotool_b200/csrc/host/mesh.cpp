// mesh.cpp — blockMesh-lite tensor-grid generator, polyMesh geometry and the RheoMeshDesc view.
//
// EXT-OF9 semantics restated here (OpenFOAM-9 is not under /root/reference):
//   * upper-triangular face order: internal faces sorted by owner, then neighbour
//   * primitiveMesh face centres/areas (triangle fan about the point average) and cell
//     centres/volumes (pyramids about the face-centre average)
//   * surfaceInterpolation::makeWeights: w = |Sf.(C_N-Cf)| / (|Sf.(Cf-C_P)| + |Sf.(C_N-Cf)|)
// Reference consumers: gaussDefCmpwConvectionScheme.C:88-91,232 ; linearExtrapolationFvPatchField.C:136-144.
#include "host_mesh.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <numeric>

namespace rheo {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

void face_centre_area(const double (*p)[3], int n, double* fC, double* fS) {
    if (n == 3) {
        for (int d = 0; d < 3; ++d) fC[d] = (1.0 / 3.0) * (p[0][d] + p[1][d] + p[2][d]);
        double a[3], b[3];
        for (int d = 0; d < 3; ++d) { a[d] = p[1][d] - p[0][d]; b[d] = p[2][d] - p[0][d]; }
        fS[0] = 0.5 * (a[1] * b[2] - a[2] * b[1]);
        fS[1] = 0.5 * (a[2] * b[0] - a[0] * b[2]);
        fS[2] = 0.5 * (a[0] * b[1] - a[1] * b[0]);
        return;
    }
    double sumN[3] = {0, 0, 0}, sumA = 0, sumAc[3] = {0, 0, 0}, est[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) est[d] += p[i][d];
    for (int d = 0; d < 3; ++d) est[d] /= n;
    for (int i = 0; i < n; ++i) {
        const double* q = p[i];
        const double* r = p[(i + 1) % n];
        double c[3], a[3], b[3], nn[3];
        for (int d = 0; d < 3; ++d) { c[d] = q[d] + r[d] + est[d]; a[d] = r[d] - q[d]; b[d] = est[d] - q[d]; }
        nn[0] = a[1] * b[2] - a[2] * b[1];
        nn[1] = a[2] * b[0] - a[0] * b[2];
        nn[2] = a[0] * b[1] - a[1] * b[0];
        double mag = std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
        for (int d = 0; d < 3; ++d) { sumN[d] += nn[d]; sumAc[d] += mag * c[d]; }
        sumA += mag;
    }
    if (sumA < 1e-150) {
        for (int d = 0; d < 3; ++d) { fC[d] = est[d]; fS[d] = 0.0; }
    } else {
        for (int d = 0; d < 3; ++d) { fC[d] = (1.0 / 3.0) * sumAc[d] / sumA; fS[d] = 0.5 * sumN[d]; }
    }
}

void cell_centres_volumes(int n_cells, int n_faces, int n_internal, const int32_t* own,
                          const int32_t* nei, const double* fC, const double* fS,
                          std::vector<double>& C, std::vector<double>& V) {
    C.assign(3 * (size_t)n_cells, 0.0);
    V.assign((size_t)n_cells, 0.0);
    std::vector<double> est(3 * (size_t)n_cells, 0.0);
    std::vector<int32_t> cnt((size_t)n_cells, 0);
    for (int f = 0; f < n_faces; ++f) {
        for (int d = 0; d < 3; ++d) est[3 * (size_t)own[f] + d] += fC[3 * (size_t)f + d];
        cnt[own[f]]++;
    }
    for (int f = 0; f < n_internal; ++f) {
        for (int d = 0; d < 3; ++d) est[3 * (size_t)nei[f] + d] += fC[3 * (size_t)f + d];
        cnt[nei[f]]++;
    }
    for (int c = 0; c < n_cells; ++c)
        for (int d = 0; d < 3; ++d) est[3 * (size_t)c + d] /= cnt[c];
    auto pyramid = [&](int f, int c, double sgn) {
        const double* s = fS + 3 * (size_t)f;
        const double* x = fC + 3 * (size_t)f;
        const double* e = &est[3 * (size_t)c];
        double pyr3 = sgn * (s[0] * (x[0] - e[0]) + s[1] * (x[1] - e[1]) + s[2] * (x[2] - e[2]));
        for (int d = 0; d < 3; ++d) C[3 * (size_t)c + d] += pyr3 * (0.75 * x[d] + 0.25 * e[d]);
        V[c] += pyr3;
    };
    for (int f = 0; f < n_faces; ++f) pyramid(f, own[f], 1.0);
    for (int f = 0; f < n_internal; ++f) pyramid(f, nei[f], -1.0);
    for (int c = 0; c < n_cells; ++c) {
        if (std::fabs(V[c]) > 1e-300)
            for (int d = 0; d < 3; ++d) C[3 * (size_t)c + d] /= V[c];
        else
            for (int d = 0; d < 3; ++d) C[3 * (size_t)c + d] = est[3 * (size_t)c + d];
        V[c] *= (1.0 / 3.0);
    }
}

void linear_weights(RheoHostMesh& m) {
    m.weights.assign((size_t)m.n_faces, 1.0);
    for (int f = 0; f < m.n_internal; ++f) {
        const double* s = &m.Sf[3 * (size_t)f];
        const double* x = &m.Cf[3 * (size_t)f];
        const double* cp = &m.C[3 * (size_t)m.owner[f]];
        const double* cn = &m.C[3 * (size_t)m.neighbour[f]];
        double so = std::fabs(s[0] * (x[0] - cp[0]) + s[1] * (x[1] - cp[1]) + s[2] * (x[2] - cp[2]));
        double sn = std::fabs(s[0] * (cn[0] - x[0]) + s[1] * (cn[1] - x[1]) + s[2] * (cn[2] - x[2]));
        m.weights[f] = sn / (so + sn);
    }
}

void grid_face_points(const RheoHostMesh& m, const int32_t* ijk, int dir, double (*p)[3]) {
    const int i = ijk[0], j = ijk[1], k = ijk[2];
    const double x0 = m.xs[i], x1 = m.xs[i + 1], y0 = m.ys[j], y1 = m.ys[j + 1], z0 = m.zs[k], z1 = m.zs[k + 1];
    auto set = [&](int n, double x, double y, double z) { p[n][0] = x; p[n][1] = y; p[n][2] = z; };
    switch (dir) {
        case 1: set(0, x1, y0, z0); set(1, x1, y1, z0); set(2, x1, y1, z1); set(3, x1, y0, z1); break;
        case 0: set(0, x0, y0, z0); set(1, x0, y0, z1); set(2, x0, y1, z1); set(3, x0, y1, z0); break;
        case 3: set(0, x0, y1, z0); set(1, x0, y1, z1); set(2, x1, y1, z1); set(3, x1, y1, z0); break;
        case 2: set(0, x0, y0, z0); set(1, x1, y0, z0); set(2, x1, y0, z1); set(3, x0, y0, z1); break;
        case 5: set(0, x0, y0, z1); set(1, x1, y0, z1); set(2, x1, y1, z1); set(3, x0, y1, z1); break;
        default: set(0, x0, y0, z0); set(1, x0, y1, z0); set(2, x1, y1, z0); set(3, x1, y0, z0); break;
    }
}

namespace {

struct GridSpec {
    int nx, ny, nz;
    std::vector<std::array<int32_t, 6>> boxes;
    // rows: merged active i-intervals per (j,k) and the global id of the first active cell
    std::vector<std::vector<std::pair<int32_t, int32_t>>> row_iv;
    std::vector<int64_t> row_start;

    void build_rows() {
        row_iv.assign((size_t)ny * nz, {});
        row_start.assign((size_t)ny * nz + 1, 0);
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j) {
                std::vector<std::pair<int32_t, int32_t>> iv;
                for (auto& b : boxes)
                    if (j >= b[2] && j < b[3] && k >= b[4] && k < b[5] && b[1] > b[0]) iv.push_back({b[0], b[1]});
                std::sort(iv.begin(), iv.end());
                std::vector<std::pair<int32_t, int32_t>> mg;
                for (auto& v : iv) {
                    if (!mg.empty() && v.first <= mg.back().second)
                        mg.back().second = std::max(mg.back().second, v.second);
                    else
                        mg.push_back(v);
                }
                size_t r = (size_t)j + (size_t)ny * k;
                int64_t n = 0;
                for (auto& v : mg) n += v.second - v.first;
                row_iv[r] = std::move(mg);
                row_start[r + 1] = row_start[r] + n;
            }
    }
    int64_t n_active() const { return row_start.back(); }
    // global id or -1
    int64_t gid(int i, int j, int k) const {
        if (i < 0 || j < 0 || k < 0 || i >= nx || j >= ny || k >= nz) return -1;
        size_t r = (size_t)j + (size_t)ny * k;
        int64_t off = 0;
        for (auto& v : row_iv[r]) {
            if (i < v.first) return -1;
            if (i < v.second) return row_start[r] + off + (i - v.first);
            off += v.second - v.first;
        }
        return -1;
    }
};

// plane-aligned split of [0,n) into p parts balancing the marginal active-cell counts
std::vector<int32_t> balanced_splits(const std::vector<int64_t>& marginal, int p) {
    int n = (int)marginal.size();
    std::vector<int32_t> s(p + 1, 0);
    int64_t total = 0;
    for (auto v : marginal) total += v;
    int64_t cum = 0;
    int part = 1;
    for (int i = 0; i < n && part < p; ++i) {
        cum += marginal[i];
        // smallest plane index where the cumulative count reaches part/p of the total
        while (part < p && cum * p >= total * part) { s[part] = i + 1; ++part; }
    }
    for (; part < p; ++part) s[part] = n;
    s[p] = n;
    return s;
}

const int DI[6] = {-1, 1, 0, 0, 0, 0}, DJ[6] = {0, 0, -1, 1, 0, 0}, DK[6] = {0, 0, 0, 0, -1, 1};

RheoHostMesh* build_grid(int nx, int ny, int nz, const double* xs, const double* ys, const double* zs,
                         int n_boxes, const int32_t* boxes6, int n_patches, const RheoPatchSpec* pspec,
                         int n_rules, const RheoPatchRule* rules, int default_patch, double tol, int two_d,
                         int px, int py, int pz, int rank) {
    if (nx < 1 || ny < 1 || nz < 1 || n_boxes < 1 || n_patches < 1 || default_patch < 0 || default_patch >= n_patches) {
        set_error("rheo_mesh_tensor_grid: bad arguments");
        return nullptr;
    }
    GridSpec g;
    g.nx = nx; g.ny = ny; g.nz = nz;
    for (int b = 0; b < n_boxes; ++b) {
        std::array<int32_t, 6> bx;
        for (int q = 0; q < 6; ++q) bx[q] = boxes6[6 * b + q];
        if (bx[0] < 0 || bx[1] > nx || bx[2] < 0 || bx[3] > ny || bx[4] < 0 || bx[5] > nz) {
            set_error("rheo_mesh_tensor_grid: box out of range");
            return nullptr;
        }
        g.boxes.push_back(bx);
    }
    g.build_rows();
    if (g.n_active() <= 0 || g.n_active() > 2000000000LL) {
        set_error("rheo_mesh_tensor_grid: no cells or too many cells for int32 labels");
        return nullptr;
    }
    const int n_ranks = px * py * pz;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) { set_error("rheo_mesh_tensor_grid_part: bad rank"); return nullptr; }

    // sub-domain index ranges (whole grid if n_ranks == 1)
    std::vector<int32_t> sx{0, nx}, sy{0, ny}, sz{0, nz};
    if (n_ranks > 1) {
        std::vector<int64_t> mx(nx, 0), my(ny, 0), mz(nz, 0);
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j) {
                size_t r = (size_t)j + (size_t)ny * k;
                for (auto& v : g.row_iv[r]) {
                    for (int i = v.first; i < v.second; ++i) mx[i]++;
                    my[j] += v.second - v.first;
                    mz[k] += v.second - v.first;
                }
            }
        sx = balanced_splits(mx, px);
        sy = balanced_splits(my, py);
        sz = balanced_splits(mz, pz);
    }
    auto rank_of = [&](int i, int j, int k) {
        int ix = int(std::upper_bound(sx.begin(), sx.end(), i) - sx.begin()) - 1;
        int iy = int(std::upper_bound(sy.begin(), sy.end(), j) - sy.begin()) - 1;
        int iz = int(std::upper_bound(sz.begin(), sz.end(), k) - sz.begin()) - 1;
        return ix + px * iy + px * py * iz;
    };
    const int rx = rank % px, ry = (rank / px) % py, rz = rank / (px * py);
    const int i0 = sx[rx], i1 = sx[rx + 1], j0 = sy[ry], j1 = sy[ry + 1], k0 = sz[rz], k1 = sz[rz + 1];

    auto* m = new RheoHostMesh();
    m->has_grid = true;
    m->xs.assign(xs, xs + nx + 1);
    m->ys.assign(ys, ys + ny + 1);
    m->zs.assign(zs, zs + nz + 1);
    if (two_d) { m->solved[2] = 0; m->solved[4] = 0; }

    // local cells in global order (x fastest), local id via a dense map over the sub-box
    const int bx = i1 - i0, by = j1 - j0, bz = k1 - k0;
    std::vector<int32_t> lid((size_t)std::max(bx, 0) * std::max(by, 0) * std::max(bz, 0), -1);
    auto lidx = [&](int i, int j, int k) { return (size_t)(i - i0) + (size_t)bx * ((size_t)(j - j0) + (size_t)by * (k - k0)); };
    int32_t nc = 0;
    for (int k = k0; k < k1; ++k)
        for (int j = j0; j < j1; ++j) {
            size_t r = (size_t)j + (size_t)ny * k;
            int64_t off = 0;
            for (auto& v : g.row_iv[r]) {
                for (int i = std::max(v.first, i0); i < std::min(v.second, i1); ++i) {
                    lid[lidx(i, j, k)] = nc++;
                    m->cell_ijk.push_back(i); m->cell_ijk.push_back(j); m->cell_ijk.push_back(k);
                    m->global_cell.push_back((int32_t)(g.row_start[r] + off + (i - v.first)));
                }
                off += v.second - v.first;
            }
        }
    m->n_cells = nc;
    if (nc == 0) { set_error("rheo_mesh_tensor_grid_part: empty sub-domain"); delete m; return nullptr; }
    auto local = [&](int i, int j, int k) -> int32_t {
        if (i < i0 || i >= i1 || j < j0 || j >= j1 || k < k0 || k >= k1) return -1;
        return lid[lidx(i, j, k)];
    };

    // ---- internal faces, upper-triangular order (+x,+y,+z neighbours have ascending ids) ----
    for (int32_t c = 0; c < nc; ++c) {
        const int i = m->cell_ijk[3 * (size_t)c], j = m->cell_ijk[3 * (size_t)c + 1], k = m->cell_ijk[3 * (size_t)c + 2];
        for (int d = 1; d < 6; d += 2) {
            int32_t n = local(i + DI[d], j + DJ[d], k + DK[d]);
            if (n >= 0) { m->owner.push_back(c); m->neighbour.push_back(n); m->face_dir.push_back((int8_t)d); }
        }
    }
    m->n_internal = (int32_t)m->neighbour.size();

    // ---- boundary faces: physical patches (by rule), then processor patches by neighbour rank ----
    struct BFace { int32_t cell; int8_t dir; int32_t key; int64_t g_lo, g_hi; };
    std::vector<std::vector<BFace>> phys(n_patches);
    std::vector<std::vector<BFace>> proc(n_ranks);
    for (int32_t c = 0; c < nc; ++c) {
        const int32_t* ijk = &m->cell_ijk[3 * (size_t)c];
        for (int d = 0; d < 6; ++d) {
            const int ii = ijk[0] + DI[d], jj = ijk[1] + DJ[d], kk = ijk[2] + DK[d];
            if (local(ii, jj, kk) >= 0) continue;
            int64_t og = g.gid(ii, jj, kk);
            if (og >= 0) {  // active cell on another rank -> processor face
                int r = rank_of(ii, jj, kk);
                int64_t mg = m->global_cell[c];
                proc[r].push_back({c, (int8_t)d, 0, std::min(mg, og), std::max(mg, og)});
                continue;
            }
            double p[4][3], fC[3], fS[3];
            grid_face_points(*m, ijk, d, p);
            face_centre_area(p, 4, fC, fS);
            int patch = default_patch;
            for (int r = 0; r < n_rules; ++r) {
                bool in = true;
                for (int q = 0; q < 3; ++q)
                    if (fC[q] < rules[r].lo[q] - tol || fC[q] > rules[r].hi[q] + tol) in = false;
                if (in) { patch = rules[r].patch; break; }
            }
            if (patch < 0 || patch >= n_patches) { set_error("rheo_mesh_tensor_grid: rule names a bad patch"); delete m; return nullptr; }
            phys[patch].push_back({c, (int8_t)d, 0, 0, 0});
        }
    }
    for (int p = 0; p < n_patches; ++p) {
        RheoPatchDesc pd;
        pd.type = pspec[p].type; pd.theta_bc = pspec[p].theta_bc; pd.tau_bc = pspec[p].tau_bc;
        pd.nbr_rank = -1; pd.start = (int32_t)m->owner.size(); pd.size = (int32_t)phys[p].size();
        for (auto& b : phys[p]) { m->owner.push_back(b.cell); m->face_dir.push_back(b.dir); }
        m->patches.push_back(pd);
    }
    std::vector<std::pair<int32_t, int32_t>> proc_faces;  // (face, other ijk index) bookkeeping below
    for (int r = 0; r < n_ranks; ++r) {
        if (proc[r].empty()) continue;
        std::stable_sort(proc[r].begin(), proc[r].end(), [](const BFace& a, const BFace& b) {
            return a.g_lo != b.g_lo ? a.g_lo < b.g_lo : a.g_hi < b.g_hi;
        });
        RheoPatchDesc pd;
        pd.type = RHEO_PATCH_PROCESSOR; pd.theta_bc = RHEO_BC_PROCESSOR; pd.tau_bc = RHEO_BC_PROCESSOR;
        pd.nbr_rank = r; pd.start = (int32_t)m->owner.size(); pd.size = (int32_t)proc[r].size();
        for (auto& b : proc[r]) { m->owner.push_back(b.cell); m->face_dir.push_back(b.dir); }
        m->patches.push_back(pd);
    }
    m->n_faces = (int32_t)m->owner.size();

    // ---- geometry ----
    m->Sf.resize(3 * (size_t)m->n_faces);
    m->Cf.resize(3 * (size_t)m->n_faces);
    for (int32_t f = 0; f < m->n_faces; ++f) {
        double p[4][3];
        grid_face_points(*m, &m->cell_ijk[3 * (size_t)m->owner[f]], m->face_dir[f], p);
        face_centre_area(p, 4, &m->Cf[3 * (size_t)f], &m->Sf[3 * (size_t)f]);
    }
    cell_centres_volumes(m->n_cells, m->n_faces, m->n_internal, m->owner.data(), m->neighbour.data(),
                         m->Cf.data(), m->Sf.data(), m->C, m->V);
    linear_weights(*m);

    // processor faces: centre of the cell on the other side (a one-cell geometry evaluation) and the
    // internal-face weight formula seen from the local (owner) side
    const int nb = m->n_boundary_faces();
    m->nbr_C.assign(3 * (size_t)nb, 0.0);
    for (auto& pd : m->patches) {
        if (pd.type != RHEO_PATCH_PROCESSOR) continue;
        for (int32_t f = pd.start; f < pd.start + pd.size; ++f) {
            const int32_t* ijk = &m->cell_ijk[3 * (size_t)m->owner[f]];
            const int d = m->face_dir[f];
            int32_t o[3] = {ijk[0] + DI[d], ijk[1] + DJ[d], ijk[2] + DK[d]};
            // geometry of the single neighbour cell (same arithmetic as cell_centres_volumes)
            double fC[6][3], fS[6][3], est[3] = {0, 0, 0};
            for (int q = 0; q < 6; ++q) {
                double p[4][3];
                grid_face_points(*m, o, q, p);
                face_centre_area(p, 4, fC[q], fS[q]);
                for (int e = 0; e < 3; ++e) est[e] += fC[q][e];
            }
            for (int e = 0; e < 3; ++e) est[e] /= 6.0;
            double cc[3] = {0, 0, 0}, vv = 0;
            for (int q = 0; q < 6; ++q) {
                double pyr3 = fS[q][0] * (fC[q][0] - est[0]) + fS[q][1] * (fC[q][1] - est[1]) + fS[q][2] * (fC[q][2] - est[2]);
                for (int e = 0; e < 3; ++e) cc[e] += pyr3 * (0.75 * fC[q][e] + 0.25 * est[e]);
                vv += pyr3;
            }
            double* nc3 = &m->nbr_C[3 * (size_t)(f - m->n_internal)];
            for (int e = 0; e < 3; ++e) nc3[e] = cc[e] / vv;
            const double* s = &m->Sf[3 * (size_t)f];
            const double* x = &m->Cf[3 * (size_t)f];
            const double* cp = &m->C[3 * (size_t)m->owner[f]];
            double so = std::fabs(s[0] * (x[0] - cp[0]) + s[1] * (x[1] - cp[1]) + s[2] * (x[2] - cp[2]));
            double sn = std::fabs(s[0] * (nc3[0] - x[0]) + s[1] * (nc3[1] - x[1]) + s[2] * (nc3[2] - x[2]));
            m->weights[f] = sn / (so + sn);
        }
    }
    return m;
}

}  // namespace
}  // namespace rheo

extern "C" {

const char* rheo_mesh_last_error(void) { return rheo::g_err.c_str(); }

RheoHostMesh* rheo_mesh_tensor_grid(int32_t nx, int32_t ny, int32_t nz, const double* xs, const double* ys,
                                    const double* zs, int32_t n_boxes, const int32_t* boxes6, int32_t n_patches,
                                    const RheoPatchSpec* patches, int32_t n_rules, const RheoPatchRule* rules,
                                    int32_t default_patch, double tol, int32_t two_d) {
    return rheo::build_grid(nx, ny, nz, xs, ys, zs, n_boxes, boxes6, n_patches, patches, n_rules, rules,
                            default_patch, tol, two_d, 1, 1, 1, 0);
}

RheoHostMesh* rheo_mesh_tensor_grid_part(int32_t nx, int32_t ny, int32_t nz, const double* xs, const double* ys,
                                         const double* zs, int32_t n_boxes, const int32_t* boxes6,
                                         int32_t n_patches, const RheoPatchSpec* patches, int32_t n_rules,
                                         const RheoPatchRule* rules, int32_t default_patch, double tol,
                                         int32_t two_d, int32_t px, int32_t py, int32_t pz, int32_t rank) {
    return rheo::build_grid(nx, ny, nz, xs, ys, zs, n_boxes, boxes6, n_patches, patches, n_rules, rules,
                            default_patch, tol, two_d, px, py, pz, rank);
}

RheoHostMesh* rheo_mesh_from_desc(const RheoMeshDesc* d) {
    if (!d || d->n_cells <= 0 || d->n_faces < d->n_internal_faces) { rheo::set_error("rheo_mesh_from_desc: bad desc"); return nullptr; }
    auto* m = new RheoHostMesh();
    m->n_cells = d->n_cells; m->n_faces = d->n_faces; m->n_internal = d->n_internal_faces;
    m->owner.assign(d->owner, d->owner + d->n_faces);
    m->neighbour.assign(d->neighbour, d->neighbour + d->n_internal_faces);
    m->Sf.assign(d->Sf, d->Sf + 3 * (size_t)d->n_faces);
    m->Cf.assign(d->Cf, d->Cf + 3 * (size_t)d->n_faces);
    m->C.assign(d->C, d->C + 3 * (size_t)d->n_cells);
    m->V.assign(d->V, d->V + d->n_cells);
    m->weights.assign(d->weights, d->weights + d->n_faces);
    const size_t nb = (size_t)(d->n_faces - d->n_internal_faces);
    if (d->nbr_C) m->nbr_C.assign(d->nbr_C, d->nbr_C + 3 * nb); else m->nbr_C.assign(3 * nb, 0.0);
    m->patches.assign(d->patches, d->patches + d->n_patches);
    for (int q = 0; q < 6; ++q) m->solved[q] = d->solved_components[q];
    m->global_cell.resize(m->n_cells);
    std::iota(m->global_cell.begin(), m->global_cell.end(), 0);
    return m;
}

void rheo_mesh_free(RheoHostMesh* m) { delete m; }

int rheo_mesh_desc(const RheoHostMesh* m, RheoMeshDesc* out) {
    if (!m || !out) return 1;
    out->n_cells = m->n_cells; out->n_faces = m->n_faces; out->n_internal_faces = m->n_internal;
    out->n_patches = (int32_t)m->patches.size();
    out->owner = m->owner.data(); out->neighbour = m->neighbour.data();
    out->Sf = m->Sf.data(); out->Cf = m->Cf.data(); out->C = m->C.data(); out->V = m->V.data();
    out->weights = m->weights.data(); out->nbr_C = m->nbr_C.data(); out->patches = m->patches.data();
    for (int q = 0; q < 6; ++q) out->solved_components[q] = m->solved[q];
    return 0;
}

int rheo_mesh_n_boundary_faces(const RheoHostMesh* m) { return m ? m->n_boundary_faces() : -1; }

int rheo_mesh_proc_addressing(const RheoHostMesh* sub, int32_t* cell_addr, int32_t* face_addr) {
    if (!sub) return 1;
    if (cell_addr) {
        const auto& src = sub->cell_addr.empty() ? sub->global_cell : sub->cell_addr;
        std::copy(src.begin(), src.end(), cell_addr);
    }
    if (face_addr) {
        if (sub->face_addr.empty()) return 2;
        std::copy(sub->face_addr.begin(), sub->face_addr.end(), face_addr);
    }
    return 0;
}

double rheo_mesh_max_courant_rate(const RheoHostMesh* m, const double* phi) {
    std::vector<double> out((size_t)m->n_cells, 0.0);
    for (int f = 0; f < m->n_faces; ++f) {
        if (phi[f] > 0) out[m->owner[f]] += phi[f];
        else if (f < m->n_internal) out[m->neighbour[f]] -= phi[f];
    }
    double mx = 0;
    for (int c = 0; c < m->n_cells; ++c) mx = std::max(mx, out[c] / m->V[c]);
    return mx;
}

}  // extern "C"
