// synth.cpp — synthetic benchmark inputs for the stress step.
//
// U, phi come from an analytic vector potential Psi; the face flux is the circulation of Psi round
// the face (sum over edges of Psi(edge midpoint).edge), so that the flux field is DISCRETELY
// divergence-free on every cell — what rheoFoam's pressure correction guarantees for the phi that
// constitutiveEq::correct() receives (of90/src/solvers/rheoFoam/rheoFoam.C:147-153).
// theta0 is a smooth symmetric field plus hash noise keyed by the GLOBAL cell id, so that every
// decomposition sees the same field.
#include <cmath>
#include <cstdint>

#include "host_mesh.hpp"

namespace {

const double PI = 3.14159265358979323846;

struct Flow {
    int kind;
    double A, h_up, h_down, x_ramp;
    double lo[3], len[3];

    double hx(double x, double* dh) const {
        double t = -x / x_ramp;
        if (t <= 0) { *dh = 0; return h_down; }
        if (t >= 1) { *dh = 0; return h_up; }
        double s = t * t * (3 - 2 * t), ds = 6 * t * (1 - t);
        *dh = (h_up - h_down) * ds * (-1.0 / x_ramp);
        return h_down + (h_up - h_down) * s;
    }
    // contraction stream function and its gradient
    double psi2d(double x, double y, double* dpx, double* dpy) const {
        const double Q = 2.0 * h_down * A;
        double dh, h = hx(x, &dh);
        double eta = y / h;
        if (eta >= 1.0) { *dpx = *dpy = 0; return 0.5 * Q; }
        if (eta <= -1.0) { *dpx = *dpy = 0; return -0.5 * Q; }
        double g = 0.25 * (3 * eta - eta * eta * eta), dg = 0.75 * (1 - eta * eta);
        *dpy = Q * dg / h;
        *dpx = Q * dg * (-y * dh / (h * h));
        return Q * g;
    }
    double zmod(double z, double* dz) const {
        double zh = (z - lo[2]) / len[2];
        *dz = 0.3 * PI / len[2] * std::cos(PI * zh);
        return 1.0 + 0.3 * std::sin(PI * zh);
    }
    // Psi_z and its gradient
    double Psi(const double* x, double* grad) const {
        if (kind == RHEO_FLOW_CONTRACTION_2D) {
            grad[2] = 0;
            return psi2d(x[0], x[1], &grad[0], &grad[1]);
        }
        if (kind == RHEO_FLOW_CONTRACTION_3D) {
            double dpx, dpy, dz, p = psi2d(x[0], x[1], &dpx, &dpy), mz = zmod(x[2], &dz);
            grad[0] = dpx * mz; grad[1] = dpy * mz; grad[2] = p * dz;
            return p * mz;
        }
        // vortex on the bounding box
        double xh = (x[0] - lo[0]) / len[0], yh = (x[1] - lo[1]) / len[1];
        double L = std::fmin(len[0], len[1]);
        double dz, mz = zmod(x[2], &dz);
        double s = A * L / PI;
        double sx = std::sin(PI * xh), sy = std::sin(PI * yh), cx = std::cos(PI * xh), cy = std::cos(PI * yh);
        grad[0] = s * PI / len[0] * cx * sy * mz;
        grad[1] = s * PI / len[1] * sx * cy * mz;
        grad[2] = s * sx * sy * dz;
        return s * sx * sy * mz;
    }
    void velocity(const double* x, double* u) const {
        double g[3];
        Psi(x, g);
        u[0] = g[1]; u[1] = -g[0]; u[2] = 0.0;   // curl (0,0,Psi_z)
    }
};

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

}  // namespace

extern "C" int rheo_synth_fields(const RheoHostMesh* m, const RheoSynthSpec* spec, const int32_t* global_ids,
                                 double* U, double* U_b, double* phi, double* theta0) {
    if (!m || !spec) { rheo::set_error("rheo_synth_fields: null argument"); return 1; }
    if (!m->has_grid) { rheo::set_error("rheo_synth_fields: mesh has no tensor-grid provenance"); return 2; }
    Flow fl;
    fl.kind = spec->flow; fl.A = spec->amplitude; fl.h_up = spec->h_up; fl.h_down = spec->h_down;
    fl.x_ramp = spec->x_ramp > 0 ? spec->x_ramp : 1.0;
    fl.lo[0] = m->xs.front(); fl.lo[1] = m->ys.front(); fl.lo[2] = m->zs.front();
    fl.len[0] = m->xs.back() - fl.lo[0]; fl.len[1] = m->ys.back() - fl.lo[1]; fl.len[2] = m->zs.back() - fl.lo[2];
    const bool two_d = (m->solved[2] == 0 && m->solved[4] == 0);

    if (phi) {
        for (int32_t f = 0; f < m->n_faces; ++f) {
            double p[4][3];
            rheo::grid_face_points(*m, &m->cell_ijk[3 * (size_t)m->owner[f]], m->face_dir[f], p);
            double circ = 0;
            for (int e = 0; e < 4; ++e) {
                const double* a = p[e];
                const double* b = p[(e + 1) & 3];
                const double dz = b[2] - a[2];
                if (dz == 0.0) continue;   // Psi has only a z component
                double mid[3] = {0.5 * (a[0] + b[0]), 0.5 * (a[1] + b[1]), 0.5 * (a[2] + b[2])}, g[3];
                circ += fl.Psi(mid, g) * dz;
            }
            phi[f] = circ;
        }
    }
    if (U)
        for (int32_t c = 0; c < m->n_cells; ++c) fl.velocity(&m->C[3 * (size_t)c], &U[3 * (size_t)c]);
    if (U_b)
        for (int32_t f = m->n_internal; f < m->n_faces; ++f)
            fl.velocity(&m->Cf[3 * (size_t)f], &U_b[3 * (size_t)(f - m->n_internal)]);
    if (theta0) {
        const double a = spec->theta_amp;
        for (int32_t c = 0; c < m->n_cells; ++c) {
            const double* x = &m->C[3 * (size_t)c];
            double xh = (x[0] - fl.lo[0]) / fl.len[0], yh = (x[1] - fl.lo[1]) / fl.len[1], zh = (x[2] - fl.lo[2]) / fl.len[2];
            double t[6];
            t[0] = a * std::sin(2 * PI * xh) * std::cos(PI * yh);
            t[1] = 0.5 * a * std::sin(PI * xh + 1.0) * std::sin(2 * PI * yh);
            t[2] = two_d ? 0.0 : 0.3 * a * std::sin(PI * zh + 0.5) * std::cos(PI * xh);
            t[3] = a * std::cos(2 * PI * yh) * std::sin(PI * xh + 0.3);
            t[4] = two_d ? 0.0 : 0.3 * a * std::sin(PI * yh) * std::sin(PI * zh + 0.2);
            t[5] = 0.5 * a * std::cos(PI * xh + PI * yh);
            const uint64_t gid = (uint64_t)(global_ids ? global_ids[c] : (m->global_cell.empty() ? c : m->global_cell[c]));
            for (int q = 0; q < 6; ++q) {
                if (two_d && (q == 2 || q == 4)) continue;
                uint64_t h = splitmix64(spec->seed * 0x100000001B3ULL + gid * 6 + q);
                double u01 = (double)(h >> 11) * (1.0 / 9007199254740992.0);
                t[q] += spec->noise * (2.0 * u01 - 1.0);
            }
            for (int q = 0; q < 6; ++q) theta0[6 * (size_t)c + q] = t[q];
        }
    }
    return 0;
}

// thermoFunctions of the reference (of90/src/libs/thermo/thermoFunctions/{Arrhenius,ArrheniusModified,WLF,VFT,Constant}): the
// factor that createField() / multiply() apply to lambda and etaP
extern "C" int rheo_thermo_factor(int32_t kind, const double* p, int64_t n, const double* T, double* out) {
    if (!out || (n > 0 && !T && kind != RHEO_THERMO_CONSTANT) || (kind != RHEO_THERMO_CONSTANT && !p)) return 1;
    for (int64_t i = 0; i < n; ++i) {
        switch (kind) {
            case RHEO_THERMO_CONSTANT: out[i] = 1.0; break;
            case RHEO_THERMO_ARRHENIUS: out[i] = std::exp(p[0] * (1. / T[i] - 1. / p[1])); break;
            case RHEO_THERMO_ARRHENIUS_MODIFIED: out[i] = std::exp(-p[0] * (T[i] - p[1])); break;
            case RHEO_THERMO_WLF: out[i] = std::pow(10.0, -p[0] * (T[i] - p[2]) / (p[1] + (T[i] - p[2]))); break;
            case RHEO_THERMO_VFT: out[i] = std::pow(10.0, p[1] + p[0] / (T[i] - p[2])); break;
            default: return 2;
        }
    }
    return 0;
}

