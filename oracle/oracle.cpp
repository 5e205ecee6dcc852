// oracle.cpp — CPU restatement of rheoTool's log-conformation stress step.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle and the timed CPU baseline.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it; the product (rheotool_b200/)
// never does.
//
// PARITY PIN (what is pinned on the reference itself, and what is not):
//   PINNED on the reference's own text, compiled from /root/reference into oracle/_ref/libref_stress.so by
//   `make -C oracle ref` (recipe oracle/Makefile, shells oracle/ref_shim/; outputs committed as
//   tests/golden/reference_{cell,correct}.npz by tools/make_golden_reference.py; tests/test_reference_pin.py):
//   utils/jacobi.H, utils/boilerLog.H, constitutiveEq::decomposeGradU / innerP, the whole correct() bodies of
//   Oldroyd_BLog, GiesekusLog, PTTLog (linear / exponential / generalized, zeta != 0), FENE_PLog, FENE_CRLog,
//   WhiteMetznerCYLog, RoliePolyLog, XPomPomLog, SaramitoLog, gaussDefCmpwConvectionScheme::{fvmDiv, phifDefC, lims} with
//   every limiter row of limiters.H (its coupled-patch branches too, on emulated ranks),
//   linearExtrapolationFvPatchField::updateCoeffs (the gradient branch and, round 2, the useRegression branch: two fixture
//   cases, 6e-16).  Measured agreement of theta, tau and their boundary fields after one and three correct() calls:
//   <= 4e-15 relative L2 (test bar 1e-12).
//   NOT PINNED (not in /root/reference, restated from published OpenFOAM-9 / Eigen semantics):
//   the OpenFOAM-9 layer under that text — gaussGrad/linear, EulerDdtScheme / backwardDdtScheme, fvMatrix::relax,
//   solveSegregated, PBiCG / PBiCGStab / DILU iteration histories (the pin compares the SOLUTION of the assembled
//   system, solved by the harness with a different method to round-off) — and Eigen 3.2.9's
//   SelfAdjointEigenSolver (the reference's own jacobi.H alternative, constitutiveEq.C:418-426, stands in; theta
//   and tau do not depend on the order or sign of the eigen-pairs).  Round 2 added: BMPLog — PINNED: oracle/_ref compiles the whole
//   BMPLog::correct (BMPLog.C:142-201, fluidity equation included: PhiEqn is a scalar fvMatrix of the stand-in types, its
//   convection term the reference's own scheme instantiated for a scalar); fluidity 4e-16, theta / tau 4e-15 after three chained
//   calls (tests/test_bmp_log.py, tests/golden/reference_bmp.npz).  Equally unpinned: the explicit part of constitutiveEq::divTau
//   (constitutiveEq.C:72-132 over EXT-OF9 fvc::div / gaussGrad boundary values; analytic identities, tests/test_div_tau.py),
//   steadyState / bounded / CrankNicolson (tests/test_ddt_schemes.py, test_steady_bounded_thermo.py) and the caller-supplied
//   gradU of correct(alpha, gradU).  Those parts stay pinned only by (a) analytic
//   material functions and algebraic identities (tests/test_oracle_*.py), (b) cross-file steady states
//   (RoliePoly.C, XPomPom.C; tests/test_oracle_analytic.py), (c) partition invariance on tensor grids and on a piece
//   of the polyhedral polyMesh the reference ships (tests/test_unstructured.py).
//
// Conventions follow OpenFOAM: symmTensor = (xx,xy,xz,yy,yz,zz); tensor row-major; fields AoS;
// face loops in face order; one scalar Krylov solve per valid component (segregated).
// Paths below are relative to /root/reference/of90/src/libs/ ; CE = constitutiveEquations/constitutiveEqs.
// "EXT-OF9" marks OpenFOAM-9 behaviour that is not in /root/reference and is restated from its
// published semantics.
//
// Multi-rank runs are emulated in-process: a Case holds R sub-domain meshes with processor patches;
// "messages" are copies; global reductions sum per-rank partial sums in rank order.  Each phase is an
// OpenMP loop over ranks, which is also how the multi-core CPU baseline is timed.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <array>
#include <vector>

#include <omp.h>

#include "rheo_gpu.h"
#include "rheo_mesh.h"

namespace {

typedef std::vector<double> dvec;

std::string g_err;

// ------------------------------------------------------------------ small tensor algebra (EXT-OF9)
struct T9 { double v[9]; };
inline T9 mul(const T9& a, const T9& b) {   // A & B
    T9 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r.v[3 * i + j] = a.v[3 * i] * b.v[j] + a.v[3 * i + 1] * b.v[3 + j] + a.v[3 * i + 2] * b.v[6 + j];
    return r;
}
inline T9 transpose(const T9& a) { return T9{{a.v[0], a.v[3], a.v[6], a.v[1], a.v[4], a.v[7], a.v[2], a.v[5], a.v[8]}}; }
inline T9 add(const T9& a, const T9& b) { T9 r; for (int i = 0; i < 9; ++i) r.v[i] = a.v[i] + b.v[i]; return r; }
inline T9 sub(const T9& a, const T9& b) { T9 r; for (int i = 0; i < 9; ++i) r.v[i] = a.v[i] - b.v[i]; return r; }
inline T9 scale(double s, const T9& a) { T9 r; for (int i = 0; i < 9; ++i) r.v[i] = s * a.v[i]; return r; }
inline T9 identity() { return T9{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
inline T9 from_sym(const double* s) { return T9{{s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5]}}; }
inline void symm(const T9& t, double* s) {   // EXT-OF9 symm(T) = 1/2 (T + T^T)
    s[0] = t.v[0]; s[1] = 0.5 * (t.v[1] + t.v[3]); s[2] = 0.5 * (t.v[2] + t.v[6]);
    s[3] = t.v[4]; s[4] = 0.5 * (t.v[5] + t.v[7]); s[5] = t.v[8];
}
inline double tr(const T9& t) { return t.v[0] + t.v[4] + t.v[8]; }
inline T9 inv(const T9& t) {   // EXT-OF9 inv(tensor): cofactors / det (general formula, also for diagonal input)
    const double xx = t.v[0], xy = t.v[1], xz = t.v[2], yx = t.v[3], yy = t.v[4], yz = t.v[5], zx = t.v[6], zy = t.v[7], zz = t.v[8];
    const double det = xx * yy * zz + xy * yz * zx + xz * yx * zy - xx * yz * zy - xy * yx * zz - xz * yy * zx;
    T9 r = {{yy * zz - zy * yz, xz * zy - xy * zz, xy * yz - xz * yy,
             zx * yz - yx * zz, xx * zz - xz * zx, yx * xz - xx * yz,
             yx * zy - yy * zx, xy * zx - xx * zy, xx * yy - yx * xy}};
    for (int i = 0; i < 9; ++i) r.v[i] /= det;
    return r;
}
// CE/constitutiveEq/constitutiveEq.C:471-518  innerP: (t1^T & t2) & t1  or  (t1 & t2) & t1^T
inline T9 innerP(const T9& t1, const T9& t2, bool firstT) {
    return firstT ? mul(mul(transpose(t1), t2), t1) : mul(mul(t1, t2), transpose(t1));
}
inline double pos(double x) { return x >= 0 ? 1.0 : 0.0; }

// ------------------------------------------------------------------ mesh / state
struct Patch { int type, start, size, nbr_rank, theta_bc, tau_bc, nbr_patch; };

struct Mesh {
    int nCells = 0, nFaces = 0, nInt = 0;
    std::vector<int> own, nei, losort;
    dvec Sf, Cf, C, V, w, nbrC;
    std::vector<Patch> patches;
    int solved[6];
    int nB() const { return nFaces - nInt; }
};

struct Model {
    RheoModelDesc d;
    double phiCell = 0;   // BMPLog: this cell's fluidity (set by Mode::at)
    dvec gammaVals;   // CE/PTT/PTTLog/PTTLog.C:143-170
    int mlMaxIter = 0;
};

struct Mode {
    Model model;
    dvec theta, thetaOld, thetaOldOld, tau, eigVals, eigVecs, thetaB, tauB;
    dvec ddt0;                // CrankNicolson: the scheme's ddt0 field (EXT-OF9 CrankNicolsonDdtScheme::ddt0_)
    int ddt0TimeIndex = 0;    //                 time step at which ddt0 was last evaluated
    dvec lambdaCell, etaPCell;   // thermo-dependent lambda / etaP per cell (Oldroyd_BLog.C:133-135: createField); empty = the scalars
    // BMPLog (BMPLog.C:142-201).  The fluidity equation is a scalar transport equation with theta's convection scheme and
    // solver; EXT-OF9 solves a symmTensor matrix component by component with the very algorithm it applies to a scalar one, so
    // Phi is carried as component xx of a hidden mode (model RHEO_MODEL_BMP_FLUIDITY) whose other components are identically 0
    // (zero source, zero initial residual, no iterations): its theta = [Phi, 0, 0, 0, 0, 0].
    int fluidityMode = -1;    // BMPLog mode: index of its hidden fluidity mode
    int fluidityOf = -1;      // hidden fluidity mode: index of the BMPLog mode it belongs to
    const dvec* fluidity = nullptr;   // BMPLog mode during correct(): the hidden mode's theta (Phi = [6 c])
    bool per_cell() const { return !lambdaCell.empty() || fluidity != nullptr; }
    // the model with this cell's lambda / etaP / fluidity
    Model at(int c) const {
        Model q = model;
        if (!lambdaCell.empty()) { q.d.lambda = lambdaCell[c]; q.d.etaP = etaPCell[c]; }
        if (fluidity) q.phiCell = (*fluidity)[(size_t)6 * c];
        return q;
    }
};

struct Rank {
    Mesh mesh;
    std::vector<Mode> modes;
    dvec U, Ub, phi;
    dvec gradUext;   // caller-supplied gradU (correct(alpha, gradU), boilerLog.H:1), OpenFOAM tensor order; empty: fvc::grad(U)
    // per-step work (one mode at a time, like the reference)
    dvec L, rhs, fFene, diag, lower, upper, source, iC, bC, gradTheta;
};

struct Case {
    std::vector<Rank> ranks;
    RheoSchemeCtl ctl;
    // Alternative reading of `tau_ = ...` (off by default): GeometricField::operator= assigns the boundary value of
    // the right-hand expression to every non-fixed patch before correctBoundaryConditions(); with the boundary values of
    // eigVals_/eigVecs_ left at their construction value I that is 0, or -etaP/lambda I where the expression goes through
    // innerP (Oldroyd_BLog.C:176: the temporary's boundary is zero).  It is what a linearExtrapolation patch sees on a
    // zeroGradient patch that comes LATER in the patch list.
    bool tauAssign = false;
    bool sortEig = true;   // order eigenpairs ascending like Eigen::SelfAdjointEigenSolver (CE/constitutiveEq/constitutiveEq.C:390-414)
    int lastIters = 0;
    bool finalized = false;
    // time levels (EXT-OF9 Time::deltaT0Value, GeometricField::nOldTimes): set by orc_store_old_time / orc_step
    int nOldTimes = 0;       // store_old_time calls since the state was set
    double dtNow = 0, dt0 = 0;
};

// patchNeighbourField of a cell field with nc components per cell (EXT-OF9 processorFvPatchField)
template <class Get>
void patch_neighbour(const Case& cs, int r, const Patch& p, int nc, Get get, double* out) {
    const Rank& o = cs.ranks[p.nbr_rank];
    const Patch& q = o.mesh.patches[p.nbr_patch];
    const double* fld = get(p.nbr_rank);
    for (int i = 0; i < p.size; ++i) {
        const int cell = o.mesh.own[q.start + i];
        for (int k = 0; k < nc; ++k) out[(size_t)nc * i + k] = fld[(size_t)nc * cell + k];
    }
    (void)r;
}

// ------------------------------------------------------------------ CE/utils/jacobi.H:7-158
// Cyclic Jacobi, 50 fixed sweeps, thresholds exactly as in the reference (incl. tresh = 0.2*sm*sm
// for the first three sweeps).  Returns V (columns = eigenvectors) and D (eigenvalues, unsorted).
void jacobi_ref(const double* At, double* D, double* V) {
    const int N = 3;
    double A[3][3] = {{At[0], At[1], At[2]}, {At[1], At[3], At[4]}, {At[2], At[4], At[5]}};
    double B[3], Z[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[3 * i + j] = (i == j) ? 1.0 : 0.0;
    for (int ip = 0; ip < N; ++ip) { B[ip] = A[ip][ip]; D[ip] = B[ip]; Z[ip] = 0; }
    for (int i = 1; i <= 50; ++i) {
        double sm = 0;
        for (int ip = 0; ip < N - 1; ++ip)
            for (int iq = ip + 1; iq <= N - 1; ++iq) sm = sm + std::fabs(A[ip][iq]);
        double tresh = (i < 4) ? 0.2 * sm * sm : 0.0;
        for (int ip = 0; ip < N - 1; ++ip) {
            for (int iq = ip + 1; iq <= N - 1; ++iq) {
                double g = 100 * std::fabs(A[ip][iq]);
                if ((i > 4) && (std::fabs(D[ip]) + g == std::fabs(D[ip])) && (std::fabs(D[iq]) + g == std::fabs(D[iq])))
                    A[ip][iq] = 0;
                else if (std::fabs(A[ip][iq]) > tresh) {
                    double h = D[iq] - D[ip], t;
                    if (std::fabs(h) + g == std::fabs(h))
                        t = A[ip][iq] / h;
                    else {
                        double theta = 0.5 * h / A[ip][iq];
                        t = 1 / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
                        if (theta < 0) t = -t;
                    }
                    double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c, tau = s / (1.0 + c);
                    h = t * A[ip][iq];
                    Z[ip] -= h; Z[iq] += h; D[ip] -= h; D[iq] += h;
                    A[ip][iq] = 0;
                    for (int j = 0; j < ip; ++j) {
                        g = A[j][ip]; h = A[j][iq];
                        A[j][ip] = g - s * (h + g * tau); A[j][iq] = h + s * (g - h * tau);
                    }
                    for (int j = ip + 1; j < iq; ++j) {
                        g = A[ip][j]; h = A[j][iq];
                        A[ip][j] = g - s * (h + g * tau); A[j][iq] = h + s * (g - h * tau);
                    }
                    for (int j = iq + 1; j <= N - 1; ++j) {
                        g = A[ip][j]; h = A[iq][j];
                        A[ip][j] = g - s * (h + g * tau); A[iq][j] = h + s * (g - h * tau);
                    }
                    for (int j = 0; j <= N - 1; ++j) {
                        g = V[3 * j + ip]; h = V[3 * j + iq];
                        V[3 * j + ip] = g - s * (h + g * tau); V[3 * j + iq] = h + s * (g - h * tau);
                    }
                }
            }
        }
        for (int ip = 0; ip <= N - 1; ++ip) { B[ip] += Z[ip]; D[ip] = B[ip]; Z[ip] = 0; }
    }
}

// CE/constitutiveEq/constitutiveEq.C:360-416  calcEig: vecs columns = eigenvectors,
// vals = diag(exp(eig)), off-diagonals zeroed.  Back-end = jacobi.H restatement (the in-tree
// alternative, constitutiveEq.C:418-426); the active Eigen-3.2.9 QR back-end returns ascending
// eigenvalues, which `sortEig` reproduces (tau, theta, Omega, B are invariant to order and sign).
void calc_eig_cell(const double* theta6, double* vals9, double* vecs9, bool sortEig) {
    double D[3], V[9];
    jacobi_ref(theta6, D, V);
    int idx[3] = {0, 1, 2};
    if (sortEig) std::stable_sort(idx, idx + 3, [&](int a, int b) { return D[a] < D[b]; });
    for (int i = 0; i < 9; ++i) vals9[i] = 0.0;
    for (int c = 0; c < 3; ++c) {
        vals9[4 * c] = std::exp(D[idx[c]]);
        for (int r = 0; r < 3; ++r) vecs9[3 * r + c] = V[3 * r + idx[c]];
    }
}

// ------------------------------------------------------------------ Gauss-linear gradient (EXT-OF9)
// gaussGrad::gradf + linear interpolation; call sites CE/utils/boilerLog.H:1,
// gaussDefCmpwConvectionScheme/gaussDefCmpwConvectionScheme.C:254,
// boundaryConditions/linearExtrapolation/linearExtrapolationFvPatchField.C:130.
// psi: nc components per cell; psiFaceB: face values on boundary faces (nc per boundary face);
// out: 3*nc per cell, out[(3*k? )] layout: for component k, gradient d = out[cell*3*nc + 3*k + d]
void gauss_grad(const Mesh& m, int nc, const double* psi, const double* psiFaceB, double* out) {
    std::fill(out, out + (size_t)3 * nc * m.nCells, 0.0);
    for (int f = 0; f < m.nInt; ++f) {
        const int P = m.own[f], N = m.nei[f];
        const double* S = &m.Sf[3 * (size_t)f];
        for (int k = 0; k < nc; ++k) {
            const double pf = m.w[f] * (psi[(size_t)nc * P + k] - psi[(size_t)nc * N + k]) + psi[(size_t)nc * N + k];
            for (int d = 0; d < 3; ++d) {
                const double sp = S[d] * pf;
                out[(size_t)3 * nc * P + 3 * k + d] += sp;
                out[(size_t)3 * nc * N + 3 * k + d] -= sp;
            }
        }
    }
    for (const Patch& p : m.patches) {
        if (p.type == RHEO_PATCH_EMPTY) continue;
        for (int f = p.start; f < p.start + p.size; ++f) {
            const int P = m.own[f];
            const double* S = &m.Sf[3 * (size_t)f];
            const double* pb = &psiFaceB[(size_t)nc * (f - m.nInt)];
            for (int k = 0; k < nc; ++k)
                for (int d = 0; d < 3; ++d) out[(size_t)3 * nc * P + 3 * k + d] += S[d] * pb[k];
        }
    }
    for (int c = 0; c < m.nCells; ++c)
        for (int q = 0; q < 3 * nc; ++q) out[(size_t)3 * nc * c + q] /= m.V[c];
}

// boundary FACE values used by the linear scheme: non-coupled = patch value, coupled =
// w*psi_P + (1-w)*psi_N (EXT-OF9 surfaceInterpolationScheme::interpolate).
template <class Get>
void face_values_boundary(const Case& cs, int r, int nc, Get get, const double* patchVals, double* out) {
    const Mesh& m = cs.ranks[r].mesh;
    const double* psi = get(r);
    dvec nbr;
    for (const Patch& p : m.patches) {
        if (p.type == RHEO_PATCH_EMPTY) continue;
        if (p.type == RHEO_PATCH_PROCESSOR) {
            nbr.resize((size_t)nc * p.size);
            patch_neighbour(cs, r, p, nc, get, nbr.data());
            for (int i = 0; i < p.size; ++i) {
                const int f = p.start + i, P = m.own[f];
                for (int k = 0; k < nc; ++k)
                    out[(size_t)nc * (f - m.nInt) + k] = m.w[f] * psi[(size_t)nc * P + k] + (1.0 - m.w[f]) * nbr[(size_t)nc * i + k];
            }
        } else {
            for (int f = p.start; f < p.start + p.size; ++f)
                for (int k = 0; k < nc; ++k) out[(size_t)nc * (f - m.nInt) + k] = patchVals[(size_t)nc * (f - m.nInt) + k];
        }
    }
}

// ------------------------------------------------------------------ per-cell algebra
// CE/utils/boilerLog.H:26-34 + CE/constitutiveEq/constitutiveEq.C:323-358 (decomposeGradU)
void decompose_gradU_cell(const T9& L, const T9& R, const T9& Lam, double zeta, bool ptt, T9& omega, T9& B) {
    T9 X = transpose(L);
    if (ptt) {   // L.T() - zeta*symm(L)
        double s[6];
        symm(L, s);
        X = sub(X, scale(zeta, from_sym(s)));
    }
    const T9 M = innerP(R, X, true);
    T9 b = {{M.v[0], 0, 0, 0, M.v[4], 0, 0, 0, M.v[8]}};
    T9 o = {{0, 0, 0, 0, 0, 0, 0, 0, 0}};
    const double lx = Lam.v[0], ly = Lam.v[4], lz = Lam.v[8];
    o.v[1] = (ly * M.v[1] + lx * M.v[3]) / (ly - lx + 1e-16);
    o.v[2] = (lz * M.v[2] + lx * M.v[6]) / (lz - lx + 1e-16);
    o.v[5] = (lz * M.v[5] + ly * M.v[7]) / (lz - ly + 1e-16);
    o.v[3] = -o.v[1]; o.v[6] = -o.v[2]; o.v[7] = -o.v[5];
    omega = innerP(R, o, false);
    B = innerP(R, b, false);
}

// Mittag-Leffler series, CE/PTT/PTTLog/PTTLog.C:202-236
double mittag_leffler(const Model& mo, double zi) {
    double sum = 0, sumOld = 0, error = 1;
    int k = 0;
    while (k < mo.mlMaxIter && error > mo.d.ml_rtol) {
        double Eabk = std::pow(zi, k) / mo.gammaVals[k + 1];
        sumOld = sum;
        sum += Eabk;
        error = std::fabs((sumOld - sum) / (sumOld + 1e-12));
        k++;
    }
    return sum;
}

// right-hand side of the theta equation for one cell; returns the FENE-P / FENE-CR f (else 0)
//   Oldroyd_BLog.C:146-163, GiesekusLog.C:142-157, PTTLog.C:190-251, FENE_PLog.C:142-163, FENE_CRLog.C:141-163
//   SaramitoLog.C:150-238 (tau6 = the model's CURRENT tau of the cell; unused by the other models)
double model_rhs_cell(const Model& mo, const T9& L, const double* theta6, const T9& R, const T9& Lam, double* rhs6, const double* tau6 = nullptr) {
    const RheoModelDesc& d = mo.d;
    if (d.model == RHEO_MODEL_BMP_FLUIDITY) {   // BMPLog.C:158-160: Phi0/lambda + k (PhiInf - Phi) (tau && symm(L)); -Sp(1/lambda) is on the diagonal
        double s6[6];
        symm(L, s6);
        const double tD = tau6[0] * s6[0] + tau6[3] * s6[3] + tau6[5] * s6[5] + 2.0 * (tau6[1] * s6[1] + tau6[2] * s6[2] + tau6[4] * s6[4]);
        rhs6[0] = d.bmp_Phi0 / d.lambda + d.bmp_k * (d.bmp_PhiInf - theta6[0]) * tD;
        for (int q = 1; q < 6; ++q) rhs6[q] = 0.0;
        return 0.0;
    }
    T9 omega, B;
    // boilerLog.H:26: the zeta-branch is compiled for PTTLog and SaramitoLog
    decompose_gradU_cell(L, R, Lam, d.zeta, d.model == RHEO_MODEL_PTT_LOG || d.model == RHEO_MODEL_SARAMITO_LOG, omega, B);
    const T9 I = identity();
    const T9 th = from_sym(theta6);
    T9 acc = add(sub(mul(omega, th), mul(th, omega)), scale(2.0, B));
    double f = 0;
    switch (d.model) {
        case RHEO_MODEL_WM_CY_LOG: {   // WhiteMetznerCYLog.C:155-196
            double s6[6];
            symm(L, s6);
            const double magD = std::sqrt(s6[0] * s6[0] + s6[3] * s6[3] + s6[5] * s6[5] + 2.0 * (s6[1] * s6[1] + s6[2] * s6[2] + s6[4] * s6[4]));
            const double cy = std::pow(1.0 + std::pow(d.wm_K * std::sqrt(2.0) * magD, d.wm_a), (d.wm_n - 1.0) / d.wm_a);
            const double lamC = d.lambda * cy, etaC = d.etaP * cy;
            f = etaC / lamC;   // carried to theta -> tau (etaP/lambda fields of before the solve, :207)
            acc = add(acc, scale(1.0 / lamC, innerP(R, sub(inv(Lam), I), false)));
            break;
        }
        case RHEO_MODEL_OLDROYD_B_LOG:
            acc = add(acc, scale(1.0 / d.lambda, innerP(R, sub(inv(Lam), I), false)));
            break;
        case RHEO_MODEL_BMP_LOG:   // BMPLog.C:177-185: (Phi G0) (eigVecs & (inv(eigVals) - I) & eigVecs.T()), Phi of AFTER PhiEqn.solve()
            acc = add(acc, scale(mo.phiCell * d.bmp_G0, mul(mul(R, sub(inv(Lam), I)), transpose(R))));
            break;
        case RHEO_MODEL_GIESEKUS_LOG: {
            const T9 trhs = mul(mul(R, sub(inv(Lam), I)), transpose(R));
            const T9 A = mul(mul(R, Lam), transpose(R));
            acc = add(acc, scale(1.0 / d.lambda, sub(trhs, scale(d.alpha, mul(A, mul(trhs, trhs))))));
            break;
        }
        case RHEO_MODEL_PTT_LOG: {
            T9 ext = scale(1.0 / d.lambda, mul(mul(R, sub(inv(Lam), I)), transpose(R)));
            const T9 A = mul(mul(R, Lam), transpose(R));
            const double z = (d.epsilon / (1 - d.zeta)) * (tr(A) - 3.);
            if (d.ptt_function == RHEO_PTT_LINEAR) ext = scale(1. + z, ext);
            else if (d.ptt_function == RHEO_PTT_EXPONENTIAL) ext = scale(std::exp(z), ext);
            else ext = scale(mo.gammaVals[0] * mittag_leffler(mo, z), ext);
            acc = add(acc, ext);
            break;
        }
        case RHEO_MODEL_FENE_P_LOG: {
            const T9 A = mul(mul(R, Lam), transpose(R));
            f = d.L2 / (d.L2 - tr(A));
            const double a = d.L2 / (d.L2 - 3.);
            acc = add(acc, scale(1.0 / d.lambda, mul(mul(R, sub(scale(a, inv(Lam)), scale(f, I))), transpose(R))));
            break;
        }
        case RHEO_MODEL_ROLIE_POLY_LOG: {   // RoliePolyLog.C:144-186
            double a6[6];
            symm(mul(mul(R, Lam), transpose(R)), a6);
            const T9 A = from_sym(a6);
            const double trA = tr(A);
            double M1 = 2. * (1. - std::sqrt(3. / trA)) / d.rp_lambdaR;
            if (d.rp_chiMax > 1.) {
                const double c2 = d.rp_chiMax * d.rp_chiMax;
                M1 *= ((3. - (trA / 3.) / c2) * (1. - 1. / c2)) / ((1. - (trA / 3.) / c2) * (3. - 1. / c2));
            }
            const T9 inner = add(sub(A, I), scale(M1 * d.lambda, add(A, scale(d.rp_beta * std::pow(trA / 3., d.rp_delta), sub(A, I)))));
            acc = sub(acc, scale(1.0 / d.lambda, mul(mul(mul(R, inv(Lam)), transpose(R)), inner)));
            break;
        }
        case RHEO_MODEL_XPOMPOM_LOG: {   // XPomPomLog.C:148-183
            double a6[6];
            symm(mul(mul(R, Lam), transpose(R)), a6);
            const T9 A = from_sym(a6);
            const double trA = tr(A);
            const double ls = std::sqrt(trA / 3.);
            const T9 AA = mul(A, A);
            const double stretch = d.xpp_n == 0 ? (1. - 1. / ls) : (1. - 1. / std::pow(ls, d.xpp_n + 1.));
            const double fx = 2. * (d.lambda / d.xpp_lambdaS) * std::exp((2. / d.xpp_q) * (ls - 1.)) * stretch +
                              (1. / (ls * ls)) * (1. - d.alpha - (d.alpha / 3.) * (tr(AA) - 2. * trA));
            const T9 inner = add(add(scale(fx - 2. * d.alpha, A), scale(d.alpha, AA)), scale(d.alpha - 1., I));
            acc = sub(acc, scale(1.0 / d.lambda, mul(mul(mul(R, inv(Lam)), transpose(R)), inner)));
            break;
        }
        case RHEO_MODEL_FENE_CR_LOG: {   // FENE_CRLog.C:141-163
            const T9 A = mul(mul(R, Lam), transpose(R));
            f = d.L2 / (d.L2 - tr(A));
            acc = add(acc, scale((1.0 / d.lambda) * f, mul(mul(R, sub(inv(Lam), I)), transpose(R))));
            break;
        }
        case RHEO_MODEL_SARAMITO_LOG: {   // SaramitoLog.C:153-238
            // 2nd invariant of the deviatoric stress: mag(tau - ItensorCorr*tr(tau)/nDims)/sqrt(2), ItensorCorr = diag(dims)
            static const double zero6[6] = {0, 0, 0, 0, 0, 0};
            if (!tau6) tau6 = zero6;   // stand-alone per-cell entry point (orc_model_rhs) without a stress state
            const double nDims = d.sar_dims[0] + d.sar_dims[1] + d.sar_dims[2];
            const double trT = tau6[0] + tau6[3] + tau6[5];
            const double t6[6] = {tau6[0] - d.sar_dims[0] * trT / nDims, tau6[1], tau6[2], tau6[3] - d.sar_dims[1] * trT / nDims, tau6[4],
                                  tau6[5] - d.sar_dims[2] * trT / nDims};
            const double tauDMag = std::sqrt(t6[0] * t6[0] + t6[3] * t6[3] + t6[5] * t6[5] + 2.0 * (t6[1] * t6[1] + t6[2] * t6[2] + t6[4] * t6[4])) / std::sqrt(2.);
            double fac;
            if (d.sar_n == 1.) fac = std::max(0., (tauDMag - d.sar_tau0) / (d.sar_k * tauDMag + 1e-16));
            else fac = std::pow(std::max(0., (tauDMag - d.sar_tau0) / (d.sar_k * std::pow(tauDMag, d.sar_n) + 1e-16)), 1. / d.sar_n);
            double Y = 1.;   // PTT function (only with n == 1, SaramitoLog.C:133-165)
            if (d.sar_n == 1. && d.sar_ptt != 0) {
                const double z = (d.epsilon / (1 - d.zeta)) * (tr(mul(mul(R, Lam), transpose(R))) - 3.);
                Y = d.sar_ptt == 1 ? 1. + z : std::exp(z);
            }
            // thetaEqn -= symm(...)  ->  source += V * symm(...): the term joins the right-hand side with a plus sign
            double g6[6], a6[6];
            symm(acc, a6);
            symm(scale((fac * d.etaP / d.lambda) * Y, mul(mul(R, sub(inv(Lam), I)), transpose(R))), g6);
            for (int q = 0; q < 6; ++q) rhs6[q] = a6[q] + g6[q];
            return 0.;
        }
    }
    symm(acc, rhs6);
    return f;
}

// theta -> tau for one cell: Oldroyd_BLog.C:175, GiesekusLog.C:172, PTTLog.C:264, FENE_PLog.C:178
void tau_cell(const Model& mo, const T9& R, const T9& Lam, double fOld, double* tau6) {
    const RheoModelDesc& d = mo.d;
    const T9 A = innerP(R, Lam, false);
    const T9 I = identity();
    double s[6];
    double coef = d.etaP / d.lambda;
    if (d.model == RHEO_MODEL_BMP_LOG) {   // BMPLog.C:196: tau = G0 symm((eigVecs & eigVals & eigVecs.T()) - I)
        symm(sub(mul(mul(R, Lam), transpose(R)), I), s);
        coef = d.bmp_G0;
    } else if (d.model == RHEO_MODEL_FENE_P_LOG) {
        const double a = d.L2 / (d.L2 - 3.);
        symm(sub(scale(fOld, A), scale(a, I)), s);
    } else {
        if (d.model == RHEO_MODEL_PTT_LOG || d.model == RHEO_MODEL_SARAMITO_LOG) coef = d.etaP / (d.lambda * (1 - d.zeta));   // SaramitoLog.C:242
        if (d.model == RHEO_MODEL_FENE_CR_LOG) coef = (d.etaP / d.lambda) * fOld;   // FENE_CRLog.C:174 (f of before the solve)
        if (d.model == RHEO_MODEL_WM_CY_LOG) coef = fOld;                            // WhiteMetznerCYLog.C:207
        if (d.model == RHEO_MODEL_ROLIE_POLY_LOG && d.rp_chiMax > 1.) {              // RoliePolyLog.C:203-212 (updated tr A)
            double a6[6];
            symm(A, a6);
            const double trA = a6[0] + a6[3] + a6[5], c2 = d.rp_chiMax * d.rp_chiMax;
            coef *= ((3. - (trA / 3.) / c2) * (1. - 1. / c2)) / ((1. - (trA / 3.) / c2) * (3. - 1. / c2));
        }
        symm(sub(A, I), s);
    }
    for (int q = 0; q < 6; ++q) tau6[q] = coef * s[q];
}

// ------------------------------------------------------------------ limiter table
// gaussDefCmpwConvectionScheme/limiters.H:48-98
bool limiter_table(int lim, double* alpha, double* beta, double* bounds) {
    switch (lim) {
        case RHEO_LIMITER_CUBISTA: alpha[0] = 7. / 4.; alpha[1] = 3. / 4.; alpha[2] = 1. / 4.; beta[0] = 0.; beta[1] = 3. / 8.; beta[2] = 3. / 4.; bounds[0] = 3. / 8.; bounds[1] = 3. / 4.; return true;
        case RHEO_LIMITER_MINMOD: alpha[0] = 1.5; alpha[1] = .5; alpha[2] = .5; beta[0] = 0.; beta[1] = .5; beta[2] = .5; bounds[0] = .5; bounds[1] = 1.; return true;
        case RHEO_LIMITER_SMART: alpha[0] = 3.; alpha[1] = 3. / 4.; alpha[2] = 0.; beta[0] = 0.; beta[1] = 3. / 8.; beta[2] = 1.; bounds[0] = 1. / 6.; bounds[1] = 5. / 6.; return true;
        case RHEO_LIMITER_WACEB: alpha[0] = 2.; alpha[1] = 3. / 4.; alpha[2] = 0.; beta[0] = 0.; beta[1] = 3. / 8.; beta[2] = 1.; bounds[0] = 3. / 10.; bounds[1] = 5. / 6.; return true;
        case RHEO_LIMITER_SUPERBEE: alpha[0] = 0.5; alpha[1] = 1.5; alpha[2] = 0.; beta[0] = 0.5; beta[1] = 0.; beta[2] = 1.; bounds[0] = 1. / 2.; bounds[1] = 2. / 3.; return true;
        default: return false;
    }
}

// deferred-correction face value, gaussDefCmpwConvectionScheme.C:259-274 (swit = 1) and :304-317
inline double phif_defc(double vP, double vN, const double* gP, const double* gN, const double* dlt, double upw,
                        const double* aL, const double* bL, const double* bnd) {
    const double gd_up = (gP[0] * upw + (1.0 - upw) * gN[0]) * dlt[0] + (gP[1] * upw + (1.0 - upw) * gN[1]) * dlt[1] +
                         (gP[2] * upw + (1.0 - upw) * gN[2]) * dlt[2];
    const double phitc = 1.0 - ((vN - vP) / (2.0 * gd_up + 1e-18));
    double alpha, beta;
    if (phitc <= 0. || phitc >= 1.) { alpha = 1.; beta = 0.; }
    else if (phitc < bnd[0]) { alpha = aL[0]; beta = bL[0]; }
    else if (phitc < bnd[1]) { alpha = aL[1]; beta = bL[1]; }
    else { alpha = aL[2]; beta = bL[2]; }
    const double gPd = gP[0] * dlt[0] + gP[1] * dlt[1] + gP[2] * dlt[2];
    const double gNd = gN[0] * dlt[0] + gN[1] * dlt[1] + gN[2] * dlt[2];
    return (1.0 - alpha - beta) * (vN - 2.0 * gPd) * upw + (1.0 - alpha - beta) * (vP + 2.0 * gNd) * (1.0 - upw) +
           ((alpha - 1.0) * upw + beta * (1.0 - upw)) * vP + (beta * upw + (alpha - 1.0) * (1.0 - upw)) * vN;
}

// ------------------------------------------------------------------ distributed vectors + LDU ops (EXT-OF9)
struct DVec { std::vector<dvec> v; };

struct Ldu {   // one scalar system per component on the shared LDU structure
    Case* cs;
    // per rank: diag (incl. internalCoeffs of this component), lower, upper, interface boundary coeffs
    std::vector<const double*> lower, upper;
    std::vector<dvec> diag;
    std::vector<dvec> ifBou, ifInt;   // per boundary face (only processor faces used)
};

// threads of the rank loop; 0 = OpenMP's default (OMP_NUM_THREADS).  bench.py sets it explicitly: under torchrun the
// environment carries OMP_NUM_THREADS=1, which would silently time the CPU arm on one core.
int g_threads = 0;
template <class F> void for_ranks(Case& cs, F f) {
    const int R = (int)cs.ranks.size();
    const int nt = g_threads > 0 ? g_threads : omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nt) if (R > 1)
    for (int r = 0; r < R; ++r) f(r);
}

double gsum(Case& cs, const std::vector<double>& part) { double s = 0; for (size_t r = 0; r < cs.ranks.size(); ++r) s += part[r]; return s; }

// lduMatrix::Amul / Tmul incl. processor interfaces (result -= coeff * psi_nbr)
void amul(Ldu& A, DVec& y, const DVec& x, bool transposeA) {
    Case& cs = *A.cs;
    for_ranks(cs, [&](int r) {
        const Mesh& m = cs.ranks[r].mesh;
        double* yy = y.v[r].data();
        const double* xx = x.v[r].data();
        const double* lo = transposeA ? A.upper[r] : A.lower[r];
        const double* up = transposeA ? A.lower[r] : A.upper[r];
        for (int c = 0; c < m.nCells; ++c) yy[c] = A.diag[r][c] * xx[c];
        for (int f = 0; f < m.nInt; ++f) {
            yy[m.nei[f]] += lo[f] * xx[m.own[f]];
            yy[m.own[f]] += up[f] * xx[m.nei[f]];
        }
        const dvec& co = transposeA ? A.ifInt[r] : A.ifBou[r];
        for (const Patch& p : m.patches) {
            if (p.type != RHEO_PATCH_PROCESSOR) continue;
            const Rank& o = cs.ranks[p.nbr_rank];
            const Patch& q = o.mesh.patches[p.nbr_patch];
            const double* xo = x.v[p.nbr_rank].data();
            for (int i = 0; i < p.size; ++i) {
                const int f = p.start + i;
                yy[m.own[f]] -= co[f - m.nInt] * xo[o.mesh.own[q.start + i]];
            }
        }
    });
}

// lduMatrix::sumA
void sum_a(Ldu& A, DVec& s) {
    Case& cs = *A.cs;
    for_ranks(cs, [&](int r) {
        const Mesh& m = cs.ranks[r].mesh;
        double* ss = s.v[r].data();
        for (int c = 0; c < m.nCells; ++c) ss[c] = A.diag[r][c];
        for (int f = 0; f < m.nInt; ++f) { ss[m.nei[f]] += A.lower[r][f]; ss[m.own[f]] += A.upper[r][f]; }
        for (const Patch& p : m.patches)
            if (p.type == RHEO_PATCH_PROCESSOR)
                for (int f = p.start; f < p.start + p.size; ++f) ss[m.own[f]] -= A.ifBou[r][f - m.nInt];
    });
}

// DILUPreconditioner (rank-local)
struct Dilu {
    std::vector<dvec> rD;
    void init(Ldu& A) {
        Case& cs = *A.cs;
        rD.resize(cs.ranks.size());
        for_ranks(cs, [&](int r) {
            const Mesh& m = cs.ranks[r].mesh;
            rD[r] = A.diag[r];
            double* d = rD[r].data();
            for (int f = 0; f < m.nInt; ++f) d[m.nei[f]] -= A.upper[r][f] * A.lower[r][f] / d[m.own[f]];
            for (int c = 0; c < m.nCells; ++c) d[c] = 1.0 / d[c];
        });
    }
    void precondition(Ldu& A, DVec& w, const DVec& rr, bool transposeA) {
        Case& cs = *A.cs;
        for_ranks(cs, [&](int r) {
            const Mesh& m = cs.ranks[r].mesh;
            const double* d = rD[r].data();
            double* ww = w.v[r].data();
            const double* rv = rr.v[r].data();
            const double* lo = A.lower[r];
            const double* up = A.upper[r];
            for (int c = 0; c < m.nCells; ++c) ww[c] = d[c] * rv[c];
            if (!transposeA) {
                for (int q = 0; q < m.nInt; ++q) { const int f = m.losort[q]; ww[m.nei[f]] -= d[m.nei[f]] * lo[f] * ww[m.own[f]]; }
                for (int f = m.nInt - 1; f >= 0; --f) ww[m.own[f]] -= d[m.own[f]] * up[f] * ww[m.nei[f]];
            } else {
                for (int f = 0; f < m.nInt; ++f) ww[m.nei[f]] -= d[m.nei[f]] * up[f] * ww[m.own[f]];
                for (int q = m.nInt - 1; q >= 0; --q) { const int f = m.losort[q]; ww[m.own[f]] -= d[m.own[f]] * lo[f] * ww[m.nei[f]]; }
            }
        });
    }
};

struct Perf { double init = 0, fin = 0; int iters = 0; bool conv = false; bool singular = false; };

bool check_conv(const Perf& p, double tol, double relTol) {
    return p.fin < tol || (relTol > 1e-20 && p.fin < relTol * p.init);
}

template <class F> double reduce_ranks(Case& cs, F f) {
    std::vector<double> part(cs.ranks.size(), 0.0);
    for_ranks(cs, [&](int r) { part[r] = f(r); });
    return gsum(cs, part);
}

DVec make_vec(Case& cs) {
    DVec v;
    v.v.resize(cs.ranks.size());
    for (size_t r = 0; r < cs.ranks.size(); ++r) v.v[r].assign(cs.ranks[r].mesh.nCells, 0.0);
    return v;
}

// lduMatrix::solver::normFactor (restated in-repo: sparseMatrixSolvers/segregated/sparseSolver.C:152-179)
double norm_factor(Ldu& A, const DVec& psi, const DVec& source, const DVec& Apsi, DVec& tmp) {
    Case& cs = *A.cs;
    sum_a(A, tmp);
    double tot = reduce_ranks(cs, [&](int r) { double s = 0; for (double v : psi.v[r]) s += v; return s; });
    long n = 0;
    for (auto& rk : cs.ranks) n += rk.mesh.nCells;
    const double xRef = tot / (double)n;   // gAverage
    return reduce_ranks(cs, [&](int r) {
               double s = 0;
               const size_t nc = psi.v[r].size();
               for (size_t c = 0; c < nc; ++c) {
                   const double t = tmp.v[r][c] * xRef;
                   s += std::fabs(Apsi.v[r][c] - t) + std::fabs(source.v[r][c] - t);
               }
               return s;
           }) + 1e-20;
}

// EXT-OF9 PBiCGStab::solve (van der Vorst, right-preconditioned)
Perf pbicgstab(Ldu& A, DVec& psi, const DVec& source, const RheoSchemeCtl& ctl) {
    Case& cs = *A.cs;
    Perf perf;
    DVec yA = make_vec(cs), rA = make_vec(cs), pA = make_vec(cs);
    amul(A, yA, psi, false);
    for_ranks(cs, [&](int r) { for (size_t c = 0; c < rA.v[r].size(); ++c) rA.v[r][c] = source.v[r][c] - yA.v[r][c]; });
    const double normFactor = norm_factor(A, psi, source, yA, pA);
    auto sumMag = [&](const DVec& v) { return reduce_ranks(cs, [&](int r) { double s = 0; for (double x : v.v[r]) s += std::fabs(x); return s; }); };
    auto sumProd = [&](const DVec& a, const DVec& b) { return reduce_ranks(cs, [&](int r) { double s = 0; for (size_t c = 0; c < a.v[r].size(); ++c) s += a.v[r][c] * b.v[r][c]; return s; }); };
    perf.init = sumMag(rA) / normFactor;
    perf.fin = perf.init;
    if (ctl.min_iter > 0 || !check_conv(perf, ctl.tolerance, ctl.rel_tol)) {
        DVec AyA = make_vec(cs), sA = make_vec(cs), zA = make_vec(cs), tA = make_vec(cs);
        const DVec rA0 = rA;
        double rA0rA = 0, alpha = 0, omega = 0;
        Dilu pre;
        pre.init(A);
        do {
            const double rA0rAold = rA0rA;
            rA0rA = sumProd(rA0, rA);
            if (!(std::fabs(rA0rA) > 1e-300)) { perf.singular = true; break; }
            if (perf.iters == 0) {
                pA = rA;
            } else {
                if (!(std::fabs(omega) > 1e-300)) { perf.singular = true; break; }
                const double beta = (rA0rA / rA0rAold) * (alpha / omega);
                for_ranks(cs, [&](int r) { for (size_t c = 0; c < pA.v[r].size(); ++c) pA.v[r][c] = rA.v[r][c] + beta * (pA.v[r][c] - omega * AyA.v[r][c]); });
            }
            pre.precondition(A, yA, pA, false);
            amul(A, AyA, yA, false);
            const double rA0AyA = sumProd(rA0, AyA);
            alpha = rA0rA / rA0AyA;
            for_ranks(cs, [&](int r) { for (size_t c = 0; c < sA.v[r].size(); ++c) sA.v[r][c] = rA.v[r][c] - alpha * AyA.v[r][c]; });
            perf.fin = sumMag(sA) / normFactor;
            if (check_conv(perf, ctl.tolerance, ctl.rel_tol)) {
                for_ranks(cs, [&](int r) { for (size_t c = 0; c < psi.v[r].size(); ++c) psi.v[r][c] += alpha * yA.v[r][c]; });
                perf.iters++;
                perf.conv = true;
                return perf;
            }
            pre.precondition(A, zA, sA, false);
            amul(A, tA, zA, false);
            const double tAtA = sumProd(tA, tA);
            omega = sumProd(tA, sA) / tAtA;
            for_ranks(cs, [&](int r) {
                for (size_t c = 0; c < psi.v[r].size(); ++c) {
                    psi.v[r][c] += alpha * yA.v[r][c] + omega * zA.v[r][c];
                    rA.v[r][c] = sA.v[r][c] - omega * tA.v[r][c];
                }
            });
            perf.fin = sumMag(rA) / normFactor;
            // EXT-OF9 PBiCGStab.C: `(++solverPerf.nIterations() < maxIter_ && !converged) || nIterations() < minIter_` — a
            // PRE-increment, unlike PBiCG / PCG below (`nIterations()++ < maxIter_`)
        } while ((++perf.iters < ctl.max_iter && !check_conv(perf, ctl.tolerance, ctl.rel_tol)) || perf.iters < ctl.min_iter);
    }
    perf.conv = check_conv(perf, ctl.tolerance, ctl.rel_tol);
    return perf;
}

// EXT-OF9 PBiCG::solve (the solver every theta tutorial selects, e.g.
// of90/tutorials/rheoFoam/Cylinder/Oldroyd-BLog/system/fvSolution:32-45)
Perf pbicg(Ldu& A, DVec& psi, const DVec& source, const RheoSchemeCtl& ctl) {
    Case& cs = *A.cs;
    Perf perf;
    DVec pA = make_vec(cs), wA = make_vec(cs), pT = make_vec(cs), wT = make_vec(cs), rA = make_vec(cs), rT = make_vec(cs);
    amul(A, wA, psi, false);
    amul(A, wT, psi, true);
    for_ranks(cs, [&](int r) {
        for (size_t c = 0; c < rA.v[r].size(); ++c) { rA.v[r][c] = source.v[r][c] - wA.v[r][c]; rT.v[r][c] = source.v[r][c] - wT.v[r][c]; }
    });
    const double normFactor = norm_factor(A, psi, source, wA, pA);
    auto sumMag = [&](const DVec& v) { return reduce_ranks(cs, [&](int r) { double s = 0; for (double x : v.v[r]) s += std::fabs(x); return s; }); };
    auto sumProd = [&](const DVec& a, const DVec& b) { return reduce_ranks(cs, [&](int r) { double s = 0; for (size_t c = 0; c < a.v[r].size(); ++c) s += a.v[r][c] * b.v[r][c]; return s; }); };
    perf.init = sumMag(rA) / normFactor;
    perf.fin = perf.init;
    if (ctl.min_iter > 0 || !check_conv(perf, ctl.tolerance, ctl.rel_tol)) {
        Dilu pre;
        pre.init(A);
        double wArT = 1e300;   // solverPerf.great_
        do {
            const double wArTold = wArT;
            pre.precondition(A, wA, rA, false);
            pre.precondition(A, wT, rT, true);
            wArT = sumProd(wA, rT);
            if (perf.iters == 0) {
                pA = wA; pT = wT;
            } else {
                const double beta = wArT / wArTold;
                for_ranks(cs, [&](int r) {
                    for (size_t c = 0; c < pA.v[r].size(); ++c) { pA.v[r][c] = wA.v[r][c] + beta * pA.v[r][c]; pT.v[r][c] = wT.v[r][c] + beta * pT.v[r][c]; }
                });
            }
            amul(A, wA, pA, false);
            amul(A, wT, pT, true);
            const double wApT = sumProd(wA, pT);
            if (!(std::fabs(wApT) / normFactor > 1e-300)) { perf.singular = true; break; }
            const double alpha = wArT / wApT;
            for_ranks(cs, [&](int r) {
                for (size_t c = 0; c < psi.v[r].size(); ++c) {
                    psi.v[r][c] += alpha * pA.v[r][c];
                    rA.v[r][c] -= alpha * wA.v[r][c];
                    rT.v[r][c] -= alpha * wT.v[r][c];
                }
            });
            perf.fin = sumMag(rA) / normFactor;
        } while ((perf.iters++ < ctl.max_iter && !check_conv(perf, ctl.tolerance, ctl.rel_tol)) || perf.iters < ctl.min_iter);
    }
    perf.conv = check_conv(perf, ctl.tolerance, ctl.rel_tol);
    return perf;
}

// ------------------------------------------------------------------ one constitutiveEq::correct() of one mode
int correct_mode(Case& cs, int mi, double dt, RheoStepStats* st) {
    const int R = (int)cs.ranks.size();
    const RheoSchemeCtl& ctl = cs.ctl;
    if (ctl.ddt != RHEO_DDT_EULER && ctl.ddt != RHEO_DDT_BACKWARD && ctl.ddt != RHEO_DDT_CRANK_NICOLSON && ctl.ddt != RHEO_DDT_STEADY_STATE) { g_err = "oracle: only the Euler, backward, CrankNicolson and steadyState ddt schemes are restated"; return 3; }
    double aL[3] = {1, 1, 1}, bL[3] = {0, 0, 0}, bnd[2] = {1, 1};
    const bool hrs = limiter_table(ctl.limiter, aL, bL, bnd);
    const bool noConv = (ctl.limiter == RHEO_LIMITER_NONE);

    // --- boilerLog.H:1  L = fvc::grad(U)   (EXT-OF9 Gauss linear; processor faces interpolate with the neighbour cell)
    for_ranks(cs, [&](int r) {
        Rank& rk = cs.ranks[r];
        const Mesh& m = rk.mesh;
        dvec fb((size_t)3 * m.nB(), 0.0);
        face_values_boundary(cs, r, 3, [&](int q) { return cs.ranks[q].U.data(); }, rk.Ub.data(), fb.data());
        rk.L.resize((size_t)9 * m.nCells);
        if (!rk.gradUext.empty()) { rk.L = rk.gradUext; return; }   // boilerLog.H:1: L(gradU == nullptr ? fvc::grad(U)() : *gradU)
        gauss_grad(m, 3, rk.U.data(), fb.data(), rk.L.data());
        // gauss_grad stores, for component k of U, gradient d at [3k+d]; OpenFOAM's L_ij = d_i U_j = [3i+j]
        for (int c = 0; c < m.nCells; ++c) {
            double* l = &rk.L[(size_t)9 * c];
            std::swap(l[1], l[3]); std::swap(l[2], l[6]); std::swap(l[5], l[7]);
        }
    });

    // --- per-cell: Omega/B split and model source (uses the eigen-pairs of the PREVIOUS theta)
    for_ranks(cs, [&](int r) {
        Rank& rk = cs.ranks[r];
        Mode& mo = rk.modes[mi];
        const int n = rk.mesh.nCells;
        rk.rhs.resize((size_t)6 * n);
        rk.fFene.assign(n, 0.0);
        for (int c = 0; c < n; ++c) {
            T9 L, Rm, Lam;
            std::memcpy(L.v, &rk.L[(size_t)9 * c], 72);
            std::memcpy(Rm.v, &mo.eigVecs[(size_t)9 * c], 72);
            std::memcpy(Lam.v, &mo.eigVals[(size_t)9 * c], 72);
            Model tmp;
            const Model* pm = &mo.model;
            if (mo.per_cell()) { tmp = mo.at(c); pm = &tmp; }
            // the fluidity equation's source reads the stress of the BMPLog mode it belongs to (BMPLog.C:160: tau_ && symm(L))
            const dvec& tauSrc = mo.fluidityOf >= 0 ? rk.modes[mo.fluidityOf].tau : mo.tau;
            rk.fFene[c] = model_rhs_cell(*pm, L, &mo.theta[(size_t)6 * c], Rm, Lam, &rk.rhs[(size_t)6 * c], &tauSrc[(size_t)6 * c]);
        }
    });

    // --- fvm::ddt(theta) + fvm::div(phi,theta)  (EXT-OF9 EulerDdtScheme; gaussDefCmpwConvectionScheme.C:70-170)
    const double rDeltaT = 1.0 / dt;
    for_ranks(cs, [&](int r) {
        Rank& rk = cs.ranks[r];
        Mode& mo = rk.modes[mi];
        const Mesh& m = rk.mesh;
        const int n = m.nCells, nB = m.nB();
        rk.diag.assign(n, 0.0);
        rk.source.assign((size_t)6 * n, 0.0);
        rk.lower.assign(m.nInt, 0.0);
        rk.upper.assign(m.nInt, 0.0);
        rk.iC.assign(nB, 0.0);
        rk.bC.assign((size_t)6 * nB, 0.0);
        if (ctl.ddt == RHEO_DDT_BACKWARD) {
            // EXT-OF9 backwardDdtScheme<Type>::fvmDdt (static mesh): deltaT0 = great while the field has < 2 old times
            const double deltaT0 = cs.nOldTimes < 2 ? 1e15 : cs.dt0;
            const double coefft = 1 + dt / (dt + deltaT0);
            const double coefft00 = dt * dt / (deltaT0 * (dt + deltaT0));
            const double coefft0 = coefft + coefft00;
            const dvec& oo = mo.thetaOldOld.empty() ? mo.thetaOld : mo.thetaOldOld;
            for (int c = 0; c < n; ++c) {
                rk.diag[c] = (coefft * rDeltaT) * m.V[c];
                for (int q = 0; q < 6; ++q)
                    rk.source[(size_t)6 * c + q] = rDeltaT * m.V[c] * (coefft0 * mo.thetaOld[(size_t)6 * c + q] - coefft00 * oo[(size_t)6 * c + q]);
            }
        } else if (ctl.ddt == RHEO_DDT_CRANK_NICOLSON) {
            // EXT-OF9 CrankNicolsonDdtScheme<Type>::fvmDdt (static mesh), fresh start (the ddt0 field is created by the first
            // call, so its startTimeIndex is the first step's time index): k = time steps since the state was set,
            //   coef_  = (k > 1) ? 1 + psi : 1        rDtCoef  = coef_/deltaT          (first step: Euler)
            //   coef0_ = (k > 2) ? 1 + psi : 1        rDtCoef0 = coef0_/deltaT0
            //   once per time index:  ddt0 = rDtCoef0 (theta_old - theta_oldold) - offCentre(ddt0)
            //   diag = rDtCoef V;   source = (rDtCoef theta_old + offCentre(ddt0)) V;   offCentre(x) = psi < 1 ? psi x : x
            const double psi = ctl.cn_psi;
            const int k = std::max(1, cs.nOldTimes);
            const double off = psi < 1 ? psi : 1.0;
            if (mo.ddt0.size() != (size_t)6 * n) { mo.ddt0.assign((size_t)6 * n, 0.0); mo.ddt0TimeIndex = 0; }
            if (k > mo.ddt0TimeIndex) {
                if (k > 1) {
                    const double rDtCoef0 = (k > 2 ? 1 + psi : 1.0) / cs.dt0;
                    for (size_t i = 0; i < mo.ddt0.size(); ++i) mo.ddt0[i] = rDtCoef0 * (mo.thetaOld[i] - mo.thetaOldOld[i]) - off * mo.ddt0[i];
                }
                mo.ddt0TimeIndex = k;
            }
            const double rDtCoef = (k > 1 ? 1 + psi : 1.0) / dt;
            for (int c = 0; c < n; ++c) {
                rk.diag[c] = rDtCoef * m.V[c];
                for (int q = 0; q < 6; ++q)
                    rk.source[(size_t)6 * c + q] = (rDtCoef * mo.thetaOld[(size_t)6 * c + q] + off * mo.ddt0[(size_t)6 * c + q]) * m.V[c];
            }
        } else if (ctl.ddt == RHEO_DDT_STEADY_STATE) {
            // EXT-OF9 steadyStateDdtScheme<Type>::fvmDdt: an empty matrix (diag 0, source 0)
        } else
        for (int c = 0; c < n; ++c) {
            rk.diag[c] = rDeltaT * m.V[c];
            for (int q = 0; q < 6; ++q) rk.source[(size_t)6 * c + q] = rDeltaT * mo.thetaOld[(size_t)6 * c + q] * m.V[c];
        }
        if (!noConv) {
            dvec ddiag(n, 0.0);
            for (int f = 0; f < m.nInt; ++f) {
                const double upw = pos(rk.phi[f]);
                rk.lower[f] = -upw * rk.phi[f];
                rk.upper[f] = (1.0 - upw) * rk.phi[f];
            }
            for (int f = 0; f < m.nInt; ++f) { ddiag[m.own[f]] -= rk.lower[f]; ddiag[m.nei[f]] -= rk.upper[f]; }   // negSumDiag
            for (int c = 0; c < n; ++c) rk.diag[c] += ddiag[c];
            if (ctl.bounded) {
                // EXT-OF9 boundedConvectionScheme<Type>::fvmDiv: scheme.fvmDiv(phi, vf) - fvm::Sp(fvc::surfaceIntegrate(phi), vf);
                // fvm::Sp(sp, vf): diag += V sp, and V surfaceIntegrate(phi) = sum of the outward face fluxes of the cell
                // (internal faces in face order, then the patches: EXT-OF9 fvc::surfaceIntegrate; empty patches hold no faces)
                dvec net(n, 0.0);
                for (int f = 0; f < m.nInt; ++f) { net[m.own[f]] += rk.phi[f]; net[m.nei[f]] -= rk.phi[f]; }
                for (const Patch& p : m.patches) {
                    if (p.type == RHEO_PATCH_EMPTY) continue;
                    for (int f = p.start; f < p.start + p.size; ++f) net[m.own[f]] += rk.phi[f];
                }
                for (int c = 0; c < n; ++c) rk.diag[c] -= net[c];
            }
            for (const Patch& p : m.patches) {
                if (p.type == RHEO_PATCH_EMPTY) continue;
                for (int f = p.start; f < p.start + p.size; ++f) {
                    const int b = f - m.nInt;
                    const double ph = rk.phi[f];
                    double vIC, vBC[6];
                    if (p.theta_bc == RHEO_BC_PROCESSOR) {
                        const double plim = pos(ph);
                        vIC = plim;
                        for (int q = 0; q < 6; ++q) vBC[q] = 1.0 - plim;
                    } else if (p.theta_bc == RHEO_BC_ZERO_GRADIENT) {
                        vIC = 1.0;
                        for (int q = 0; q < 6; ++q) vBC[q] = 0.0;
                    } else {   // fixedValue
                        vIC = 0.0;
                        for (int q = 0; q < 6; ++q) vBC[q] = mo.thetaB[(size_t)6 * b + q];
                    }
                    rk.iC[b] = ph * vIC;
                    for (int q = 0; q < 6; ++q) rk.bC[(size_t)6 * b + q] = -ph * vBC[q];
                }
            }
        }
    });

    if (hrs) {
        // phifDefC (gaussDefCmpwConvectionScheme.C:195-328): one Gauss-linear gradient per component
        for_ranks(cs, [&](int r) {
            Rank& rk = cs.ranks[r];
            Mode& mo = rk.modes[mi];
            const Mesh& m = rk.mesh;
            dvec fb((size_t)6 * m.nB(), 0.0);
            face_values_boundary(cs, r, 6, [&](int q) { return cs.ranks[q].modes[mi].theta.data(); }, mo.thetaB.data(), fb.data());
            rk.gradTheta.resize((size_t)18 * m.nCells);
            gauss_grad(m, 6, mo.theta.data(), fb.data(), rk.gradTheta.data());
        });
        for_ranks(cs, [&](int r) {
            Rank& rk = cs.ranks[r];
            Mode& mo = rk.modes[mi];
            const Mesh& m = rk.mesh;
            dvec souT((size_t)6 * m.nCells, 0.0);
            for (int f = 0; f < m.nInt; ++f) {
                const int P = m.own[f], N = m.nei[f];
                const double upw = pos(rk.phi[f]);
                const double dl[3] = {m.C[3 * (size_t)N] - m.C[3 * (size_t)P], m.C[3 * (size_t)N + 1] - m.C[3 * (size_t)P + 1], m.C[3 * (size_t)N + 2] - m.C[3 * (size_t)P + 2]};
                for (int q = 0; q < 6; ++q) {
                    const double v = phif_defc(mo.theta[(size_t)6 * P + q], mo.theta[(size_t)6 * N + q], &rk.gradTheta[(size_t)18 * P + 3 * q],
                                               &rk.gradTheta[(size_t)18 * N + 3 * q], dl, upw, aL, bL, bnd);
                    souT[(size_t)6 * P + q] += v * rk.phi[f];
                    souT[(size_t)6 * N + q] -= v * rk.phi[f];
                }
            }
            dvec nTh, nGr;
            for (const Patch& p : m.patches) {
                if (p.type != RHEO_PATCH_PROCESSOR) continue;
                nTh.resize((size_t)6 * p.size);
                nGr.resize((size_t)18 * p.size);
                patch_neighbour(cs, r, p, 6, [&](int q) { return cs.ranks[q].modes[mi].theta.data(); }, nTh.data());
                patch_neighbour(cs, r, p, 18, [&](int q) { return cs.ranks[q].gradTheta.data(); }, nGr.data());
                for (int i = 0; i < p.size; ++i) {
                    const int f = p.start + i, P = m.own[f], b = f - m.nInt;
                    const double upw = pos(rk.phi[f]);
                    const double dl[3] = {m.nbrC[3 * (size_t)b] - m.C[3 * (size_t)P], m.nbrC[3 * (size_t)b + 1] - m.C[3 * (size_t)P + 1], m.nbrC[3 * (size_t)b + 2] - m.C[3 * (size_t)P + 2]};
                    for (int q = 0; q < 6; ++q) {
                        const double v = phif_defc(mo.theta[(size_t)6 * P + q], nTh[(size_t)6 * i + q], &rk.gradTheta[(size_t)18 * P + 3 * q],
                                                   &nGr[(size_t)18 * i + 3 * q], dl, upw, aL, bL, bnd);
                        souT[(size_t)6 * P + q] += v * rk.phi[f];   // only contributes once (to owner cell)
                    }
                }
            }
            for (size_t q = 0; q < souT.size(); ++q) rk.source[q] += -souT[q];
        });
    }

    // --- `== symm(...)`  : source += V*rhs (EXT-OF9 fvMatrix operator==)
    for_ranks(cs, [&](int r) {
        Rank& rk = cs.ranks[r];
        for (int c = 0; c < rk.mesh.nCells; ++c)
            for (int q = 0; q < 6; ++q) rk.source[(size_t)6 * c + q] += rk.mesh.V[c] * rk.rhs[(size_t)6 * c + q];
    });

    // --- thetaEqn.relax()  (EXT-OF9 fvMatrix::relax; only with a relaxation factor in fvSolution)
    const bool isFluidity = cs.ranks[0].modes[mi].model.d.model == RHEO_MODEL_BMP_FLUIDITY;
    if (isFluidity) {   // == - fvm::Sp(1/lambda, Phi) (BMPLog.C:158): EXT-OF9 fvm::Sp adds V sp to the diagonal; `A == B` is A - B
        for_ranks(cs, [&](int r) {
            Rank& rk = cs.ranks[r];
            const double rl = 1.0 / rk.modes[mi].model.d.lambda;
            for (int c = 0; c < rk.mesh.nCells; ++c) rk.diag[c] += rl * rk.mesh.V[c];
        });
    }
    const double relaxFactor = isFluidity ? cs.ranks[0].modes[mi].model.d.bmp_relax : ctl.relax;   // PhiEqn.relax() / thetaEqn.relax()
    if (relaxFactor > 0) {
        const double alpha = relaxFactor;
        for_ranks(cs, [&](int r) {
            Rank& rk = cs.ranks[r];
            Mode& mo = rk.modes[mi];
            const Mesh& m = rk.mesh;
            dvec D0 = rk.diag, sumOff(m.nCells, 0.0);
            dvec& D = rk.diag;
            for (int f = 0; f < m.nInt; ++f) { sumOff[m.nei[f]] += std::fabs(rk.lower[f]); sumOff[m.own[f]] += std::fabs(rk.upper[f]); }
            for (const Patch& p : m.patches) {
                if (p.type == RHEO_PATCH_EMPTY) continue;
                for (int f = p.start; f < p.start + p.size; ++f) {
                    const int b = f - m.nInt;
                    if (p.type == RHEO_PATCH_PROCESSOR) { D[m.own[f]] += rk.iC[b]; sumOff[m.own[f]] += std::fabs(rk.bC[(size_t)6 * b]); }
                    else D[m.own[f]] += std::fabs(rk.iC[b]);   // cmptMax(cmptMag(iCoeffs))
                }
            }
            for (int c = 0; c < m.nCells; ++c) D[c] = std::max(std::fabs(D[c]), sumOff[c]);
            for (int c = 0; c < m.nCells; ++c) D[c] /= alpha;
            for (const Patch& p : m.patches) {
                if (p.type == RHEO_PATCH_EMPTY) continue;
                for (int f = p.start; f < p.start + p.size; ++f) D[m.own[f]] -= rk.iC[f - m.nInt];   // coupled: component 0; else cmptMin
            }
            for (int c = 0; c < m.nCells; ++c)
                for (int q = 0; q < 6; ++q) rk.source[(size_t)6 * c + q] += (D[c] - D0[c]) * mo.theta[(size_t)6 * c + q];
        });
    }

    // --- thetaEqn.solve()  (EXT-OF9 fvMatrix::solveSegregated; in-repo restatement
    //     sparseMatrixSolvers/segregated/sparseSolver.C:72-127 and eigenSolver/eigenSolver.C:409-563)
    for_ranks(cs, [&](int r) {   // addBoundarySource: non-coupled only (coupled part cancels)
        Rank& rk = cs.ranks[r];
        const Mesh& m = rk.mesh;
        for (const Patch& p : m.patches) {
            if (p.type == RHEO_PATCH_EMPTY || p.type == RHEO_PATCH_PROCESSOR) continue;
            for (int f = p.start; f < p.start + p.size; ++f)
                for (int q = 0; q < 6; ++q) rk.source[(size_t)6 * m.own[f] + q] += rk.bC[(size_t)6 * (f - m.nInt) + q];
        }
    });
    int maxIters = 0;
    for (int cmpt = 0; cmpt < 6; ++cmpt) {
        if (st) { st->initial_residual[cmpt] = 0; st->final_residual[cmpt] = 0; st->n_iterations[cmpt] = 0; st->converged[cmpt] = 1; }
        if (!cs.ranks[0].mesh.solved[cmpt]) continue;
        Ldu A;
        A.cs = &cs;
        A.lower.resize(R); A.upper.resize(R); A.diag.resize(R); A.ifBou.resize(R); A.ifInt.resize(R);
        DVec psi = make_vec(cs), src = make_vec(cs);
        for_ranks(cs, [&](int r) {
            Rank& rk = cs.ranks[r];
            const Mesh& m = rk.mesh;
            A.lower[r] = rk.lower.data();
            A.upper[r] = rk.upper.data();
            A.diag[r] = rk.diag;
            A.ifBou[r].assign(m.nB(), 0.0);
            A.ifInt[r] = rk.iC;
            for (const Patch& p : m.patches) {   // addBoundaryDiag
                if (p.type == RHEO_PATCH_EMPTY) continue;
                for (int f = p.start; f < p.start + p.size; ++f) {
                    A.diag[r][m.own[f]] += rk.iC[f - m.nInt];
                    A.ifBou[r][f - m.nInt] = rk.bC[(size_t)6 * (f - m.nInt) + cmpt];
                }
            }
            for (int c = 0; c < m.nCells; ++c) { psi.v[r][c] = rk.modes[mi].theta[(size_t)6 * c + cmpt]; src.v[r][c] = rk.source[(size_t)6 * c + cmpt]; }
        });
        Perf pf = (ctl.solver == RHEO_SOLVER_PBICG) ? pbicg(A, psi, src, ctl) : pbicgstab(A, psi, src, ctl);
        for_ranks(cs, [&](int r) {
            Rank& rk = cs.ranks[r];
            for (int c = 0; c < rk.mesh.nCells; ++c) rk.modes[mi].theta[(size_t)6 * c + cmpt] = psi.v[r][c];
        });
        if (st) { st->initial_residual[cmpt] = pf.init; st->final_residual[cmpt] = pf.fin; st->n_iterations[cmpt] = pf.iters; st->converged[cmpt] = pf.conv ? 1 : 0; }
        maxIters = std::max(maxIters, pf.iters);
    }
    cs.lastIters = std::max(cs.lastIters, maxIters);

    // --- theta.correctBoundaryConditions(); calcEig; tau; tau.correctBoundaryConditions()
    for_ranks(cs, [&](int r) {
        Rank& rk = cs.ranks[r];
        Mode& mo = rk.modes[mi];
        const Mesh& m = rk.mesh;
        for (const Patch& p : m.patches)
            if (p.theta_bc == RHEO_BC_ZERO_GRADIENT && p.type != RHEO_PATCH_EMPTY)
                for (int f = p.start; f < p.start + p.size; ++f)
                    for (int q = 0; q < 6; ++q) mo.thetaB[(size_t)6 * (f - m.nInt) + q] = mo.theta[(size_t)6 * m.own[f] + q];
        if (isFluidity) return;   // the fluidity has no eigen-decomposition and no stress of its own
        for (int c = 0; c < m.nCells; ++c) {
            calc_eig_cell(&mo.theta[(size_t)6 * c], &mo.eigVals[(size_t)9 * c], &mo.eigVecs[(size_t)9 * c], cs.sortEig);
            T9 Rm, Lam;
            std::memcpy(Rm.v, &mo.eigVecs[(size_t)9 * c], 72);
            std::memcpy(Lam.v, &mo.eigVals[(size_t)9 * c], 72);
            Model tmp;
            const Model* pm = &mo.model;
            if (mo.per_cell()) { tmp = mo.at(c); pm = &tmp; }
            tau_cell(*pm, Rm, Lam, rk.fFene[c], &mo.tau[(size_t)6 * c]);
        }
    });
    // tau BCs: processor values first (all sends complete before any evaluate), then the physical
    // patches in patch order; linearExtrapolation (linearExtrapolationFvPatchField.C:101-151) uses a
    // full Gauss-linear gradient per component with the boundary values current at that moment.
    if (isFluidity) return 0;
    for_ranks(cs, [&](int r) {
        Rank& rk = cs.ranks[r];
        Mode& mo = rk.modes[mi];
        const Mesh& m = rk.mesh;
        dvec fb((size_t)6 * m.nB()), g((size_t)18 * m.nCells);
        if (cs.tauAssign) {
            const RheoModelDesc& d = mo.model.d;
            const double e = d.model == RHEO_MODEL_OLDROYD_B_LOG ? -d.etaP / d.lambda : 0.0;
            for (const Patch& p : m.patches)
                if (p.type != RHEO_PATCH_EMPTY && p.type != RHEO_PATCH_PROCESSOR && p.tau_bc == RHEO_BC_ZERO_GRADIENT)
                    for (int f = p.start; f < p.start + p.size; ++f) {
                        double* t = &mo.tauB[(size_t)6 * (f - m.nInt)];
                        t[0] = e; t[1] = 0; t[2] = 0; t[3] = e; t[4] = 0; t[5] = e;
                    }
        }
        for (const Patch& p : m.patches) {
            if (p.type == RHEO_PATCH_EMPTY || p.type == RHEO_PATCH_PROCESSOR) continue;
            if (p.tau_bc == RHEO_BC_ZERO_GRADIENT) {
                for (int f = p.start; f < p.start + p.size; ++f)
                    for (int q = 0; q < 6; ++q) mo.tauB[(size_t)6 * (f - m.nInt) + q] = mo.tau[(size_t)6 * m.own[f] + q];
            } else if (p.tau_bc == RHEO_BC_LINEAR_EXTRAPOLATION) {
                if (p.size == 0) continue;
                face_values_boundary(cs, r, 6, [&](int q) { return cs.ranks[q].modes[mi].tau.data(); }, mo.tauB.data(), fb.data());
                gauss_grad(m, 6, mo.tau.data(), fb.data(), g.data());
                dvec varp((size_t)6 * p.size);
                for (int i = 0; i < p.size; ++i) {
                    const int f = p.start + i, cellA = m.own[f];
                    const double CtoF[3] = {m.Cf[3 * (size_t)f] - m.C[3 * (size_t)cellA], m.Cf[3 * (size_t)f + 1] - m.C[3 * (size_t)cellA + 1], m.Cf[3 * (size_t)f + 2] - m.C[3 * (size_t)cellA + 2]};
                    for (int q = 0; q < 6; ++q) {
                        const double* gq = &g[(size_t)18 * cellA + 3 * q];
                        varp[(size_t)6 * i + q] = mo.tau[(size_t)6 * cellA + q] + (gq[0] * CtoF[0] + gq[1] * CtoF[1] + gq[2] * CtoF[2]);
                    }
                }
                std::copy(varp.begin(), varp.end(), mo.tauB.begin() + (size_t)6 * (p.start - m.nInt));
            } else if (p.tau_bc == RHEO_BC_LINEAR_EXTRAPOLATION_REG) {
                // linearExtrapolationFvPatchField.C:152-219 (useRegression true): least-squares line y(x) through the values linearly
                // interpolated to the wall cell's INTERNAL faces (coupled faces are skipped, :181-183) and the cell value, x = wall
                // distance of the face / cell centre along the patch-face normal; wall value = y(0) = yav - xav num/den.
                // Faces in the order of EXT-OF9 primitiveMesh::cells(): the cell's owner faces, then its neighbour faces, ascending.
                std::vector<int> listOf(m.nCells, -1);
                std::vector<std::vector<int>> cellFaces;
                for (int i = 0; i < p.size; ++i) {
                    const int c = m.own[p.start + i];
                    if (listOf[c] < 0) { listOf[c] = (int)cellFaces.size(); cellFaces.emplace_back(); }
                }
                for (int f = 0; f < m.nInt; ++f) if (listOf[m.own[f]] >= 0) cellFaces[listOf[m.own[f]]].push_back(f);
                for (int f = 0; f < m.nInt; ++f) if (listOf[m.nei[f]] >= 0) cellFaces[listOf[m.nei[f]]].push_back(f);
                for (int i = 0; i < p.size; ++i) {
                    const int fp = p.start + i, cellA = m.own[fp];
                    const double* Sp = &m.Sf[3 * (size_t)fp];
                    const double magSp = std::sqrt(Sp[0] * Sp[0] + Sp[1] * Sp[1] + Sp[2] * Sp[2]);
                    const double n[3] = {Sp[0] / magSp, Sp[1] / magSp, Sp[2] / magSp};
                    const double* fx = &m.Cf[3 * (size_t)fp];
                    std::vector<double> x;
                    std::vector<std::array<double, 6>> y;
                    double xav = 0;
                    std::array<double, 6> yav{};
                    auto add_face = [&](int f) {
                        const int P = m.own[f], N = m.nei[f];
                        const double* S = &m.Sf[3 * (size_t)f];
                        const double* cf = &m.Cf[3 * (size_t)f];
                        double so = 0, sn = 0, xd = 0;
                        for (int d = 0; d < 3; ++d) {
                            so += S[d] * (cf[d] - m.C[3 * (size_t)P + d]);
                            sn += S[d] * (m.C[3 * (size_t)N + d] - cf[d]);
                            xd += n[d] * (fx[d] - cf[d]);
                        }
                        const double SfdOwn = std::fabs(so), SfdNei = std::fabs(sn);
                        const double w = SfdOwn / (SfdOwn + SfdNei);
                        std::array<double, 6> yy;
                        for (int q = 0; q < 6; ++q) { yy[q] = w * mo.tau[(size_t)6 * N + q] + (1. - w) * mo.tau[(size_t)6 * P + q]; yav[q] += yy[q]; }
                        y.push_back(yy);
                        x.push_back(std::fabs(xd));
                        xav += x.back();
                    };
                    for (int f : cellFaces[listOf[cellA]]) add_face(f);
                    {   // last pair: the cell itself
                        std::array<double, 6> yy;
                        double xd = 0;
                        for (int q = 0; q < 6; ++q) { yy[q] = mo.tau[(size_t)6 * cellA + q]; yav[q] += yy[q]; }
                        for (int d = 0; d < 3; ++d) xd += n[d] * (fx[d] - m.C[3 * (size_t)cellA + d]);
                        y.push_back(yy);
                        x.push_back(std::fabs(xd));
                        xav += x.back();
                    }
                    const int id = (int)x.size();
                    for (int q = 0; q < 6; ++q) yav[q] /= id;
                    xav /= id;
                    double den = 0;
                    std::array<double, 6> num{};
                    for (int k = 0; k < id; ++k) {
                        for (int q = 0; q < 6; ++q) num[q] += (x[k] - xav) * (y[k][q] - yav[q]);
                        den += (x[k] - xav) * (x[k] - xav);
                    }
                    for (int q = 0; q < 6; ++q) mo.tauB[(size_t)6 * (fp - m.nInt) + q] = yav[q] - xav * num[q] / den;
                }
            }   // fixedValue: unchanged
        }
    });
    return 0;
}

void finalize(Case& cs) {
    for (size_t r = 0; r < cs.ranks.size(); ++r) {
        Mesh& m = cs.ranks[r].mesh;
        for (Patch& p : m.patches) {
            p.nbr_patch = -1;
            if (p.type != RHEO_PATCH_PROCESSOR) continue;
            const Mesh& o = cs.ranks[p.nbr_rank].mesh;
            for (size_t q = 0; q < o.patches.size(); ++q)
                if (o.patches[q].type == RHEO_PATCH_PROCESSOR && o.patches[q].nbr_rank == (int)r) p.nbr_patch = (int)q;
        }
        // lduAddressing::losort: faces ordered by neighbour (upper) cell, stable
        m.losort.resize(m.nInt);
        for (int f = 0; f < m.nInt; ++f) m.losort[f] = f;
        std::stable_sort(m.losort.begin(), m.losort.end(), [&](int a, int b) { return m.nei[a] < m.nei[b]; });
    }
    cs.finalized = true;
}

void init_model(Model& mo) {
    mo.mlMaxIter = mo.d.ml_max_iter;
    mo.gammaVals.clear();
    if (mo.d.model == RHEO_MODEL_PTT_LOG && mo.d.ptt_function == RHEO_PTT_GENERALIZED) {   // PTTLog.C:143-170
        mo.gammaVals.push_back(std::tgamma(mo.d.ml_beta));
        int k = 0;
        while (k < mo.mlMaxIter && mo.gammaVals.back() < 1e+100) {
            mo.gammaVals.push_back(std::tgamma(mo.d.ml_alpha * k + mo.d.ml_beta));
            k++;
        }
        mo.mlMaxIter = k;
    }
}

}  // namespace

// ==================================================================== C API (ctypes)
// ------------------------------------------------------------------ explicit part of constitutiveEq::divTau
// CE/constitutiveEq/constitutiveEq.C:72-132 (stabilization none / BSD / coupling) summed over the modes as multiMode.C:143-157
// does; the divSchemes div(tau) and div(grad(U)) are `Gauss linear` in every tutorial.  EXT-OF9 (builder's reading, not in
// /root/reference): gaussDivScheme::fvcDiv = fvc::surfaceIntegrate(Sf & linearInterpolate(vf)); gaussGrad::calcGrad ends with
// correctBoundaryConditions: g_b += n (snGrad(U)_b - (n & g_b)) on non-coupled patches, g_b starting from the patch-internal
// value; fvPatchField::snGrad = deltaCoeffs (U_b - U_c), deltaCoeffs = 1/|delta|, fvPatch::delta = n (n & (Cf - Cn)).
// Fields here: X[9*cell + 3i + j] = X_ij with i the index contracted with Sf.
static void gauss_div9(const Mesh& m, const double* X, const double* XfB, double sign, double* out) {
    dvec tmp((size_t)3 * m.nCells, 0.0);
    for (int f = 0; f < m.nInt; ++f) {
        const int P = m.own[f], N = m.nei[f];
        const double* S = &m.Sf[3 * (size_t)f];
        for (int j = 0; j < 3; ++j) {
            double v = 0;
            for (int i = 0; i < 3; ++i) {
                const double xf = m.w[f] * (X[(size_t)9 * P + 3 * i + j] - X[(size_t)9 * N + 3 * i + j]) + X[(size_t)9 * N + 3 * i + j];
                v += S[i] * xf;
            }
            tmp[(size_t)3 * P + j] += v;
            tmp[(size_t)3 * N + j] -= v;
        }
    }
    for (const Patch& p : m.patches) {
        if (p.type == RHEO_PATCH_EMPTY) continue;
        for (int f = p.start; f < p.start + p.size; ++f) {
            const int P = m.own[f];
            const double* S = &m.Sf[3 * (size_t)f];
            const double* xb = &XfB[(size_t)9 * (f - m.nInt)];
            for (int j = 0; j < 3; ++j) tmp[(size_t)3 * P + j] += S[0] * xb[j] + S[1] * xb[3 + j] + S[2] * xb[6 + j];
        }
    }
    for (int c = 0; c < m.nCells; ++c)
        for (int j = 0; j < 3; ++j) out[(size_t)3 * c + j] += sign * (tmp[(size_t)3 * c + j] / m.V[c]);
}

// out[r]: [3 * nCells] per rank
static int div_tau_explicit(Case& cs, int stab, std::vector<dvec>& out) {
    const int R = (int)cs.ranks.size();
    out.assign(R, dvec());
    for (int r = 0; r < R; ++r) out[r].assign((size_t)3 * cs.ranks[r].mesh.nCells, 0.0);
    const int nModes = (int)cs.ranks[0].modes.size();
    std::vector<dvec> X(R), XB(R);
    auto sym9 = [](const double* t, double s, double* x) {
        x[0] = s * t[0]; x[1] = s * t[1]; x[2] = s * t[2]; x[3] = s * t[1]; x[4] = s * t[3]; x[5] = s * t[4]; x[6] = s * t[2]; x[7] = s * t[4]; x[8] = s * t[5];
    };
    for (int mi = 0; mi < nModes; ++mi) {
        if (cs.ranks[0].modes[mi].fluidityOf >= 0) continue;   // BMPLog's hidden fluidity mode has no stress and no etaP of its own
        // ---- fvc::div(tau/rho)
        for_ranks(cs, [&](int r) {
            const Rank& rk = cs.ranks[r];
            const Mode& mo = rk.modes[mi];
            const double rr = 1.0 / mo.model.d.rho;
            X[r].resize((size_t)9 * rk.mesh.nCells); XB[r].resize((size_t)9 * rk.mesh.nB());
            for (int c = 0; c < rk.mesh.nCells; ++c) sym9(&mo.tau[(size_t)6 * c], rr, &X[r][(size_t)9 * c]);
            for (int b = 0; b < rk.mesh.nB(); ++b) sym9(&mo.tauB[(size_t)6 * b], rr, &XB[r][(size_t)9 * b]);
        });
        for_ranks(cs, [&](int r) {
            const Mesh& m = cs.ranks[r].mesh;
            dvec fb((size_t)9 * m.nB(), 0.0);
            face_values_boundary(cs, r, 9, [&](int q) { return X[q].data(); }, XB[r].data(), fb.data());
            gauss_div9(m, X[r].data(), fb.data(), 1.0, out[r].data());
        });
        if (stab != RHEO_STAB_COUPLING) continue;
        // ---- - fvc::div((etaP/rho) fvc::grad(U))
        for_ranks(cs, [&](int r) {
            Rank& rk = cs.ranks[r];
            const Mesh& m = rk.mesh;
            const Mode& mo = rk.modes[mi];
            const double coef = mo.model.d.etaP / mo.model.d.rho;
            dvec fb((size_t)3 * m.nB(), 0.0), g((size_t)9 * m.nCells);
            face_values_boundary(cs, r, 3, [&](int q) { return cs.ranks[q].U.data(); }, rk.Ub.data(), fb.data());
            gauss_grad(m, 3, rk.U.data(), fb.data(), g.data());   // g[9c + 3k + d] = d_d U_k
            X[r].resize((size_t)9 * m.nCells); XB[r].assign((size_t)9 * m.nB(), 0.0);
            for (int c = 0; c < m.nCells; ++c)
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) X[r][(size_t)9 * c + 3 * i + j] = g[(size_t)9 * c + 3 * j + i];   // (grad U)_ij = d_i U_j, not yet scaled
            for (const Patch& p : m.patches) {
                if (p.type == RHEO_PATCH_EMPTY || p.type == RHEO_PATCH_PROCESSOR) continue;
                for (int f = p.start; f < p.start + p.size; ++f) {
                    const int c = m.own[f], b = f - m.nInt;
                    const double* S = &m.Sf[3 * (size_t)f];
                    const double magS = std::sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2]);
                    const double n[3] = {S[0] / magS, S[1] / magS, S[2] / magS};
                    double nd = 0;
                    for (int d = 0; d < 3; ++d) nd += n[d] * (m.Cf[3 * (size_t)f + d] - m.C[3 * (size_t)c + d]);
                    const double deltaCoeff = 1.0 / std::fabs(nd);
                    for (int j = 0; j < 3; ++j) {
                        const double sn = deltaCoeff * (rk.Ub[(size_t)3 * b + j] - rk.U[(size_t)3 * c + j]);
                        double ng = 0;
                        for (int k = 0; k < 3; ++k) ng += n[k] * X[r][(size_t)9 * c + 3 * k + j];
                        for (int i = 0; i < 3; ++i) XB[r][(size_t)9 * b + 3 * i + j] = X[r][(size_t)9 * c + 3 * i + j] + n[i] * (sn - ng);
                    }
                }
            }
            for (double& x : X[r]) x *= coef;
            for (double& x : XB[r]) x *= coef;
        });
        for_ranks(cs, [&](int r) {
            const Mesh& m = cs.ranks[r].mesh;
            dvec fb((size_t)9 * m.nB(), 0.0);
            face_values_boundary(cs, r, 9, [&](int q) { return X[q].data(); }, XB[r].data(), fb.data());
            gauss_div9(m, X[r].data(), fb.data(), -1.0, out[r].data());
        });
    }
    return 0;
}

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

// number of threads the rank loop uses (0 restores OMP_NUM_THREADS); returns the count a parallel region actually gets
int orc_set_num_threads(int n) {
    g_threads = n > 0 ? n : 0;
    omp_set_dynamic(0);
    int got = 1;
    const int nt = g_threads > 0 ? g_threads : omp_get_max_threads();
#pragma omp parallel num_threads(nt)
    {
#pragma omp single
        got = omp_get_num_threads();
    }
    return got;
}

void* orc_create(int n_ranks) {
    auto* cs = new Case();
    cs->ranks.resize(n_ranks);
    cs->ctl = RheoSchemeCtl{RHEO_LIMITER_CUBISTA, RHEO_DDT_EULER, RHEO_SOLVER_PBICG, 1e-10, 0.0, 0, 1000, 0.0};
    return cs;
}
void orc_destroy(void* h) { delete (Case*)h; }

int orc_set_mesh(void* h, int rank, const RheoMeshDesc* d) {
    Case& cs = *(Case*)h;
    if (rank < 0 || rank >= (int)cs.ranks.size()) { g_err = "orc_set_mesh: bad rank"; return 1; }
    Mesh& m = cs.ranks[rank].mesh;
    m.nCells = d->n_cells; m.nFaces = d->n_faces; m.nInt = d->n_internal_faces;
    m.own.assign(d->owner, d->owner + d->n_faces);
    m.nei.assign(d->neighbour, d->neighbour + d->n_internal_faces);
    m.Sf.assign(d->Sf, d->Sf + 3 * (size_t)d->n_faces);
    m.Cf.assign(d->Cf, d->Cf + 3 * (size_t)d->n_faces);
    m.C.assign(d->C, d->C + 3 * (size_t)d->n_cells);
    m.V.assign(d->V, d->V + d->n_cells);
    m.w.assign(d->weights, d->weights + d->n_faces);
    const size_t nb = (size_t)m.nB();
    if (d->nbr_C) m.nbrC.assign(d->nbr_C, d->nbr_C + 3 * nb); else m.nbrC.assign(3 * nb, 0.0);
    m.patches.clear();
    for (int p = 0; p < d->n_patches; ++p) {
        const RheoPatchDesc& q = d->patches[p];
        m.patches.push_back(Patch{q.type, q.start, q.size, q.nbr_rank, q.theta_bc, q.tau_bc, -1});
    }
    for (int q = 0; q < 6; ++q) m.solved[q] = d->solved_components[q];
    cs.finalized = false;
    return 0;
}

int orc_add_mode(void* h, const RheoModelDesc* d) {
    Case& cs = *(Case*)h;
    for (Rank& rk : cs.ranks) {
        Mode mo;
        mo.model.d = *d;
        init_model(mo.model);
        const size_t n = rk.mesh.nCells, nb = rk.mesh.nB();
        mo.theta.assign(6 * n, 0.0); mo.thetaOld.assign(6 * n, 0.0); mo.tau.assign(6 * n, 0.0);
        mo.eigVals.assign(9 * n, 0.0); mo.eigVecs.assign(9 * n, 0.0);
        for (size_t c = 0; c < n; ++c)
            for (int q = 0; q < 3; ++q) { mo.eigVals[9 * c + 4 * q] = 1.0; mo.eigVecs[9 * c + 4 * q] = 1.0; }
        mo.thetaB.assign(6 * nb, 0.0); mo.tauB.assign(6 * nb, 0.0);
        rk.modes.push_back(std::move(mo));
    }
    if (d->model == RHEO_MODEL_BMP_LOG) {   // its fluidity equation: a hidden mode behind the public ones
        if (cs.ranks[0].modes.size() != 1) { g_err = "oracle: BMPLog is restated as a single-mode model"; return 3; }
        RheoModelDesc f = *d;
        f.model = RHEO_MODEL_BMP_FLUIDITY;
        const int rc = orc_add_mode(h, &f);
        if (rc) return rc;
        for (Rank& rk : cs.ranks) { rk.modes[0].fluidityMode = 1; rk.modes[1].fluidityOf = 0; }
    } else if (d->model != RHEO_MODEL_BMP_FLUIDITY && cs.ranks[0].modes[0].fluidityMode >= 0) { g_err = "oracle: BMPLog is restated as a single-mode model"; return 3; }
    return 0;
}

int orc_set_state(void* h, int rank, int mode, const double* theta, const double* tau, const double* eigvals,
                  const double* eigvecs, const double* theta_b, const double* tau_b);
// correct(alpha, gradU) with a caller-supplied gradient (NULL: back to fvc::grad(U))
int orc_set_grad_u(void* h, int rank, const double* gradU9) {
    Case& cs = *(Case*)h;
    Rank& rk = cs.ranks[rank];
    if (!gradU9) rk.gradUext.clear();
    else rk.gradUext.assign(gradU9, gradU9 + 9 * (size_t)rk.mesh.nCells);
    return 0;
}

// BMPLog: the fluidity field of `mode` (Phi, MUST_READ in BMPLog.C:112-122)
int orc_set_fluidity(void* h, int rank, int mode, const double* Phi, const double* Phi_b) {
    Case& cs = *(Case*)h;
    Rank& rk = cs.ranks[rank];
    if (mode < 0 || mode >= (int)rk.modes.size() || rk.modes[mode].fluidityMode < 0) { g_err = "orc_set_fluidity: not a BMPLog mode"; return 3; }
    const size_t n = rk.mesh.nCells, nb = rk.mesh.nB();
    dvec th(6 * n, 0.0), thb(6 * nb, 0.0);
    for (size_t c = 0; c < n; ++c) th[6 * c] = Phi[c];
    if (Phi_b) for (size_t b = 0; b < nb; ++b) thb[6 * b] = Phi_b[b];
    return orc_set_state(h, rank, rk.modes[mode].fluidityMode, th.data(), nullptr, nullptr, nullptr, thb.data(), nullptr);
}

int orc_set_schemes(void* h, const RheoSchemeCtl* c) { ((Case*)h)->ctl = *c; return 0; }
int orc_set_sort_eig(void* h, int on) { ((Case*)h)->sortEig = on != 0; return 0; }
int orc_set_tau_assignment(void* h, int on) { ((Case*)h)->tauAssign = on != 0; return 0; }

int orc_set_state(void* h, int rank, int mode, const double* theta, const double* tau, const double* eigvals,
                  const double* eigvecs, const double* theta_b, const double* tau_b) {
    Case& cs = *(Case*)h;
    Rank& rk = cs.ranks[rank];
    Mode& mo = rk.modes[mode];
    const size_t n = rk.mesh.nCells, nb = rk.mesh.nB();
    if (theta) { mo.theta.assign(theta, theta + 6 * n); mo.thetaOld = mo.theta; mo.thetaOldOld = mo.theta; cs.nOldTimes = 0; mo.ddt0.clear(); mo.ddt0TimeIndex = 0; }
    if (tau) mo.tau.assign(tau, tau + 6 * n);
    if (eigvals) mo.eigVals.assign(eigvals, eigvals + 9 * n);
    if (eigvecs) mo.eigVecs.assign(eigvecs, eigvecs + 9 * n);
    // EXT-OF9 zeroGradientFvPatchField(p, iF, dict): the constructor evaluates (value = patchInternalField), whatever the
    // file's `value` entry says — so a zeroGradient patch never keeps caller-supplied boundary values
    if (theta_b) mo.thetaB.assign(theta_b, theta_b + 6 * nb);
    for (const Patch& p : rk.mesh.patches)
        if (p.theta_bc == RHEO_BC_ZERO_GRADIENT && p.type != RHEO_PATCH_EMPTY)
            for (int f = p.start; f < p.start + p.size; ++f)
                for (int q = 0; q < 6; ++q) mo.thetaB[6 * (size_t)(f - rk.mesh.nInt) + q] = mo.theta[6 * (size_t)rk.mesh.own[f] + q];
    if (tau_b) mo.tauB.assign(tau_b, tau_b + 6 * nb);
    else {
        for (const Patch& p : rk.mesh.patches)
            if (p.tau_bc == RHEO_BC_ZERO_GRADIENT && p.type != RHEO_PATCH_EMPTY)
                for (int f = p.start; f < p.start + p.size; ++f)
                    for (int q = 0; q < 6; ++q) mo.tauB[6 * (size_t)(f - rk.mesh.nInt) + q] = mo.tau[6 * (size_t)rk.mesh.own[f] + q];
    }
    return 0;
}

// thermo-dependent lambda / etaP per cell (NULL, NULL: back to the scalars of the model)
int orc_set_thermo(void* h, int rank, int mode, const double* lambdaCell, const double* etaPCell) {
    Case& cs = *(Case*)h;
    if (rank < 0 || rank >= (int)cs.ranks.size() || mode < 0 || mode >= (int)cs.ranks[rank].modes.size()) { g_err = "orc_set_thermo: bad rank/mode"; return 1; }
    Mode& mo = cs.ranks[rank].modes[mode];
    const int n = cs.ranks[rank].mesh.nCells;
    if (!lambdaCell || !etaPCell) { mo.lambdaCell.clear(); mo.etaPCell.clear(); return 0; }
    mo.lambdaCell.assign(lambdaCell, lambdaCell + n);
    mo.etaPCell.assign(etaPCell, etaPCell + n);
    return 0;
}

int orc_set_velocity(void* h, int rank, const double* U, const double* Ub, const double* phi) {
    Case& cs = *(Case*)h;
    Rank& rk = cs.ranks[rank];
    rk.U.assign(U, U + 3 * (size_t)rk.mesh.nCells);
    rk.Ub.assign(Ub, Ub + 3 * (size_t)rk.mesh.nB());
    rk.phi.assign(phi, phi + rk.mesh.nFaces);
    return 0;
}

int orc_store_old_time(void* h) {
    Case& cs = *(Case*)h;
    for (Rank& rk : cs.ranks)
        for (Mode& mo : rk.modes) { mo.thetaOldOld = mo.thetaOld; mo.thetaOld = mo.theta; }
    cs.nOldTimes++;
    cs.dt0 = cs.dtNow;   // Time::operator++: deltaT0_ = deltaT_
    return 0;
}

// multiMode::correct (CE/multiMode/multiMode.C:247-260): modes one after the other
int orc_step(void* h, double dt, RheoStepStats* stats) {
    Case& cs = *(Case*)h;
    if (!cs.finalized) finalize(cs);
    cs.lastIters = 0;
    cs.dtNow = dt;
    const int nm = (int)cs.ranks[0].modes.size();
    for (int mi = 0; mi < nm; ++mi) {
        if (cs.ranks[0].modes[mi].fluidityOf >= 0) continue;   // hidden: solved with the BMPLog mode it belongs to
        const int fm = cs.ranks[0].modes[mi].fluidityMode;
        if (fm >= 0) {   // BMPLog.C:151-166: the fluidity equation first; theta then sees the new Phi
            int rc = correct_mode(cs, fm, dt, nullptr);
            if (rc) return rc;
            for (Rank& rk : cs.ranks) rk.modes[mi].fluidity = &rk.modes[fm].theta;
        }
        int rc = correct_mode(cs, mi, dt, stats ? &stats[mi] : nullptr);
        if (rc) return rc;
    }
    return 0;
}
int orc_last_iterations(void* h) { return ((Case*)h)->lastIters; }

int orc_get(void* h, int rank, int mode, int field, double* out) {
    Case& cs = *(Case*)h;
    Rank& rk = cs.ranks[rank];
    const dvec* src = nullptr;
    dvec tot;
    switch (field) {
        case RHEO_FIELD_THETA: src = &rk.modes[mode].theta; break;
        case RHEO_FIELD_TAU: src = &rk.modes[mode].tau; break;
        case RHEO_FIELD_EIGVALS: src = &rk.modes[mode].eigVals; break;
        case RHEO_FIELD_EIGVECS: src = &rk.modes[mode].eigVecs; break;
        case RHEO_FIELD_THETA_B: src = &rk.modes[mode].thetaB; break;
        case RHEO_FIELD_TAU_B: src = &rk.modes[mode].tauB; break;
        case RHEO_FIELD_THETA_OLD: src = &rk.modes[mode].thetaOld; break;
        case RHEO_FIELD_FLUIDITY: case RHEO_FIELD_FLUIDITY_B: {
            if (rk.modes[mode].fluidityMode < 0) { g_err = "orc_get: not a BMPLog mode"; return 3; }
            const dvec& f = field == RHEO_FIELD_FLUIDITY ? rk.modes[rk.modes[mode].fluidityMode].theta : rk.modes[rk.modes[mode].fluidityMode].thetaB;
            for (size_t i = 0; i < f.size() / 6; ++i) out[i] = f[6 * i];
            return 0;
        }
        case RHEO_FIELD_TAU_TOTAL:   // multiMode::tau(), multiMode.C:216-226
            tot.assign(rk.modes[0].tau.size(), 0.0);
            for (Mode& mo : rk.modes) for (size_t q = 0; q < tot.size(); ++q) tot[q] += mo.tau[q];
            src = &tot;
            break;
        default: g_err = "orc_get: unknown field"; return 1;
    }
    std::copy(src->begin(), src->end(), out);
    return 0;
}

// ---- stand-alone pieces for unit tests
void orc_jacobi(int n, const double* theta6, double* D3, double* V9) {
    for (int c = 0; c < n; ++c) jacobi_ref(theta6 + 6 * (size_t)c, D3 + 3 * (size_t)c, V9 + 9 * (size_t)c);
}
void orc_calc_eig(int n, const double* theta6, double* vals9, double* vecs9, int sortEig) {
    for (int c = 0; c < n; ++c) calc_eig_cell(theta6 + 6 * (size_t)c, vals9 + 9 * (size_t)c, vecs9 + 9 * (size_t)c, sortEig != 0);
}
// Omega and B of decomposeGradU for n cells
void orc_decompose_gradU(int n, const double* L9, const double* R9, const double* Lam9, double zeta, int ptt, double* omega9, double* B9) {
    for (int c = 0; c < n; ++c) {
        T9 L, R, Lam, o, B;
        std::memcpy(L.v, L9 + 9 * (size_t)c, 72); std::memcpy(R.v, R9 + 9 * (size_t)c, 72); std::memcpy(Lam.v, Lam9 + 9 * (size_t)c, 72);
        decompose_gradU_cell(L, R, Lam, zeta, ptt != 0, o, B);
        std::memcpy(omega9 + 9 * (size_t)c, o.v, 72); std::memcpy(B9 + 9 * (size_t)c, B.v, 72);
    }
}
void orc_model_rhs(const RheoModelDesc* d, int n, const double* L9, const double* theta6, const double* R9, const double* Lam9, double* rhs6, double* f) {
    Model mo;
    mo.d = *d;
    init_model(mo);
    for (int c = 0; c < n; ++c) {
        T9 L, R, Lam;
        std::memcpy(L.v, L9 + 9 * (size_t)c, 72); std::memcpy(R.v, R9 + 9 * (size_t)c, 72); std::memcpy(Lam.v, Lam9 + 9 * (size_t)c, 72);
        double ff = model_rhs_cell(mo, L, theta6 + 6 * (size_t)c, R, Lam, rhs6 + 6 * (size_t)c);
        if (f) f[c] = ff;
    }
}
// the same with the model's current tau of each cell (read by SaramitoLog only)
void orc_model_rhs_tau(const RheoModelDesc* d, int n, const double* L9, const double* theta6, const double* R9, const double* Lam9, const double* tau6,
                       double* rhs6, double* f) {
    Model mo;
    mo.d = *d;
    init_model(mo);
    for (int c = 0; c < n; ++c) {
        T9 L, R, Lam;
        std::memcpy(L.v, L9 + 9 * (size_t)c, 72); std::memcpy(R.v, R9 + 9 * (size_t)c, 72); std::memcpy(Lam.v, Lam9 + 9 * (size_t)c, 72);
        double ff = model_rhs_cell(mo, L, theta6 + 6 * (size_t)c, R, Lam, rhs6 + 6 * (size_t)c, tau6 ? tau6 + 6 * (size_t)c : nullptr);
        if (f) f[c] = ff;
    }
}
void orc_tau(const RheoModelDesc* d, int n, const double* R9, const double* Lam9, const double* f, double* tau6) {
    Model mo;
    mo.d = *d;
    init_model(mo);
    for (int c = 0; c < n; ++c) {
        T9 R, Lam;
        std::memcpy(R.v, R9 + 9 * (size_t)c, 72); std::memcpy(Lam.v, Lam9 + 9 * (size_t)c, 72);
        tau_cell(mo, R, Lam, f ? f[c] : 0.0, tau6 + 6 * (size_t)c);
    }
}
// Gauss-linear gradient of a scalar cell field on rank 0's mesh with given boundary face values
// explicit part of divTau for `rank` (3 per cell); every rank is evaluated (the processor faces need the neighbours)
int orc_div_tau(void* h, int rank, int stabilization, double* out) {
    Case& cs = *(Case*)h;
    if (stabilization != RHEO_STAB_NONE && stabilization != RHEO_STAB_BSD && stabilization != RHEO_STAB_COUPLING) { g_err = "orc_div_tau: unknown stabilization"; return 3; }
    std::vector<dvec> all;
    if (div_tau_explicit(cs, stabilization, all)) return 3;
    std::copy(all[rank].begin(), all[rank].end(), out);
    return 0;
}

int orc_gauss_grad(void* h, int rank, int nc, const double* psi, const double* faceB, double* out) {
    Case& cs = *(Case*)h;
    gauss_grad(cs.ranks[rank].mesh, nc, psi, faceB, out);
    return 0;
}

}  // extern "C"
