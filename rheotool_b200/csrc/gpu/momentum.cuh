// momentum.cuh — the explicit part of constitutiveEq::divTau on the device (first half), so that the
// stress never has to leave HBM for the momentum predictor: per cell 3 doubles go to the host instead of 6.
//
// Reference (of90/src/libs/constitutiveEquations/constitutiveEqs/constitutiveEq/constitutiveEq.C:72-132; multiMode.C:143-157 sums
// the modes' matrices):
//     stabilization none      fvc::div(tau/rho, "div(tau)")                                                  + fvm::laplacian(etaS/rho, U)
//     stabilization BSD       fvc::div(tau/rho) - fvc::laplacian(etaP/rho, U)                                + fvm::laplacian((etaP+etaS)/rho, U)
//     stabilization coupling  fvc::div(tau/rho) - fvc::div((etaP/rho) fvc::grad(U), "div(grad(U))")          + fvm::laplacian((etaP+etaS)/rho, U)
// The fvm:: term (an implicit matrix in U) and BSD's fvc::laplacian (it follows the case's laplacian / snGrad schemes) stay with
// the caller.  What is evaluated here is, with both divSchemes `Gauss linear` (every tutorial: 57 x div(tau), 55 x div(grad(U))),
//     out = sum_modes fvc::div(tau_m / rho_m)  -  [coupling]  fvc::div((sum_modes etaP_m / rho_m) fvc::grad(U))
// as ONE Gauss-linear divergence of  X = sum_m tau_m / rho_m - coef grad(U)  (fvc::div is linear in its argument; the reference
// evaluates the terms one by one, so the two agree to rounding, not bit by bit).  EXT-OF9 gaussDivScheme::fvcDiv =
// surfaceIntegrate(Sf & interpolate(X)):  internal face  Sf & (w (X_P - X_N) + X_N)  added to the owner, subtracted from the
// neighbour;  processor face  Sf & (w X_P + (1 - w) X_nbr);  other patch faces  Sf & X_b;  then / V.
//
// Boundary values of X: tau_b are the stress patch values (k_tau_bc_*); the patch values of fvc::grad(U) are those EXT-OF9
// gaussGrad::correctBoundaryConditions leaves:  g_b = g_c + n (snGrad(U) - n & g_c),  snGrad(U) = deltaCoeffs (U_b - U_c),
// deltaCoeffs = 1 / |delta|, delta = n (n & (Cf - C_c))  (builder's reading of OpenFOAM-9's fvPatch::delta; on the orthogonal
// wall cells of every tutorial mesh |Cf - C_c| is the same number); processor patches hold the neighbour cell's gradient.
#pragma once
#include "kernels.cuh"

namespace rk {

constexpr int MAX_MODES_DIV = 8;
struct DivTauArgs {
    int nModes;
    const double* tau[MAX_MODES_DIV];    // [6][NP] per mode (ghost values current: the step ends with the tau halo swap)
    const double* tauB[MAX_MODES_DIV];   // [6][nB]
    double rRho[MAX_MODES_DIV];          // 1 / rho_m
    double coefGrad;                     // sum_m etaP_m / rho_m  (coupling), 0 otherwise
    const double* gradU;                 // [9][NP], g[3k+d] = d_d U_k  (ghost values current when coefGrad != 0)
    const double* gradUb;                // [9][nB] patch values of fvc::grad(U), same component order (k_gradU_patch)
    const double* U; const double* Ub;   // [3][NP], [3][nB]
    const int* perm;                     // new -> caller's cell number
    const int* tileOrder; int nTiles;    // block ordering: 32-cell tiles in geometric order (host/ordering.hpp), or null — a cell's
                                         // out-of-tile neighbours live in another colour, i.e. far away in memory; walking the tiles
                                         // in space order lets the CTAs in flight share them in L2 (ncu r2f: 43 GB read for 7.7 GB of fields)
    double* out;                         // [N][3] in the caller's numbering
};

// symmTensor component of (i, j): xx xy xz yy yz zz
__device__ __forceinline__ int symIdx(int i, int j) { return i == j ? (i == 0 ? 0 : (i == 1 ? 3 : 5)) : (i + j == 1 ? 1 : (i + j == 2 ? 2 : 4)); }

// patch values of fvc::grad(U) on the non-coupled, non-empty boundary faces (one thread per boundary face)
__global__ void k_gradU_patch(MeshView m, const double* __restrict__ gradU, const double* __restrict__ U, const double* __restrict__ Ub,
                              double* __restrict__ gradUb) {
    pdl_sync();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= m.nB) return;
    const int kind = m.bkind[b];
    if (kind == RHEO_PATCH_EMPTY || kind == RHEO_PATCH_PROCESSOR) return;
    const int c = m.bcell[b];
    const size_t f = (size_t)m.nInt + b;
    double n[3] = {m.Sf[f], m.Sf[(size_t)m.nF + f], m.Sf[2 * (size_t)m.nF + f]};
    const double magS = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
#pragma unroll
    for (int d = 0; d < 3; ++d) n[d] /= magS;
    double nd = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) nd += n[d] * (m.CfB[(size_t)d * m.nB + b] - m.C[(size_t)d * m.NP + c]);
    const double deltaCoeff = 1.0 / fabs(nd);
#pragma unroll
    for (int k = 0; k < 3; ++k) {   // component U_k
        double g[3], ng = 0;
#pragma unroll
        for (int d = 0; d < 3; ++d) { g[d] = gradU[(size_t)(3 * k + d) * m.NP + c]; ng += n[d] * g[d]; }
        const double sn = deltaCoeff * (Ub[(size_t)k * m.nB + b] - U[(size_t)k * m.NP + c]);
#pragma unroll
        for (int d = 0; d < 3; ++d) gradUb[(size_t)(3 * k + d) * m.nB + b] = g[d] + n[d] * (sn - ng);
    }
}

// X_ij at a cell (i = row: the index contracted with Sf) — tau is symmetric, (grad U)_ij = d_i U_j = g[3j+i]
__device__ __forceinline__ void div_tau_X(const DivTauArgs& a, size_t stride, size_t idx, bool boundary, double (&X)[9]) {
    double t[6] = {0, 0, 0, 0, 0, 0};
    for (int mI = 0; mI < a.nModes; ++mI) {
        const double* src = boundary ? a.tauB[mI] : a.tau[mI];
#pragma unroll
        for (int q = 0; q < 6; ++q) t[q] += src[(size_t)q * stride + idx] * a.rRho[mI];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) X[3 * i + j] = t[symIdx(i, j)];
    if (a.coefGrad != 0.0) {
        const double* g = boundary ? a.gradUb : a.gradU;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) X[3 * i + j] -= a.coefGrad * g[(size_t)(3 * j + i) * stride + idx];
    }
}

__global__ void __launch_bounds__(BLOCK) k_div_tau(MeshView m, DivTauArgs a) {
    pdl_sync();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (a.tileOrder) {
        const int t = c >> 5;
        if (t >= a.nTiles) return;
        c = a.tileOrder[t] * 32 + (c & 31);
    }
    if (c >= m.N) return;
    double Xo[9], acc[3] = {0, 0, 0};
    div_tau_X(a, (size_t)m.NP, (size_t)c, false, Xo);
    for (int s = 0; s < m.K; ++s) {
        const int nb = m.nbr[(size_t)s * m.NS + c];
        if (nb == -1) continue;
        const int fi = m.fidx[(size_t)s * m.NS + c];
        const size_t f = fi >= 0 ? fi : ~fi;
        const double sg = fi >= 0 ? 1.0 : -1.0;
        const double S[3] = {sg * m.Sf[f], sg * m.Sf[(size_t)m.nF + f], sg * m.Sf[2 * (size_t)m.nF + f]};
        double Xn[9], Xf[9];
        if (nb >= 0) {
            div_tau_X(a, (size_t)m.NP, (size_t)nb, false, Xn);
            const double w = m.w[f];
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                if (nb >= m.N) Xf[q] = w * Xo[q] + (1.0 - w) * Xn[q];
                else Xf[q] = fi >= 0 ? w * (Xo[q] - Xn[q]) + Xn[q] : w * (Xn[q] - Xo[q]) + Xo[q];
            }
        } else {
            div_tau_X(a, (size_t)m.nB, (size_t)(-nb - 2), true, Xf);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[j] += S[0] * Xf[j] + S[1] * Xf[3 + j] + S[2] * Xf[6 + j];
    }
    const double V = m.V[c];
    const size_t o = (size_t)a.perm[c];
#pragma unroll
    for (int j = 0; j < 3; ++j) a.out[o * 3 + j] = acc[j] / V;
}

}  // namespace rk
