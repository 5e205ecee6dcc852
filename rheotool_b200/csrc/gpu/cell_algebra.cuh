// cell_algebra.cuh — per-cell 3x3 FP64 algebra of the log-conformation models, register resident.
//
// Device restatement (one thread per cell, everything in registers, no temporaries in memory) of
//   CE/utils/boilerLog.H:26-34, CE/constitutiveEq/constitutiveEq.C:323-358 (Omega/B split),
//   CE/utils/jacobi.H:7-158 (cyclic Jacobi; same rotation formulas and thresholds, early exit once all
//   off-diagonals are exactly zero — the remaining reference sweeps are then no-ops),
//   the model sources Oldroyd_BLog.C:146-163, GiesekusLog.C:142-157, PTTLog.C:190-251,
//   FENE_PLog.C:142-163 and the theta->tau maps Oldroyd_BLog.C:175, GiesekusLog.C:172, PTTLog.C:264,
//   FENE_PLog.C:178.   (CE = of90/src/libs/constitutiveEquations/constitutiveEqs)
// Tensors are row-major double[9]; symmTensors double[6] = xx,xy,xz,yy,yz,zz.
#pragma once
#include "rheo_gpu.h"

namespace rk {

struct ModelParams {
    int model, ptt_function, ml_max_iter;
    double etaP, lambda, alpha, epsilon, zeta, L2, ml_rtol, gamma_beta, wmK, wmN, wmA,
        rpLambdaR, rpBeta, rpDelta, rpChiMax, xppLambdaS, xppQ, xppN,
        sarTau0, sarK, sarN, sarD0, sarD1, sarD2;   // SaramitoLog.C:113-130
    int sarPtt;                                     // 0 none, 1 linear, 2 exponential (n == 1 only)
    double bmpK, bmpPhi0, bmpPhiInf;                // BMPLog.C:129-136 (the fluidity equation; theta runs as Oldroyd-BLog on per-cell rates)
    const double* gamma_vals;   // device table Gamma(alpha k + beta), PTTLog.C:143-170
};

__device__ __forceinline__ void mat_mul(const double* a, const double* b, double* r) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
// r = a^T b
__device__ __forceinline__ void mat_tmul(const double* a, const double* b, double* r) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
// r = a b^T
__device__ __forceinline__ void mat_mult(const double* a, const double* b, double* r) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
}
// s = symm(R diag(d) R^T)  (exactly symmetric by construction)
__device__ __forceinline__ void rdrt_sym(const double* R, double d0, double d1, double d2, double* s) {
    s[0] = d0 * R[0] * R[0] + d1 * R[1] * R[1] + d2 * R[2] * R[2];
    s[1] = d0 * R[0] * R[3] + d1 * R[1] * R[4] + d2 * R[2] * R[5];
    s[2] = d0 * R[0] * R[6] + d1 * R[1] * R[7] + d2 * R[2] * R[8];
    s[3] = d0 * R[3] * R[3] + d1 * R[4] * R[4] + d2 * R[5] * R[5];
    s[4] = d0 * R[3] * R[6] + d1 * R[4] * R[7] + d2 * R[5] * R[8];
    s[5] = d0 * R[6] * R[6] + d1 * R[7] * R[7] + d2 * R[8] * R[8];
}
__device__ __forceinline__ void sym_to_full(const double* s, double* t) {
    t[0] = s[0]; t[1] = s[1]; t[2] = s[2]; t[3] = s[1]; t[4] = s[3]; t[5] = s[4]; t[6] = s[2]; t[7] = s[4]; t[8] = s[5];
}

__device__ __forceinline__ double mittag_leffler(const ModelParams& mp, double z) {   // PTTLog.C:202-236
    double sum = 0, sumOld = 0, err = 1;
    int k = 0;
    while (k < mp.ml_max_iter && err > mp.ml_rtol) {
        double e = pow(z, (double)k) / mp.gamma_vals[k + 1];
        sumOld = sum;
        sum += e;
        err = fabs((sumOld - sum) / (sumOld + 1e-12));
        k++;
    }
    return sum;
}

// rhs6 = symm(Omega.theta - theta.Omega + 2B + G);  returns the FENE-P / FENE-CR f (0 otherwise).
// MODEL is a compile-time constant: every model gets its own instance of the source kernel (no dead code, fewer registers).
// L: grad(U), L_ij = d_i U_j.  R: eigenvectors in columns.  lam: exp(eigenvalues).
// tau6: the model's current tau of the cell (SaramitoLog only; nullptr otherwise).
template <int MODEL>
__device__ __forceinline__ double model_rhs(const ModelParams& mp, const double* L, const double* th6, const double* R,
                                            const double* lam, double* rhs6, const double* tau6 = nullptr) {
    if (MODEL == RHEO_MODEL_BMP_FLUIDITY) {
        // BMPLog.C:158-160: Phi0/lambda + k (PhiInf - Phi) (tau && symm(L)); th6[0] = Phi, tau6 = the BMPLog mode's stress.
        // (-Sp(1/lambda, Phi) is on the diagonal: the assembly gets 1/dt + 1/lambda.)
        const double sxy = 0.5 * (L[1] + L[3]), sxz = 0.5 * (L[2] + L[6]), syz = 0.5 * (L[5] + L[7]);
        const double tD = tau6[0] * L[0] + tau6[3] * L[4] + tau6[5] * L[8] + 2.0 * (tau6[1] * sxy + tau6[2] * sxz + tau6[4] * syz);
        rhs6[0] = mp.bmpPhi0 / mp.lambda + mp.bmpK * (mp.bmpPhiInf - th6[0]) * tD;
#pragma unroll
        for (int q = 1; q < 6; ++q) rhs6[q] = 0.0;
        return 0.0;
    }
    // X = L^T (- zeta symm(L) for PTT and Saramito)          boilerLog.H:26-32
    double X[9] = {L[0], L[3], L[6], L[1], L[4], L[7], L[2], L[5], L[8]};
    if (MODEL == RHEO_MODEL_PTT_LOG || MODEL == RHEO_MODEL_SARAMITO_LOG) {
        const double z = mp.zeta;
        const double sxy = 0.5 * (L[1] + L[3]), sxz = 0.5 * (L[2] + L[6]), syz = 0.5 * (L[5] + L[7]);
        X[0] -= z * L[0]; X[4] -= z * L[4]; X[8] -= z * L[8];
        X[1] -= z * sxy; X[3] -= z * sxy; X[2] -= z * sxz; X[6] -= z * sxz; X[5] -= z * syz; X[7] -= z * syz;
    }
    double T[9], M[9];
    mat_tmul(R, X, T);      // R^T X
    mat_mul(T, R, M);       // (R^T X) R                        constitutiveEq.C:506
    // decomposeGradU, constitutiveEq.C:333-350
    const double lx = lam[0], ly = lam[1], lz = lam[2];
    const double oxy = (ly * M[1] + lx * M[3]) / (ly - lx + 1e-16);
    const double oxz = (lz * M[2] + lx * M[6]) / (lz - lx + 1e-16);
    const double oyz = (lz * M[5] + ly * M[7]) / (lz - ly + 1e-16);
    const double om[9] = {0, oxy, oxz, -oxy, 0, oyz, -oxz, -oyz, 0};
    double Om[9];
    mat_mul(R, om, T);
    mat_mult(T, R, Om);     // Omega = (R omega) R^T            constitutiveEq.C:355,513
    double B6[6];
    rdrt_sym(R, M[0], M[4], M[8], B6);   // B = R diag(M) R^T   constitutiveEq.C:356
    double th[9], A1[9], A2[9];
    sym_to_full(th6, th);
    mat_mul(Om, th, A1);    // omega & theta
    mat_mul(th, Om, A2);    // theta & omega
    // model term G (symmetric): eigen-frame diagonal g_k, G = R diag(g) R^T, except Giesekus' quadratic term
    double f = 0.0;
    double g0, g1, g2;
    double il = 1.0 / mp.lambda;
    if (MODEL == RHEO_MODEL_WM_CY_LOG) {
        // WhiteMetznerCYLog.C:155-163: etaP, lambda *= (1 + (K sqrt(2) |symm L|)^a)^((n-1)/a); f carries etaP/lambda to theta->tau
        const double sxy = 0.5 * (L[1] + L[3]), sxz = 0.5 * (L[2] + L[6]), syz = 0.5 * (L[5] + L[7]);
        const double magD = sqrt(L[0] * L[0] + L[4] * L[4] + L[8] * L[8] + 2.0 * (sxy * sxy + sxz * sxz + syz * syz));
        const double cy = pow(1.0 + pow(mp.wmK * sqrt(2.0) * magD, mp.wmA), (mp.wmN - 1.0) / mp.wmA);
        const double lamC = mp.lambda * cy, etaC = mp.etaP * cy;
        il = 1.0 / lamC;
        f = etaC / lamC;
    }
    const double i0 = 1.0 / lx, i1 = 1.0 / ly, i2 = 1.0 / lz;
    double G6[6];
    if (MODEL == RHEO_MODEL_OLDROYD_B_LOG || MODEL == RHEO_MODEL_WM_CY_LOG) {
        g0 = il * (i0 - 1.0); g1 = il * (i1 - 1.0); g2 = il * (i2 - 1.0);
        rdrt_sym(R, g0, g1, g2, G6);
    } else if (MODEL == RHEO_MODEL_GIESEKUS_LOG) {
        // trhs - alpha A trhs^2 with A, trhs co-diagonal in R:  (1/L-1) - alpha L (1/L-1)^2
        const double t0 = i0 - 1.0, t1 = i1 - 1.0, t2 = i2 - 1.0;
        g0 = il * (t0 - mp.alpha * (lx * (t0 * t0)));
        g1 = il * (t1 - mp.alpha * (ly * (t1 * t1)));
        g2 = il * (t2 - mp.alpha * (lz * (t2 * t2)));
        rdrt_sym(R, g0, g1, g2, G6);
    } else if (MODEL == RHEO_MODEL_PTT_LOG) {
        double A6[6];
        rdrt_sym(R, lx, ly, lz, A6);
        const double z = (mp.epsilon / (1.0 - mp.zeta)) * ((A6[0] + A6[3] + A6[5]) - 3.0);
        double Y;
        if (mp.ptt_function == RHEO_PTT_LINEAR) Y = 1.0 + z;
        else if (mp.ptt_function == RHEO_PTT_EXPONENTIAL) Y = exp(z);
        else Y = mp.gamma_beta * mittag_leffler(mp, z);
        g0 = il * (i0 - 1.0) * Y; g1 = il * (i1 - 1.0) * Y; g2 = il * (i2 - 1.0) * Y;
        rdrt_sym(R, g0, g1, g2, G6);
    } else if (MODEL == RHEO_MODEL_SARAMITO_LOG) {
        // SaramitoLog.C:153-238: `thetaEqn -= symm((fac etaP/lambda) Y R (1/Lambda - I) R^T)` adds the term to the right-hand
        // side; fac = max(0, (|tau_d| - tau0)/(k |tau_d|^n + 1e-16))^(1/n) of the CURRENT tau, |tau_d| = mag(dev tau)/sqrt(2)
        const double nDims = mp.sarD0 + mp.sarD1 + mp.sarD2;
        const double trT = tau6[0] + tau6[3] + tau6[5];
        const double d0 = tau6[0] - mp.sarD0 * trT / nDims, d3 = tau6[3] - mp.sarD1 * trT / nDims, d5 = tau6[5] - mp.sarD2 * trT / nDims;
        const double tauDMag = sqrt(d0 * d0 + d3 * d3 + d5 * d5 + 2.0 * (tau6[1] * tau6[1] + tau6[2] * tau6[2] + tau6[4] * tau6[4])) / sqrt(2.0);
        double fac;
        if (mp.sarN == 1.0) fac = fmax(0.0, (tauDMag - mp.sarTau0) / (mp.sarK * tauDMag + 1e-16));
        else fac = pow(fmax(0.0, (tauDMag - mp.sarTau0) / (mp.sarK * pow(tauDMag, mp.sarN) + 1e-16)), 1.0 / mp.sarN);
        double Y = 1.0;
        if (mp.sarN == 1.0 && mp.sarPtt != 0) {
            double A6[6];
            rdrt_sym(R, lx, ly, lz, A6);
            const double z = (mp.epsilon / (1.0 - mp.zeta)) * ((A6[0] + A6[3] + A6[5]) - 3.0);
            Y = mp.sarPtt == 1 ? 1.0 + z : exp(z);
        }
        const double cf = (fac * mp.etaP / mp.lambda) * Y;
        g0 = cf * (i0 - 1.0); g1 = cf * (i1 - 1.0); g2 = cf * (i2 - 1.0);
        rdrt_sym(R, g0, g1, g2, G6);
    } else if (MODEL == RHEO_MODEL_FENE_CR_LOG) {   // FENE_CRLog.C:143-163: (f/lambda) R (1/Lambda - I) R^T
        double A6[6];
        rdrt_sym(R, lx, ly, lz, A6);
        f = mp.L2 / (mp.L2 - (A6[0] + A6[3] + A6[5]));
        g0 = il * f * (i0 - 1.0); g1 = il * f * (i1 - 1.0); g2 = il * f * (i2 - 1.0);
        rdrt_sym(R, g0, g1, g2, G6);
    } else if (MODEL == RHEO_MODEL_ROLIE_POLY_LOG) {
        // RoliePolyLog.C:144-186: -(1/lambdaD) A^-1 ((A - I) + M1 lambdaD (A + beta (trA/3)^delta (A - I))); every factor is
        // a function of A, so the product is R diag(g) R^T with the scalar expression per eigenvalue
        double A6[6];
        rdrt_sym(R, lx, ly, lz, A6);
        const double trA = A6[0] + A6[3] + A6[5];
        double M1 = 2.0 * (1.0 - sqrt(3.0 / trA)) / mp.rpLambdaR;
        if (mp.rpChiMax > 1.0) {
            const double c2 = mp.rpChiMax * mp.rpChiMax;
            M1 *= ((3.0 - (trA / 3.0) / c2) * (1.0 - 1.0 / c2)) / ((1.0 - (trA / 3.0) / c2) * (3.0 - 1.0 / c2));
        }
        const double bt = mp.rpBeta * pow(trA / 3.0, mp.rpDelta);
        const double m1l = M1 * mp.lambda;
        g0 = -il * i0 * ((lx - 1.0) + m1l * (lx + bt * (lx - 1.0)));
        g1 = -il * i1 * ((ly - 1.0) + m1l * (ly + bt * (ly - 1.0)));
        g2 = -il * i2 * ((lz - 1.0) + m1l * (lz + bt * (lz - 1.0)));
        rdrt_sym(R, g0, g1, g2, G6);
    } else if (MODEL == RHEO_MODEL_XPOMPOM_LOG) {
        // XPomPomLog.C:148-183: -(1/lambdaB) A^-1 (A (f - 2 alpha) + alpha A.A + (alpha - 1) I)
        double A6[6];
        rdrt_sym(R, lx, ly, lz, A6);
        const double trA = A6[0] + A6[3] + A6[5];
        const double trAA = A6[0] * A6[0] + A6[3] * A6[3] + A6[5] * A6[5] + 2.0 * (A6[1] * A6[1] + A6[2] * A6[2] + A6[4] * A6[4]);
        const double ls = sqrt(trA / 3.0);
        const double stretch = mp.xppN == 0.0 ? (1.0 - 1.0 / ls) : (1.0 - 1.0 / pow(ls, mp.xppN + 1.0));
        const double fx = 2.0 * (mp.lambda / mp.xppLambdaS) * exp((2.0 / mp.xppQ) * (ls - 1.0)) * stretch +
                          (1.0 / (ls * ls)) * (1.0 - mp.alpha - (mp.alpha / 3.0) * (trAA - 2.0 * trA));
        g0 = -il * i0 * (lx * (fx - 2.0 * mp.alpha) + mp.alpha * (lx * lx) + (mp.alpha - 1.0));
        g1 = -il * i1 * (ly * (fx - 2.0 * mp.alpha) + mp.alpha * (ly * ly) + (mp.alpha - 1.0));
        g2 = -il * i2 * (lz * (fx - 2.0 * mp.alpha) + mp.alpha * (lz * lz) + (mp.alpha - 1.0));
        rdrt_sym(R, g0, g1, g2, G6);
    } else {   // FENE-P
        double A6[6];
        rdrt_sym(R, lx, ly, lz, A6);
        f = mp.L2 / (mp.L2 - (A6[0] + A6[3] + A6[5]));
        const double a = mp.L2 / (mp.L2 - 3.0);
        g0 = il * (a * i0 - f); g1 = il * (a * i1 - f); g2 = il * (a * i2 - f);
        rdrt_sym(R, g0, g1, g2, G6);
    }
    // symm(A1 - A2 + 2B + G)
    rhs6[0] = (A1[0] - A2[0]) + 2.0 * B6[0] + G6[0];
    rhs6[1] = 0.5 * ((A1[1] - A2[1]) + (A1[3] - A2[3])) + 2.0 * B6[1] + G6[1];
    rhs6[2] = 0.5 * ((A1[2] - A2[2]) + (A1[6] - A2[6])) + 2.0 * B6[2] + G6[2];
    rhs6[3] = (A1[4] - A2[4]) + 2.0 * B6[3] + G6[3];
    rhs6[4] = 0.5 * ((A1[5] - A2[5]) + (A1[7] - A2[7])) + 2.0 * B6[4] + G6[4];
    rhs6[5] = (A1[8] - A2[8]) + 2.0 * B6[5] + G6[5];
    return f;
}

// one Jacobi rotation in the (p,q) plane; r is the third index.  a_pr / a_qr are the couplings to r.
// Rotation formulas of CE/utils/jacobi.H:69-113.
#define RK_ROT(app, aqq, apq, apr, aqr, zp, zq, vp0, vq0, vp1, vq1, vp2, vq2, sweep, tresh)        \
    {                                                                                              \
        double g_ = 100.0 * fabs(apq);                                                             \
        if ((sweep > 4) && (fabs(app) + g_ == fabs(app)) && (fabs(aqq) + g_ == fabs(aqq))) {       \
            apq = 0.0;                                                                             \
        } else if (fabs(apq) > tresh) {                                                            \
            double h_ = aqq - app, t_;                                                             \
            if (fabs(h_) + g_ == fabs(h_)) {                                                       \
                t_ = apq / h_;                                                                     \
            } else {                                                                               \
                double th_ = 0.5 * h_ / apq;                                                       \
                t_ = 1.0 / (fabs(th_) + sqrt(1.0 + th_ * th_));                                    \
                if (th_ < 0) t_ = -t_;                                                             \
            }                                                                                      \
            double c_ = 1.0 / sqrt(1.0 + t_ * t_), s_ = t_ * c_, tau_ = s_ / (1.0 + c_);          \
            h_ = t_ * apq;                                                                         \
            zp -= h_; zq += h_; app -= h_; aqq += h_;                                              \
            apq = 0.0;                                                                             \
            double gg_ = apr, hh_ = aqr;                                                           \
            apr = gg_ - s_ * (hh_ + gg_ * tau_); aqr = hh_ + s_ * (gg_ - hh_ * tau_);              \
            gg_ = vp0; hh_ = vq0; vp0 = gg_ - s_ * (hh_ + gg_ * tau_); vq0 = hh_ + s_ * (gg_ - hh_ * tau_); \
            gg_ = vp1; hh_ = vq1; vp1 = gg_ - s_ * (hh_ + gg_ * tau_); vq1 = hh_ + s_ * (gg_ - hh_ * tau_); \
            gg_ = vp2; hh_ = vq2; vp2 = gg_ - s_ * (hh_ + gg_ * tau_); vq2 = hh_ + s_ * (gg_ - hh_ * tau_); \
        }                                                                                          \
    }

// Eigen-decomposition of the symmetric 3x3 th6: eigenvalues ascending in d[3] (the order of
// Eigen::SelfAdjointEigenSolver, constitutiveEq.C:390-414), eigenvectors in the columns of V.
__device__ __forceinline__ void jacobi_eig(const double* th6, double* d, double* V) {
    double a01 = th6[1], a02 = th6[2], a12 = th6[4];
    double d0 = th6[0], d1 = th6[3], d2 = th6[5];
    double b0 = d0, b1 = d1, b2 = d2, z0 = 0, z1 = 0, z2 = 0;
    double v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;
    for (int sweep = 1; sweep <= 50; ++sweep) {
        const double sm = fabs(a01) + fabs(a02) + fabs(a12);
        if (sm == 0.0) break;
        const double tresh = (sweep < 4) ? 0.2 * sm * sm : 0.0;
        RK_ROT(d0, d1, a01, a02, a12, z0, z1, v00, v01, v10, v11, v20, v21, sweep, tresh)   // (0,1), third = 2
        RK_ROT(d0, d2, a02, a01, a12, z0, z2, v00, v02, v10, v12, v20, v22, sweep, tresh)   // (0,2), third = 1
        RK_ROT(d1, d2, a12, a01, a02, z1, z2, v01, v02, v11, v12, v21, v22, sweep, tresh)   // (1,2), third = 0
        b0 += z0; b1 += z1; b2 += z2;
        d0 = b0; d1 = b1; d2 = b2;
        z0 = z1 = z2 = 0;
    }
    // sort ascending (stable), permuting the columns of V
    double e0 = d0, e1 = d1, e2 = d2;
    double c00 = v00, c10 = v10, c20 = v20, c01 = v01, c11 = v11, c21 = v21, c02 = v02, c12 = v12, c22 = v22;
#define RK_SWAP(x, y) { double t_ = x; x = y; y = t_; }
    if (e1 < e0) { RK_SWAP(e0, e1) RK_SWAP(c00, c01) RK_SWAP(c10, c11) RK_SWAP(c20, c21) }
    if (e2 < e1) { RK_SWAP(e1, e2) RK_SWAP(c01, c02) RK_SWAP(c11, c12) RK_SWAP(c21, c22) }
    if (e1 < e0) { RK_SWAP(e0, e1) RK_SWAP(c00, c01) RK_SWAP(c10, c11) RK_SWAP(c20, c21) }
#undef RK_SWAP
    d[0] = e0; d[1] = e1; d[2] = e2;
    V[0] = c00; V[1] = c01; V[2] = c02; V[3] = c10; V[4] = c11; V[5] = c12; V[6] = c20; V[7] = c21; V[8] = c22;
}

// tau from (R, Lambda); fOld = FENE-P / FENE-CR f computed before the solve (FENE_PLog.C:142,178; FENE_CRLog.C:141,174)
template <int MODEL>
__device__ __forceinline__ void tau_from_eig(const ModelParams& mp, const double* R, const double* lam, double fOld, double* tau6) {
    double A6[6];
    rdrt_sym(R, lam[0], lam[1], lam[2], A6);
    double coef = mp.etaP / mp.lambda;
    if (MODEL == RHEO_MODEL_FENE_P_LOG) {
        const double a = mp.L2 / (mp.L2 - 3.0);
        tau6[0] = coef * (fOld * A6[0] - a); tau6[1] = coef * (fOld * A6[1]); tau6[2] = coef * (fOld * A6[2]);
        tau6[3] = coef * (fOld * A6[3] - a); tau6[4] = coef * (fOld * A6[4]); tau6[5] = coef * (fOld * A6[5] - a);
    } else {
        if (MODEL == RHEO_MODEL_PTT_LOG || MODEL == RHEO_MODEL_SARAMITO_LOG) coef = mp.etaP / (mp.lambda * (1.0 - mp.zeta));   // SaramitoLog.C:242
        if (MODEL == RHEO_MODEL_FENE_CR_LOG) coef = (mp.etaP / mp.lambda) * fOld;   // FENE_CRLog.C:174: f of BEFORE the solve
        if (MODEL == RHEO_MODEL_WM_CY_LOG) coef = fOld;                              // WhiteMetznerCYLog.C:207: etaP/lambda of BEFORE the solve
        if (MODEL == RHEO_MODEL_ROLIE_POLY_LOG && mp.rpChiMax > 1.0) {               // RoliePolyLog.C:203-212: finite extensibility, NEW tr(A)
            const double trA = A6[0] + A6[3] + A6[5], c2 = mp.rpChiMax * mp.rpChiMax;
            coef *= ((3.0 - (trA / 3.0) / c2) * (1.0 - 1.0 / c2)) / ((1.0 - (trA / 3.0) / c2) * (3.0 - 1.0 / c2));
        }
        tau6[0] = coef * (A6[0] - 1.0); tau6[1] = coef * A6[1]; tau6[2] = coef * A6[2];
        tau6[3] = coef * (A6[3] - 1.0); tau6[4] = coef * A6[4]; tau6[5] = coef * (A6[5] - 1.0);
    }
}

}  // namespace rk
