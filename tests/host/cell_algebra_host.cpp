// cell_algebra_host.cpp — TEST INFRASTRUCTURE: compiles the DEVICE per-cell source (rheotool_b200/csrc/gpu/cell_algebra.cuh:
// model_rhs<MODEL>, tau_from_eig<MODEL>, jacobi_eig) for the host, unchanged, so that tests/test_device_cell_algebra_on_host.py
// can check the very text the kernels k_cell_source2 / k_eig_tau run against the oracle without a GPU.
// Only the two CUDA function qualifiers are defined away; everything else is the header as it is.
#include <math.h>
#include <string.h>
#include <vector>

#define __device__
#define __forceinline__ inline
#include "../../rheotool_b200/csrc/gpu/cell_algebra.cuh"

namespace {

// RheoModelDesc -> ModelParams: the mapping of rheo_gpu_create (engine.cu, "ModelParams& mp = md.mp; ...")
void params(const RheoModelDesc& q, rk::ModelParams& mp, std::vector<double>& gv) {
    memset(&mp, 0, sizeof(mp));
    mp.model = q.model; mp.ptt_function = q.ptt_function; mp.ml_max_iter = q.ml_max_iter;
    mp.etaP = q.etaP; mp.lambda = q.lambda; mp.alpha = q.alpha; mp.epsilon = q.epsilon; mp.zeta = q.zeta; mp.L2 = q.L2;
    mp.ml_rtol = q.ml_rtol; mp.gamma_beta = 1.0; mp.gamma_vals = nullptr;
    mp.wmK = q.wm_K; mp.wmN = q.wm_n; mp.wmA = q.wm_a;
    mp.rpLambdaR = q.rp_lambdaR; mp.rpBeta = q.rp_beta; mp.rpDelta = q.rp_delta; mp.rpChiMax = q.rp_chiMax;
    mp.xppLambdaS = q.xpp_lambdaS; mp.xppQ = q.xpp_q; mp.xppN = q.xpp_n;
    mp.sarTau0 = q.sar_tau0; mp.sarK = q.sar_k; mp.sarN = q.sar_n; mp.sarD0 = q.sar_dims[0]; mp.sarD1 = q.sar_dims[1]; mp.sarD2 = q.sar_dims[2];
    mp.sarPtt = q.sar_n == 1.0 ? q.sar_ptt : 0;
    if (q.model == RHEO_MODEL_PTT_LOG && q.ptt_function == RHEO_PTT_GENERALIZED) {
        gv.assign(1, tgamma(q.ml_beta));
        int k = 0;
        while (k < q.ml_max_iter && gv.back() < 1e+100) { gv.push_back(tgamma(q.ml_alpha * k + q.ml_beta)); k++; }
        mp.ml_max_iter = k;
        mp.gamma_beta = gv[0];
        mp.gamma_vals = gv.data();
    }
}

template <int MODEL>
void run_rhs(const rk::ModelParams& mp, int n, const double* L9, const double* th6, const double* R9, const double* lam3, const double* tau6,
             double* rhs6, double* f) {
    for (int c = 0; c < n; ++c)
        f[c] = rk::model_rhs<MODEL>(mp, L9 + 9 * (size_t)c, th6 + 6 * (size_t)c, R9 + 9 * (size_t)c, lam3 + 3 * (size_t)c, rhs6 + 6 * (size_t)c,
                                    MODEL == RHEO_MODEL_SARAMITO_LOG ? tau6 + 6 * (size_t)c : nullptr);
}
template <int MODEL>
void run_tau(const rk::ModelParams& mp, int n, const double* R9, const double* lam3, const double* f, double* tau6) {
    for (int c = 0; c < n; ++c) rk::tau_from_eig<MODEL>(mp, R9 + 9 * (size_t)c, lam3 + 3 * (size_t)c, f[c], tau6 + 6 * (size_t)c);
}

}  // namespace

#define DISPATCH(FN, ...)                                                                         \
    switch (d->model) {                                                                           \
        case RHEO_MODEL_OLDROYD_B_LOG: FN<RHEO_MODEL_OLDROYD_B_LOG>(__VA_ARGS__); break;          \
        case RHEO_MODEL_GIESEKUS_LOG: FN<RHEO_MODEL_GIESEKUS_LOG>(__VA_ARGS__); break;            \
        case RHEO_MODEL_PTT_LOG: FN<RHEO_MODEL_PTT_LOG>(__VA_ARGS__); break;                      \
        case RHEO_MODEL_FENE_P_LOG: FN<RHEO_MODEL_FENE_P_LOG>(__VA_ARGS__); break;                \
        case RHEO_MODEL_FENE_CR_LOG: FN<RHEO_MODEL_FENE_CR_LOG>(__VA_ARGS__); break;              \
        case RHEO_MODEL_WM_CY_LOG: FN<RHEO_MODEL_WM_CY_LOG>(__VA_ARGS__); break;                  \
        case RHEO_MODEL_ROLIE_POLY_LOG: FN<RHEO_MODEL_ROLIE_POLY_LOG>(__VA_ARGS__); break;        \
        case RHEO_MODEL_SARAMITO_LOG: FN<RHEO_MODEL_SARAMITO_LOG>(__VA_ARGS__); break;            \
        default: FN<RHEO_MODEL_XPOMPOM_LOG>(__VA_ARGS__); break;                                  \
    }

extern "C" {

// the body of k_cell_source2's per-cell call: lam3 = (exp(eig_0), exp(eig_1), exp(eig_2)), L9[3i+j] = d_i U_j
void hca_model_rhs(const RheoModelDesc* d, int n, const double* L9, const double* th6, const double* R9, const double* lam3, const double* tau6,
                   double* rhs6, double* f) {
    rk::ModelParams mp; std::vector<double> gv;
    params(*d, mp, gv);
    DISPATCH(run_rhs, mp, n, L9, th6, R9, lam3, tau6, rhs6, f)
}
// the body of k_eig_tau after the eigen-decomposition
void hca_tau(const RheoModelDesc* d, int n, const double* R9, const double* lam3, const double* f, double* tau6) {
    rk::ModelParams mp; std::vector<double> gv;
    params(*d, mp, gv);
    DISPATCH(run_tau, mp, n, R9, lam3, f, tau6)
}
// jacobi_eig: eigenvalues ascending (NOT exponentiated), eigenvectors in the columns of V (row-major)
void hca_eig(int n, const double* th6, double* d3, double* V9) {
    for (int c = 0; c < n; ++c) rk::jacobi_eig(th6 + 6 * (size_t)c, d3 + 3 * (size_t)c, V9 + 9 * (size_t)c);
}

}
