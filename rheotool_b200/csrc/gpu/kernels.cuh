// kernels.cuh — sm_100a kernels of the stress step.  All of them are HBM-bound FP64 streaming /
// gather kernels (no dense contraction anywhere => no tensor cores): one thread per cell, SoA planes,
// slot-major ELL connectivity so that every warp-level load is a contiguous 256-byte segment, cell
// numbering sorted by colour so that the DILU sweeps are a few fully parallel phases.
//
// Reference loops restated (of90/src/libs/...):
//   k_flux_assemble, k_cell_source2 (assembly.cuh)   gaussDefCmpwConvectionScheme.C:93-167, :242-319, boilerLog.H:1-36,
//                     the model sources, EXT-OF9 fvMatrix::relax, addBoundaryDiag/addBoundarySource
//   k_spmv*, k_sweep* EXT-OF9 lduMatrix::Amul, DILUPreconditioner::precondition
//   k_eig_tau         constitutiveEq.C:360-416 (calcEig) + theta->tau maps
//   k_tau_bc_linext   boundaryConditions/linearExtrapolation/linearExtrapolationFvPatchField.C:101-151
#pragma once
#include <cstdint>
#include "cell_algebra.cuh"

namespace rk {

// Programmatic dependent launch (sm_90+): every kernel of this library is launched with the programmatic-stream-
// serialization attribute and starts with this pair.  `wait` blocks until the previous kernel in the stream has completed
// and its memory is visible (a no-op without the attribute); `launch_dependents` lets the NEXT kernel's CTAs become
// resident as soon as every CTA of this one has started, so launch latency and ramp-up overlap this kernel's tail.
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

constexpr int BLOCK = 256;
constexpr int MAX_RHS = 24;
constexpr int MAX_RED = 3 * MAX_RHS;

// ---------------------------------------------------------------- device views
struct MeshView {
    int N, H, NT, NS, NP, K, nInt, nF, nB;
    const int* nbr;    // [K*NS] topological: >=0 cell/ghost, -1 empty, <=-2 boundary face -(b+2)
    const int* nbrA;   // [K*NS] algebraic: neighbour cell/ghost, self for empty/boundary slots; TILE-MAJOR, index ell_t(K, s, c)
    const int* fidx;   // [K*NS] face id, ~face when this cell is the face's neighbour
    const double* Sf;  // [3*nF] planes (x|y|z), stride nF, orientation of the (renumbered) owner
    const double* w;   // [nF]
    const double* C;   // [3*NP] planes, incl. ghost cell centres
    const double* V;   // [N]
    const double* rV;  // [N] 1/V
    const int* bcell;  // [nB] owner cell of boundary face
    const int* bkind;  // [nB] patch type
    const int* bthetaBC;  // [nB]
    const int* btauBC;    // [nB]
    const double* CfB;    // [3*nB] planes
    // block ordering only (host/ordering.hpp, blocksweep.cuh): levels of the in-chunk dependency graphs
    const uint16_t* lev;       // [NS] forward level | backward level << 8
    const uint16_t* chunkLev;  // [NS / 256] max forward level | max backward level << 8 of the chunk
};

struct RhsPtrs {      // one batch of right-hand sides sharing the matrix
    int n;
    double* psi[MAX_RHS];       // theta planes (solution, in place)
    const double* b[MAX_RHS];   // source planes
    const double* corr[MAX_RHS];   // [K*NS] deferred face values handed over by the upwind neighbours (assembly.cuh), or null
};

// ---------------------------------------------------------------- reductions
// Block-reduce NV per-thread values and store the block partials; the last block to finish sums
// the partials of all blocks in a fixed order (deterministic) into out[].
template <int NV>
__device__ __forceinline__ void block_reduce_to_partials(double (&v)[NV], double* partials, int slotBase, int nSlots) {
    __shared__ double sm[BLOCK / 32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) sm[warp][i] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double x = 0;
#pragma unroll
        for (int wv = 0; wv < BLOCK / 32; ++wv) x += sm[wv][threadIdx.x];
        partials[(size_t)blockIdx.x * nSlots + slotBase + threadIdx.x] = x;
    }
    __syncthreads();
}

// ---------------------------------------------------------------- layout transforms (upload / download)
// AoS in the caller's numbering -> SoA planes in device numbering
__global__ void k_aos_to_soa(int n, int nc, const int* __restrict__ perm, const double* __restrict__ aos, double* __restrict__ soa, int stride) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int o = perm ? perm[c] : c;
    for (int k = 0; k < nc; ++k) soa[(size_t)k * stride + c] = aos[(size_t)o * nc + k];
}
__global__ void k_soa_to_aos(int n, int nc, const int* __restrict__ perm, const double* __restrict__ soa, double* __restrict__ aos, int stride) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int o = perm ? perm[c] : c;
    for (int k = 0; k < nc; ++k) aos[(size_t)o * nc + k] = soa[(size_t)k * stride + c];
}
// eigVals tensor (9, diagonal used) <-> Lam[3]; accum != 0 adds instead of overwriting (sum over modes)
__global__ void k_soa_to_aos_acc(int n, int nc, const int* __restrict__ perm, const double* __restrict__ soa, double* __restrict__ aos, int stride) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int o = perm ? perm[c] : c;
    for (int k = 0; k < nc; ++k) aos[(size_t)o * nc + k] += soa[(size_t)k * stride + c];
}
__global__ void k_lam_from_tensor(int n, const int* __restrict__ perm, const double* __restrict__ aos9, double* __restrict__ lam, int stride) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int o = perm[c];
    lam[c] = aos9[(size_t)o * 9]; lam[stride + c] = aos9[(size_t)o * 9 + 4]; lam[2 * (size_t)stride + c] = aos9[(size_t)o * 9 + 8];
}
__global__ void k_lam_to_tensor(int n, const int* __restrict__ perm, const double* __restrict__ lam, double* __restrict__ aos9, int stride) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int o = perm[c];
    double* t = aos9 + (size_t)o * 9;
    t[0] = lam[c]; t[1] = 0; t[2] = 0; t[3] = 0; t[4] = lam[stride + c]; t[5] = 0; t[6] = 0; t[7] = 0; t[8] = lam[2 * (size_t)stride + c];
}
// phi: caller's face order -> device face order (flip sign where the renumbered owner changed)
__global__ void k_phi_in(int nF, const int* __restrict__ faceOld, const double* __restrict__ src, double* __restrict__ dst) {
    pdl_sync();
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    const int o = faceOld[f];   // old face + 1, negative if flipped
    dst[f] = o > 0 ? src[o - 1] : -src[-o - 1];
}
// CrankNicolson ddt0 update (EXT-OF9 CrankNicolsonDdtScheme::fvmDdt): ddt0 = a (theta_old - theta_oldold) - off ddt0
__global__ void k_cn_ddt0(size_t n, double a, double off, const double* __restrict__ thOld, const double* __restrict__ thOldOld, double* __restrict__ ddt0) {
    pdl_sync();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ddt0[i] = a * (thOld[i] - thOldOld[i]) - off * ddt0[i];
}
__global__ void k_fill(size_t n, double* p, double v) {
    pdl_sync();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// ---------------------------------------------------------------- halo pack / unpack
// record layout: buf[h * n + p] for ghost h and plane p (a neighbour's segment is a contiguous range of h)
struct PlaneList { int n; double* p[MAX_RHS]; };
__global__ void k_halo_pack(int H, PlaneList pl, const int* __restrict__ haloCell, double* __restrict__ buf) {
    pdl_sync();
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    const int c = haloCell[h];
    for (int p = 0; p < pl.n; ++p) buf[(size_t)h * pl.n + p] = pl.p[p][c];
}
__global__ void k_halo_unpack(int H, int N, PlaneList pl, const double* __restrict__ buf) {
    pdl_sync();
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    for (int p = 0; p < pl.n; ++p) pl.p[p][N + h] = buf[(size_t)h * pl.n + p];
}

// ---------------------------------------------------------------- boundary values
// theta.correctBoundaryConditions(): zeroGradient patch value = internal value
__global__ void k_bc_zero_gradient(MeshView m, const int* __restrict__ bc, const double* __restrict__ fld, double* __restrict__ fldB, int nc) {
    pdl_sync();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= m.nB) return;
    if (bc[b] != RHEO_BC_ZERO_GRADIENT || m.bkind[b] == RHEO_PATCH_EMPTY) return;
    const int c = m.bcell[b];
    for (int k = 0; k < nc; ++k) fldB[(size_t)k * m.nB + b] = fld[(size_t)k * m.NP + c];
}

// ---------------------------------------------------------------- Gauss-linear gradient of NC fields
// face value: internal w(P-N)+N in the face's owner/neighbour frame; processor wP+(1-w)N; boundary = patch value
// KT = compile-time slot count (4: 2-D hex, 6: 3-D hex) so that all index / geometry loads of a cell are issued
// together and then all field gathers (two dependent memory latencies per cell instead of 2*K);
// KT = 0: run-time K (unstructured meshes).  Slot order and arithmetic are identical in both paths.
template <int NC, int KT>
__device__ __forceinline__ void gauss_grad_cell(const MeshView& m, int c, const double* __restrict__ fld, const double* __restrict__ fldB,
                                                const double* own, double* g /*[3*NC]: g[3k+d]*/) {
#pragma unroll
    for (int i = 0; i < 3 * NC; ++i) g[i] = 0.0;
    if constexpr (KT == 0) {
        for (int s = 0; s < m.K; ++s) {
            const int nb = m.nbr[(size_t)s * m.NS + c];
            if (nb == -1) continue;
            const int fi = m.fidx[(size_t)s * m.NS + c];
            const int f = fi >= 0 ? fi : ~fi;
            const double sg = fi >= 0 ? 1.0 : -1.0;
            const double Sx = sg * m.Sf[f], Sy = sg * m.Sf[(size_t)m.nF + f], Sz = sg * m.Sf[2 * (size_t)m.nF + f];
            if (nb >= 0) {
                const double w = m.w[f];
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    const double vn = fld[(size_t)k * m.NP + nb];
                    double vf;
                    if (nb >= m.N) vf = w * own[k] + (1.0 - w) * vn;
                    else vf = fi >= 0 ? w * (own[k] - vn) + vn : w * (vn - own[k]) + own[k];
                    g[3 * k] += Sx * vf; g[3 * k + 1] += Sy * vf; g[3 * k + 2] += Sz * vf;
                }
            } else {
                const int b = -nb - 2;
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    const double vf = fldB[(size_t)k * m.nB + b];
                    g[3 * k] += Sx * vf; g[3 * k + 1] += Sy * vf; g[3 * k + 2] += Sz * vf;
                }
            }
        }
    } else {
        int nb[KT], fi[KT];
#pragma unroll
        for (int s = 0; s < KT; ++s) { nb[s] = m.nbr[(size_t)s * m.NS + c]; fi[s] = m.fidx[(size_t)s * m.NS + c]; }
        double Sx[KT], Sy[KT], Sz[KT], w[KT];
#pragma unroll
        for (int s = 0; s < KT; ++s) {
            const bool used = nb[s] != -1;
            const int f = used ? (fi[s] >= 0 ? fi[s] : ~fi[s]) : 0;
            const double sg = used ? (fi[s] >= 0 ? 1.0 : -1.0) : 0.0;   // unused slot: contributes S = 0
            Sx[s] = sg * m.Sf[f]; Sy[s] = sg * m.Sf[(size_t)m.nF + f]; Sz[s] = sg * m.Sf[2 * (size_t)m.nF + f];
            w[s] = m.w[f];
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            double v[KT];
#pragma unroll
            for (int s = 0; s < KT; ++s) {   // one load per slot through a selected pointer: cell / ghost / patch value / (unused: own)
                const double* src = nb[s] >= 0 ? fld + (size_t)k * m.NP + nb[s]
                                  : (nb[s] == -1 ? fld + (size_t)k * m.NP + c : fldB + (size_t)k * m.nB + (-nb[s] - 2));
                v[s] = *src;
            }
#pragma unroll
            for (int s = 0; s < KT; ++s) {
                double vf = v[s];
                if (nb[s] >= m.N) vf = w[s] * own[k] + (1.0 - w[s]) * v[s];
                else if (nb[s] >= 0) vf = fi[s] >= 0 ? w[s] * (own[k] - v[s]) + v[s] : w[s] * (v[s] - own[k]) + own[k];
                g[3 * k] += Sx[s] * vf; g[3 * k + 1] += Sy[s] * vf; g[3 * k + 2] += Sz[s] * vf;
            }
        }
    }
    const double rv = m.rV[c];
#pragma unroll
    for (int i = 0; i < 3 * NC; ++i) g[i] *= rv;
}

// ---------------------------------------------------------------- tiles
// assembly (assembly.cuh): 32 consecutive cells x (solved components + velocity components), one warp each;
// Krylov gather kernels (krylov.cuh): 256 consecutive cells, one thread each.  Unsolved components (xz, yz in 2-D: theta
// stays 0 there) are skipped altogether.
struct CompList { int n; int c[6]; };
constexpr int TILE = 32;
constexpr int RT = 256;   // cells per row tile of the matrix

// Matrix rows (neighbour table nbrA and coefficients A) are stored tile-major: the K x 256 words of the 256 consecutive
// cells of a row tile are contiguous, so that one cp.async.bulk brings the tile's rows (krylov.cuh); a warp still reads
// 32 consecutive words per slot.  NS is a multiple of 256.
__host__ __device__ __forceinline__ size_t ell_t(int K, int s, int c) { return ((size_t)(c >> 8) * K + s) * RT + (c & (RT - 1)); }

// ---------------------------------------------------------------- convection: upwind LDU + deferred HRS + boundary folding + relax
struct Limiter { int hrs; double a0, a1, a2, b0, b1, b2, bnd0, bnd1; };

// Deferred-correction face value of gaussDefCmpwConvectionScheme.C:259-274 (swit = 1) for one face and
// component, in the face's owner(P) -> neighbour(N) frame.  upw in {0,1} makes the reference's blend
//   (1-a-b)(vN - 2gPd) upw + (1-a-b)(vP + 2gNd)(1-upw) + ((a-1)upw + b(1-upw)) vP + (b upw + (a-1)(1-upw)) vN
// a selection (x*1 = x, y*0 = 0 exactly), which is how it is evaluated here — no divergent branches.
// The normalised variable phi~_C = 1 - r, r = (vN - vP) / (2 g.d + 1e-18), only SELECTS the (alpha, beta) row, so the FP64
// division of the reference is replaced by sign-corrected comparisons of the numerator with multiples of the denominator:
//   phi~ <= 0  <=>  r >= 1 ;   phi~ >= 1  <=>  r <= 0 ;   phi~ < b  <=>  r > 1 - b .
// The selected row can differ from the reference's only when phi~ is within rounding of a breakpoint, where the limiters are
// continuous (except the reference's superbee row at 0, see tests/test_gpu_parity.py); 0/0 keeps the reference's
// "every comparison false" outcome.
__device__ __forceinline__ double phif_defc(double vP, double vN, double gPd, double gNd, bool upw, const Limiter& lm) {
    const double gd_up = upw ? gPd : gNd;
    const double den = 2.0 * gd_up + 1e-18;
    const double num = vN - vP;
    const bool neg = den < 0.0;
    const double n = neg ? -num : num, d = fabs(den);
    const bool nan = (d == 0.0) && (n == 0.0);
    const bool out = !nan && ((n >= d) || (n <= 0.0));
    const bool s0 = !nan && (n > (1.0 - lm.bnd0) * d), s1 = !nan && (n > (1.0 - lm.bnd1) * d);
    const double alpha = out ? 1.0 : (s0 ? lm.a0 : (s1 ? lm.a1 : lm.a2));
    const double beta = out ? 0.0 : (s0 ? lm.b0 : (s1 ? lm.b1 : lm.b2));
    const double oab = 1.0 - alpha - beta;
    const double far = upw ? (vN - 2.0 * gPd) : (vP + 2.0 * gNd);
    const double cP = upw ? (alpha - 1.0) : beta, cN = upw ? beta : (alpha - 1.0);
    return oab * far + cP * vP + cN * vN;
}

enum { SLOT_CELL = 1, SLOT_OWNER = 2, SLOT_GHOST = 8, SLOT_PATCH = 16, SLOT_PATCH_ZG = 32 };   // static slot kinds (tile records, assembly.cuh)

// ---------------------------------------------------------------- Krylov control block (kernels in krylov.cuh)
struct KrylovCtl {   // one per RHS, device resident
    double rho, rhoOld, alpha, omega, beta, normFactor, initRes, finRes;
    int state;   // 0 active, 1 converged at the half step (needs psi += alpha y), 2 done
    int iters;
    int singular;
    int pad;
};

// ---------------------------------------------------------------- eig + exp + tau
template <int MODEL>
__global__ void __launch_bounds__(BLOCK) k_eig_tau(int N, int NP, ModelParams mp, const double* __restrict__ theta, const double* __restrict__ fFene,
                                                    double* __restrict__ lam, double* __restrict__ R, double* __restrict__ tau,
                                                    const double* __restrict__ lamCell, const double* __restrict__ etaCell) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    if (lamCell) { mp.lambda = lamCell[c]; mp.etaP = etaCell[c]; }   // thermo-dependent parameters (Oldroyd_BLog.C:133-135)
    double th[6], d[3], V[9], l[3], t6[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) th[k] = theta[(size_t)k * NP + c];
    jacobi_eig(th, d, V);
    l[0] = exp(d[0]); l[1] = exp(d[1]); l[2] = exp(d[2]);
    tau_from_eig<MODEL>(mp, V, l, fFene[c], t6);
#pragma unroll
    for (int k = 0; k < 3; ++k) lam[(size_t)k * NP + c] = l[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[(size_t)k * NP + c] = V[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) tau[(size_t)k * NP + c] = t6[k];
}
// stand-alone calcEig on AoS host-layout arrays (unit parity test entry point)
__global__ void k_eig_exp_aos(int n, const double* __restrict__ th6, double* __restrict__ vals9, double* __restrict__ vecs9) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double th[6], d[3], V[9];
    for (int k = 0; k < 6; ++k) th[k] = th6[(size_t)c * 6 + k];
    jacobi_eig(th, d, V);
    double* o = vals9 + (size_t)c * 9;
    for (int k = 0; k < 9; ++k) o[k] = 0.0;
    o[0] = exp(d[0]); o[4] = exp(d[1]); o[8] = exp(d[2]);
    for (int k = 0; k < 9; ++k) vecs9[(size_t)c * 9 + k] = V[k];
}

// ---------------------------------------------------------------- tau wall BC: linearExtrapolation on one patch
// one thread per patch face; gradient of the 6 tau components at the wall-adjacent cell only
__global__ void k_tau_bc_linext(MeshView m, int start, int size, const double* __restrict__ tau, const double* __restrict__ tauB, double* __restrict__ tmp) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    const int b = start + i;
    const int c = m.bcell[b];
    double own[6], g[18];
#pragma unroll
    for (int k = 0; k < 6; ++k) own[k] = tau[(size_t)k * m.NP + c];
    gauss_grad_cell<6, 0>(m, c, tau, tauB, own, g);
    const double dx = m.CfB[b] - m.C[c], dy = m.CfB[(size_t)m.nB + b] - m.C[(size_t)m.NP + c], dz = m.CfB[2 * (size_t)m.nB + b] - m.C[2 * (size_t)m.NP + c];
#pragma unroll
    for (int k = 0; k < 6; ++k) tmp[(size_t)k * size + i] = own[k] + (g[3 * k] * dx + g[3 * k + 1] * dy + g[3 * k + 2] * dz);
}
// linearExtrapolation with `useRegression true` (linearExtrapolationFvPatchField.C:152-219): per patch face the least-squares
// line through (wall distance, value) of the wall cell's internal faces (values linearly interpolated with weights recomputed
// from the face centres, :186-190; faces on coupled patches are skipped, :181-183) and of the cell centre; wall value =
// yav - xav num/den.  Two passes over the cell's slots (averages, then the sums) instead of the reference's lists; reads cell
// values only, so the patch order does not matter.  CfI: [3][nInt] internal face centres in device face order.
__global__ void k_tau_bc_regress(MeshView m, const double* __restrict__ CfI, int start, int size, const double* __restrict__ tau, double* __restrict__ tauB) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    const int b = start + i;
    const int c = m.bcell[b];
    const size_t fp = (size_t)m.nInt + b;
    double n[3] = {m.Sf[fp], m.Sf[(size_t)m.nF + fp], m.Sf[2 * (size_t)m.nF + fp]};
    const double magS = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    double fx[3], Cc[3], own[6];
#pragma unroll
    for (int d = 0; d < 3; ++d) { n[d] /= magS; fx[d] = m.CfB[(size_t)d * m.nB + b]; Cc[d] = m.C[(size_t)d * m.NP + c]; }
#pragma unroll
    for (int q = 0; q < 6; ++q) own[q] = tau[(size_t)q * m.NP + c];
    double xav = 0, yav[6], den = 0, num[6];
    int id = 0;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            xav /= id;
#pragma unroll
            for (int q = 0; q < 6; ++q) { yav[q] /= id; num[q] = 0.0; }
        } else {
#pragma unroll
            for (int q = 0; q < 6; ++q) yav[q] = 0.0;
        }
        for (int s = 0; s <= m.K; ++s) {
            double x, y[6];
            if (s < m.K) {
                const int nb = m.nbr[(size_t)s * m.NS + c];
                if (nb < 0 || nb >= m.N) continue;   // boundary / empty slot, or a processor face
                const int fi = m.fidx[(size_t)s * m.NS + c];
                const size_t f = fi >= 0 ? fi : ~fi;
                const int P = fi >= 0 ? c : nb, N = fi >= 0 ? nb : c;
                double so = 0, sn = 0, xd = 0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const double S = m.Sf[(size_t)d * m.nF + f], cf = CfI[(size_t)d * m.nInt + f];
                    so += S * (cf - m.C[(size_t)d * m.NP + P]);
                    sn += S * (m.C[(size_t)d * m.NP + N] - cf);
                    xd += n[d] * (fx[d] - cf);
                }
                const double SfdOwn = fabs(so), SfdNei = fabs(sn);
                const double w = SfdOwn / (SfdOwn + SfdNei);
                x = fabs(xd);
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const double vn = tau[(size_t)q * m.NP + nb];
                    y[q] = fi >= 0 ? w * vn + (1. - w) * own[q] : w * own[q] + (1. - w) * vn;   // w var[neighbour] + (1 - w) var[owner]
                }
            } else {   // last pair: the cell itself
                double xd = 0;
#pragma unroll
                for (int d = 0; d < 3; ++d) xd += n[d] * (fx[d] - Cc[d]);
                x = fabs(xd);
#pragma unroll
                for (int q = 0; q < 6; ++q) y[q] = own[q];
            }
            if (pass == 0) {
                xav += x; ++id;
#pragma unroll
                for (int q = 0; q < 6; ++q) yav[q] += y[q];
            } else {
                den += (x - xav) * (x - xav);
#pragma unroll
                for (int q = 0; q < 6; ++q) num[q] += (x - xav) * (y[q] - yav[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) tauB[(size_t)q * m.nB + b] = yav[q] - xav * num[q] / den;
}
__global__ void k_tau_bc_commit(int nB, int start, int size, const double* __restrict__ tmp, double* __restrict__ tauB) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    for (int k = 0; k < 6; ++k) tauB[(size_t)k * nB + start + i] = tmp[(size_t)k * size + i];
}

}  // namespace rk
