// assembly.cuh — "upwind cell computes" assembly of the theta transport equation.
//
// Reference loops restated (of90/src/libs/...):
//   gaussDefCmpwConvectionScheme/gaussDefCmpwConvectionScheme.C:93-167  upwind LDU, boundary coefficients, deferred source
//   gaussDefCmpwConvectionScheme.C:242-319                              phifDefC: per component fvc::grad + face loop
//   constitutiveEquations/constitutiveEqs/utils/boilerLog.H:1           L = fvc::grad(U)
//   EXT-OF9 fvMatrix::relax, addBoundaryDiag / addBoundarySource, EulerDdtScheme::fvmDdt
//
// Observation that shapes the kernel: in phifDefC (:257-274) the gradient of the NON-upwind cell is multiplied by
// exactly 0 (upw is 0 or 1), so the deferred face value needs (theta_P, theta_N, grad(theta)_upwind . d) only.  The
// upwind cell of a face therefore has everything in registers right after it has computed its own Gauss gradient
// (which gathers theta of all neighbours anyway): grad(theta) never goes to memory.  Each face value is computed ONCE,
// by its upwind cell, which adds v*F to its own deferred source and hands v to the downwind cell through `corr`, an
// ELL-shaped array indexed by the downwind cell's own (slot, cell) — the downwind cell reads its row of `corr` with
// coalesced loads when the first Krylov residual is formed (k_krylov_init folds  b -= sum_inflow A * corr).  Processor faces: v goes to the send buffer of the halo exchange instead (the
// former halo of the 18 grad(theta) planes becomes one value per face and component).
//
// Compared with the first version (k_grad_theta -> 18 planes -> k_convect gathering theta + 3 gradient planes from all
// neighbours on both sides of every face): no gradient planes (-288 B/cell of HBM traffic), 6 instead of 24 neighbour
// gathers per (cell, component), half the limiter evaluations (ncu r1c: k_convect was L1-throughput bound at 87 %).
#pragma once
#include "kernels.cuh"
#include "tma.cuh"

namespace rk {


struct FluxArgs {
    CompList cl;          // solved components of theta
    int nU;               // 3: also compute grad(U) (first mode of a step), 0: not
    Limiter lim;
    int noConv;           // limiter `none`: no convection term at all
    double rDeltaT, relax;   // rDeltaT: ddt diagonal coefficient per unit volume (Euler 1/dt; backward coefft/dt)
    int writeMatrix;      // first mode: write A, diag, rD
    int bounded;          // `bounded GaussDefCmpw`: - fvm::Sp(div(phi)) (EXT-OF9 boundedConvectionScheme): diag -= sum of the cell's outward fluxes
    const double* Fell;   // [nTiles][K][32] signed outflow flux per slot (patch slots: phi_b)
    const double* theta; const double* thetaB;
    const double* U; const double* Ub;
    double* bsrc;         // [6*NP] own part of the source (ddt and model terms are added by k_cell_source)
    double* diag; double* rD; double* Fs;   // Fs = A, tile-major (ell_t)
    double* FsT;          // A^T in the same layout: A[nb][c] = min(-F, 0) (PBiCG only, pbicg.cuh); null otherwise
    double* corr;         // [nComp][K*NS] face values handed to the downwind cell
    double* ghostCorr;    // send buffer: [(h * ghostStride) + ghostOffset + comp]
    int ghostStride, ghostOffset;
    double* gradU;        // [9*NP] g[3k+d] = d_d U_k
    // k_flux3 only (assembly3.cuh): what the fused source + first-residual kernel needs instead of gathers
    double* acc;          // [nComp][NP] A theta (diag theta_P + sum_s A_s theta_N), this mode
    double* rowsum;       // [N] diag + sum_s A_s           (written with the matrix)
    unsigned* inflow;     // [N] bit s: A_s < 0             (written with the matrix)
    double* sumPartials; double* sumOut; unsigned* counter;   // sum(theta) per solved component -> sumOut[nComp]
    const int* tileOrder; // [nTiles] tiles in geometric (super-block) order, or null: the tiles a wave of CTAs works on share their neighbours in L2
};

// Tile record of the mesh (static, built once per mesh; one contiguous block per 32 consecutive cells so that ONE bulk copy
// brings everything the tile needs): for K slots and 32 lanes
//   int    nbr [K][32]    neighbour (>=0 cell / ghost, -1 unused, <=-2 patch face -(b+2))
//   int    meta[K][32]    SLOT_CELL | SLOT_OWNER | SLOT_GHOST | SLOT_PATCH | SLOT_PATCH_ZG | reverse slot << 8
//   double S[3][K][32]    face area vector pointing out of the cell        double W[K][32]  linear weight of the face
//   double D[3][K][32]    C_N - C_P in the face's owner -> neighbour frame  double rV[32], V[32]
__host__ __device__ constexpr size_t tile_record_bytes(int K) { return (size_t)K * TILE * (2 * sizeof(int) + 7 * sizeof(double)) + 2 * TILE * sizeof(double); }
__host__ __device__ constexpr size_t tile_flux_bytes(int K) { return (size_t)K * TILE * sizeof(double); }

// Persistent tile kernel: CTA = (nComp theta warps + nU velocity warps) x 32 lanes, looping over tiles of 32 consecutive cells.
//   producer  one thread streams the NEXT tile's mesh record and its K x 32 face fluxes into the other shared-memory stage
//             with two cp.async.bulk copies signalled on an mbarrier (the whole connectivity / geometry stream of the
//             assembly goes through the TMA unit: no registers, no L1, perfectly sequential HBM reads);
//   consumers theta warp g: Gauss gradient of component comps[g] in registers (6 neighbour gathers), deferred face values
//             of the faces this cell is the upwind cell of, diagonal / relax / boundary source;
//             velocity warp u: Gauss gradient of U_u;  warp g also writes A = min(F,0) of slot g.
// KT = compile-time slot count (neighbour values stay in registers between the gradient and the face pass); KT = 0:
// run-time K, the face pass re-reads the neighbour value (L1 hit).
#ifndef RK_FLUX_MINB
#define RK_FLUX_MINB 4   // resident CTAs per SM k_flux_assemble is compiled for (register cap 65536 / (288 * MINB))
#endif
template <int KT>
__global__ void __launch_bounds__(TILE * 9, RK_FLUX_MINB) k_flux_assemble(MeshView m, FluxArgs a, const unsigned char* __restrict__ tileRec, int nTiles) {
    pdl_sync();
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const int K = KT > 0 ? KT : m.K;
    const int KTL = K * TILE;
    const size_t recBytes = tile_record_bytes(K), fluxBytes = tile_flux_bytes(K), stageBytes = recBytes + fluxBytes;
    __shared__ uint64_t full[2];
    const int lane = threadIdx.x & (TILE - 1), grp = threadIdx.x / TILE, nGrp = blockDim.x / TILE;
    const bool hrs = a.lim.hrs && !a.noConv;
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && (int)blockIdx.x < nTiles) {
        mbar_expect_tx(&full[0], (uint32_t)stageBytes);
        bulk_g2s(smemRaw, tileRec + (size_t)blockIdx.x * recBytes, (uint32_t)recBytes, &full[0]);
        bulk_g2s(smemRaw + recBytes, (const unsigned char*)a.Fell + (size_t)blockIdx.x * fluxBytes, (uint32_t)fluxBytes, &full[0]);
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++it) {
        const int st = it & 1;
        const int next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < nTiles) {   // stage st^1 was released by the __syncthreads that ended the previous iteration
            unsigned char* dst = smemRaw + (size_t)(st ^ 1) * stageBytes;
            mbar_expect_tx(&full[st ^ 1], (uint32_t)stageBytes);
            bulk_g2s(dst, tileRec + (size_t)next * recBytes, (uint32_t)recBytes, &full[st ^ 1]);
            bulk_g2s(dst + recBytes, (const unsigned char*)a.Fell + (size_t)next * fluxBytes, (uint32_t)fluxBytes, &full[st ^ 1]);
        }
        const unsigned char* base = smemRaw + (size_t)st * stageBytes;
        const int* pNb = (const int*)base + lane;
        const int* pMeta = pNb + KTL;
        const double* pS = (const double*)(base + (size_t)2 * KTL * sizeof(int)) + lane;   // [3][K][TILE]
        const double* pW = pS + 3 * KTL;
        const double* pD = pW + KTL;                                                        // [3][K][TILE]
        const double* pRV = pD + 3 * KTL;                                                   // rV[32], V[32]
        const double* pF = (const double*)(base + recBytes) + lane;
        const int c = tile * TILE + lane;
        const bool active = c < m.N;
        mbar_wait(&full[st], (uint32_t)((it >> 1) & 1));

        if (active) {
            if (a.writeMatrix)   // row coefficients A[c][nb] = min(F,0), slot-major for the Krylov kernels
                for (int s = grp; s < K; s += nGrp)
                    a.Fs[ell_t(m.K, s, c)] = ((pMeta[s * TILE] & SLOT_CELL) && !a.noConv) ? fmin(pF[s * TILE], 0.0) : 0.0;
            if (a.writeMatrix && a.FsT)
                for (int s = grp; s < K; s += nGrp)
                    a.FsT[ell_t(m.K, s, c)] = ((pMeta[s * TILE] & SLOT_CELL) && !a.noConv) ? fmin(-pF[s * TILE], 0.0) : 0.0;
            if (grp >= a.cl.n) {   // ---------------- velocity warp: grad(U_u)
                const int u = grp - a.cl.n;
                const double* fk = a.U + (size_t)u * m.NP;
                const double own = fk[c];
                const double* uB = a.Ub + (size_t)u * m.nB;
                double gx = 0, gy = 0, gz = 0;
                double un[KT > 0 ? KT : 1];
                if constexpr (KT > 0) {
#pragma unroll
                    for (int s = 0; s < KT; ++s) {   // all neighbour gathers in flight before the first use
                        const int nb = pNb[s * TILE];
                        const double* src = nb >= 0 ? fk + nb : (nb == -1 ? fk + c : uB + (-nb - 2));
                        un[s] = *src;
                    }
                }
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    const int nb = pNb[s * TILE];
                    if (nb == -1) continue;
                    double v;
                    if constexpr (KT > 0) v = un[s];
                    else v = nb >= 0 ? fk[nb] : uB[-nb - 2];
                    double vf = v;
                    if (nb >= 0) {
                        const double w = pW[s * TILE];
                        if (nb >= m.N) vf = w * own + (1.0 - w) * v;
                        else vf = nb > c ? w * (own - v) + v : w * (v - own) + own;
                    }
                    gx += pS[s * TILE] * vf; gy += pS[KTL + s * TILE] * vf; gz += pS[2 * KTL + s * TILE] * vf;
                }
                const double rv = pRV[0];
                a.gradU[(size_t)(3 * u) * m.NP + c] = gx * rv;
                a.gradU[(size_t)(3 * u + 1) * m.NP + c] = gy * rv;
                a.gradU[(size_t)(3 * u + 2) * m.NP + c] = gz * rv;
            } else {               // ---------------- theta warp
                const int k = a.cl.c[grp];
                const double* tk = a.theta + (size_t)k * m.NP;
                const double* tB = a.thetaB + (size_t)k * m.nB;
                const double tP = tk[c];
                double* corrRow = a.corr + (size_t)grp * m.K * m.NS;
                // neighbour values first: K independent gathers in flight before anything else is computed
                double vn[KT > 0 ? KT : 1];
                if constexpr (KT > 0) {
                    if (hrs) {
#pragma unroll
                        for (int s = 0; s < KT; ++s) {   // one load per slot through a selected pointer: cell / ghost / patch value / (unused: own)
                            const int nb = pNb[s * TILE];
                            const double* src = nb >= 0 ? tk + nb : (nb == -1 ? tk + c : tB + (-nb - 2));
                            vn[s] = *src;
                        }
                    }
                }
                // diagonal: only needed by the warp that writes it, or by every component when relax() is active
                double D = 0, sumOff = 0, iCcoupled = 0, iCplainAbs = 0, iCplain = 0;
                if (a.relax > 0 || (a.writeMatrix && grp == 0)) {
                    D = a.rDeltaT * pRV[TILE];   // ddt diag + negSumDiag
                    for (int s = 0; s < K; ++s) {
                        const int meta = pMeta[s * TILE];
                        const double F = a.noConv ? 0.0 : pF[s * TILE];
                        if (a.bounded) D -= F;   // unused slots and faces of empty patches carry F = 0
                        if (meta & SLOT_CELL) {
                            if (!(meta & SLOT_GHOST)) { D += fmax(F, 0.0); sumOff += fmax(-F, 0.0); }
                            else { iCcoupled += (F >= 0 ? F : 0.0); sumOff += fmax(-F, 0.0); }
                        } else if (meta & SLOT_PATCH_ZG) { iCplain += F; iCplainAbs += fabs(F); }
                    }
                }
                double sou = 0, bnd = 0;
                if (hrs) {
                    // Gauss-linear gradient of theta_k (gaussDefCmpwConvectionScheme.C:254), kept in registers
                    double gx = 0, gy = 0, gz = 0;
#pragma unroll
                    for (int s = 0; s < K; ++s) {
                        const int nb = pNb[s * TILE];
                        if (nb == -1) continue;
                        double v;
                        if constexpr (KT > 0) v = vn[s];
                        else v = nb >= 0 ? tk[nb] : tB[-nb - 2];
                        double vf = v;
                        if (nb >= 0) {
                            const double w = pW[s * TILE];
                            if (nb >= m.N) vf = w * tP + (1.0 - w) * v;
                            else vf = nb > c ? w * (tP - v) + v : w * (v - tP) + tP;
                        }
                        gx += pS[s * TILE] * vf; gy += pS[KTL + s * TILE] * vf; gz += pS[2 * KTL + s * TILE] * vf;
                    }
                    const double rv = pRV[0];
                    gx *= rv; gy *= rv; gz *= rv;
                    // faces this cell is the upwind cell of: deferred face value, own source, hand-over to the downwind cell
                    const Limiter L = a.lim;
#pragma unroll
                    for (int s = 0; s < K; ++s) {
                        const int meta = pMeta[s * TILE];
                        if (meta & SLOT_CELL) {
                            const int nb = pNb[s * TILE];
                            const double F = pF[s * TILE];
                            const bool own = meta & SLOT_OWNER;
                            const bool upwFace = own ? (F >= 0) : (-F >= 0);   // pos(phi), phi = own ? F : -F
                            double v = 0.0;
                            if (own == upwFace) {   // this cell is the upwind cell of the face
                                double tn;
                                if constexpr (KT > 0) tn = vn[s];
                                else tn = tk[nb];
                                const double gd = gx * pD[s * TILE] + gy * pD[KTL + s * TILE] + gz * pD[2 * KTL + s * TILE];
                                v = phif_defc(own ? tP : tn, own ? tn : tP, gd, gd, upwFace, L);
                                sou += v * F;   // souT[own] += v*phi ; souT[nei] -= v*phi
                                if (!(meta & SLOT_GHOST)) corrRow[(unsigned)(meta >> 8) * (unsigned)m.NS + (unsigned)nb] = v;   // K*NS < 2^31 (checked at create)
                            }
                            if (meta & SLOT_GHOST) {
                                // processor face: the value travels with the halo exchange (0 where the other side is upwind, so
                                // that the receiver never reads an unwritten word); my own ghost slot of `corr` is cleared
                                a.ghostCorr[(size_t)(nb - m.N) * a.ghostStride + a.ghostOffset + grp] = v;
                                corrRow[(unsigned)s * (unsigned)m.NS + (unsigned)c] = 0.0;
                            }
                        } else if ((meta & (SLOT_PATCH | SLOT_PATCH_ZG)) == SLOT_PATCH) {
                            bnd += -pF[s * TILE] * tB[-pNb[s * TILE] - 2];
                        }
                    }
                } else {
                    for (int s = 0; s < K; ++s) {
                        const int meta = pMeta[s * TILE];
                        if ((meta & (SLOT_PATCH | SLOT_PATCH_ZG)) == SLOT_PATCH) bnd += -(a.noConv ? 0.0 : pF[s * TILE]) * tB[-pNb[s * TILE] - 2];
                    }
                }
                double add = 0;
                if (a.relax > 0) {   // EXT-OF9 fvMatrix::relax
                    const double D0 = D;
                    double Dn = D + iCcoupled + iCplainAbs;
                    Dn = fmax(fabs(Dn), sumOff);
                    Dn /= a.relax;
                    Dn -= iCcoupled;
                    Dn -= iCplain;
                    add = (Dn - D0) * tP;
                    D = Dn;
                }
                a.bsrc[(size_t)k * m.NP + c] = (-sou + add) + bnd;
                if (a.writeMatrix && grp == 0) {
                    const double Dfull = D + iCcoupled + iCplain;   // addBoundaryDiag
                    a.diag[c] = Dfull;
                    a.rD[c] = 1.0 / Dfull;   // DILU: upper*lower == 0 on every face of an upwind matrix
                }
            }
        }
        __syncthreads();   // every warp is done with stage st before the next iteration's prefetch overwrites it
    }
}

// ---------------------------------------------------------------- per-cell source: Omega/B split, model term, Euler ddt
// One thread per cell, pure streaming (no gathers): L comes from gradU (k_flux_assemble); the result is added to the
// own-face part k_flux_assemble left in bsrc.  The faces the cell is DOWNWIND of are folded in by k_krylov_init, which
// reads the cell's matrix row anyway: b -= sum_{inflow slots} A * corr  (krylov.cuh).
struct SourceArgs {
    ModelParams mp;
    double rDeltaT;
    int solvedIdx[6];       // component -> index among the solved components, -1 if not solved (2-D: xz, yz)
    const double* gradU; const double* theta; const double* thetaOld; const double* lam; const double* R;
    double* bsrc; double* fFene;
    const double* tau;      // the model's current tau planes (read by SaramitoLog only)
    const double* lamCell; const double* etaCell;   // thermo-dependent lambda / etaP per cell (null: the scalars of mp)
    // EXT-OF9 backwardDdtScheme: source = (1/dt) V (c0 theta_old - c00 theta_oldold); Euler: (1/dt) theta_old V
    int backward; double c0, c00; const double* thetaOldOld;
    // sum of theta over the cells per solved component (gAverage(psi) of the solver's normFactor): theta is in registers
    // here anyway, so the reduction rides along instead of being a kernel of its own
    double* sumPartials; double* sumOut; unsigned* counter;
};

constexpr int SRC_BLOCK = 128;
template <int MODEL>
__global__ void __launch_bounds__(SRC_BLOCK, 3) k_cell_source2(MeshView m, SourceArgs a) {
    pdl_sync();
    double sum[6] = {0, 0, 0, 0, 0, 0};
    // persistent grid (one resident wave): the reduction epilogue is paid once per CTA, not once per 128 cells
    for (int c = blockIdx.x * SRC_BLOCK + threadIdx.x; c < m.N; c += gridDim.x * SRC_BLOCK) {
        double g[9], thO[6], own6[6], th[6], Rm[9], lm[3], rhs[6];
#pragma unroll
        for (int i = 0; i < 9; ++i) g[i] = a.gradU[(size_t)i * m.NP + c];
#pragma unroll
        for (int k = 0; k < 6; ++k) th[k] = a.theta[(size_t)k * m.NP + c];
#pragma unroll
        for (int k = 0; k < 9; ++k) Rm[k] = a.R[(size_t)k * m.NP + c];
#pragma unroll
        for (int k = 0; k < 3; ++k) lm[k] = a.lam[(size_t)k * m.NP + c];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            thO[k] = a.thetaOld[(size_t)k * m.NP + c];
            own6[k] = a.solvedIdx[k] >= 0 ? a.bsrc[(size_t)k * m.NP + c] : 0.0;
        }
        const double V = m.V[c];
        // g[3k+d] = d_d U_k  ->  L_ij = d_i U_j = g[3j+i]
        const double L[9] = {g[0], g[3], g[6], g[1], g[4], g[7], g[2], g[5], g[8]};
        double f;
        ModelParams mp = a.mp;
        if (a.lamCell) { mp.lambda = a.lamCell[c]; mp.etaP = a.etaCell[c]; }   // Oldroyd_BLog.C:133-135: createField(lambda_), createField(etaP_)
        if constexpr (MODEL == RHEO_MODEL_SARAMITO_LOG || MODEL == RHEO_MODEL_BMP_FLUIDITY) {
            double tc[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) tc[k] = a.tau[(size_t)k * m.NP + c];
            f = model_rhs<MODEL>(mp, L, th, Rm, lm, rhs, tc);
        } else {
            f = model_rhs<MODEL>(mp, L, th, Rm, lm, rhs);
        }
        a.fFene[c] = f;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double ddtSrc;
            if (a.backward) ddtSrc = a.rDeltaT * V * (a.c0 * thO[k] - a.c00 * a.thetaOldOld[(size_t)k * m.NP + c]);
            else ddtSrc = a.rDeltaT * thO[k] * V;
            a.bsrc[(size_t)k * m.NP + c] = (ddtSrc + V * rhs[k]) + own6[k];
            sum[k] += th[k];
        }
    }
    // ---- sum(theta) per component: warp shuffle, block partials, last block sums the partials in block order
    __shared__ double sm[SRC_BLOCK / 32][6];
    __shared__ bool isLast;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double x = sum[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) sm[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double x = 0;
#pragma unroll
        for (int wv = 0; wv < SRC_BLOCK / 32; ++wv) x += sm[wv][threadIdx.x];
        a.sumPartials[(size_t)blockIdx.x * 6 + threadIdx.x] = x;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) isLast = atomicAdd(a.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    for (int k = warp; k < 6; k += SRC_BLOCK / 32) {
        double x = 0;
        for (unsigned b = lane; b < gridDim.x; b += 32) x += a.sumPartials[(size_t)b * 6 + k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0 && a.solvedIdx[k] >= 0) a.sumOut[a.solvedIdx[k]] = x;
    }
    if (threadIdx.x == 0) *a.counter = 0;
}

// theta.correctBoundaryConditions() (theta != nullptr: every zeroGradient face) and the zeroGradient faces of tau in the
// boundary-face range [t0, t1) in one launch: zeroGradient patch value = internal value.  tau's patches are evaluated in
// patch order (EXT-OF9 GeometricBoundaryField::evaluate), so zeroGradient patches that FOLLOW a linearExtrapolation patch
// are updated after it — the caller passes the ranges accordingly.
__global__ void k_bc_zero_gradient2(MeshView m, const double* __restrict__ theta, double* __restrict__ thetaB, const double* __restrict__ tau,
                                    double* __restrict__ tauB, int t0, int t1) {
    pdl_sync();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= m.nB) return;
    if (m.bkind[b] == RHEO_PATCH_EMPTY) return;
    const int c = m.bcell[b];
    if (theta != nullptr && m.bthetaBC[b] == RHEO_BC_ZERO_GRADIENT)
        for (int k = 0; k < 6; ++k) thetaB[(size_t)k * m.nB + b] = theta[(size_t)k * m.NP + c];
    if (b >= t0 && b < t1 && m.btauBC[b] == RHEO_BC_ZERO_GRADIENT)
        for (int k = 0; k < 6; ++k) tauB[(size_t)k * m.nB + b] = tau[(size_t)k * m.NP + c];
}

// BMPLog.C:177-196 with the fluidity of AFTER PhiEqn.solve(): theta relaxes at Phi G0 and tau = G0 (c - I), i.e. the Oldroyd-BLog
// source and theta -> tau map with lambda = 1 / (Phi G0), etaP = 1 / Phi per cell (the arrays rheo_gpu_upload_thermo fills otherwise)
__global__ void k_bmp_rates(int N, const double* __restrict__ Phi, double G0, double* __restrict__ lamCell, double* __restrict__ etaCell) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const double f = Phi[c];
    lamCell[c] = 1.0 / (f * G0);
    etaCell[c] = 1.0 / f;
}

// Optional (rheo_gpu_set_tau_assignment, off by default): what `tau_ = ...` leaves on the non-fixed tau
// patches before correctBoundaryConditions() under the alternative reading — e on the diagonal, 0 off it
__global__ void k_tau_b_assign(MeshView m, double e, double* __restrict__ tauB) {
    pdl_sync();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= m.nB) return;
    if (m.bkind[b] == RHEO_PATCH_EMPTY || m.bkind[b] == RHEO_PATCH_PROCESSOR || m.btauBC[b] != RHEO_BC_ZERO_GRADIENT) return;
    tauB[b] = e; tauB[(size_t)m.nB + b] = 0.0; tauB[(size_t)2 * m.nB + b] = 0.0;
    tauB[(size_t)3 * m.nB + b] = e; tauB[(size_t)4 * m.nB + b] = 0.0; tauB[(size_t)5 * m.nB + b] = e;
}

// processor faces: deferred values received from the upwind side, for the cells that own ghost slots
//   b[c] -= F * v  for inflow ghost slots;  recv layout: [(h * stride) + offset + comp]
__global__ void k_ghost_corr(MeshView m, int nBcells, const int* __restrict__ bcells, CompList cl, const double* __restrict__ Fs,
                             const double* __restrict__ recv, int stride, int offset, double* __restrict__ bsrc) {
    pdl_sync();
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 >= nBcells) return;
    const int c = bcells[i0];
    for (int s = 0; s < m.K; ++s) {
        const int nb = m.nbrA[ell_t(m.K, s, c)];
        if (nb < m.N) continue;
        const double A = Fs[ell_t(m.K, s, c)];
        if (!(A < 0.0)) continue;
        for (int j = 0; j < cl.n; ++j) bsrc[(size_t)cl.c[j] * m.NP + c] -= A * recv[(size_t)(nb - m.N) * stride + offset + j];
    }
}

// phi in device face order -> signed outflow flux per (tile, slot, lane), the layout k_flux_assemble streams;
// patch slots hold phi_b, unused slots 0
__global__ void k_flux_ell(MeshView m, const double* __restrict__ phi, double* __restrict__ Fell) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m.NS) return;
    const int tile = c / TILE, lane = c % TILE;
    for (int s = 0; s < m.K; ++s) {
        double F = 0.0;
        if (c < m.N) {
            const size_t e = (size_t)s * m.NS + c;
            const int nb = m.nbr[e];
            if (nb != -1) {
                const int fi = m.fidx[e];
                const double ph = phi[fi >= 0 ? fi : ~fi];
                F = (nb >= 0 && fi < 0) ? -ph : ph;
            }
        }
        Fell[((size_t)tile * m.K + s) * TILE + lane] = F;
    }
}

}  // namespace rk
