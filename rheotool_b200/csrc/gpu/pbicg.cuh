// pbicg.cuh — batched multi-RHS PBiCG with colour-parallel DILU (the solver every Log tutorial's fvSolution selects).
//
// EXT-OF9 semantics restated: PBiCG::solve, DILUPreconditioner::precondition / preconditionT,
// lduMatrix::Amul / Tmul.  Per iteration, for the system A and its transpose side by side:
//     wA = M^-1 rA,  wT = M^-T rT;  wArT = wA.rT;  beta = wArT / wArT_old
//     pA = wA + beta pA,  pT = wT + beta pT        (first iteration: pA = wA, pT = wT)
//     wA = A pA,  wT = A^T pT;  wApT = wA.pT  (singular if |wApT| / normFactor <= 1e-300);  alpha = wArT / wApT
//     psi += alpha pA;  rA -= alpha wA;  rT -= alpha wT;  residual = sum|rA| / normFactor
//
// The transposed matrix needs no second mesh structure: with A[c][nb] = min(F, 0) (F = signed outflow of c towards nb) the
// transposed coefficient is A[nb][c] = min(-F, 0), written by k_flux_assemble into a second coefficient array (FsT) in the
// same tile-major slot layout; diag and rD = 1/diag are shared (upper*lower == 0 on every face of an upwind matrix, so
// DILU's reciprocal diagonal is the same for A and A^T).  preconditionT is precondition with lower and upper swapped.
//
// This is the CORRECTNESS-FIRST version: one thread per cell, run-time slot loops, one launch per colour and direction,
// separate reduction kernels with the scalar control in their last-block epilogue (same KrylovCtl block as PBiCGStab).
// It reads each matrix row twice per product instead of streaming it through the row-tile pipeline of krylov.cuh: PBiCGStab
// remains the tuned path.  Several ranks (solve.inl: solve_batch_pbicg): the products skip the ghost slots here; pA and pT are
// halo-swapped before them and k_ghost (krylov.cuh) adds the processor-patch columns of A and of A^T afterwards (EXT-OF9
// lduMatrix::Tmul uses interfaceIntCoeffs, which are the A^T slots); DILU / DILU^T stay rank-local; dots are all-reduced.
// STATUS: run on B200 in round 2 (tests/test_gpu_pbicg.py): parity with the oracle's PBiCG, with PBiCGStab on the device and
// with the reference fixtures; same iteration counts as the oracle on the renumbered mesh.
#pragma once
#include "krylov.cuh"

namespace rk {

// one colour of the forward (FWD = 1) or backward (FWD = 0) substitution of BOTH systems, cells [c0, c1)
template <int NR, int FWD>
__global__ void __launch_bounds__(BLOCK) k_pb_sweep(MeshView m, int c0, int c1, int nModes, const KrylovShared* __restrict__ ks,
                                                     const double* __restrict__ rD, const double* __restrict__ A, const double* __restrict__ AT,
                                                     const double* __restrict__ rA, const double* __restrict__ rT, double* wA, double* wT) {
    pdl_sync();
    if (ks->nActive == 0) return;
    const int stride = gridDim.x * BLOCK;
    for (int c = c0 + blockIdx.x * BLOCK + threadIdx.x; c < c1; c += stride) {
        const double rd = rD[c];
        for (int md = 0; md < nModes; ++md) {
            const size_t base = (size_t)md * m.NP;
            double accA[NR], accT[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j) { accA[j] = 0.0; accT[j] = 0.0; }
            for (int s = 0; s < m.K; ++s) {
                const size_t e = ell_t(m.K, s, c);
                const int nb = m.nbrA[e];
                const bool sel = FWD ? (nb < c) : (nb > c && nb < m.N);
                if (!sel) continue;
                const double a = A[e], at = AT[e];
                double ya[NR], yt[NR];
                ldv<NR>(wA, base + nb, ya);
                ldv<NR>(wT, base + nb, yt);
#pragma unroll
                for (int j = 0; j < NR; ++j) { accA[j] += a * ya[j]; accT[j] += at * yt[j]; }
            }
            double oa[NR], ot[NR];
            if (FWD) {   // w[c] = rD (r[c] - sum_{nb<c} coeff w[nb])
                ldv<NR>(rA, base + c, oa);
                ldv<NR>(rT, base + c, ot);
#pragma unroll
                for (int j = 0; j < NR; ++j) { oa[j] = rd * (oa[j] - accA[j]); ot[j] = rd * (ot[j] - accT[j]); }
            } else {     // w[c] -= rD sum_{nb>c} coeff w[nb]
                ldv<NR>(wA, base + c, oa);
                ldv<NR>(wT, base + c, ot);
#pragma unroll
                for (int j = 0; j < NR; ++j) { oa[j] -= rd * accA[j]; ot[j] -= rd * accT[j]; }
            }
            stv<NR>(wA, base + c, oa);
            stv<NR>(wT, base + c, ot);
        }
    }
}

// Initial transpose residual.  EXT-OF9 PBiCG::solve starts from rT = source - A^T psi (Tmul), not from rT = rA: with
// rA = b - A psi already formed by k_krylov_init (b includes the deferred inflow terms folded there),
//     rT = rA + (A - A^T) psi = rA + sum_{local slots} (A[s] - AT[s]) psi[nb]     (the diagonal is shared)
template <int NR>
__global__ void __launch_bounds__(BLOCK) k_pb_init_rT(MeshView m, int nModes, RhsPtrs rp, const double* __restrict__ A, const double* __restrict__ AT,
                                                       const double* __restrict__ rA, double* __restrict__ rT) {
    pdl_sync();
    const int stride = gridDim.x * BLOCK;
    for (int md = 0; md < nModes; ++md)
        for (int c = blockIdx.x * BLOCK + threadIdx.x; c < m.N; c += stride) {
            double acc[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j) acc[j] = 0.0;
            for (int s = 0; s < m.K; ++s) {
                const size_t e = ell_t(m.K, s, c);
                const int nb = m.nbrA[e];
                if (nb == c) continue;   // ghost columns included: psi holds the neighbour ranks' values (halo swap at the start of the step)
                const double d = A[e] - AT[e];
#pragma unroll
                for (int j = 0; j < NR; ++j) acc[j] += d * rp.psi[md * NR + j][nb];
            }
            double r[NR];
            ldv<NR>(rA, (size_t)md * m.NP + c, r);
#pragma unroll
            for (int j = 0; j < NR; ++j) r[j] += acc[j];
            stv<NR>(rT, (size_t)md * m.NP + c, r);
        }
}

// wArT = wA . rT per RHS; epilogue: beta
template <int NR>
__global__ void __launch_bounds__(BLOCK) k_pb_dot(int N, int NP, int nModes, KrylovShared* ks, const double* __restrict__ wA, const double* __restrict__ rT,
                                                   double* partials, double* out, unsigned* counter, int ctlWhat, SolveCtl sc) {
    pdl_sync();
    if (ks->nActive == 0) return;
    const int stride = gridDim.x * BLOCK;
    for (int md = 0; md < nModes; ++md) {
        double red[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) red[j] = 0.0;
        for (int c = blockIdx.x * BLOCK + threadIdx.x; c < N; c += stride) {
            double a[NR], b[NR];
            ldv<NR>(wA, (size_t)md * NP + c, a);
            ldv<NR>(rT, (size_t)md * NP + c, b);
#pragma unroll
            for (int j = 0; j < NR; ++j) red[j] += a[j] * b[j];
        }
        block_reduce_to_partials<NR>(red, partials, md * NR, nModes * NR);
    }
    finalize_ctl(partials, gridDim.x, nModes * NR, out, counter, gridDim.x, ctlWhat, ks, nModes * NR, sc);   // CTL_PB_BETA, or CTL_NONE before an all-reduce
}

// pA = wA + beta pA, pT = wT + beta pT (first iteration of a right-hand side: plain copies)
template <int NR>
__global__ void __launch_bounds__(BLOCK) k_pb_update_p(int N, int NP, int nModes, const KrylovShared* __restrict__ ks, const double* __restrict__ wA,
                                                        const double* __restrict__ wT, double* __restrict__ pA, double* __restrict__ pT) {
    pdl_sync();
    if (ks->nActive == 0) return;
    const int stride = gridDim.x * BLOCK;
    for (int md = 0; md < nModes; ++md) {
        double beta[NR];
        int st[NR], first[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            const KrylovCtl& k = ks->ctl[md * NR + j];
            beta[j] = k.beta; st[j] = k.state; first[j] = k.iters == 0;
        }
        for (int c = blockIdx.x * BLOCK + threadIdx.x; c < N; c += stride) {
            const size_t i = (size_t)md * NP + c;
            double a[NR], t[NR], pa[NR], pt[NR];
            ldv<NR>(wA, i, a); ldv<NR>(wT, i, t); ldv<NR>(pA, i, pa); ldv<NR>(pT, i, pt);
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                if (st[j] != 0) continue;
                pa[j] = first[j] ? a[j] : a[j] + beta[j] * pa[j];
                pt[j] = first[j] ? t[j] : t[j] + beta[j] * pt[j];
            }
            stv<NR>(pA, i, pa); stv<NR>(pT, i, pt);
        }
    }
}

// wA = A pA, wT = A^T pT (lduMatrix::Amul / Tmul), wApT = wA . pT per RHS; epilogue: singularity check, alpha
template <int NR>
__global__ void __launch_bounds__(BLOCK) k_pb_spmv(MeshView m, int nModes, KrylovShared* ks, const double* __restrict__ diag, const double* __restrict__ A,
                                                    const double* __restrict__ AT, const double* __restrict__ pA, const double* __restrict__ pT,
                                                    double* __restrict__ wA, double* __restrict__ wT, double* partials, double* out, unsigned* counter, int ctlWhat, SolveCtl sc) {
    pdl_sync();
    if (ks->nActive == 0) return;
    const int stride = gridDim.x * BLOCK;
    for (int md = 0; md < nModes; ++md) {
        const size_t base = (size_t)md * m.NP;
        double red[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) red[j] = 0.0;
        for (int c = blockIdx.x * BLOCK + threadIdx.x; c < m.N; c += stride) {
            const double d = diag[c];
            double accA[NR], accT[NR], pc[NR];
            ldv<NR>(pA, base + c, accA);
            ldv<NR>(pT, base + c, pc);
#pragma unroll
            for (int j = 0; j < NR; ++j) { accA[j] *= d; accT[j] = d * pc[j]; }
            for (int s = 0; s < m.K; ++s) {
                const size_t e = ell_t(m.K, s, c);
                const int nb = m.nbrA[e];
                if (nb == c || nb >= m.N) continue;
                const double a = A[e], at = AT[e];
                double ya[NR], yt[NR];
                ldv<NR>(pA, base + nb, ya);
                ldv<NR>(pT, base + nb, yt);
#pragma unroll
                for (int j = 0; j < NR; ++j) { accA[j] += a * ya[j]; accT[j] += at * yt[j]; }
            }
            stv<NR>(wA, base + c, accA);
            stv<NR>(wT, base + c, accT);
#pragma unroll
            for (int j = 0; j < NR; ++j) red[j] += accA[j] * pc[j];
        }
        block_reduce_to_partials<NR>(red, partials, md * NR, nModes * NR);
    }
    finalize_ctl(partials, gridDim.x, nModes * NR, out, counter, gridDim.x, ctlWhat, ks, nModes * NR, sc);   // CTL_PB_ALPHA, or CTL_NONE (ghost columns + all-reduce follow)
}

// psi += alpha pA; rA -= alpha wA; rT -= alpha wT; sum|rA| per RHS; epilogue: residual, iteration count, convergence
template <int NR>
__global__ void __launch_bounds__(BLOCK) k_pb_update_x_r(int N, int NP, int nModes, RhsPtrs rp, KrylovShared* ks, const double* __restrict__ pA,
                                                          const double* __restrict__ wA, const double* __restrict__ wT, double* __restrict__ rA,
                                                          double* __restrict__ rT, double* partials, double* out, unsigned* counter, int ctlWhat, SolveCtl sc) {
    pdl_sync();
    if (ks->nActive == 0) return;
    const int stride = gridDim.x * BLOCK;
    for (int md = 0; md < nModes; ++md) {
        double red[NR], alpha[NR];
        int st[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            red[j] = 0.0;
            const KrylovCtl& k = ks->ctl[md * NR + j];
            st[j] = k.state; alpha[j] = k.alpha;
        }
        for (int c = blockIdx.x * BLOCK + threadIdx.x; c < N; c += stride) {
            const size_t i = (size_t)md * NP + c;
            double p[NR], a[NR], t[NR], ra[NR], rt[NR];
            ldv<NR>(pA, i, p); ldv<NR>(wA, i, a); ldv<NR>(wT, i, t); ldv<NR>(rA, i, ra); ldv<NR>(rT, i, rt);
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                if (st[j] != 0) continue;
                rp.psi[md * NR + j][c] += alpha[j] * p[j];
                ra[j] -= alpha[j] * a[j];
                rt[j] -= alpha[j] * t[j];
                red[j] += fabs(ra[j]);
            }
            stv<NR>(rA, i, ra); stv<NR>(rT, i, rt);
        }
        block_reduce_to_partials<NR>(red, partials, md * NR, nModes * NR);
    }
    finalize_ctl(partials, gridDim.x, nModes * NR, out, counter, gridDim.x, ctlWhat, ks, nModes * NR, sc);   // CTL_PB_END, or CTL_NONE before an all-reduce
}

}  // namespace rk
