// krylov.cuh — batched multi-RHS PBiCGStab with colour-parallel DILU, fused for HBM traffic.
//
// EXT-OF9 semantics restated: PBiCGStab::solve, DILUPreconditioner::precondition, lduMatrix::Amul /
// sumA / normFactor (in-repo restatement of the residual norm:
// of90/src/libs/sparseMatrixSolvers/segregated/sparseSolver.C:152-179).
//
// Layout: the NR right-hand sides of one mode (NR = 4 in 2-D, 6 in 3-D: the valid components of
// theta) are INTERLEAVED per cell, X[(mode*NP + cell)*NR + j], so one 16-byte load fetches two RHS
// and a neighbour gather uses whole 32-byte sectors.  theta itself (psi) stays in SoA planes.
// Matrix: diag[c], rD[c] = 1/diag[c] (DILU of an upwind matrix: upper*lower == 0 on every face),
// A[s][c] = min(signed outflow flux, 0) in slot-major ELL with neighbour table nbrA.
//
// Cells are numbered colour by colour.  One preconditioned product v = A M^-1 p is
//     colour 0           :  p = r + beta (p - omega v), y = rD p         (k_update_p / k_make_s on the colour-0 range)
//     k_sweep<UPD, fwd>  colours 1..nc-1 :  the same vector update of the cell, then y[c] -= rD[c] sum_{nb<c} A y[nb]
//     k_sweep<0, bwd>    colours nc-2..1 :  y[c] -= rD[c] sum_{c<nb<N} A y[nb]
//     k_spmv<FUSE=1> colour 0 :  S = sum_{local nb} A y[nb]; y[c] -= rD[c] S; v[c] = diag[c] y[c] + S
//     k_spmv<FUSE=0> the rest :  v[c] = diag[c] y[c] + sum A y[nb]
// (for the 2 colours of a hex mesh: 3 half-size gather launches).  Ghost (processor) neighbours are added by
// k_ghost after the halo exchange so that the interior work never waits for NCCL.
// The gather kernels are persistent over row tiles of 256 cells: one thread streams the NEXT tile's matrix rows (nbrA, A,
// rD, diag: contiguous blocks) into the other shared-memory stage with cp.async.bulk + mbarrier while the CTA gathers for
// the current tile, so the only exposed memory latency is the gather itself.
// Scalar control (alpha, omega, beta, convergence per RHS) runs in the last-block epilogue of the
// reductions on one GPU, or in k_ctl after the all-reduce on several.
#pragma once
#include "kernels.cuh"
#include "tma.cuh"

namespace rk {

struct KrylovShared {      // device resident
    KrylovCtl ctl[MAX_RHS];
    int nActive;
    int pad[3];
};

struct SolveCtl { double tol, relTol; int minIter, maxIter; };

__device__ __forceinline__ bool conv_check(double fin, double init, const SolveCtl& sc) {
    return fin < sc.tol || (sc.relTol > 1e-20 && fin < sc.relTol * init);
}

template <int NR> __device__ __forceinline__ void ldv(const double* __restrict__ p, size_t cell, double (&o)[NR]) {
    const double2* q = reinterpret_cast<const double2*>(p + cell * NR);
#pragma unroll
    for (int j = 0; j < NR / 2; ++j) { const double2 t = q[j]; o[2 * j] = t.x; o[2 * j + 1] = t.y; }
}
template <int NR> __device__ __forceinline__ void stv(double* __restrict__ p, size_t cell, const double (&o)[NR]) {
    double2* q = reinterpret_cast<double2*>(p + cell * NR);
#pragma unroll
    for (int j = 0; j < NR / 2; ++j) q[j] = make_double2(o[2 * j], o[2 * j + 1]);
}


#ifndef RK_ROW_MINB
#define RK_ROW_MINB 2   // resident CTAs per SM the row-tile kernels are compiled for (register cap 65536 / (256 * MINB))
#endif
// ---------------------------------------------------------------- row-tile pipeline (TMA bulk copies, 2 stages)
struct RowSrc { const int* nbrT; const double* A; const double* rD; const double* diag; };   // tile-major nbrA / A, per-cell rD / diag
__host__ __device__ __forceinline__ size_t row_stage_bytes(int K) { return (size_t)K * RT * (sizeof(int) + sizeof(double)) + 2 * RT * sizeof(double); }

__device__ __forceinline__ void row_issue(unsigned char* dst, uint64_t* bar, const RowSrc& rs, int K, int tile) {
    const uint32_t nbB = (uint32_t)K * RT * sizeof(int), aB = (uint32_t)K * RT * sizeof(double), vB = RT * sizeof(double);
    mbar_expect_tx(bar, nbB + aB + 2 * vB);
    bulk_g2s(dst, rs.nbrT + (size_t)tile * K * RT, nbB, bar);
    bulk_g2s(dst + nbB, rs.A + (size_t)tile * K * RT, aB, bar);
    bulk_g2s(dst + nbB + aB, rs.rD + (size_t)tile * RT, vB, bar);
    bulk_g2s(dst + nbB + aB + vB, rs.diag + (size_t)tile * RT, vB, bar);
}
struct RowView {   // this thread's row in the current stage
    const int* nb;      // nb[s * RT]
    const double* a;    // a[s * RT]
    double rd, dg;
};
__device__ __forceinline__ RowView row_view(const unsigned char* stage, int K) {
    RowView v;
    v.nb = (const int*)stage + threadIdx.x;
    const double* d = (const double*)(stage + (size_t)K * RT * sizeof(int));
    v.a = d + threadIdx.x;
    v.rd = d[(size_t)K * RT + threadIdx.x];
    v.dg = d[(size_t)K * RT + RT + threadIdx.x];
    return v;
}
__device__ __forceinline__ void row_pipe_init(uint64_t* full) {
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
}

// Row gather  acc[j] = sum_s A[s][c] * y[nb(s)][j]  over the slots selected by SEL:
//   SEL 0: forward substitution (nb < c)   1: backward (c < nb < N)   2: all local columns (nb != c, nb < N)
// KT > 0: compile-time slot count — the K gathers are issued together (excluded slots read the cell's own row with
// coefficient 0, so there is no divergent control flow); KT = 0: run-time K.  Slot order and arithmetic are the same in both.
template <int SEL> __device__ __forceinline__ bool slot_selected(int nb, int c, int N) {
    return SEL == 0 ? (nb < c) : (SEL == 1 ? (nb > c && nb < N) : (nb != c && nb < N));
}
template <int NR, int KT, int SEL>
__device__ __forceinline__ void row_gather(const RowView& rv, int K, int c, int N, const double* __restrict__ y, size_t base, double (&acc)[NR]) {
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = 0.0;
    if constexpr (KT == 0) {
        for (int s = 0; s < K; ++s) {
            const int nb = rv.nb[s * RT];
            if (!slot_selected<SEL>(nb, c, N)) continue;
            const double a = rv.a[s * RT];
            double yn[NR];
            ldv<NR>(y, base + nb, yn);
#pragma unroll
            for (int j = 0; j < NR; ++j) acc[j] += a * yn[j];
        }
    } else {
        int nb[KT];
        double a[KT];
#pragma unroll
        for (int s = 0; s < KT; ++s) { nb[s] = rv.nb[s * RT]; a[s] = rv.a[s * RT]; }
#pragma unroll
        for (int s = 0; s < KT; ++s)
            if (!slot_selected<SEL>(nb[s], c, N)) { nb[s] = c; a[s] = 0.0; }
#pragma unroll
        for (int s = 0; s < KT; ++s) {
            double yn[NR];
            ldv<NR>(y, base + nb[s], yn);
#pragma unroll
            for (int j = 0; j < NR; ++j) acc[j] += a[s] * yn[j];
        }
    }
}

// ---- scalar control, shared by the epilogues (1 GPU), k_ctl (after ncclAllReduce) and the peer-memory kernels.
// One lane per right-hand side (the RHS are independent; MAX_RHS <= 32): ctl_dispatch is called by ALL 32 lanes of one warp.
__device__ __forceinline__ void ctl_init_one(KrylovCtl& k, const double* red3, const SolveCtl& sc) {
    k.normFactor = red3[0] + 1e-20;
    k.initRes = red3[1] / k.normFactor;
    k.finRes = k.initRes;
    k.rho = red3[2];
    k.rhoOld = 0; k.alpha = 0; k.omega = 0; k.beta = 0; k.iters = 0; k.singular = 0; k.pad = 0;
    k.state = (sc.minIter > 0 || !conv_check(k.finRes, k.initRes, sc)) ? 0 : 2;
    if (k.state == 0 && !(fabs(k.rho) > 1e-300)) { k.state = 2; k.singular = 1; }
}
// after r0.v : alpha
__device__ __forceinline__ void ctl_alpha_one(KrylovCtl& k, double dotV) {
    if (k.state == 0) k.alpha = k.rho / dotV;
}
// after sum|s| : half-step convergence
__device__ __forceinline__ void ctl_half_one(KrylovCtl& k, double sumS, const SolveCtl& sc) {
    if (k.state != 0) return;
    k.finRes = sumS / k.normFactor;
    if (conv_check(k.finRes, k.initRes, sc)) k.state = 1;
}
// after t.t, t.s : omega
__device__ __forceinline__ void ctl_omega_one(KrylovCtl& k, double tt, double ts) {
    if (k.state == 0) k.omega = ts / tt;
}
// end of iteration
__device__ __forceinline__ void ctl_end_one(KrylovCtl& k, double sumR, double r0r, const SolveCtl& sc) {
    if (k.state == 1) { k.iters++; k.state = 2; }
    else if (k.state == 0) {
        k.finRes = sumR / k.normFactor;
        k.rhoOld = k.rho;
        k.rho = r0r;
        // EXT-OF9 PBiCGStab::solve: `(++nIterations() < maxIter_ && !converged) || nIterations() < minIter_` — PRE-increment
        // (PBiCG and PCG post-increment: ctl_pb_end_one)
        const bool cont = ((++k.iters < sc.maxIter) && !conv_check(k.finRes, k.initRes, sc)) || k.iters < sc.minIter;
        if (!cont) k.state = 2;
        else if (!(fabs(k.rho) > 1e-300) || !(fabs(k.omega) > 1e-300)) { k.state = 2; k.singular = 1; }
        else k.beta = (k.rho / k.rhoOld) * (k.alpha / k.omega);
    }
}
// ---- PBiCG (pbicg.cuh; EXT-OF9 PBiCG.C): rho holds wArT
__device__ __forceinline__ void ctl_pb_beta_one(KrylovCtl& k, double wArT) {
    if (k.state != 0) return;
    k.rhoOld = k.rho;
    k.rho = wArT;
    k.beta = k.iters == 0 ? 0.0 : k.rho / k.rhoOld;
}
__device__ __forceinline__ void ctl_pb_alpha_one(KrylovCtl& k, double wApT) {
    if (k.state != 0) return;
    if (!(fabs(wApT) / k.normFactor > 1e-300)) { k.state = 2; k.singular = 1; return; }   // solverPerf.checkSingularity: break
    k.alpha = k.rho / wApT;
}
__device__ __forceinline__ void ctl_pb_end_one(KrylovCtl& k, double sumR, const SolveCtl& sc) {
    if (k.state != 0) return;
    k.finRes = sumR / k.normFactor;
    const bool cont = ((k.iters++ < sc.maxIter) && !conv_check(k.finRes, k.initRes, sc)) || k.iters < sc.minIter;
    if (!cont) k.state = 2;
}
enum { CTL_NONE = 0, CTL_INIT = 1, CTL_ALPHA = 2, CTL_HALF = 3, CTL_OMEGA = 4, CTL_END = 5, CTL_HALF_OMEGA = 6,
       CTL_PB_BETA = 7, CTL_PB_ALPHA = 8, CTL_PB_END = 9 };
// CTL_HALF_OMEGA: red = [t.t, t.s per RHS (2 nrhs) | sum|s| per RHS (nrhs)] — the half-step check deferred to the end of the
// second preconditioned product (peer-memory path)
__device__ __noinline__ void ctl_dispatch(int what, KrylovShared* ks, int nrhs, const double* red, const SolveCtl& sc) {
    const int q = threadIdx.x & 31;
    bool active = false;
    if (q < nrhs && what != CTL_NONE) {
        KrylovCtl k = ks->ctl[q];
        switch (what) {
            case CTL_INIT: ctl_init_one(k, red + 3 * q, sc); break;
            case CTL_ALPHA: ctl_alpha_one(k, red[q]); break;
            case CTL_HALF: ctl_half_one(k, red[q], sc); break;
            case CTL_OMEGA: ctl_omega_one(k, red[2 * q], red[2 * q + 1]); break;
            case CTL_HALF_OMEGA: ctl_half_one(k, red[2 * nrhs + q], sc); ctl_omega_one(k, red[2 * q], red[2 * q + 1]); break;
            case CTL_END: ctl_end_one(k, red[2 * q], red[2 * q + 1], sc); break;
            case CTL_PB_BETA: ctl_pb_beta_one(k, red[q]); break;
            case CTL_PB_ALPHA: ctl_pb_alpha_one(k, red[q]); break;
            case CTL_PB_END: ctl_pb_end_one(k, red[q], sc); break;
            default: break;
        }
        ks->ctl[q] = k;
        active = k.state == 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, active);
    if (q == 0 && (what == CTL_INIT || what == CTL_END || what == CTL_PB_ALPHA || what == CTL_PB_END)) ks->nActive = __popc(m);
}
__global__ void k_ctl(int what, KrylovShared* ks, int nrhs, const double* red, SolveCtl sc) {
    pdl_sync();
    if (blockIdx.x == 0 && threadIdx.x < 32) ctl_dispatch(what, ks, nrhs, red, sc);
}

// Reduction epilogue: the last block (of `expected` participating blocks, possibly spread over several
// launches that share partials/counter) sums the partials in a fixed order and optionally runs the
// scalar control step.
__device__ __forceinline__ void finalize_ctl(const double* partials, int nBlocksTotal, int nSlots, double* out, unsigned* counter,
                                             unsigned expected, int what, KrylovShared* ks, int nrhs, const SolveCtl& sc) {
    __shared__ bool isLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(counter, 1u);
        isLast = (t == expected - 1);
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = warp; s < nSlots; s += BLOCK / 32) {
        double x = 0;
        for (int b = lane; b < nBlocksTotal; b += 32) x += partials[(size_t)b * nSlots + s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) out[s] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) *counter = 0;
    if (threadIdx.x < 32 && what != CTL_NONE) {   // `what` is uniform: the whole first warp takes part (one lane per RHS)
        __threadfence();
        ctl_dispatch(what, ks, nrhs, out, sc);
    }
}

// All kernels below are persistent-style: a bounded grid (a few CTAs per SM) strides over the cells, so the
// block reduction / last-block epilogue is paid once per CTA instead of once per 256 cells.

// ---------------------------------------------------------------- initial residual
// gAverage(psi): the per-component sums of theta arrive in sumPsi (accumulated by k_cell_source2, assembly.cuh)
// v = A psi (ghost columns included: psi halos are exchanged before), r = b - v, r0 = r
// sums per RHS: [0] |v - xRef rowsum| + |b - xRef rowsum|, [1] |r|, [2] r.r
#ifndef RK_INIT_MINB
#define RK_INIT_MINB 2   // resident CTAs per SM k_krylov_init is compiled for
#endif
template <int NR, int KT>
__global__ void __launch_bounds__(BLOCK, RK_INIT_MINB) k_krylov_init(MeshView m, int nModes, RhsPtrs rp, const double* __restrict__ diag, const double* __restrict__ A,
                                                        const double* __restrict__ sumPsi, double nGlobal, double* __restrict__ r, double* __restrict__ r0v,
                                                        double* partials, double* out, unsigned* counter, int ctlWhat, KrylovShared* ks, SolveCtl sc) {
    pdl_sync();
    const int stride = gridDim.x * BLOCK;
    for (int md = 0; md < nModes; ++md) {
        double red[3 * NR];
#pragma unroll
        for (int j = 0; j < 3 * NR; ++j) red[j] = 0.0;
        for (int c = blockIdx.x * BLOCK + threadIdx.x; c < m.N; c += stride) {
            const double d = diag[c];
            double rowsum = d;
            double acc[NR];
            double aRow[KT > 0 ? KT : 1];
#pragma unroll
            for (int j = 0; j < NR; ++j) acc[j] = d * rp.psi[md * NR + j][c];
            if constexpr (KT == 0) {
                for (int s = 0; s < m.K; ++s) {
                    const double a = A[ell_t(m.K, s, c)];
                    const int nb = m.nbrA[ell_t(m.K, s, c)];
                    rowsum += a;
#pragma unroll
                    for (int j = 0; j < NR; ++j) acc[j] += a * rp.psi[md * NR + j][nb];
                }
            } else {
                int nb[KT];
#pragma unroll
                for (int s = 0; s < KT; ++s) { nb[s] = m.nbrA[ell_t(KT, s, c)]; aRow[s] = A[ell_t(KT, s, c)]; }
#pragma unroll
                for (int s = 0; s < KT; ++s) rowsum += aRow[s];
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    const double* pj = rp.psi[md * NR + j];
                    double pn[KT];
#pragma unroll
                    for (int s = 0; s < KT; ++s) pn[s] = pj[nb[s]];
#pragma unroll
                    for (int s = 0; s < KT; ++s) acc[j] += aRow[s] * pn[s];
                }
            }
            // source: own part + the deferred values of the faces this cell is downwind of (A = min(F,0) < 0 marks them)
            double bb[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j) bb[j] = rp.b[md * NR + j][c];
            if (rp.corr[md * NR] != nullptr) {
                if constexpr (KT == 0) {
                    for (int s = 0; s < m.K; ++s) {
                        const double as = A[ell_t(m.K, s, c)];
                        if (as < 0.0) {
#pragma unroll
                            for (int j = 0; j < NR; ++j) bb[j] -= as * rp.corr[md * NR + j][(size_t)s * m.NS + c];
                        }
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < KT; ++s) {
                        if (aRow[s] < 0.0) {
#pragma unroll
                            for (int j = 0; j < NR; ++j) bb[j] -= aRow[s] * rp.corr[md * NR + j][(size_t)s * m.NS + c];
                        }
                    }
                }
            }
            double rr[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                rr[j] = bb[j] - acc[j];
                const double t = rowsum * (sumPsi[md * NR + j] / nGlobal);
                red[3 * j] += fabs(acc[j] - t) + fabs(bb[j] - t);
                red[3 * j + 1] += fabs(rr[j]);
                red[3 * j + 2] += rr[j] * rr[j];
            }
            stv<NR>(r, (size_t)md * m.NP + c, rr);
            stv<NR>(r0v, (size_t)md * m.NP + c, rr);
        }
        block_reduce_to_partials<3 * NR>(red, partials, 3 * md * NR, 3 * nModes * NR);
    }
    finalize_ctl(partials, gridDim.x, 3 * nModes * NR, out, counter, gridDim.x, ctlWhat, ks, nModes * NR, sc);
}

// ---------------------------------------------------------------- p = r + beta (p - omega v);  y = rD p   on cells [c0, c1)
template <int NR>
__global__ void __launch_bounds__(BLOCK) k_update_p(int c0, int c1, int NP, int nModes, const KrylovShared* __restrict__ ks, const double* __restrict__ rD,
                                                     const double* __restrict__ r, const double* __restrict__ v, double* __restrict__ p, double* __restrict__ y) {
    pdl_sync();
    if (ks->nActive == 0) return;
    const int stride = gridDim.x * BLOCK;
    for (int md = 0; md < nModes; ++md) {
        double beta[NR], omega[NR];
        bool on[NR], first[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            const KrylovCtl& k = ks->ctl[md * NR + j];
            on[j] = k.state == 0; first[j] = k.iters == 0; beta[j] = k.beta; omega[j] = k.omega;
        }
        for (int c = c0 + blockIdx.x * BLOCK + threadIdx.x; c < c1; c += stride) {
            const double d = rD[c];
            const size_t i = (size_t)md * NP + c;
            double rr[NR], pp[NR], vv[NR], yy[NR];
            ldv<NR>(r, i, rr); ldv<NR>(p, i, pp); ldv<NR>(v, i, vv);
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                yy[j] = 0.0;   // y of a finished RHS is never read again
                if (!on[j]) continue;
                pp[j] = first[j] ? rr[j] : rr[j] + beta[j] * (pp[j] - omega[j] * vv[j]);
                yy[j] = d * pp[j];
            }
            stv<NR>(p, i, pp); stv<NR>(y, i, yy);
        }
    }
}

// ---------------------------------------------------------------- DILU forward / backward phases on a cell range
// UPD (forward sweeps only) fuses the vector update that produces the right-hand side of the substitution, so that the
// freshly computed y = rD p (or z = rD s) of the cell never makes a round trip through HBM:
//   UPD 0 : y holds rD*rhs already             UPD 1 : p = r + beta (p - omega v), y = rD p   (k_update_p of this cell)
//   UPD 2 : s = r - alpha v, z = rD s, sum|s|  (k_make_s of this cell; y is z, the partial sums share partials/counter
//           with the colour-0 launch of k_make_s and the other sweeps: blockBase / totalBlocks)
struct SweepUpd {
    const double* r; const double* v; double* p; double* sv;
    double* partials; double* out; unsigned* counter; int blockBase, totalBlocks, ctlWhat;
    SolveCtl sc;
};
template <int NR, int KT, int FWD, int UPD>
__global__ void __launch_bounds__(RT, RK_ROW_MINB) k_sweep(MeshView m, RowSrc rs, int c0, int c1, int nModes, KrylovShared* ks, double* __restrict__ y, SweepUpd u) {
    pdl_sync();
    if (ks->nActive == 0) return;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ uint64_t full[2];
    const int K = KT > 0 ? KT : m.K;
    const size_t stageBytes = row_stage_bytes(K);
    const int tBeg = c0 / RT, tEnd = (c1 - 1) / RT;
    row_pipe_init(full);
    int it = 0;
    for (int md = 0; md < nModes; ++md) {
        bool on[NR];
        double ca[NR], cb[NR];   // UPD 1: beta, omega;  UPD 2: alpha
        bool first[NR];
        double red[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            const KrylovCtl& k = ks->ctl[md * NR + j];
            on[j] = k.state == 0; first[j] = k.iters == 0;
            ca[j] = UPD == 1 ? k.beta : k.alpha; cb[j] = k.omega;
        }
        if (UPD == 2) {
#pragma unroll
            for (int j = 0; j < NR; ++j) red[j] = 0.0;
        }
        int tile = tBeg + blockIdx.x;
        if (threadIdx.x == 0 && tile <= tEnd) row_issue(smemRaw + (size_t)(it & 1) * stageBytes, &full[it & 1], rs, K, tile);
        for (; tile <= tEnd; tile += gridDim.x, ++it) {
            const int st = it & 1, next = tile + gridDim.x;
            if (threadIdx.x == 0 && next <= tEnd) row_issue(smemRaw + (size_t)(st ^ 1) * stageBytes, &full[st ^ 1], rs, K, next);
            const int c = tile * RT + threadIdx.x;
            const bool valid = c >= c0 && c < c1;
            const size_t i = (size_t)md * m.NP + c;
            // own-cell vectors do not depend on the row: issue their loads before waiting for the stage
            double yy[NR], rr[NR], vv[NR], pp[NR];
            if (valid) {
                if (UPD == 0) ldv<NR>(y, i, yy);
                else { ldv<NR>(u.r, i, rr); ldv<NR>(u.v, i, vv); if (UPD == 1) ldv<NR>(u.p, i, pp); }
            }
            mbar_wait(&full[st], (uint32_t)((it >> 1) & 1));
            if (valid) {
                const RowView rv = row_view(smemRaw + (size_t)st * stageBytes, K);
                double acc[NR];
                row_gather<NR, KT, FWD ? 0 : 1>(rv, K, c, m.N, y, (size_t)md * m.NP, acc);
                if (UPD == 1) {
#pragma unroll
                    for (int j = 0; j < NR; ++j) {
                        yy[j] = 0.0;
                        if (!on[j]) continue;
                        pp[j] = first[j] ? rr[j] : rr[j] + ca[j] * (pp[j] - cb[j] * vv[j]);
                        yy[j] = rv.rd * pp[j];
                    }
                    stv<NR>(u.p, i, pp);
                } else if (UPD == 2) {
                    double ss[NR];
#pragma unroll
                    for (int j = 0; j < NR; ++j) {
                        ss[j] = 0.0; yy[j] = 0.0;
                        if (!on[j]) continue;
                        ss[j] = rr[j] - ca[j] * vv[j];
                        yy[j] = rv.rd * ss[j];
                        red[j] += fabs(ss[j]);
                    }
                    stv<NR>(u.sv, i, ss);
                }
#pragma unroll
                for (int j = 0; j < NR; ++j)
                    if (on[j]) yy[j] -= rv.rd * acc[j];
                stv<NR>(y, i, yy);
            }
            __syncthreads();   // stage st is free for the prefetch issued in the next iteration
        }
        if constexpr (UPD == 2) block_reduce_to_partials<NR>(red, u.partials + (size_t)u.blockBase * nModes * NR, md * NR, nModes * NR);
    }
    if constexpr (UPD == 2) finalize_ctl(u.partials, u.totalBlocks, nModes * NR, u.out, u.counter, (unsigned)u.totalBlocks, u.ctlWhat, ks, nModes * NR, u.sc);
}

// ---------------------------------------------------------------- v = A y (local columns) with fused dots
//   FUSE = 1 : the range is colour 0 — first finish the backward substitution of these cells
//   MODE 0   : dot[q]            = other . v   (other = r0)
//   MODE 1   : dot[2q], [2q+1]   = v . v , v . other   (other = s)
// Several launches (cell ranges) share partials; the last launched range finalises (blockBase/totalBlocks).
template <int NR, int KT, int MODE, int FUSE>
__global__ void __launch_bounds__(RT, RK_ROW_MINB) k_spmv(MeshView m, RowSrc rs, int c0, int c1, int nModes, KrylovShared* ks, double* __restrict__ y, double* __restrict__ v,
                                              const double* __restrict__ other, double* partials, double* out, unsigned* counter, int blockBase,
                                              int totalBlocks, int ctlWhat, SolveCtl sc) {
    pdl_sync();
    if (ks->nActive == 0) return;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ uint64_t full[2];
    constexpr int ND = MODE == 0 ? 1 : 2;
    const int K = KT > 0 ? KT : m.K;
    const size_t stageBytes = row_stage_bytes(K);
    const int tBeg = c0 / RT, tEnd = (c1 - 1) / RT;
    row_pipe_init(full);
    int it = 0;
    for (int md = 0; md < nModes; ++md) {
        double red[ND * NR];
        bool on[NR];
#pragma unroll
        for (int j = 0; j < ND * NR; ++j) red[j] = 0.0;
#pragma unroll
        for (int j = 0; j < NR; ++j) on[j] = ks->ctl[md * NR + j].state == 0;
        int tile = tBeg + blockIdx.x;
        if (threadIdx.x == 0 && tile <= tEnd) row_issue(smemRaw + (size_t)(it & 1) * stageBytes, &full[it & 1], rs, K, tile);
        for (; tile <= tEnd; tile += gridDim.x, ++it) {
            const int st = it & 1, next = tile + gridDim.x;
            if (threadIdx.x == 0 && next <= tEnd) row_issue(smemRaw + (size_t)(st ^ 1) * stageBytes, &full[st ^ 1], rs, K, next);
            const int c = tile * RT + threadIdx.x;
            const bool valid = c >= c0 && c < c1;
            const size_t i = (size_t)md * m.NP + c;
            double yy[NR], oo[NR];
            if (valid) { ldv<NR>(y, i, yy); ldv<NR>(other, i, oo); }
            mbar_wait(&full[st], (uint32_t)((it >> 1) & 1));
            if (valid) {
                const RowView rv = row_view(smemRaw + (size_t)st * stageBytes, K);
                double acc[NR], vv[NR];
                row_gather<NR, KT, 2>(rv, K, c, m.N, y, (size_t)md * m.NP, acc);
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    vv[j] = 0.0;   // v of a finished RHS is never read again
                    if (!on[j]) continue;
                    if (FUSE) yy[j] -= rv.rd * acc[j];
                    vv[j] = rv.dg * yy[j] + acc[j];
                    if (MODE == 0) red[j] += oo[j] * vv[j];
                    else { red[2 * j] += vv[j] * vv[j]; red[2 * j + 1] += vv[j] * oo[j]; }
                }
                if (FUSE) stv<NR>(y, i, yy);
                stv<NR>(v, i, vv);
            }
            __syncthreads();
        }
        block_reduce_to_partials<ND * NR>(red, partials + (size_t)blockBase * ND * nModes * NR, ND * md * NR, ND * nModes * NR);
    }
    finalize_ctl(partials, totalBlocks, ND * nModes * NR, out, counter, (unsigned)totalBlocks, ctlWhat, ks, nModes * NR, sc);
}

// ghost (processor-patch) columns, after the halo exchange of y (NCCL fallback path; the peer-memory path does this in
// k_peer_ghost): grid-stride over the cells that own ghost slots
//   v[c] += sum_{ghost slots} A y[ghost]; the dots are corrected for the change of v — per-CTA partial sums, added to the
//   dots by the last CTA in block order (deterministic: no atomics on doubles)
template <int NR, int MODE>
__global__ void __launch_bounds__(BLOCK) k_ghost(MeshView m, int nBcells, const int* __restrict__ bcells, int nModes, const KrylovShared* __restrict__ ks,
                                                  const double* __restrict__ A, const double* __restrict__ y, double* __restrict__ v,
                                                  const double* __restrict__ other, double* dots, double* partials, unsigned* counter) {
    pdl_sync();
    if (ks->nActive == 0) return;
    constexpr int ND = MODE == 0 ? 1 : 2;
    const int nrhs = nModes * NR;
    for (int md = 0; md < nModes; ++md) {
        double red[ND * NR];
#pragma unroll
        for (int j = 0; j < ND * NR; ++j) red[j] = 0.0;
        for (int i0 = blockIdx.x * BLOCK + threadIdx.x; i0 < nBcells; i0 += gridDim.x * BLOCK) {
            const int c = bcells[i0];
            double acc[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j) acc[j] = 0.0;
            for (int s = 0; s < m.K; ++s) {
                const int nb = m.nbrA[ell_t(m.K, s, c)];
                if (nb < m.N) continue;
                const double a = A[ell_t(m.K, s, c)];
                double yn[NR];
                ldv<NR>(y, (size_t)md * m.NP + nb, yn);
#pragma unroll
                for (int j = 0; j < NR; ++j) acc[j] += a * yn[j];
            }
            const size_t i = (size_t)md * m.NP + c;
            double vv[NR], oo[NR];
            ldv<NR>(v, i, vv); ldv<NR>(other, i, oo);
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                if (ks->ctl[md * NR + j].state != 0) continue;
                const double vn = vv[j] + acc[j];
                if (MODE == 0) red[j] += oo[j] * acc[j];
                else { red[2 * j] += vn * vn - vv[j] * vv[j]; red[2 * j + 1] += acc[j] * oo[j]; }
                vv[j] = vn;
            }
            stv<NR>(v, i, vv);
        }
        block_reduce_to_partials<ND * NR>(red, partials, ND * md * NR, ND * nrhs);
    }
    __shared__ bool isLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) isLast = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = warp; q < ND * nrhs; q += BLOCK / 32) {
        double t = 0;
        for (unsigned b = lane; b < gridDim.x; b += 32) t += __ldcg(&partials[(size_t)b * ND * nrhs + q]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (lane == 0) dots[q] += t;
    }
    if (threadIdx.x == 0) *counter = 0;
}

// ---------------------------------------------------------------- s = r - alpha v ; z = rD s ; sum|s|   on cells [c0, c1)
// (the colour-0 range; the other colours do this inside k_sweep<UPD = 2> and share partials / counter: totalBlocks)
template <int NR>
__global__ void __launch_bounds__(BLOCK) k_make_s(int c0, int c1, int NP, int nModes, KrylovShared* ks, const double* __restrict__ rD, const double* __restrict__ r,
                                                   const double* __restrict__ v, double* __restrict__ sv, double* __restrict__ z, double* partials,
                                                   double* out, unsigned* counter, int totalBlocks, int ctlWhat, SolveCtl sc) {
    pdl_sync();
    if (ks->nActive == 0) return;
    const int stride = gridDim.x * BLOCK;
    for (int md = 0; md < nModes; ++md) {
        double red[NR], alpha[NR];
        bool on[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) { red[j] = 0.0; const KrylovCtl& k = ks->ctl[md * NR + j]; on[j] = k.state == 0; alpha[j] = k.alpha; }
        for (int c = c0 + blockIdx.x * BLOCK + threadIdx.x; c < c1; c += stride) {
            const double d = rD[c];
            const size_t i = (size_t)md * NP + c;
            double rr[NR], vv[NR], ss[NR], zz[NR];
            ldv<NR>(r, i, rr); ldv<NR>(v, i, vv);
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                ss[j] = 0.0; zz[j] = 0.0;
                if (!on[j]) continue;
                ss[j] = rr[j] - alpha[j] * vv[j];
                zz[j] = d * ss[j];
                red[j] += fabs(ss[j]);
            }
            stv<NR>(sv, i, ss); stv<NR>(z, i, zz);
        }
        block_reduce_to_partials<NR>(red, partials, md * NR, nModes * NR);
    }
    finalize_ctl(partials, totalBlocks, nModes * NR, out, counter, (unsigned)totalBlocks, ctlWhat, ks, nModes * NR, sc);
}

// ---------------------------------------------------------------- psi += alpha y + omega z ; r = s - omega t ; sum|r| , r0.r
template <int NR>
__global__ void __launch_bounds__(BLOCK, 2) k_update_x_r(int N, int NP, int nModes, RhsPtrs rp, KrylovShared* ks, const double* __restrict__ y,
                                                       const double* __restrict__ z, const double* __restrict__ sv, const double* __restrict__ t,
                                                       const double* __restrict__ r0v, double* __restrict__ r, double* partials, double* out,
                                                       unsigned* counter, int ctlWhat, SolveCtl sc) {
    pdl_sync();
    if (ks->nActive == 0) return;
    const int stride = gridDim.x * BLOCK;
    for (int md = 0; md < nModes; ++md) {
        double red[2 * NR], alpha[NR], omega[NR];
        int st[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            red[2 * j] = 0.0; red[2 * j + 1] = 0.0;
            const KrylovCtl& k = ks->ctl[md * NR + j];
            st[j] = k.state; alpha[j] = k.alpha; omega[j] = k.omega;
        }
        for (int c = blockIdx.x * BLOCK + threadIdx.x; c < N; c += stride) {
            const size_t i = (size_t)md * NP + c;
            double yy[NR], zz[NR], ss[NR], tt[NR], r0[NR], rr[NR];
            ldv<NR>(y, i, yy); ldv<NR>(z, i, zz); ldv<NR>(sv, i, ss); ldv<NR>(t, i, tt); ldv<NR>(r0v, i, r0);
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                rr[j] = 0.0;
                if (st[j] == 2) continue;
                double* psi = rp.psi[md * NR + j];
                if (st[j] == 1) { psi[c] += alpha[j] * yy[j]; continue; }
                psi[c] += alpha[j] * yy[j] + omega[j] * zz[j];
                rr[j] = ss[j] - omega[j] * tt[j];
                red[2 * j] += fabs(rr[j]);
                red[2 * j + 1] += r0[j] * rr[j];
            }
            stv<NR>(r, i, rr);
        }
        block_reduce_to_partials<2 * NR>(red, partials, 2 * md * NR, 2 * nModes * NR);
    }
    finalize_ctl(partials, gridDim.x, 2 * nModes * NR, out, counter, gridDim.x, ctlWhat, ks, nModes * NR, sc);
}

}  // namespace rk
