// assembly3.cuh — round-2 assembly of the theta transport equation on lattice-sized rows (K = 4 / 6 slots): ONE THREAD PER
// CELL, every solved component in turn, the cell's static geometry in registers.
//
// Same reference loops as assembly.cuh (gaussDefCmpwConvectionScheme.C:93-167, :242-319; boilerLog.H:1; EXT-OF9
// fvMatrix::relax / addBoundaryDiag / addBoundarySource) and the same "upwind cell computes" formulation.  What changed, and why
// (profiles/r1_final_C3_full.md, r2: k_flux_assemble was ISSUE-bound — 190 warp instructions per cell, 64 % of the issue
// slots, HBM at 38-52 %):
//   * k_flux_assemble maps one thread to one (cell, component): every theta warp re-reads the tile's geometry from shared
//     memory (540 LDS.64 per cell — the shared-memory pipe alone needs 80 % of the kernel's HBM time), re-derives the slot
//     flags and the limiter inputs per component.  k_flux3 reads the tile record ONCE per cell into registers (52 LDS) and
//     loops over the 3 velocity and the NC theta components; slot kinds, upwind predicates, A = min(F, 0), gather pointers
//     are derived once.  ~47 warp instructions per cell.
//   * The Gauss-linear face interpolation and the division by V are folded into static per-slot vectors
//         grad(f)_P = G0 f_P + sum_s G_s f_N(s),   G_s = b_s S_s / V,  G0 = sum_s a_s S_s / V
//     (a_s, b_s = the weights of the cell's and the neighbour's value in the face value: owner w, 1-w; neighbour side 1-w, w;
//     patch 0, 1): 3 (K + 1) DFMA per component instead of K x (branch + 2 + 3), and 8 B less per slot in the record.
//   * One warp = one CTA = one tile of 32 consecutive cells, persistent.  The tile record + the tile's face fluxes arrive by
//     two cp.async.bulk copies on one mbarrier; the warp copies them into registers, and lane 0 re-arms the SAME stage for its
//     next tile straight away, so the copy has the whole compute time of the current tile to land (no second stage: 13 KB per
//     warp, 8-10 warps per SM at up to 255 registers, no CTA-wide barrier anywhere).
//   * While the neighbour values are in registers the kernel also forms  A theta  (the Amul of the FIRST residual: diag
//     theta_P + sum_s A_s theta_N), the row sum, the cell's inflow-slot mask and sum(theta) per component.  The initial
//     residual of the Krylov solve then needs no gather at all and is fused with the per-cell source (k_source_init below);
//     `corr` carries v F (the deferred flux itself) instead of v, so that the downwind cell adds it without reading its
//     matrix row:   b -= A_s v  with  A_s = min(-F_up, 0) = -F_up   <=>   b += v F_up.
// Step traffic (3-D, 6 components): k_flux_assemble 848 + k_cell_source2 376 + k_krylov_init 464 = 1,688 B/cell becomes
// k_flux3 876 + k_source_init 628 = 1,504 B/cell, and the 36 scalar gathers of k_krylov_init disappear.
//
// Run-time K (unstructured meshes) and the PBiCG solver keep the round-1 kernels (assembly.cuh, k_krylov_init).
#pragma once
#include "assembly.cuh"
#include "krylov.cuh"

namespace rk {

// Tile record, version 3 (static, one contiguous block per 32 consecutive cells):
//   int    nbr [K][32], meta[K][32]   as in assembly.cuh
//   double G[3][K][32]                b_s S_s / V  (S_s pointing out of the cell)
//   double D[3][K][32]                C_N - C_P in the face's owner -> neighbour frame
//   double G0[3][32]                  sum_s a_s S_s / V
//   double V[32]
__host__ __device__ constexpr size_t tile_record3_bytes(int K) { return (size_t)K * TILE * (2 * sizeof(int) + 6 * sizeof(double)) + 4 * TILE * sizeof(double); }

#ifndef RK_FLUX3_DEPTH
#define RK_FLUX3_DEPTH 1   // fields gathered ahead of the one being worked on (2 with 12 warps per SM measured the same: r3d/r3f)
#endif
#ifndef RK_FLUX3_MINB
#define RK_FLUX3_MINB 16   // resident single-warp CTAs per SM k_flux3 is compiled for (register cap 65536 / (32 * MINB); 13.3 KB of shared memory each)
#endif

// The per-slot vectors G and D stay in the shared-memory stage for the whole tile (the first version copied them to registers:
// 254 registers, 8 warps per SM, 41 % of the stall samples on the first use of a gathered value — gpurun_out/r3a_ncu); the
// neighbour values of field n + 1 (U_x, U_y, U_z, theta_0 ... theta_5) are gathered before field n is worked on, because the
// stores of field n (which may alias anything, as far as the compiler knows) would otherwise pin the loads behind them.
template <int KT, int NC>
__global__ void __launch_bounds__(TILE, RK_FLUX3_MINB) k_flux3(MeshView m, FluxArgs a, const unsigned char* __restrict__ tileRec, int nTiles) {
    pdl_sync();
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ uint64_t full;
    constexpr int KTL = KT * TILE;
    constexpr uint32_t recBytes = (uint32_t)tile_record3_bytes(KT), fluxBytes = (uint32_t)tile_flux_bytes(KT);
    const int lane = threadIdx.x;
    if (lane == 0) {
        mbar_init(&full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0 && (int)blockIdx.x < nTiles) {
        const int first = a.tileOrder ? a.tileOrder[blockIdx.x] : (int)blockIdx.x;
        mbar_expect_tx(&full, recBytes + fluxBytes);
        bulk_g2s(smemRaw, tileRec + (size_t)first * recBytes, recBytes, &full);
        bulk_g2s(smemRaw + recBytes, (const unsigned char*)a.Fell + (size_t)first * fluxBytes, fluxBytes, &full);
    }
    const bool hrs = a.lim.hrs && !a.noConv;
    const Limiter L = a.lim;
    const size_t NP = (size_t)m.NP, nB = (size_t)m.nB;
    const int* sNb = (const int*)smemRaw + lane;
    const int* sMeta = sNb + KTL;
    const double* sG = (const double*)(smemRaw + (size_t)2 * KTL * sizeof(int)) + lane;   // [3][K][TILE]
    const double* sD = sG + 3 * KTL;                                                      // [3][K][TILE]
    const double* sG0 = sD + 3 * KTL;                                                     // [3][TILE], then V[TILE]
    const double* sF = (const double*)(smemRaw + recBytes) + lane;
    double sumTh = 0.0;   // lane g < NC: this warp's running sum of theta component g (warp-reduced tile by tile: 2 registers instead of 2 NC)

    int it = 0;
    for (int t = blockIdx.x; t < nTiles; t += gridDim.x, ++it) {
        const int tile = a.tileOrder ? a.tileOrder[t] : t;
        const int tn = t + (int)gridDim.x;
        const int nextTile = tn < nTiles ? (a.tileOrder ? a.tileOrder[tn] : tn) : -1;
        if (lane == 0 && nextTile >= 0) {   // the stage is busy until this tile is done: bring the next record as far as L2 meanwhile
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(tileRec + (size_t)nextTile * recBytes), "r"(recBytes) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const unsigned char*)a.Fell + (size_t)nextTile * fluxBytes), "r"(fluxBytes) : "memory");
        }
        mbar_wait(&full, (uint32_t)(it & 1));
        const int c = tile * TILE + lane;
        const unsigned vmask = __ballot_sync(0xffffffffu, c < m.N);
        if (c < m.N) {
            // ---- per-slot quantities shared by every field, packed: bit s + base of `fl`; reverse slots 3 bits each in `rev`
            enum : unsigned { UP = 0, GHOST = 6, FIXEDB = 12, GB = 18, CELL = 24 };
            static_assert(KT <= 6, "flag packing");
            unsigned fl = 0, ownMask = 0, rev = 0;
            int gi[KT];   // gather index: neighbour cell / ghost (== the slot's neighbour id), own cell (unused slot: coefficient 0), or patch face
            double F[KT];
            double Dg = a.rDeltaT * sG0[3 * TILE], sumOff = 0, iCcoupled = 0, iCplainAbs = 0, iCplain = 0;
#pragma unroll
            for (int s = 0; s < KT; ++s) {
                const int nb = sNb[s * TILE];
                const int meta = sMeta[s * TILE];
                rev |= (unsigned)(meta >> 8) << (3 * s);
                F[s] = a.noConv ? 0.0 : sF[s * TILE];
                const bool isCell = meta & SLOT_CELL, own = meta & SLOT_OWNER, ghost = meta & SLOT_GHOST;
                ownMask |= (own ? 1u : 0u) << s;
                const bool up = isCell && (own ? (F[s] >= 0) : !(-F[s] >= 0));   // own == pos(phi): this cell is the upwind cell of the face
                fl |= (up ? 1u : 0u) << (UP + s) | (ghost ? 1u : 0u) << (GHOST + s) | (isCell ? 1u : 0u) << (CELL + s) |
                      (((meta & (SLOT_PATCH | SLOT_PATCH_ZG)) == SLOT_PATCH) ? 1u : 0u) << (FIXEDB + s) | (nb <= -2 ? 1u : 0u) << (GB + s);
                gi[s] = nb >= 0 ? nb : (nb == -1 ? c : -nb - 2);
                // diagonal (component independent)
                if (a.bounded) Dg -= F[s];   // unused slots carry F = 0
                if (isCell) {
                    if (!ghost) { Dg += fmax(F[s], 0.0); sumOff += fmax(-F[s], 0.0); }
                    else { iCcoupled += (F[s] >= 0 ? F[s] : 0.0); sumOff += fmax(-F[s], 0.0); }
                } else if (meta & SLOT_PATCH_ZG) { iCplain += F[s]; iCplainAbs += fabs(F[s]); }
            }
            auto Aof = [&](int s) { return ((fl >> (CELL + s)) & 1u) ? fmin(F[s], 0.0) : 0.0; };   // A_s = min(F, 0) on cell slots
            // fields 0..2 = U_x, U_y, U_z (first mode of a step only), 3.. = the solved theta components; the values of field
            // f + DEPTH are gathered before field f is worked on (register buffers rotate at compile time)
            constexpr int NF = 3 + NC, DEPTH = RK_FLUX3_DEPTH, NBUF = DEPTH + 1;
            double ownv[NBUF], nv[NBUF][KT];
            auto gather = [&](int f, int b) {
                const double* __restrict__ fk;
                const double* __restrict__ fB;
                if (f < 3) { fk = a.U + (size_t)f * NP; fB = a.Ub + (size_t)f * nB; }
                else { const int k = a.cl.c[f - 3]; fk = a.theta + (size_t)k * NP; fB = a.thetaB + (size_t)k * nB; }
                ownv[b] = fk[c];
#pragma unroll
                for (int s = 0; s < KT; ++s) nv[b][s] = ((fl >> (GB + s)) & 1u ? fB : fk)[gi[s]];
            };
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) {
                if (a.nU) gather(d, d % NBUF);
                else gather(3 + d, (3 + d) % NBUF);
            }

            double relaxAdd = 0.0;
            if (a.relax > 0) {   // EXT-OF9 fvMatrix::relax
                const double D0 = Dg;
                double Dn = Dg + iCcoupled + iCplainAbs;
                Dn = fmax(fabs(Dn), sumOff);
                Dn /= a.relax;
                Dn -= iCcoupled;
                Dn -= iCplain;
                relaxAdd = Dn - D0;
                Dg = Dn;
            }
            const double Dfull = Dg + iCcoupled + iCplain;   // addBoundaryDiag
            if (a.writeMatrix) {
                double rowsum = Dfull;
                unsigned mask = 0;
#pragma unroll
                for (int s = 0; s < KT; ++s) {
                    const double As = Aof(s);
                    a.Fs[ell_t(KT, s, c)] = As;
                    rowsum += As;
                    mask |= (As < 0.0 ? 1u : 0u) << s;
                }
                if (a.FsT) {
#pragma unroll
                    for (int s = 0; s < KT; ++s) a.FsT[ell_t(KT, s, c)] = ((fl >> (CELL + s)) & 1u) ? fmin(-F[s], 0.0) : 0.0;
                }
                a.diag[c] = Dfull;
                a.rD[c] = 1.0 / Dfull;   // DILU: upper*lower == 0 on every face of an upwind matrix
                a.rowsum[c] = rowsum;
                a.inflow[c] = mask;
            }
            const double G0x = sG0[0], G0y = sG0[TILE], G0z = sG0[2 * TILE];

#pragma unroll
            for (int f = 0; f < NF; ++f) {
                if (f < 3 && !a.nU) continue;
                if (f + DEPTH < NF) gather(f + DEPTH, (f + DEPTH) % NBUF);
                const double tP = ownv[f % NBUF];
                const double(&cur)[KT] = nv[f % NBUF];
                if (f < 3) {
                    // ---- grad(U)  (boilerLog.H:1; first mode of a step)
                    double gx = G0x * tP, gy = G0y * tP, gz = G0z * tP;
#pragma unroll
                    for (int s = 0; s < KT; ++s) { gx += sG[s * TILE] * cur[s]; gy += sG[KTL + s * TILE] * cur[s]; gz += sG[2 * KTL + s * TILE] * cur[s]; }
                    a.gradU[(size_t)(3 * f) * NP + c] = gx;
                    a.gradU[(size_t)(3 * f + 1) * NP + c] = gy;
                    a.gradU[(size_t)(3 * f + 2) * NP + c] = gz;
                    continue;
                }
                // ---- theta component g
                const int g = f - 3 < NC ? f - 3 : 0;
                const int k = a.cl.c[g];
                {
                    double x = tP;
                    if (vmask == 0xffffffffu) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                    } else {   // the mesh's one partial tile
                        double tot = 0.0;
                        for (int l = 0; l < 32; ++l)
                            if ((vmask >> l) & 1u) tot += __shfl_sync(vmask, x, l);
                        x = tot;
                    }
                    if (lane == g) sumTh += x;
                }
                // A theta of the first residual (lduMatrix::Amul incl. the processor interfaces; slot order as k_krylov_init)
                double ac = Dfull * tP;
#pragma unroll
                for (int s = 0; s < KT; ++s) ac += Aof(s) * cur[s];
                a.acc[(size_t)g * NP + c] = ac;
                double sou = 0.0, bnd = 0.0;
#pragma unroll
                for (int s = 0; s < KT; ++s)
                    if ((fl >> (FIXEDB + s)) & 1u) bnd += -F[s] * cur[s];   // fixedValue patch: boundaryCoeffs = -phi_b theta_b
                if (hrs) {
                    // Gauss-linear gradient of theta_k (gaussDefCmpwConvectionScheme.C:254), in registers
                    double gx = G0x * tP, gy = G0y * tP, gz = G0z * tP;
#pragma unroll
                    for (int s = 0; s < KT; ++s) { gx += sG[s * TILE] * cur[s]; gy += sG[KTL + s * TILE] * cur[s]; gz += sG[2 * KTL + s * TILE] * cur[s]; }
                    double* corrRow = a.corr + (size_t)g * KT * m.NS;
#pragma unroll
                    for (int s = 0; s < KT; ++s) {
                        double v = 0.0;
                        const bool ghost = (fl >> (GHOST + s)) & 1u;
                        if ((fl >> (UP + s)) & 1u) {   // faces this cell is the upwind cell of: deferred face value, own source, hand-over
                            const bool own = (ownMask >> s) & 1u;
                            const double gd = gx * sD[s * TILE] + gy * sD[KTL + s * TILE] + gz * sD[2 * KTL + s * TILE];
                            v = phif_defc(own ? tP : cur[s], own ? cur[s] : tP, gd, gd, own, L);
                            const double vF = v * F[s];
                            sou += vF;   // souT[own] += v*phi ; souT[nei] -= v*phi
                            if (!ghost) corrRow[((rev >> (3 * s)) & 7u) * (unsigned)m.NS + (unsigned)gi[s]] = vF;   // K*NS < 2^31 (checked at create)
                        }
                        if (ghost) {
                            // processor face: v travels with the halo exchange (0 where the other side is upwind, so that the
                            // receiver never reads an unwritten word); my own ghost slot of `corr` is cleared
                            a.ghostCorr[(size_t)(gi[s] - m.N) * a.ghostStride + a.ghostOffset + g] = v;
                            corrRow[(unsigned)s * (unsigned)m.NS + (unsigned)c] = 0.0;
                        }
                    }
                }
                a.bsrc[(size_t)k * NP + c] = (-sou + relaxAdd * tP) + bnd;
            }
        }
        __syncwarp();   // every lane is done with the stage: it may be overwritten
        if (lane == 0 && nextTile >= 0) {
            mbar_expect_tx(&full, recBytes + fluxBytes);
            bulk_g2s(smemRaw, tileRec + (size_t)nextTile * recBytes, recBytes, &full);
            bulk_g2s(smemRaw + recBytes, (const unsigned char*)a.Fell + (size_t)nextTile * fluxBytes, fluxBytes, &full);
        }
    }

    // ---- sum(theta) per solved component (gAverage(psi) of the solver's normFactor): warp sums, the last CTA adds them in CTA order
    if (a.sumOut == nullptr) return;
    if (lane < NC) a.sumPartials[(size_t)blockIdx.x * NC + lane] = sumTh;
    __threadfence();
    unsigned last = 0;
    if (lane == 0) last = atomicAdd(a.counter, 1u) == gridDim.x - 1 ? 1u : 0u;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
#pragma unroll
    for (int g = 0; g < NC; ++g) {
        double x = 0;
        for (unsigned b = lane; b < gridDim.x; b += 32) x += __ldcg(&a.sumPartials[(size_t)b * NC + g]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) a.sumOut[g] = x;
    }
    if (lane == 0) *a.counter = 0;
}

// ---------------------------------------------------------------- per-cell source + first Krylov residual, fused
// k_cell_source2 (Omega/B split, model term, ddt source; boilerLog.H:26-36, Oldroyd_BLog.C:141-163 and the other models) and
// k_krylov_init (EXT-OF9 PBiCGStab::solve: r = b - A psi, normFactor, initial residual) in one streaming pass: with A theta,
// the row sum and the inflow mask left behind by k_flux3 the residual needs no neighbour value, so the source b is formed in
// registers and never stored.  One launch per mode (the model is a compile-time constant); the launches of a batch share
// partials / counter, the last CTA of the last launch finalises the 3 NR sums per mode and runs the CTL_INIT control step.
struct SrcInitArgs {
    SourceArgs s;            // bsrc: the own-face part left by k_flux3 (+ received processor-face values), read only
    const double* corr;      // this mode's [NR][K * NS]: v F handed over by the upwind neighbours; null: upwind scheme
    const double* acc;       // this mode's [NR][NP]: A theta
    const double* rowsum;    // [N] diag + sum_s A_s
    const unsigned* inflow;  // [N] bit s: slot s is an inflow face (A_s < 0)
    const double* sumPsi;    // this mode's NR component sums of theta (all ranks)
    double nGlobal;
    double* r; double* r0;   // this mode's block of the interleaved Krylov vectors: [cell * NR + j]
    double* partials; double* out; unsigned* counter;
    int slotBase, nSlots, totalBlocks, ctlWhat, nrhs;
    KrylovShared* ks; SolveCtl sc;
};

template <int NV, int BS>
__device__ __forceinline__ void block_reduce_to_partials_bs(double (&v)[NV], double* partials, int slotBase, int nSlots) {
    __shared__ double sm[BS / 32][NV];
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
        for (int wv = 0; wv < BS / 32; ++wv) x += sm[wv][threadIdx.x];
        partials[(size_t)blockIdx.x * nSlots + slotBase + threadIdx.x] = x;
    }
    __syncthreads();
}
template <int BS>
__device__ __forceinline__ void finalize_ctl_bs(const double* partials, int nBlocksTotal, int nSlots, double* out, unsigned* counter, unsigned expected,
                                                int what, KrylovShared* ks, int nrhs, const SolveCtl& sc) {
    __shared__ bool isLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) isLast = atomicAdd(counter, 1u) == expected - 1;
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = warp; s < nSlots; s += BS / 32) {
        double x = 0;
        for (int b = lane; b < nBlocksTotal; b += 32) x += __ldcg(&partials[(size_t)b * nSlots + s]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) out[s] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) *counter = 0;
    if (threadIdx.x < 32 && what != CTL_NONE) {
        __threadfence();
        ctl_dispatch(what, ks, nrhs, out, sc);
    }
}

template <int MODEL, int NR, int KT>
__global__ void __launch_bounds__(SRC_BLOCK, 3) k_source_init(MeshView m, SrcInitArgs a) {
    pdl_sync();
    static_assert(3 * NR <= SRC_BLOCK, "one thread per reduction slot");
    // the thread's 3 NR running sums live in shared memory ([slot][thread]: conflict-free), not in 36 registers that would be
    // live across the whole cell loop
    __shared__ double sRed[3 * NR][SRC_BLOCK];
    double* red = &sRed[0][threadIdx.x];
#pragma unroll
    for (int j = 0; j < 3 * NR; ++j) red[j * SRC_BLOCK] = 0.0;
    const size_t NP = (size_t)m.NP;
    for (int c = blockIdx.x * SRC_BLOCK + threadIdx.x; c < m.N; c += gridDim.x * SRC_BLOCK) {
        double g[9], thO[6], th[6], Rm[9], lm[3], rhs[6];
#pragma unroll
        for (int i = 0; i < 9; ++i) g[i] = a.s.gradU[(size_t)i * NP + c];
#pragma unroll
        for (int k = 0; k < 6; ++k) th[k] = a.s.theta[(size_t)k * NP + c];
#pragma unroll
        for (int k = 0; k < 9; ++k) Rm[k] = a.s.R[(size_t)k * NP + c];
#pragma unroll
        for (int k = 0; k < 3; ++k) lm[k] = a.s.lam[(size_t)k * NP + c];
#pragma unroll
        for (int k = 0; k < 6; ++k) thO[k] = a.s.thetaOld[(size_t)k * NP + c];
        const double V = m.V[c];
        const double rowsum = a.rowsum[c];
        const unsigned mask = a.corr ? a.inflow[c] : 0u;
        // g[3k+d] = d_d U_k  ->  L_ij = d_i U_j = g[3j+i]
        const double Lg[9] = {g[0], g[3], g[6], g[1], g[4], g[7], g[2], g[5], g[8]};
        double f;
        ModelParams mp = a.s.mp;
        if (a.s.lamCell) { mp.lambda = a.s.lamCell[c]; mp.etaP = a.s.etaCell[c]; }   // Oldroyd_BLog.C:133-135
        if constexpr (MODEL == RHEO_MODEL_SARAMITO_LOG || MODEL == RHEO_MODEL_BMP_FLUIDITY) {
            double tc[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) tc[k] = a.s.tau[(size_t)k * NP + c];
            f = model_rhs<MODEL>(mp, Lg, th, Rm, lm, rhs, tc);
        } else {
            f = model_rhs<MODEL>(mp, Lg, th, Rm, lm, rhs);
        }
        // second batch of loads, all in flight together (they used to sit behind the per-component branches): the own-face
        // part of the source, A theta, and the deferred fluxes of the faces this cell is downwind of (predicated, slot order)
        double bs[NR], acv[NR], cin[NR][KT], thOO[NR];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            int j = k;   // NR == 6: every component is solved, in order
            if constexpr (NR != 6) {
                j = a.s.solvedIdx[k];   // uniform
                if (j < 0) continue;
            }
#pragma unroll
            for (int jj = 0; jj < NR; ++jj)
                if (jj == j) {
                    bs[jj] = a.s.bsrc[(size_t)k * NP + c];
                    acv[jj] = a.acc[(size_t)jj * NP + c];
                    thOO[jj] = a.s.backward ? a.s.thetaOldOld[(size_t)k * NP + c] : 0.0;
                    const double* cr = a.corr + (size_t)jj * KT * m.NS + c;
#pragma unroll
                    for (int s = 0; s < KT; ++s) cin[jj][s] = (mask & (1u << s)) ? cr[(size_t)s * m.NS] : 0.0;
                }
        }
        double rr[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) rr[j] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            int j = k;
            if constexpr (NR != 6) {
                j = a.s.solvedIdx[k];
                if (j < 0) continue;
            }
#pragma unroll
            for (int jj = 0; jj < NR; ++jj)
                if (jj == j) {
                    double ddtSrc;
                    if (a.s.backward) ddtSrc = a.s.rDeltaT * V * (a.s.c0 * thO[k] - a.s.c00 * thOO[jj]);
                    else ddtSrc = a.s.rDeltaT * thO[k] * V;
                    double bb = (ddtSrc + V * rhs[k]) + bs[jj];
#pragma unroll
                    for (int s = 0; s < KT; ++s)
                        if (mask & (1u << s)) bb += cin[jj][s];
                    const double ac = acv[jj];
                    const double res = bb - ac;
                    const double ta = fabs(ac - rowsum * (a.sumPsi[jj] / a.nGlobal)) + fabs(bb - rowsum * (a.sumPsi[jj] / a.nGlobal));
                    rr[jj] = res; red[(3 * jj) * SRC_BLOCK] += ta; red[(3 * jj + 1) * SRC_BLOCK] += fabs(res); red[(3 * jj + 2) * SRC_BLOCK] += res * res;
                }
        }
        a.s.fFene[c] = f;
        stv<NR>(a.r, (size_t)c, rr);
        stv<NR>(a.r0, (size_t)c, rr);
    }
    double redv[3 * NR];
#pragma unroll
    for (int j = 0; j < 3 * NR; ++j) redv[j] = red[j * SRC_BLOCK];
    block_reduce_to_partials_bs<3 * NR, SRC_BLOCK>(redv, a.partials, a.slotBase, a.nSlots);
    const SolveCtl sc = a.sc;   // a local copy: the control step takes it by reference (ctl_dispatch is not inlined), and the address of a
                                // member of `a` would move the whole parameter block to local memory
    finalize_ctl_bs<SRC_BLOCK>(a.partials, gridDim.x, a.nSlots, a.out, a.counter, (unsigned)a.totalBlocks, a.ctlWhat, a.ks, a.nrhs, sc);
}

}  // namespace rk
