// blocksweep.cuh — DILU substitutions in the BLOCK ordering (host/ordering.hpp): natural cell order inside chunks of 32
// cells (4x4x2 lattice cells in 3-D, 8x4 in 2-D), chunks coloured so that chunks of one colour are pairwise non-adjacent.
//
// EXT-OF9 DILUPreconditioner::precondition restated for that numbering.  One WARP owns one chunk, one thread one cell; the
// CTA still walks row tiles of 256 cells (8 chunks) whose matrix rows arrive through the cp.async.bulk pipeline of krylov.cuh:
//   * neighbours in OTHER chunks belong to another colour; the kernels run colour by colour, so those values are final in
//     HBM and are gathered as before (outflow faces have A = min(F, 0) = 0 and are not gathered at all);
//   * neighbours INSIDE the chunk are resolved in shared memory by a level-scheduled sweep: a cell of level l (longest chain
//     of lower-numbered in-chunk neighbours, precomputed on the host) is updated in round l, after a __syncwarp — no CTA-wide
//     barrier, so the 16-24 resident warps of an SM hide each other's latency chains (a 4x4x2 block has 8 levels, an 8x4
//     block 11).  The slot classification (in-chunk lower / higher, shared-memory index, coefficient) is done once per cell,
//     before the rounds.  The arithmetic is the sequential substitution's (same operands; in-chunk and out-of-chunk sums
//     are formed separately) — tests hold the iteration history to the oracle's on the renumbered mesh.
//   (The first version of this round used 256-cell chunks swept by the whole CTA with __syncthreads between 18 + 18 levels:
//    same iteration counts, but every round was an exposed latency chain — 4x the memory time; profiles/r2_ordering.md.)
// Per preconditioned product x = M^-1 rhs, w = A x with nc chunk colours:
//   k_bsweep<DIR 0>   colours 0 .. nc-2 : (vector update of the cell) + forward substitution
//   k_bsweep<DIR 1>   colour nc-1       : forward AND backward substitution in one pass (no higher colour exists)
//   k_bsweep<DIR 2>   colours nc-2 .. 1 : backward substitution
//   k_bspmv0          colour 0          : backward substitution + SpMV of these cells (all their out-of-chunk neighbours are
//                                         final, the in-chunk ones are in shared memory)
//   k_spmv<MODE, 0>   colours 1 .. nc-1 : SpMV (krylov.cuh)
// i.e. for the two colours of a box of whole blocks: 4 launches over half the mesh each, as in round 1's cell colouring, with
// the reference's iteration counts instead of one iteration more.
#pragma once
#include "krylov.cuh"

namespace rk {

constexpr int CH = 32;        // cells per chunk (== rk_host::CHUNK)
#ifndef RK_BLK_MINB
#define RK_BLK_MINB 2         // resident CTAs per SM the chunk kernels are compiled for
#endif
constexpr int CH_SHIFT = 5;

// nbrA holds c itself for unused / boundary slots and indices >= N for ghosts (which may share c's chunk index range)
__device__ __forceinline__ bool in_chunk(int nb, int c, int N) { return (nb >> CH_SHIFT) == (c >> CH_SHIFT) && nb < N && nb != c; }

// this cell's matrix row, classified once: out-of-chunk neighbours (gathered from HBM) and in-chunk ones (shared memory).
// A lattice cell has at most 3 lower and 3 higher neighbours, so the in-chunk ones are compacted into 3 + 3 fixed registers
// (shared-memory offset of the neighbour's record, coefficient); empty entries point at the cell's own record with
// coefficient 0.  The rounds of chunk_sweep then cost 3 x (3 LDS.128 + 6 DFMA) instead of a 6-slot loop with branches.
template <int KT> struct RowSplit {
    int nb[KT > 0 ? KT : 1];      // neighbour (global index)
    double aRem[KT > 0 ? KT : 1]; // coefficient if the slot is an OUT-of-chunk local column (nb < N), else 0
    int lo[3], hi[3];             // shared-memory record index (cell & 255) of the in-chunk lower / higher neighbours, slot order
    double aLo[3], aHi[3];
};
template <int KT>
__device__ __forceinline__ void split_row(const RowView& rv, int c, int N, RowSplit<KT>& r) {
    const int self = c & (RT - 1);
    int nl = 0, nh = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) { r.lo[i] = self; r.hi[i] = self; r.aLo[i] = 0.0; r.aHi[i] = 0.0; }
#pragma unroll
    for (int s = 0; s < KT; ++s) {
        const int nb = rv.nb[s * RT];
        const double a = rv.a[s * RT];
        const bool loc = in_chunk(nb, c, N);
        const bool rem = !loc && nb != c && nb < N;
        r.nb[s] = nb;
        r.aRem[s] = rem ? a : 0.0;
        const bool low = loc && nb < c, high = loc && nb > c;
        const int idx = nb & (RT - 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (low && nl == i) { r.lo[i] = idx; r.aLo[i] = a; }
            if (high && nh == i) { r.hi[i] = idx; r.aHi[i] = a; }
        }
        nl += low ? 1 : 0;
        nh += high ? 1 : 0;
    }
}

// acc[j] = sum over out-of-chunk slots of A[s] * y[nb][j]   WHICH 0: lower (nb < c)   1: higher (nb > c)   2: all
template <int NR, int KT, int WHICH>
__device__ __forceinline__ void gather_remote(const RowSplit<KT>& r, const RowView& rv, int K, int c, int N, const double* __restrict__ y, size_t base,
                                              double (&acc)[NR]) {
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = 0.0;
    if constexpr (KT == 0) {
        for (int s = 0; s < K; ++s) {
            const int nb = rv.nb[s * RT];
            if (nb == c || nb >= N || in_chunk(nb, c, N)) continue;
            if ((WHICH == 0 && nb > c) || (WHICH == 1 && nb < c)) continue;
            const double a = rv.a[s * RT];
            if (a == 0.0) continue;
            double yn[NR];
            ldv<NR>(y, base + nb, yn);
#pragma unroll
            for (int j = 0; j < NR; ++j) acc[j] += a * yn[j];
        }
    } else {
#pragma unroll
        for (int s = 0; s < KT; ++s) {
            // not selected, or an outflow face (A = min(F, 0) = 0: half of the faces of an upwind matrix): nothing to add
            const double a = ((WHICH == 0 && r.nb[s] > c) || (WHICH == 1 && r.nb[s] < c)) ? 0.0 : r.aRem[s];
            if (a == 0.0) continue;
            double yn[NR];
            ldv<NR>(y, base + r.nb[s], yn);
#pragma unroll
            for (int j = 0; j < NR; ++j) acc[j] += a * yn[j];
        }
    }
}

// acc[j] = sum over in-chunk slots of A[s] * ys[record][j] (shared memory)   WHICH 0: lower   1: higher   2: both
template <int NR, int KT, int WHICH>
__device__ __forceinline__ void gather_local(const RowSplit<KT>& r, const RowView& rv, int K, int c, int N, const double* ys, double (&acc)[NR]) {
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = 0.0;
    if constexpr (KT == 0) {
        for (int s = 0; s < K; ++s) {
            const int nb = rv.nb[s * RT];
            if (!in_chunk(nb, c, N)) continue;
            if ((WHICH == 0 && nb > c) || (WHICH == 1 && nb < c)) continue;
            const double a = rv.a[s * RT];
            const double* yn = ys + (size_t)(nb & (RT - 1)) * NR;
#pragma unroll
            for (int j = 0; j < NR; ++j) acc[j] += a * yn[j];
        }
    } else {
        if (WHICH != 1) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double yn[NR];
                ldv<NR>(ys, (size_t)r.lo[i], yn);
#pragma unroll
                for (int j = 0; j < NR; ++j) acc[j] += r.aLo[i] * yn[j];
            }
        }
        if (WHICH != 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                double yn[NR];
                ldv<NR>(ys, (size_t)r.hi[i], yn);
#pragma unroll
                for (int j = 0; j < NR; ++j) acc[j] += r.aHi[i] * yn[j];
            }
        }
    }
}

// level-scheduled in-chunk substitution of one warp: after round l every cell of level <= l holds its final value in ys
//   BWD 0: forward (levels of lower neighbours), BWD 1: backward (levels of higher neighbours).  Called by whole warps.
// (A right-hand side that is no longer active holds zeros in every cell, so its rounds add 0 * 0: no predicate per RHS.)
template <int NR, int KT, int BWD>
__device__ __forceinline__ void chunk_sweep(const RowSplit<KT>& r, const RowView& rv, int K, int c, int N, bool valid, int myLev, int maxLev, double rd,
                                            const bool (&on)[NR], double* ys, double (&yy)[NR]) {
    for (int l = 1; l <= maxLev; ++l) {
        __syncwarp();
        if (valid && myLev == l) {
            double acc[NR];
            gather_local<NR, KT, BWD>(r, rv, K, c, N, ys, acc);
#pragma unroll
            for (int j = 0; j < NR; ++j) yy[j] -= rd * acc[j];
            stv<NR>(ys, (size_t)threadIdx.x, yy);
        }
    }
    __syncwarp();
}

// DIR 0: forward   1: forward + backward (last colour)   2: backward (middle colours; UPD must be 0)
// FIRST: colour 0 — no out-of-chunk lower neighbour exists (forward), nothing is gathered from HBM
// UPD as in k_sweep (krylov.cuh): 0 none, 1 p = r + beta (p - omega v) / y = rD p, 2 s = r - alpha v / z = rD s / sum|s|
template <int NR, int KT, int DIR, int UPD, int FIRST>
__global__ void __launch_bounds__(RT, RK_BLK_MINB) k_bsweep(MeshView m, RowSrc rs, int c0, int c1, int nModes, KrylovShared* ks, double* __restrict__ y, SweepUpd u) {
    pdl_sync();
    if (ks->nActive == 0) return;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ uint64_t full[2];
    __shared__ __align__(16) double ys[RT * NR];
    const int K = KT > 0 ? KT : m.K;
    const size_t stageBytes = row_stage_bytes(K);
    const int tBeg = c0 / RT, tEnd = (c1 - 1) / RT;
    row_pipe_init(full);
    int it = 0;
    for (int md = 0; md < nModes; ++md) {
        bool on[NR];
        double ca[NR], cb[NR];
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
            const bool valid = c >= c0 && c < c1;   // colour ranges are chunk-aligned: a warp is in range or out of it as a whole
            const size_t i = (size_t)md * m.NP + c;
            double yy[NR], rr[NR], vv[NR], pp[NR];
            int myLev = 0, chunkLev = 0;
            if (valid) {
                if (UPD == 0) ldv<NR>(y, i, yy);
                else { ldv<NR>(u.r, i, rr); ldv<NR>(u.v, i, vv); if (UPD == 1) ldv<NR>(u.p, i, pp); }
                myLev = m.lev[c];
                chunkLev = m.chunkLev[c >> CH_SHIFT];
            }
            mbar_wait(&full[st], (uint32_t)((it >> 1) & 1));
            const RowView rv = row_view(smemRaw + (size_t)st * stageBytes, K);
            RowSplit<KT> sp;
            if constexpr (KT > 0) split_row<KT>(rv, c, m.N, sp);
            if (valid) {
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
                if (!(FIRST && DIR != 2)) {   // out-of-chunk neighbours of the direction being substituted (final in HBM)
                    double acc[NR];
                    gather_remote<NR, KT, DIR == 2 ? 1 : 0>(sp, rv, K, c, m.N, y, (size_t)md * m.NP, acc);
#pragma unroll
                    for (int j = 0; j < NR; ++j)
                        if (on[j]) yy[j] -= rv.rd * acc[j];
                }
                double* mine = ys + (size_t)threadIdx.x * NR;
#pragma unroll
                for (int j = 0; j < NR; ++j) mine[j] = yy[j];
            }
            chunkLev = __shfl_sync(0xffffffffu, chunkLev, 0);   // warp-uniform round count (lane 0 is valid whenever any lane is)
            if (DIR != 2) chunk_sweep<NR, KT, 0>(sp, rv, K, c, m.N, valid, myLev & 255, chunkLev & 255, rv.rd, on, ys, yy);
            if (DIR != 0) chunk_sweep<NR, KT, 1>(sp, rv, K, c, m.N, valid, myLev >> 8, chunkLev >> 8, rv.rd, on, ys, yy);
            if (valid) stv<NR>(y, i, yy);
            __syncthreads();   // stage st is free for the prefetch issued in the next iteration (ys is per warp)
        }
        if constexpr (UPD == 2) block_reduce_to_partials<NR>(red, u.partials + (size_t)u.blockBase * nModes * NR, md * NR, nModes * NR);
    }
    if constexpr (UPD == 2) finalize_ctl(u.partials, u.totalBlocks, nModes * NR, u.out, u.counter, (unsigned)u.totalBlocks, u.ctlWhat, ks, nModes * NR, u.sc);
}

// colour 0: backward substitution (every out-of-chunk neighbour is a higher colour) fused with the SpMV of these cells;
// dots as in k_spmv (MODE 0: other . v;  MODE 1: v . v, v . other); shares partials / counter with the k_spmv launch over
// the other colours (blockBase / totalBlocks)
template <int NR, int KT, int MODE>
__global__ void __launch_bounds__(RT, RK_BLK_MINB) k_bspmv0(MeshView m, RowSrc rs, int c0, int c1, int nModes, KrylovShared* ks, double* __restrict__ y, double* __restrict__ v,
                                                const double* __restrict__ other, double* partials, double* out, unsigned* counter, int blockBase,
                                                int totalBlocks, int ctlWhat, SolveCtl sc) {
    pdl_sync();
    if (ks->nActive == 0) return;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ uint64_t full[2];
    __shared__ __align__(16) double ys[RT * NR];
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
            double yy[NR], oo[NR], accR[NR];
            int myLev = 0, chunkLev = 0;
            if (valid) { ldv<NR>(y, i, yy); ldv<NR>(other, i, oo); myLev = m.lev[c]; chunkLev = m.chunkLev[c >> CH_SHIFT]; }
            mbar_wait(&full[st], (uint32_t)((it >> 1) & 1));
            const RowView rv = row_view(smemRaw + (size_t)st * stageBytes, K);
            RowSplit<KT> sp;
            if constexpr (KT > 0) split_row<KT>(rv, c, m.N, sp);
            if (valid) {
                gather_remote<NR, KT, 2>(sp, rv, K, c, m.N, y, (size_t)md * m.NP, accR);
                double* mine = ys + (size_t)threadIdx.x * NR;
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    if (on[j]) yy[j] -= rv.rd * accR[j];
                    mine[j] = yy[j];
                }
            }
            chunkLev = __shfl_sync(0xffffffffu, chunkLev, 0);   // warp-uniform round count (lane 0 is valid whenever any lane is)
            chunk_sweep<NR, KT, 1>(sp, rv, K, c, m.N, valid, myLev >> 8, chunkLev >> 8, rv.rd, on, ys, yy);   // ends with __syncwarp: the chunk's ys is final
            if (valid) {
                double accL[NR], vv[NR];
                gather_local<NR, KT, 2>(sp, rv, K, c, m.N, ys, accL);
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    vv[j] = 0.0;
                    if (!on[j]) continue;
                    vv[j] = rv.dg * yy[j] + (accR[j] + accL[j]);
                    if (MODE == 0) red[j] += oo[j] * vv[j];
                    else { red[2 * j] += vv[j] * vv[j]; red[2 * j + 1] += vv[j] * oo[j]; }
                }
                stv<NR>(y, i, yy);
                stv<NR>(v, i, vv);
            }
            __syncthreads();
        }
        block_reduce_to_partials<ND * NR>(red, partials + (size_t)blockBase * ND * nModes * NR, ND * md * NR, ND * nModes * NR);
    }
    finalize_ctl(partials, totalBlocks, ND * nModes * NR, out, counter, (unsigned)totalBlocks, ctlWhat, ks, nModes * NR, sc);
}

}  // namespace rk
