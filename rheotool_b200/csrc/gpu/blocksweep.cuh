// blocksweep.cuh — DILU substitutions in the BLOCK ordering (host/ordering.hpp): natural cell order inside chunks of 256
// cells, chunks coloured so that chunks of one colour are pairwise non-adjacent.
//
// EXT-OF9 DILUPreconditioner::precondition restated for that numbering.  One CTA owns one chunk (= one row tile of the
// matrix, streamed by the same cp.async.bulk pipeline as krylov.cuh), one thread one cell:
//   * neighbours in OTHER chunks belong to another colour; the kernels run colour by colour, so those values are final in
//     HBM and are gathered as before;
//   * neighbours INSIDE the chunk are resolved in shared memory by a level-scheduled sweep: a cell of level l (longest chain
//     of lower-numbered in-chunk neighbours, precomputed on the host) is updated in round l, after a __syncthreads; an
//     8x8x4 block has 18 levels, a 16x16 block 31.  The arithmetic is the sequential substitution's (same operands, the
//     in-chunk and out-of-chunk sums are formed separately) — tests hold the iteration history to the oracle's on the
//     renumbered mesh.
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

// slot classification for cell c of row tile `tile` (nbrA holds c itself for unused / boundary slots, >= N for ghosts)
__device__ __forceinline__ bool in_chunk(int nb, int c, int tile, int N) { return (nb >> 8) == tile && nb < N && nb != c; }

// acc[j] = sum over the slots selected by `sel(nb)` of A[s] * y[nb][j], neighbour values from HBM
//   WHICH 0: out-of-chunk lower (nb < c)    1: out-of-chunk higher (c < nb < N)    2: every out-of-chunk local column
template <int NR, int KT, int WHICH>
__device__ __forceinline__ void gather_remote(const RowView& rv, int K, int c, int tile, int N, const double* __restrict__ y, size_t base, double (&acc)[NR]) {
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = 0.0;
    auto selected = [&](int nb) {
        if (nb == c || nb >= N || (nb >> 8) == tile) return false;
        return WHICH == 0 ? nb < c : (WHICH == 1 ? nb > c : true);
    };
    if constexpr (KT == 0) {
        for (int s = 0; s < K; ++s) {
            const int nb = rv.nb[s * RT];
            if (!selected(nb)) continue;
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
            if (!selected(nb[s])) a[s] = 0.0;
#pragma unroll
        for (int s = 0; s < KT; ++s) {
            // unselected slot, or an outflow face (A = min(F, 0) = 0: half of the faces of an upwind matrix): nothing to add.
            // Interior cells of a block have no out-of-chunk neighbour at all and issue no gather.
            if (a[s] == 0.0) continue;
            double yn[NR];
            ldv<NR>(y, base + nb[s], yn);
#pragma unroll
            for (int j = 0; j < NR; ++j) acc[j] += a[s] * yn[j];
        }
    }
}

// acc[j] = sum over in-chunk slots (LOWER: nb < c, else nb > c; BOTH: all) of A[s] * ys[nb - tileBase][j], values from shared memory
template <int NR, int KT, int WHICH>   // WHICH 0 lower, 1 higher, 2 both
__device__ __forceinline__ void gather_local(const RowView& rv, int K, int c, int tile, int N, const double* ys, double (&acc)[NR]) {
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] = 0.0;
    const int KK = KT > 0 ? KT : K;
#pragma unroll
    for (int s = 0; s < KK; ++s) {
        const int nb = rv.nb[s * RT];
        if (!in_chunk(nb, c, tile, N)) continue;
        if (WHICH == 0 && nb > c) continue;
        if (WHICH == 1 && nb < c) continue;
        const double a = rv.a[s * RT];
        const double* yn = ys + (size_t)(nb & (RT - 1)) * NR;
#pragma unroll
        for (int j = 0; j < NR; ++j) acc[j] += a * yn[j];
    }
}

// level-scheduled in-chunk substitution: after round l every cell of level <= l holds its final value in ys
//   BWD 0: forward (levels of lower neighbours), BWD 1: backward (levels of higher neighbours)
template <int NR, int KT, int BWD>
__device__ __forceinline__ void chunk_sweep(const RowView& rv, int K, int c, int tile, int N, bool valid, int myLev, int maxLev, const bool (&on)[NR], double* ys,
                                            double (&yy)[NR]) {
    for (int l = 1; l <= maxLev; ++l) {
        __syncthreads();
        if (valid && myLev == l) {
            double acc[NR];
            gather_local<NR, KT, BWD>(rv, K, c, tile, N, ys, acc);
            double* mine = ys + (size_t)threadIdx.x * NR;
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                if (on[j]) yy[j] -= rv.rd * acc[j];
                mine[j] = yy[j];
            }
        }
    }
    __syncthreads();
}

// DIR 0: forward   1: forward + backward (last colour)   2: backward (middle colours; UPD must be 0)
// FIRST: colour 0 — no out-of-chunk lower neighbour exists (forward), nothing is gathered from HBM
// UPD as in k_sweep (krylov.cuh): 0 none, 1 p = r + beta (p - omega v) / y = rD p, 2 s = r - alpha v / z = rD s / sum|s|
template <int NR, int KT, int DIR, int UPD, int FIRST>
__global__ void __launch_bounds__(RT, RK_ROW_MINB) k_bsweep(MeshView m, RowSrc rs, int c0, int c1, int nModes, KrylovShared* ks, double* __restrict__ y, SweepUpd u) {
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
            const bool valid = c >= c0 && c < c1;
            const size_t i = (size_t)md * m.NP + c;
            double yy[NR], rr[NR], vv[NR], pp[NR];
            int myLev = 0;
            if (valid) {
                if (UPD == 0) ldv<NR>(y, i, yy);
                else { ldv<NR>(u.r, i, rr); ldv<NR>(u.v, i, vv); if (UPD == 1) ldv<NR>(u.p, i, pp); }
                myLev = m.lev[c];
            }
            const int chunkLev = m.chunkLev[tile];
            mbar_wait(&full[st], (uint32_t)((it >> 1) & 1));
            const RowView rv = row_view(smemRaw + (size_t)st * stageBytes, K);
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
                    gather_remote<NR, KT, DIR == 2 ? 1 : 0>(rv, K, c, tile, m.N, y, (size_t)md * m.NP, acc);
#pragma unroll
                    for (int j = 0; j < NR; ++j)
                        if (on[j]) yy[j] -= rv.rd * acc[j];
                }
                double* mine = ys + (size_t)threadIdx.x * NR;
#pragma unroll
                for (int j = 0; j < NR; ++j) mine[j] = yy[j];
            }
            if (DIR != 2) chunk_sweep<NR, KT, 0>(rv, K, c, tile, m.N, valid, myLev & 255, chunkLev & 255, on, ys, yy);
            if (DIR != 0) chunk_sweep<NR, KT, 1>(rv, K, c, tile, m.N, valid, myLev >> 8, chunkLev >> 8, on, ys, yy);
            if (valid) stv<NR>(y, i, yy);
            __syncthreads();   // stage st and ys are free for the next tile
        }
        if constexpr (UPD == 2) block_reduce_to_partials<NR>(red, u.partials + (size_t)u.blockBase * nModes * NR, md * NR, nModes * NR);
    }
    if constexpr (UPD == 2) finalize_ctl(u.partials, u.totalBlocks, nModes * NR, u.out, u.counter, (unsigned)u.totalBlocks, u.ctlWhat, ks, nModes * NR, u.sc);
}

// colour 0: backward substitution (every out-of-chunk neighbour is a higher colour) fused with the SpMV of these cells;
// dots as in k_spmv (MODE 0: other . v;  MODE 1: v . v, v . other); shares partials / counter with the k_spmv launch over
// the other colours (blockBase / totalBlocks)
template <int NR, int KT, int MODE>
__global__ void __launch_bounds__(RT, RK_ROW_MINB) k_bspmv0(MeshView m, RowSrc rs, int c0, int c1, int nModes, KrylovShared* ks, double* __restrict__ y, double* __restrict__ v,
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
            int myLev = 0;
            if (valid) { ldv<NR>(y, i, yy); ldv<NR>(other, i, oo); myLev = m.lev[c]; }
            const int chunkLev = m.chunkLev[tile];
            mbar_wait(&full[st], (uint32_t)((it >> 1) & 1));
            const RowView rv = row_view(smemRaw + (size_t)st * stageBytes, K);
            if (valid) {
                gather_remote<NR, KT, 2>(rv, K, c, tile, m.N, y, (size_t)md * m.NP, accR);
                double* mine = ys + (size_t)threadIdx.x * NR;
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    if (on[j]) yy[j] -= rv.rd * accR[j];
                    mine[j] = yy[j];
                }
            }
            chunk_sweep<NR, KT, 1>(rv, K, c, tile, m.N, valid, myLev >> 8, chunkLev >> 8, on, ys, yy);   // ends with a barrier: ys is final
            if (valid) {
                double accL[NR], vv[NR];
                gather_local<NR, KT, 2>(rv, K, c, tile, m.N, ys, accL);
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
