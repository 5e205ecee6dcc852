// peer.cuh — processor-patch halo swaps and the Krylov scalar reductions over NVLink peer memory.
//
// One process per GPU (like one MPI rank per sub-domain in the reference: SURVEY.md §3.5).  Every rank owns a small
// "mailbox" buffer in its own HBM that the other ranks map through CUDA IPC; NVSwitch gives every GPU a direct store path
// into every peer's memory, so
//   halo swap   = ONE kernel: each rank stores its boundary-cell records straight into the neighbours' mailboxes, publishes
//                 a sequence number per neighbour (st.release.sys) and waits for the neighbours' numbers (ld.acquire.sys);
//   all-reduce  = ONE single-CTA kernel: every rank stores its <= 72 partial sums into every peer's mailbox, waits for all
//                 ranks' numbers, adds the R contributions IN RANK ORDER (bit-identical result on every rank, so that all
//                 ranks take the same convergence decisions) and runs the scalar Krylov control step that used to be a
//                 kernel of its own.
// Both replace an NCCL call (ncclSend/ncclRecv group, ncclAllReduce) whose ~20 us launch-to-completion latency dominated
// the step at 1 M cells per GPU (profiles/: 2-GPU weak scaling 71 % with NCCL).  NCCL stays the bootstrap (the IPC handles
// travel in one ncclAllReduce) and the fallback when peer mapping is unavailable.
//
// Buffers are double-buffered on the parity of the sequence number; since every swap / reduction is symmetric (a rank
// receives from everyone it sends to) a sender can be at most one operation ahead of a receiver, so two parities suffice.
// Waits are bounded in TIME (PEER_WAIT_NS of %globaltimer, not a poll count: a neighbour that is merely slow on the host side
// must not trip it): on expiry the kernel raises an error flag instead of hanging the GPU; the host reads the flag at the end
// of every solve, of every step and before every download (engine.cu: check_peer_err) and fails the call.
#pragma once
#include "krylov.cuh"

namespace rk {

constexpr int MAX_RANKS = 16;
constexpr int AR_MAX = MAX_RED;                      // doubles per all-reduce contribution
constexpr unsigned long long PEER_WAIT_NS = 20ull * 1000 * 1000 * 1000;   // a wait gives up after 20 s of %globaltimer time

struct PeerSeg { int nbrRank, h0, len, nbrH0; };     // my ghosts [h0, h0+len) face rank nbrRank, whose matching ghosts start at nbrH0

struct PeerView {
    int rank, nRanks, nSegs, H;
    // my mailbox (local HBM, written by the peers)
    unsigned long long* haloFlag;   // [2][nRanks]
    unsigned long long* arFlag;     // [2][nRanks]
    double* arData;                 // [2][nRanks][AR_MAX]
    double* haloData;               // [2][haloCap]
    long haloCap;                   // doubles per parity of MY halo mailbox
    // the peers' mailboxes (peer-mapped; entry [rank] is my own)
    unsigned long long* pHaloFlag[MAX_RANKS];
    unsigned long long* pArFlag[MAX_RANKS];
    double* pArData[MAX_RANKS];
    double* pHaloData[MAX_RANKS];
    long pHaloCap[MAX_RANKS];
    // local helpers
    const PeerSeg* segs;            // [nSegs]
    const int* segOfGhost;          // [H]
    unsigned* blockCounter;
    int* err;                       // raised when a wait expires
    unsigned long long* stat;       // [0] ns spent waiting for neighbours' halo records, [1] for the other ranks' partial sums, [2],[3] number of such waits
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// kind 0: halo, 1: all-reduce (statistics only)
__device__ __forceinline__ void peer_wait(const unsigned long long* flag, unsigned long long seq, const PeerView& pv, int kind) {
    unsigned n = 0;
    if (*(volatile int*)pv.err) return;   // a wait has already expired on this rank: do not wait again (the host fails the call)
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < seq) {
        if ((++n & 1023u) == 0 && global_ns() - t0 > PEER_WAIT_NS) { atomicExch(pv.err, 1); break; }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) { pv.stat[kind] += global_ns() - t0; pv.stat[2 + kind] += 1; }
}

// Halo swap of `rec` doubles per processor face: send[h * rec + q] (my ghost order) -> the neighbours' mailboxes (their ghost
// order); on return of the kernel my mailbox [parity] holds the records of MY ghosts, laid out like `send`.
// grid <= SM count (all CTAs co-resident: a CTA that waits for a peer can never keep one of its own rank's CTAs from storing).
__global__ void __launch_bounds__(BLOCK) k_peer_halo(PeerView pv, int rec, unsigned long long seq, const double* __restrict__ send) {
    pdl_sync();
    __shared__ bool isLast;
    const int par = (int)(seq & 1ull);
    const long total = (long)pv.H * rec;
    for (long e = (long)blockIdx.x * BLOCK + threadIdx.x; e < total; e += (long)gridDim.x * BLOCK) {
        const int hh = (int)(e / rec), q = (int)(e % rec);
        const PeerSeg sg = pv.segs[pv.segOfGhost[hh]];
        pv.pHaloData[sg.nbrRank][(long)par * pv.pHaloCap[sg.nbrRank] + (long)(sg.nbrH0 + (hh - sg.h0)) * rec + q] = send[e];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) isLast = atomicAdd(pv.blockCounter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (isLast) {   // every CTA of this rank has stored and fenced: publish
        __threadfence_system();
        if (threadIdx.x < pv.nSegs) st_release_sys(&pv.pHaloFlag[pv.segs[threadIdx.x].nbrRank][par * pv.nRanks + pv.rank], seq);
        if (threadIdx.x == 0) *pv.blockCounter = 0;
    }
    if (threadIdx.x < pv.nSegs) peer_wait(&pv.haloFlag[par * pv.nRanks + pv.segs[threadIdx.x].nbrRank], seq, pv, 0);
    __syncthreads();
}

// All-reduce (sum) of buf[0..nd) over the ranks + the scalar control step `what` (CTL_NONE: reduction only).
__global__ void __launch_bounds__(128) k_peer_allreduce_ctl(PeerView pv, double* buf, int nd, unsigned long long seq, int what, KrylovShared* ks, int nrhs,
                                                             SolveCtl sc) {
    pdl_sync();
    const int par = (int)(seq & 1ull), R = pv.nRanks;
    for (int q = threadIdx.x; q < nd; q += blockDim.x) {
        const double v = buf[q];
        for (int r = 0; r < R; ++r) pv.pArData[r][((long)par * R + pv.rank) * AR_MAX + q] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < R) {
        st_release_sys(&pv.pArFlag[threadIdx.x][par * R + pv.rank], seq);
        peer_wait(&pv.arFlag[par * R + threadIdx.x], seq, pv, 1);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nd; q += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < R; ++r) s += __ldcg(&pv.arData[((long)par * R + r) * AR_MAX + q]);   // rank order: same bits on every rank
        buf[q] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0 && *pv.err) ks->pad[0] = 1;
    if (threadIdx.x < 32) ctl_dispatch(what, ks, nrhs, buf, sc);
}


// ---------------------------------------------------------------- what the processor patches add to one preconditioned product
// Two multi-CTA kernels (round 1 had ONE single-CTA kernel here: 87 us per launch on a 5 k-face halo, 27 % of the 8-GPU step):
//
//   k_peer_put    launched as soon as x = M^-1 p is final (before the SpMV of the remaining cells): every CTA packs the
//                 records x[boundary cell] of its share of my processor faces and stores them straight into the neighbours'
//                 mailboxes (NVLink stores are posted: nothing waits here); the last CTA to finish publishes one sequence
//                 number per neighbour (st.release.sys).  The records travel while the interior SpMV runs.
//   k_peer_ghost  after the interior SpMV: every CTA waits for the neighbours' sequence numbers (normally already there),
//                 adds the ghost columns  v[c] += sum_{ghost slots} A x[ghost]  for its share of the cells that own processor
//                 faces (x[ghost] read straight from my mailbox) and the change of the fused dots (MODE 0: r0.v;  MODE 1:
//                 t.t, t.s) as per-CTA partial sums; the LAST CTA adds the partials in block order (deterministic), then does
//                 the all-reduce of [dots | (MODE 1) sum|s| of the half step] over the ranks in rank order and the scalar
//                 control step (MODE 1: half-step convergence, then omega; MODE 0: alpha).
// MODE 1 carries the half-step sums along instead of reducing them in a round of their own: a RHS that turns out to have
// converged at the half step has had z, t computed for nothing, which update_x_r ignores (state 1) — same results.
// No early exit when every RHS has converged (speculative iterations): the sequence numbers must advance by one per executed
// operation on every rank, or the two-parity mailboxes would lose their ordering guarantee.
template <int NR>
__global__ void __launch_bounds__(BLOCK) k_peer_put(PeerView pv, unsigned long long seq, int NP, const int* __restrict__ haloCell, int nModes,
                                                     const double* __restrict__ x) {
    pdl_sync();
    __shared__ bool isLast;
    const int par = (int)(seq & 1ull), rec = nModes * NR;
    for (int hh = blockIdx.x * BLOCK + threadIdx.x; hh < pv.H; hh += gridDim.x * BLOCK) {
        const PeerSeg sg = pv.segs[pv.segOfGhost[hh]];
        double* dst = pv.pHaloData[sg.nbrRank] + (long)par * pv.pHaloCap[sg.nbrRank] + (long)(sg.nbrH0 + (hh - sg.h0)) * rec;
        const int c = haloCell[hh];
        for (int md = 0; md < nModes; ++md) {
            double t[NR];
            ldv<NR>(x, (size_t)md * NP + c, t);
            stv<NR>(dst, (size_t)md, t);   // records are 16-byte aligned (rec = NR * nModes doubles, NR even)
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) isLast = atomicAdd(pv.blockCounter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (isLast) {   // every CTA of this rank has stored and fenced: publish
        __threadfence_system();
        if (threadIdx.x < pv.nSegs) st_release_sys(&pv.pHaloFlag[pv.segs[threadIdx.x].nbrRank][par * pv.nRanks + pv.rank], seq);
        if (threadIdx.x == 0) *pv.blockCounter = 0;
    }
}

template <int NR, int MODE>
__global__ void __launch_bounds__(BLOCK) k_peer_ghost(PeerView pv, unsigned long long seqHalo, unsigned long long seqAr, MeshView m, int nBcells,
                                                       const int* __restrict__ bcells, int nModes, KrylovShared* ks, const double* __restrict__ A,
                                                       double* __restrict__ v, const double* __restrict__ other, double* dots, const double* half,
                                                       double* partials, SolveCtl sc) {
    pdl_sync();
    constexpr int ND = MODE == 0 ? 1 : 2;
    const int rec = nModes * NR, nrhs = nModes * NR;
    const int parH = (int)(seqHalo & 1ull), parA = (int)(seqAr & 1ull), R = pv.nRanks;
    __shared__ bool isLast;
    __shared__ double sCorr[MAX_RED];
    // ---- the neighbours' records of this swap
    if (threadIdx.x < pv.nSegs) peer_wait(&pv.haloFlag[parH * R + pv.segs[threadIdx.x].nbrRank], seqHalo, pv, 0);
    __syncthreads();
    // ---- ghost columns of my share of the boundary cells
    const double* box = pv.haloData + (long)parH * pv.haloCap;
    for (int md = 0; md < nModes; ++md) {
        double red[ND * NR];
        bool on[NR];
#pragma unroll
        for (int j = 0; j < ND * NR; ++j) red[j] = 0.0;
#pragma unroll
        for (int j = 0; j < NR; ++j) on[j] = ks->ctl[md * NR + j].state == 0;
        for (int i0 = blockIdx.x * BLOCK + threadIdx.x; i0 < nBcells; i0 += gridDim.x * BLOCK) {
            const int c = bcells[i0];
            double acc[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j) acc[j] = 0.0;
            for (int s = 0; s < m.K; ++s) {
                const int nb = m.nbrA[ell_t(m.K, s, c)];
                if (nb < m.N) continue;
                const double a = A[ell_t(m.K, s, c)];
                const double2* yn = reinterpret_cast<const double2*>(box + (long)(nb - m.N) * rec + md * NR);
#pragma unroll
                for (int j = 0; j < NR / 2; ++j) { const double2 t = __ldcg(yn + j); acc[2 * j] += a * t.x; acc[2 * j + 1] += a * t.y; }
            }
            const size_t i = (size_t)md * m.NP + c;
            double vv[NR], oo[NR];
            ldv<NR>(v, i, vv); ldv<NR>(other, i, oo);
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                if (!on[j]) continue;
                const double vn = vv[j] + acc[j];
                if (MODE == 0) red[j] += oo[j] * acc[j];
                else { red[2 * j] += vn * vn - vv[j] * vv[j]; red[2 * j + 1] += acc[j] * oo[j]; }
                vv[j] = vn;
            }
            stv<NR>(v, i, vv);
        }
        block_reduce_to_partials<ND * NR>(red, partials, ND * md * NR, ND * nrhs);
    }
    // ---- last CTA: partials in block order, all-reduce over the ranks in rank order, scalar control
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) isLast = atomicAdd(pv.blockCounter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    const int ndDots = ND * nrhs, nd = ndDots + (MODE == 1 ? nrhs : 0);
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int q = warp; q < ndDots; q += BLOCK / 32) {
            double t = 0;
            for (unsigned b = lane; b < gridDim.x; b += 32) t += __ldcg(&partials[(size_t)b * ndDots + q]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
            if (lane == 0) sCorr[q] = t;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *pv.blockCounter = 0;
    for (int q = threadIdx.x; q < nd; q += BLOCK) {
        const double val = q < ndDots ? dots[q] + sCorr[q] : half[q - ndDots];
        for (int r = 0; r < R; ++r) pv.pArData[r][((long)parA * R + pv.rank) * AR_MAX + q] = val;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < R) {
        st_release_sys(&pv.pArFlag[threadIdx.x][parA * R + pv.rank], seqAr);
        peer_wait(&pv.arFlag[parA * R + threadIdx.x], seqAr, pv, 1);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nd; q += BLOCK) {
        double t = 0.0;
        for (int r = 0; r < R; ++r) t += __ldcg(&pv.arData[((long)parA * R + r) * AR_MAX + q]);
        dots[q] = t;   // MODE 1: [t.t, t.s per RHS | reduced half-step sums]  (dots has room for 3 nrhs values)
    }
    __syncthreads();
    if (threadIdx.x == 0 && *pv.err) ks->pad[0] = 1;
    if (threadIdx.x < 32) ctl_dispatch(MODE == 0 ? CTL_ALPHA : CTL_HALF_OMEGA, ks, nrhs, dots, sc);
}

// ---------------------------------------------------------------- fused: pack planes + halo swap + unpack into the ghost cells
// (U, theta, tau, psi before the first residual): boundary-cell values of the planes in `pl` go straight into the
// neighbours' mailboxes; after the wait my ghosts [N, N+H) of every plane are filled from my mailbox.
__global__ void __launch_bounds__(BLOCK) k_peer_halo_planes(PeerView pv, unsigned long long seq, int N, PlaneList pl, const int* __restrict__ haloCell) {
    pdl_sync();
    __shared__ bool isLast;
    const int par = (int)(seq & 1ull), rec = pl.n;
    const unsigned long long tA = global_ns();
    for (int hh = blockIdx.x * BLOCK + threadIdx.x; hh < pv.H; hh += gridDim.x * BLOCK) {
        const PeerSeg sg = pv.segs[pv.segOfGhost[hh]];
        double* dst = pv.pHaloData[sg.nbrRank] + (long)par * pv.pHaloCap[sg.nbrRank] + (long)(sg.nbrH0 + (hh - sg.h0)) * rec;
        const int c = haloCell[hh];
        for (int p = 0; p < rec; ++p) dst[p] = pl.p[p][c];
    }
    const unsigned long long tB = global_ns();
    __threadfence_system();
    __syncthreads();
    const unsigned long long tC = global_ns();
    if (threadIdx.x == 0) isLast = atomicAdd(pv.blockCounter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (isLast) {
        __threadfence_system();
        if (threadIdx.x < pv.nSegs) st_release_sys(&pv.pHaloFlag[pv.segs[threadIdx.x].nbrRank][par * pv.nRanks + pv.rank], seq);
        if (threadIdx.x == 0) *pv.blockCounter = 0;
    }
    if (threadIdx.x < pv.nSegs) peer_wait(&pv.haloFlag[par * pv.nRanks + pv.segs[threadIdx.x].nbrRank], seq, pv, 0);
    __syncthreads();
    const unsigned long long tD = global_ns();
    const double* box = pv.haloData + (long)par * pv.haloCap;
    for (int hh = blockIdx.x * BLOCK + threadIdx.x; hh < pv.H; hh += gridDim.x * BLOCK)
        for (int p = 0; p < rec; ++p) pl.p[p][N + hh] = __ldcg(box + (long)hh * rec + p);
    if (threadIdx.x == 0 && blockIdx.x == 0) {   // phase timeline of this kernel (statistics)
        pv.stat[4] += tB - tA; pv.stat[5] += tC - tB; pv.stat[6] += tD - tC; pv.stat[7] += global_ns() - tD;
    }
}

}  // namespace rk
