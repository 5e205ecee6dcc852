// solve.inl — host orchestration of the batched PBiCGStab/DILU solve, and of PBiCG/DILU (included by engine.cu).
//
// One call solves the NR valid components of `nModes` modes on the shared LDU matrix.  Iterations are
// launched in speculative batches (as many as the previous step needed) with no host synchronisation
// inside a batch: every kernel returns immediately once the device-side nActive counter is 0.

// interleaved halo records: buf[(nModes*NR)*h + md*NR + j]
template <int NR>
__global__ void k_halo_pack_il(int H, int NP, int nModes, const int* __restrict__ haloCell, const double* __restrict__ x, double* __restrict__ buf) {
    pdl_sync();
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    const int c = haloCell[h];
    for (int md = 0; md < nModes; ++md) {
        double v[NR];
        ldv<NR>(x, (size_t)md * NP + c, v);
        stv<NR>(buf, (size_t)h * nModes + md, v);
    }
}
template <int NR>
__global__ void k_halo_unpack_il(int H, int N, int NP, int nModes, const double* __restrict__ buf, double* __restrict__ x) {
    pdl_sync();
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    for (int md = 0; md < nModes; ++md) {
        double v[NR];
        ldv<NR>(buf, (size_t)h * nModes + md, v);
        stv<NR>(x, (size_t)md * NP + N + h, v);
    }
}

template <int NR>
int halo_interleaved(RheoGpu* h, int nModes, double* x) {
    if (h->H == 0) return 0;
    LAUNCH(h, (k_halo_pack_il<NR>), cdiv(h->H, BLOCK), BLOCK, h->H, h->NP, nModes, h->d_haloCell.as<int>(), x, h->d_send.as<double>());
    const double* recv;
    if (halo_sendrecv(h, nModes * NR, &recv)) return 1;
    LAUNCH(h, (k_halo_unpack_il<NR>), cdiv(h->H, BLOCK), BLOCK, h->H, h->N, h->NP, nModes, recv, x);
    return 0;
}

// persistent grid of a row-tile kernel over `cells` cells: one resident wave for its dynamic shared memory
template <class Kern> int row_grid(RheoGpu* h, Kern kern, long cells, size_t smem) {
    auto it = h->residentBlocks.find((const void*)kern);
    int blocks;
    if (it == h->residentBlocks.end()) {
        cudaFuncAttributes fa{};   // static + dynamic shared memory above 48 KB needs the opt-in (blocksweep.cuh keeps 12 KB static)
        if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) cudaGetLastError();
        if (smem + fa.sharedSizeBytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int perSm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, RT, smem) != cudaSuccess || perSm < 1) { cudaGetLastError(); perSm = 1; }
        blocks = perSm * h->nSms;
        h->residentBlocks[(const void*)kern] = blocks;
    } else blocks = it->second;
    return std::max(1, std::min(cdiv(cells, RT) + 1, blocks));   // + 1: a range that starts mid-tile touches one more tile
}

// peer-memory path: x is final once the colour-0 backward substitution has run, so its boundary records leave for the
// neighbours BEFORE the SpMV of the remaining cells and travel over NVLink while that kernel runs (k_peer_put, peer.cuh)
template <int NR>
unsigned long long peer_put(RheoGpu* h, int nModes, const double* x) {
    const unsigned long long seq = ++h->haloSeq;
    const int grid = std::max(1, std::min(cdiv(h->H, BLOCK), 2 * h->nSms));
    LAUNCH(h, (k_peer_put<NR>), grid, BLOCK, h->pv, seq, h->NP, h->d_haloCell.as<int>(), nModes, x);
    return seq;
}
// what the processor patches add after the local SpMV: ghost columns, all-reduce of the fused dots, scalar control
template <int NR, int MODE>
int ghost_and_reduce(RheoGpu* h, int nModes, unsigned long long seqHalo, double* x, double* w, const double* other, double* redDots, double* redHalf,
                     const SolveCtl& sc) {
    if (h->nRanks <= 1) return 0;
    KrylovShared* ks = h->d_ks.as<KrylovShared>();
    double* part = h->d_partials.as<double>();
    if (h->p2p) {
        const unsigned long long sa = ++h->arSeq;
        const int grid = std::max(1, std::min(cdiv(h->nBcells, BLOCK), GRID(h, (k_peer_ghost<NR, MODE>), std::max(h->nBcells, 1))));
        LAUNCH(h, (k_peer_ghost<NR, MODE>), grid, BLOCK, h->pv, seqHalo, sa, h->mv, h->nBcells, h->d_bcells.as<int>(), nModes, ks,
               h->d_Fs.as<double>(), w, other, redDots, redHalf, part, sc);
        return 0;
    }
    const int nd = (MODE == 0 ? 1 : 2) * nModes * NR;
    if (halo_interleaved<NR>(h, nModes, x)) return 1;
    if (h->nBcells) LAUNCH(h, (k_ghost<NR, MODE>), std::min(cdiv(h->nBcells, BLOCK), 4 * h->nSms), BLOCK, h->mv, h->nBcells, h->d_bcells.as<int>(), nModes, ks,
                           h->d_Fs.as<double>(), x, w, other, redDots, part, h->d_counter.as<unsigned>());
    return all_reduce_ctl(h, redDots, nd, MODE == 0 ? CTL_ALPHA : CTL_OMEGA, nModes * NR, sc);
}

// One preconditioned product of the PBiCGStab iteration:
//   WHICH 0 :  p = r + beta (p - omega v);  y = M^-1 p;  v = A y;  dots r0.v -> alpha            (x = y, w = v)
//   WHICH 1 :  s = r - alpha v (sum|s| -> half-step convergence);  z = M^-1 s;  t = A z;  dots t.t, t.s -> omega   (x = z, w = t)
// The vector update runs inside the first pass over each cell (colour 0: streaming kernel; other colours: fused into the
// forward substitution).
template <int NR, int KT, int WHICH>
int precond_spmv(RheoGpu* h, int nModes, const SolveCtl& sc) {
    KrylovShared* ks = h->d_ks.as<KrylovShared>();
    const double* rD = h->d_rD.as<double>();
    const RowSrc rs{h->d_nbrA.as<int>(), h->d_Fs.as<double>(), rD, h->d_diag.as<double>()};
    double* part = h->d_partials.as<double>();
    double* red = h->d_red.as<double>();
    unsigned* counter = h->d_counter.as<unsigned>();
    double *r = h->d_r.as<double>(), *r0 = h->d_r0.as<double>(), *p = h->d_p.as<double>(), *y = h->d_y.as<double>(), *v = h->d_v.as<double>(),
           *sv = h->d_s.as<double>(), *z = h->d_z.as<double>(), *t = h->d_t.as<double>();
    double* x = WHICH == 0 ? y : z;
    double* w = WHICH == 0 ? v : t;
    const double* other = WHICH == 0 ? r0 : sv;
    double* redDots = WHICH == 0 ? red : red + 2 * MAX_RED;   // redA / redC
    double* redHalf = red + MAX_RED;                           // redB
    const int nc = h->nColours, N = h->N, NP = h->NP;
    const bool multi = h->nRanks > 1;
    const size_t smem = 2 * row_stage_bytes(h->K);
    constexpr int UPD = WHICH == 0 ? 1 : 2;
    const int n0 = h->colourStart[1];

    // ---- vector update + forward substitution
    std::vector<int> grids(nc, 0);
    int totalBlocks = 0;
    grids[0] = WHICH == 0 ? GRID(h, (k_update_p<NR>), n0) : GRID(h, (k_make_s<NR>), n0);
    totalBlocks = grids[0];
    for (int k = 1; k < nc; ++k) {
        const int cells = h->colourStart[k + 1] - h->colourStart[k];
        grids[k] = cells > 0 ? row_grid(h, (k_sweep<NR, KT, 1, UPD>), cells, smem) : 0;
        totalBlocks += grids[k];
    }
    const int halfWhat = multi ? CTL_NONE : CTL_HALF;
    if (WHICH == 0) LAUNCH(h, (k_update_p<NR>), grids[0], BLOCK, 0, n0, NP, nModes, ks, rD, r, v, p, y);
    else LAUNCH(h, (k_make_s<NR>), grids[0], BLOCK, 0, n0, NP, nModes, ks, rD, r, v, sv, z, part, redHalf, counter, totalBlocks, halfWhat, sc);
    int base = grids[0];
    for (int k = 1; k < nc; ++k) {
        const int c0 = h->colourStart[k], c1 = h->colourStart[k + 1];
        if (c1 <= c0) continue;
        SweepUpd u{r, v, p, sv, part, redHalf, counter, base, totalBlocks, halfWhat, sc};
        LAUNCH_SM(h, (k_sweep<NR, KT, 1, UPD>), grids[k], RT, smem, h->mv, rs, c0, c1, nModes, ks, x, u);
        base += grids[k];
    }
    // (peer-memory path: the half-step sums are reduced together with the dots of this product, see k_peer_ghost)
    if (WHICH == 1 && multi && !h->p2p && all_reduce_ctl(h, redHalf, nModes * NR, CTL_HALF, nModes * NR, sc)) return 1;
    // ---- backward substitution of the middle colours
    for (int k = nc - 2; k >= 1; --k) {
        const int c0 = h->colourStart[k], c1 = h->colourStart[k + 1];
        if (c1 <= c0) continue;
        SweepUpd u{};
        LAUNCH_SM(h, (k_sweep<NR, KT, 0, 0>), row_grid(h, (k_sweep<NR, KT, 0, 0>), c1 - c0, smem), RT, smem, h->mv, rs, c0, c1, nModes, ks, x, u);
    }
    // ---- colour 0: last backward substitution fused into the SpMV; the rest: SpMV
    constexpr int MODE = WHICH;
    const int what = multi ? CTL_NONE : (MODE == 0 ? CTL_ALPHA : CTL_OMEGA);
    const int g0 = (nc >= 2) ? row_grid(h, (k_spmv<NR, KT, MODE, 1>), n0, smem) : row_grid(h, (k_spmv<NR, KT, MODE, 0>), n0, smem);
    const int g1 = (N > n0) ? row_grid(h, (k_spmv<NR, KT, MODE, 0>), N - n0, smem) : 0;
    const bool peer = multi && h->p2p;
    unsigned long long seqHalo = 0;
    if (nc >= 2) {
        LAUNCH_SM(h, (k_spmv<NR, KT, MODE, 1>), g0, RT, smem, h->mv, rs, 0, n0, nModes, ks, x, w, other, part, redDots, counter, 0, g0 + g1, what, sc);
        if (peer) seqHalo = peer_put<NR>(h, nModes, x);
        if (g1) LAUNCH_SM(h, (k_spmv<NR, KT, MODE, 0>), g1, RT, smem, h->mv, rs, n0, N, nModes, ks, x, w, other, part, redDots, counter, g0, g0 + g1, what, sc);
    } else {
        if (peer) seqHalo = peer_put<NR>(h, nModes, x);
        LAUNCH_SM(h, (k_spmv<NR, KT, MODE, 0>), g0, RT, smem, h->mv, rs, 0, N, nModes, ks, x, w, other, part, redDots, counter, 0, g0, what, sc);
    }
    return ghost_and_reduce<NR, MODE>(h, nModes, seqHalo, x, w, other, redDots, redHalf, sc);
}

// The same preconditioned product in the BLOCK ordering (host/ordering.hpp, blocksweep.cuh): chunks of 256 cells in natural
// order, coloured; colour k's chunks are independent of each other.
template <int NR, int KT, int WHICH>
int precond_spmv_blocks(RheoGpu* h, int nModes, const SolveCtl& sc) {
    KrylovShared* ks = h->d_ks.as<KrylovShared>();
    const double* rD = h->d_rD.as<double>();
    const RowSrc rs{h->d_nbrA.as<int>(), h->d_Fs.as<double>(), rD, h->d_diag.as<double>()};
    double* part = h->d_partials.as<double>();
    double* red = h->d_red.as<double>();
    unsigned* counter = h->d_counter.as<unsigned>();
    double *r = h->d_r.as<double>(), *r0 = h->d_r0.as<double>(), *p = h->d_p.as<double>(), *y = h->d_y.as<double>(), *v = h->d_v.as<double>(),
           *sv = h->d_s.as<double>(), *z = h->d_z.as<double>(), *t = h->d_t.as<double>();
    double* x = WHICH == 0 ? y : z;
    double* w = WHICH == 0 ? v : t;
    const double* other = WHICH == 0 ? r0 : sv;
    double* redDots = WHICH == 0 ? red : red + 2 * MAX_RED;
    double* redHalf = red + MAX_RED;
    const int nc = h->nColours, N = h->N;
    const bool multi = h->nRanks > 1;
    const bool peer = multi && h->p2p;
    const size_t smem = 2 * row_stage_bytes(h->K);
    constexpr int UPD = WHICH == 0 ? 1 : 2;
    constexpr int MODE = WHICH;
    const std::vector<int>& cs = h->colourStart;

    // ---- (vector update +) forward substitution, colour by colour; the last colour runs its backward substitution too
    std::vector<int> grids(nc, 0);
    int totalBlocks = 0;
    for (int k = 0; k < nc; ++k) {
        const int cells = cs[k + 1] - cs[k];
        if (cells <= 0) continue;
        if (k == nc - 1) grids[k] = k == 0 ? row_grid(h, (k_bsweep<NR, KT, 1, UPD, 1>), cells, smem) : row_grid(h, (k_bsweep<NR, KT, 1, UPD, 0>), cells, smem);
        else grids[k] = k == 0 ? row_grid(h, (k_bsweep<NR, KT, 0, UPD, 1>), cells, smem) : row_grid(h, (k_bsweep<NR, KT, 0, UPD, 0>), cells, smem);
        totalBlocks += grids[k];
    }
    const int halfWhat = multi ? CTL_NONE : CTL_HALF;
    int base = 0;
    for (int k = 0; k < nc; ++k) {
        if (!grids[k]) continue;
        SweepUpd u{r, v, p, sv, part, redHalf, counter, base, totalBlocks, halfWhat, sc};
        if (k == nc - 1) {
            if (k == 0) LAUNCH_SM(h, (k_bsweep<NR, KT, 1, UPD, 1>), grids[k], RT, smem, h->mv, rs, cs[k], cs[k + 1], nModes, ks, x, u);
            else LAUNCH_SM(h, (k_bsweep<NR, KT, 1, UPD, 0>), grids[k], RT, smem, h->mv, rs, cs[k], cs[k + 1], nModes, ks, x, u);
        } else {
            if (k == 0) LAUNCH_SM(h, (k_bsweep<NR, KT, 0, UPD, 1>), grids[k], RT, smem, h->mv, rs, cs[k], cs[k + 1], nModes, ks, x, u);
            else LAUNCH_SM(h, (k_bsweep<NR, KT, 0, UPD, 0>), grids[k], RT, smem, h->mv, rs, cs[k], cs[k + 1], nModes, ks, x, u);
        }
        base += grids[k];
    }
    if (WHICH == 1 && multi && !peer && all_reduce_ctl(h, redHalf, nModes * NR, CTL_HALF, nModes * NR, sc)) return 1;
    // ---- backward substitution of the middle colours
    for (int k = nc - 2; k >= 1; --k) {
        if (cs[k + 1] <= cs[k]) continue;
        SweepUpd u{};
        LAUNCH_SM(h, (k_bsweep<NR, KT, 2, 0, 0>), row_grid(h, (k_bsweep<NR, KT, 2, 0, 0>), cs[k + 1] - cs[k], smem), RT, smem, h->mv, rs, cs[k], cs[k + 1], nModes, ks, x, u);
    }
    // ---- colour 0: backward substitution + SpMV; the other colours: SpMV
    const int what = multi ? CTL_NONE : (MODE == 0 ? CTL_ALPHA : CTL_OMEGA);
    unsigned long long seqHalo = 0;
    if (nc >= 2) {
        const int n0 = cs[1];
        const int g0 = row_grid(h, (k_bspmv0<NR, KT, MODE>), n0, smem);
        const int g1 = (N > n0) ? row_grid(h, (k_spmv<NR, KT, MODE, 0>), N - n0, smem) : 0;
        LAUNCH_SM(h, (k_bspmv0<NR, KT, MODE>), g0, RT, smem, h->mv, rs, 0, n0, nModes, ks, x, w, other, part, redDots, counter, 0, g0 + g1, what, sc);
        if (peer) seqHalo = peer_put<NR>(h, nModes, x);
        if (g1) LAUNCH_SM(h, (k_spmv<NR, KT, MODE, 0>), g1, RT, smem, h->mv, rs, n0, N, nModes, ks, x, w, other, part, redDots, counter, g0, g0 + g1, what, sc);
    } else {
        const int g0 = row_grid(h, (k_spmv<NR, KT, MODE, 0>), N, smem);
        if (peer) seqHalo = peer_put<NR>(h, nModes, x);
        LAUNCH_SM(h, (k_spmv<NR, KT, MODE, 0>), g0, RT, smem, h->mv, rs, 0, N, nModes, ks, x, w, other, part, redDots, counter, 0, g0, what, sc);
    }
    return ghost_and_reduce<NR, MODE>(h, nModes, seqHalo, x, w, other, redDots, redHalf, sc);
}

template <int NR, int KT>
int solve_batch(RheoGpu* h, const RhsPtrs& rp, int firstMode, int nModes, int* itersOut, bool initDone) {
    const int nrhs = nModes * NR, N = h->N, NP = h->NP;
    KrylovShared* ks = h->d_ks.as<KrylovShared>();
    double* part = h->d_partials.as<double>();
    double* red = h->d_red.as<double>();
    double *redB = red + MAX_RED, *redD = red + 3 * MAX_RED;
    unsigned* counter = h->d_counter.as<unsigned>();
    double *r = h->d_r.as<double>(), *r0 = h->d_r0.as<double>(), *y = h->d_y.as<double>(),
           *sv = h->d_s.as<double>(), *z = h->d_z.as<double>(), *t = h->d_t.as<double>();
    const SolveCtl sc{h->ctl.tolerance, h->ctl.rel_tol, h->ctl.min_iter, h->ctl.max_iter};
    const double* diag = h->d_diag.as<double>();
    const double* A = h->d_Fs.as<double>();
    const bool multi = h->nRanks > 1;

    // gAverage(psi), initial residual + normFactor.  (psi = theta: its processor-patch values were swapped at the start of
    // the step and nothing has written theta since, so the ghost cells are current.)
    // gAverage(psi): the per-component sums were accumulated by k_cell_source2 while it had theta in registers
    // initDone: k_source_init (assembly3.cuh) has done all of this, fused with the per-cell source
    if (!initDone) {
        double* sumPsi = h->d_sumPsi.as<double>() + (size_t)firstMode * NR;
        if (all_reduce(h, sumPsi, nrhs)) return 1;
        LAUNCH(h, (k_krylov_init<NR, KT>), GRID(h, (k_krylov_init<NR, KT>), N), BLOCK, h->mv, nModes, rp, diag, A, sumPsi, (double)h->nGlobalCells, r, r0, part, redB, counter,
               multi ? CTL_NONE : CTL_INIT, ks, sc);
        if (multi && all_reduce_ctl(h, redB, 3 * nrhs, CTL_INIT, nrhs, sc)) return 1;
    }
    int launched = 0;
    int spec = std::max(1, h->specIters);
    for (;;) {
        for (int it = 0; it < spec; ++it) {
            if (h->blockMode) {
                if (precond_spmv_blocks<NR, KT, 0>(h, nModes, sc)) return 1;
                if (precond_spmv_blocks<NR, KT, 1>(h, nModes, sc)) return 1;
            } else {
                if (precond_spmv<NR, KT, 0>(h, nModes, sc)) return 1;
                if (precond_spmv<NR, KT, 1>(h, nModes, sc)) return 1;
            }
            LAUNCH(h, (k_update_x_r<NR>), GRID(h, (k_update_x_r<NR>), N), BLOCK, N, NP, nModes, rp, ks, y, z, sv, t, r0, r, part, redD, counter, multi ? CTL_NONE : CTL_END, sc);
            if (multi && all_reduce_ctl(h, redD, 2 * nrhs, CTL_END, nrhs, sc)) return 1;
            ++launched;
        }
        CK(cudaMemcpyAsync(h->h_ks, ks, sizeof(KrylovShared), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (h->h_ks->pad[0]) return fail("peer-memory wait expired: a neighbour rank did not reach the halo swap / reduction (peer.cuh)");
        if (h->h_ks->nActive == 0 || launched > h->ctl.max_iter + 2) break;
        spec = 1;
    }
    int iters = 0;
    for (int q = 0; q < nrhs; ++q) iters = std::max(iters, h->h_ks->ctl[q].iters);
    *itersOut = iters;
    return 0;
}

// PBiCG + DILU (pbicg.cuh): correctness-first, one rank.  Vectors: rA = r, rT = r0, pA = p, pT = y, wA = v, wT = t.
template <int NR, int KT>
int solve_batch_pbicg(RheoGpu* h, const RhsPtrs& rp, int firstMode, int nModes, int* itersOut) {
    const int nrhs = nModes * NR, N = h->N, NP = h->NP;
    KrylovShared* ks = h->d_ks.as<KrylovShared>();
    double* part = h->d_partials.as<double>();
    double* red = h->d_red.as<double>();
    double* redB = red + MAX_RED;
    unsigned* counter = h->d_counter.as<unsigned>();
    double *rA = h->d_r.as<double>(), *rT = h->d_r0.as<double>(), *pA = h->d_p.as<double>(), *pT = h->d_y.as<double>(),
           *wA = h->d_v.as<double>(), *wT = h->d_t.as<double>();
    const SolveCtl sc{h->ctl.tolerance, h->ctl.rel_tol, h->ctl.min_iter, h->ctl.max_iter};
    const double* diag = h->d_diag.as<double>();
    const double* rD = h->d_rD.as<double>();
    const double* A = h->d_Fs.as<double>();
    const double* AT = h->d_FsT.as<double>();
    const int nc = h->nColours;

    // Several ranks (EXT-OF9 lduMatrix::Amul / Tmul with processor interfaces, gSumProd / gSumMag): DILU and DILU^T stay rank-local
    // (block-Jacobi across ranks, as in OpenFOAM); each product is preceded by the halo swap of pA and pT and followed by the
    // ghost columns (k_ghost, twice: A with pA, A^T with pT — Tmul's interfaceIntCoeffs are exactly the A^T slots); every dot and
    // residual sum goes through all_reduce_ctl, which also runs the scalar control step on the summed values.
    const bool multi = h->nRanks > 1;
    double* redScratch = red + 3 * MAX_RED;   // dot correction of the transposed ghost columns: not used by anything

    // gAverage(psi), rA = b - A psi, rT = rA, normFactor, initial residual: the same kernel as PBiCGStab (r0 = r)
    double* sumPsi = h->d_sumPsi.as<double>() + (size_t)firstMode * NR;
    if (all_reduce(h, sumPsi, nrhs)) return 1;
    LAUNCH(h, (k_krylov_init<NR, KT>), GRID(h, (k_krylov_init<NR, KT>), N), BLOCK, h->mv, nModes, rp, diag, A, sumPsi, (double)h->nGlobalCells, rA, rT, part, redB, counter,
           multi ? CTL_NONE : CTL_INIT, ks, sc);
    if (multi && all_reduce_ctl(h, redB, 3 * nrhs, CTL_INIT, nrhs, sc)) return 1;
    // EXT-OF9 PBiCG::solve: the transpose residual starts from source - A^T psi
    LAUNCH(h, (k_pb_init_rT<NR>), GRID(h, (k_pb_init_rT<NR>), N), BLOCK, h->mv, nModes, rp, A, AT, rA, rT);
    int launched = 0;
    int spec = std::max(1, h->specIters);
    for (;;) {
        for (int it = 0; it < spec; ++it) {
            // wA = M^-1 rA, wT = M^-T rT: forward substitution colour by colour, backward in reverse (the last colour has no
            // higher-numbered neighbour)
            for (int k = 0; k < nc; ++k) {
                const int c0 = h->colourStart[k], c1 = h->colourStart[k + 1];
                if (c1 > c0) LAUNCH(h, (k_pb_sweep<NR, 1>), GRID(h, (k_pb_sweep<NR, 1>), c1 - c0), BLOCK, h->mv, c0, c1, nModes, ks, rD, A, AT, rA, rT, wA, wT);
            }
            for (int k = nc - 2; k >= 0; --k) {
                const int c0 = h->colourStart[k], c1 = h->colourStart[k + 1];
                if (c1 > c0) LAUNCH(h, (k_pb_sweep<NR, 0>), GRID(h, (k_pb_sweep<NR, 0>), c1 - c0), BLOCK, h->mv, c0, c1, nModes, ks, rD, A, AT, rA, rT, wA, wT);
            }
            LAUNCH(h, (k_pb_dot<NR>), GRID(h, (k_pb_dot<NR>), N), BLOCK, N, NP, nModes, ks, wA, rT, part, red, counter, multi ? CTL_NONE : CTL_PB_BETA, sc);
            if (multi && all_reduce_ctl(h, red, nrhs, CTL_PB_BETA, nrhs, sc)) return 1;
            LAUNCH(h, (k_pb_update_p<NR>), GRID(h, (k_pb_update_p<NR>), N), BLOCK, N, NP, nModes, ks, wA, wT, pA, pT);
            if (multi && (halo_interleaved<NR>(h, nModes, pA) || halo_interleaved<NR>(h, nModes, pT))) return 1;
            LAUNCH(h, (k_pb_spmv<NR>), GRID(h, (k_pb_spmv<NR>), N), BLOCK, h->mv, nModes, ks, diag, A, AT, pA, pT, wA, wT, part, red, counter, multi ? CTL_NONE : CTL_PB_ALPHA, sc);
            if (multi) {
                if (h->nBcells) {
                    const int gg = std::min(cdiv(h->nBcells, BLOCK), 4 * h->nSms);
                    LAUNCH(h, (k_ghost<NR, 0>), gg, BLOCK, h->mv, h->nBcells, h->d_bcells.as<int>(), nModes, ks, A, pA, wA, pT, red, part, counter);          // wA += A_ghost pA; wA.pT corrected
                    LAUNCH(h, (k_ghost<NR, 0>), gg, BLOCK, h->mv, h->nBcells, h->d_bcells.as<int>(), nModes, ks, AT, pT, wT, pA, redScratch, part, counter);   // wT += A^T_ghost pT
                }
                if (all_reduce_ctl(h, red, nrhs, CTL_PB_ALPHA, nrhs, sc)) return 1;
            }
            LAUNCH(h, (k_pb_update_x_r<NR>), GRID(h, (k_pb_update_x_r<NR>), N), BLOCK, N, NP, nModes, rp, ks, pA, wA, wT, rA, rT, part, red, counter, multi ? CTL_NONE : CTL_PB_END, sc);
            if (multi && all_reduce_ctl(h, red, nrhs, CTL_PB_END, nrhs, sc)) return 1;
            ++launched;
        }
        CK(cudaMemcpyAsync(h->h_ks, ks, sizeof(KrylovShared), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (h->h_ks->pad[0]) return fail("peer-memory wait expired: a neighbour rank did not reach the halo swap / reduction (peer.cuh)");
        if (h->h_ks->nActive == 0 || launched > h->ctl.max_iter + 2) break;
        spec = 1;
    }
    int iters = 0;
    for (int q = 0; q < nrhs; ++q) iters = std::max(iters, h->h_ks->ctl[q].iters);
    *itersOut = iters;
    return 0;
}
