// ordering.hpp — block ordering of lattice (blockMesh-like, tensor-product) meshes for the DILU substitutions.
//
// EXT-OF9 DILUPreconditioner is (D+L) D^-1 (D+U) with L / U split by the CELL NUMBERING; for the upwind matrix of the theta
// equation (gaussDefCmpwConvectionScheme.C:98-100: on every face either `lower` or `upper` is zero) the numbering decides
// how much of the matrix the forward substitution solves exactly.  The reference runs in blockMesh's natural order (i
// fastest); round 1's cell-wise red-black order made every substitution parallel but cost one extra Krylov iteration on
// every configuration (C2 1 -> 2, C3/C5 2 -> 3, i.e. 25-30 % of the step).  This ordering keeps the natural
// order INSIDE blocks of 32 cells (4x4x2 cells in 3-D, 8x4 in 2-D: one warp) and colours the BLOCKS:
//
//   1. lattice indices (i,j,k) of every cell from its centre coordinates; every internal face must join lattice neighbours;
//   2. cells sorted block by block (blocks in lexicographic order, natural order inside), the sequence cut into chunks of
//      exactly 32 cells (for box sizes that are multiples of the block a chunk IS a block, otherwise chunk boundaries drift
//      across clipped blocks);
//   3. greedy colouring of the chunk graph in chunk order (2 colours on a box of whole blocks), chunks of one colour are
//      pairwise non-adjacent => independent in the substitutions; the only partial chunk (N mod 32 cells) goes last;
//   4. new numbering = colour by colour, chunk by chunk.
//
// On the device one WARP owns one chunk: neighbours in other chunks are gathered from HBM (they belong to another colour and
// are final), neighbours inside the chunk are resolved in shared memory by a level-scheduled sweep synchronised with
// __syncwarp only — the per-cell levels of the in-chunk dependency graph are computed here (a 4x4x2 block has 8 levels).
// (First version of this round: 256-cell chunks = 8x8x4 blocks swept by a whole CTA with __syncthreads between the 18 + 18
// levels: right iteration counts, but the barrier-separated latency chains made the sweeps 4x slower than the memory time.)  With the oracle's sequential DILU on the same numbering the arithmetic is
// the same (tests/test_gpu_parity.py: same iteration counts), and the iteration counts are those of the reference's order
// (tools/ordering_experiment.py; profiles/r2_ordering.md).
//
// The integer contract (perm, colour starts, levels) is restated in oracle/mesh_ref.py::block_renumber and compared
// bit-exactly (tests/test_mesh_integers.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace rk_host {

constexpr int CHUNK = 32;   // one warp; the row tiles of the Krylov kernels (RT = 256, kernels.cuh) hold 8 chunks

struct BlockOrdering {
    std::vector<int> perm;          // perm[new] = old
    std::vector<int> colourStart;   // [nColours + 1], in cells; every entry but the last is a multiple of CHUNK (32)
    int nColours = 0;
    int tile[3] = {0, 0, 0};        // block shape in lattice cells
    int dims[3] = {0, 0, 0};        // lattice extents
    std::vector<int> sweepOrder;    // [nChunks] chunk positions (new cell number / CHUNK) in geometric order: super-blocks of 8x8x8 blocks in
                                    // lexicographic order, blocks in lexicographic order inside.  Kernels without an ordering constraint
                                    // (the assembly) walk the chunks in this order, so that the chunks in flight at any time are
                                    // neighbours in space and find each other's values in L2 — the numbering itself puts every
                                    // neighbour of a chunk into ANOTHER colour, i.e. (with 2 colours) half the mesh away in memory.
};

// lattice index of every cell along axis d (sorted unique centre coordinates within tol), or false if there are more than
// `cap` distinct planes (not a lattice mesh)
inline bool lattice_axis(int N, const double* C, int d, double tol, std::vector<int>& idx, int& extent) {
    const size_t cap = 8192;
    std::vector<double> u;
    auto find = [&](double x) -> int {   // index of the plane within tol of x, or -1
        auto it = std::lower_bound(u.begin(), u.end(), x - tol);
        if (it != u.end() && std::fabs(*it - x) <= tol) return (int)(it - u.begin());
        return -1;
    };
    int last = -1;
    for (int c = 0; c < N; ++c) {
        const double x = C[3 * (size_t)c + d];
        if (last >= 0 && std::fabs(u[last] - x) <= tol) continue;
        last = find(x);
        if (last < 0) {
            if (u.size() >= cap) return false;
            auto it = std::lower_bound(u.begin(), u.end(), x);
            last = (int)(it - u.begin());
            u.insert(it, x);
        }
    }
    idx.resize(N);
    last = 0;
    for (int c = 0; c < N; ++c) {
        const double x = C[3 * (size_t)c + d];
        if (!(std::fabs(u[last] - x) <= tol)) last = find(x);
        idx[c] = last;
    }
    extent = (int)u.size();
    return true;
}

// Returns false when the mesh is not a lattice mesh (the caller falls back to the cell colouring).
inline bool block_renumber(int N, int nInt, const int32_t* own, const int32_t* nei, const double* C, BlockOrdering& out) {
    if (N < 1) return false;
    std::vector<int> ijk[3];
    double ext = 0;   // planes closer than 1e-10 of the largest extent of the bounding box are one plane (rounding of the centres)
    for (int d = 0; d < 3; ++d) {
        double lo = 1e300, hi = -1e300;
        for (int c = 0; c < N; ++c) { const double x = C[3 * (size_t)c + d]; lo = std::min(lo, x); hi = std::max(hi, x); }
        ext = std::max(ext, hi - lo);
    }
    const double tol = 1e-10 * std::max(ext, 1e-300);
    for (int d = 0; d < 3; ++d)
        if (!lattice_axis(N, C, d, tol, ijk[d], out.dims[d])) return false;
    for (int f = 0; f < nInt; ++f) {
        const int o = own[f], n = nei[f];
        const int dist = std::abs(ijk[0][o] - ijk[0][n]) + std::abs(ijk[1][o] - ijk[1][n]) + std::abs(ijk[2][o] - ijk[2][n]);
        if (dist != 1) return false;
    }
    // block shape: 32 cells over the axes that have more than one layer
    int thick[3], nThick = 0;
    for (int d = 0; d < 3; ++d) { thick[d] = out.dims[d] > 1; nThick += thick[d]; }
    int* t = out.tile;
    t[0] = t[1] = t[2] = 1;
    if (nThick == 3) { t[0] = 4; t[1] = 4; t[2] = 2; }
    else if (nThick == 2) { int w = 8; for (int d = 0; d < 3; ++d) if (thick[d]) { t[d] = w; w = 4; } }
    else if (nThick == 1) { for (int d = 0; d < 3; ++d) if (thick[d]) t[d] = CHUNK; }
    const long nT[3] = {(out.dims[0] + t[0] - 1) / t[0], (out.dims[1] + t[1] - 1) / t[1], (out.dims[2] + t[2] - 1) / t[2]};
    const long nTiles = nT[0] * nT[1] * nT[2];
    if (nTiles >= (1L << 31)) return false;
    // ---- cells block by block, natural order (i fastest) inside a block: bucket by block, then order each bucket by local index
    std::vector<int> tileOf(N), local(N);
    std::vector<int> start((size_t)nTiles + 1, 0);
    for (int c = 0; c < N; ++c) {
        const int i = ijk[0][c], j = ijk[1][c], k = ijk[2][c];
        const long tid = ((long)(k / t[2]) * nT[1] + (j / t[1])) * nT[0] + (i / t[0]);
        tileOf[c] = (int)tid;
        local[c] = ((k % t[2]) * t[1] + (j % t[1])) * t[0] + (i % t[0]);
        start[tid + 1]++;
    }
    for (long q = 0; q < nTiles; ++q) start[q + 1] += start[q];
    std::vector<int> seq(N), fill(start.begin(), start.end() - 1);
    for (int c = 0; c < N; ++c) seq[fill[tileOf[c]]++] = c;
    for (long q = 0; q < nTiles; ++q)
        std::sort(seq.begin() + start[q], seq.begin() + start[q + 1], [&](int a, int b) { return local[a] < local[b] || (local[a] == local[b] && a < b); });
    // ---- chunks of CHUNK consecutive cells of that sequence; chunk graph; greedy colouring in chunk order
    const int nChunks = (N + CHUNK - 1) / CHUNK;
    std::vector<long> chunkKey(nChunks);   // geometric key of the chunk = that of the block its first cell lies in
    {
        constexpr long SB = 8;
        const long nS[2] = {(nT[0] + SB - 1) / SB, (nT[1] + SB - 1) / SB};
        for (int q = 0; q < nChunks; ++q) {
            const long tid = tileOf[seq[(size_t)q * CHUNK]];
            const long bi = tid % nT[0], bj = (tid / nT[0]) % nT[1], bk = tid / (nT[0] * nT[1]);
            chunkKey[q] = ((((bk / SB) * nS[1] + bj / SB) * nS[0] + bi / SB) * SB * SB * SB) + ((bk % SB) * SB + bj % SB) * SB + bi % SB;
        }
    }
    std::vector<int>& chunkOf = tileOf;   // reuse
    for (int p = 0; p < N; ++p) chunkOf[seq[p]] = p / CHUNK;
    std::vector<int> adjStart((size_t)nChunks + 1, 0);
    for (int f = 0; f < nInt; ++f) {
        const int a = chunkOf[own[f]], b = chunkOf[nei[f]];
        if (a != b) adjStart[std::max(a, b) + 1]++;   // only lower-numbered neighbours matter to the greedy colouring
    }
    for (int q = 0; q < nChunks; ++q) adjStart[q + 1] += adjStart[q];
    std::vector<int> adj((size_t)adjStart[nChunks]), afill(adjStart.begin(), adjStart.end() - 1);
    for (int f = 0; f < nInt; ++f) {
        const int a = chunkOf[own[f]], b = chunkOf[nei[f]];
        if (a != b) adj[afill[std::max(a, b)]++] = std::min(a, b);
    }
    std::vector<int> colour(nChunks);
    int nCol = 0;
    for (int q = 0; q < nChunks; ++q) {
        uint64_t used = 0;
        for (int e = adjStart[q]; e < adjStart[q + 1]; ++e) used |= (uint64_t)1 << colour[adj[e]];
        int col = 0;
        while (used & ((uint64_t)1 << col)) ++col;
        if (col >= 63) return false;
        colour[q] = col;
        nCol = std::max(nCol, col + 1);
    }
    // the partial chunk (if any) is the last chunk of the sequence: its colour goes last, so that every other chunk keeps a
    // 32-aligned position in the new numbering
    if (N % CHUNK != 0) {
        const int cp = colour[nChunks - 1], cl = nCol - 1;
        if (cp != cl)
            for (int q = 0; q < nChunks; ++q) colour[q] = colour[q] == cp ? cl : (colour[q] == cl ? cp : colour[q]);
    }
    // ---- new numbering: colour by colour, chunk by chunk
    out.nColours = nCol;
    out.colourStart.assign(nCol + 1, 0);
    for (int q = 0; q < nChunks; ++q) out.colourStart[colour[q] + 1] += std::min(CHUNK, N - q * CHUNK);
    for (int c = 0; c < nCol; ++c) out.colourStart[c + 1] += out.colourStart[c];
    std::vector<int> pos(out.colourStart.begin(), out.colourStart.end() - 1);
    out.perm.resize(N);
    std::vector<int> newChunk(nChunks);
    for (int q = 0; q < nChunks; ++q) {
        const int n = std::min(CHUNK, N - q * CHUNK);
        int& p = pos[colour[q]];
        for (int e = 0; e < n; ++e) out.perm[p + e] = seq[(size_t)q * CHUNK + e];
        newChunk[q] = p / CHUNK;
        p += n;
    }
    std::vector<int> byKey(nChunks);
    for (int q = 0; q < nChunks; ++q) byKey[q] = q;
    std::stable_sort(byKey.begin(), byKey.end(), [&](int a, int b) { return chunkKey[a] < chunkKey[b]; });
    out.sweepOrder.resize(nChunks);
    for (int q = 0; q < nChunks; ++q) out.sweepOrder[q] = newChunk[byKey[q]];
    return true;
}

// Levels of the in-chunk dependency graphs in the NEW numbering (nbr: slot-major neighbour table with row stride NS, >= 0
// cell / ghost, negative otherwise): fwd[c] = longest chain of lower-numbered neighbours inside c's chunk, bwd[c] the same
// over higher-numbered ones (at most 31).  lev[c] = fwd | bwd << 8; chunkLev[q] = max fwd | max bwd << 8 per chunk.
inline bool chunk_levels(int N, int NS, int K, const std::vector<int>& nbr, std::vector<uint16_t>& lev, std::vector<uint16_t>& chunkLev) {
    std::vector<int> fwd(N, 0), bwd(N, 0);
    for (int c = 0; c < N; ++c) {
        const int base = c & ~(CHUNK - 1);
        int l = 0;
        for (int s = 0; s < K; ++s) {
            const int nb = nbr[(size_t)s * NS + c];
            if (nb >= base && nb < c) l = std::max(l, fwd[nb] + 1);
        }
        fwd[c] = l;
    }
    for (int c = N - 1; c >= 0; --c) {
        const int end = std::min(N, (c & ~(CHUNK - 1)) + CHUNK);
        int l = 0;
        for (int s = 0; s < K; ++s) {
            const int nb = nbr[(size_t)s * NS + c];
            if (nb > c && nb < end) l = std::max(l, bwd[nb] + 1);
        }
        bwd[c] = l;
    }
    lev.assign(NS, 0);
    chunkLev.assign(NS / CHUNK, 0);
    for (int c = 0; c < N; ++c) {
        if (fwd[c] > 255 || bwd[c] > 255) return false;
        lev[c] = (uint16_t)(fwd[c] | (bwd[c] << 8));
        uint16_t& q = chunkLev[c / CHUNK];
        q = (uint16_t)(std::max<int>(q & 255, fwd[c]) | (std::max<int>(q >> 8, bwd[c]) << 8));
    }
    return true;
}

}  // namespace rk_host
