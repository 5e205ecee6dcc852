// engine.cu — device-resident stress step behind the C-ABI of include/rheo_gpu.h.
//
// One RheoGpu == one constitutiveEq object (or a multiMode set) on one GPU.  Layout in HBM: FP64 SoA
// planes of stride NP, cells renumbered by colour (greedy multi-colouring, then colour-major order)
// so that the DILU forward/backward substitutions of EXT-OF9 DILUPreconditioner become one fully
// parallel kernel per colour; faces re-sorted to upper-triangular order of the new numbering;
// connectivity as a slot-major ELL table; processor-patch neighbours appended as ghost cells
// [N, N+H) that the halo exchange (NCCL send/recv) fills.  See DESIGN.md.
//
// There is no CPU fallback: every compute entry point needs a CUDA device and fails otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <map>
#include <cstring>
#include <string>
#include <vector>

#include "krylov.cuh"
#include "blocksweep.cuh"
#include "ordering.hpp"
#include "pbicg.cuh"
#include "assembly.cuh"
#include "assembly3.cuh"
#include "momentum.cuh"
#include "peer.cuh"
#include "rheo_gpu.h"

using namespace rk;

namespace {

thread_local std::string g_err;
// Programmatic dependent launch (kernels.cuh: pdl_sync) lets the next kernel's CTAs become resident under the tail of the
// running one.  Measured (profiles/r2_pdl.md): +4 % on 1 M cells (27 launches of 20-170 us per step), neutral at 8-32 M, and
// -10 % on 64 M cells per GPU, where the early-resident CTAs of the dependent kernels cost more than the ramp they hide.
// So it is on for meshes up to RK_PDL_MAX_CELLS cells per rank; RHEO_PDL=0 / 1 forces it off / on.
constexpr int RK_PDL_MAX_CELLS = 4 * 1000 * 1000;
inline bool pdl_policy(int nCells) {
    const char* e = getenv("RHEO_PDL");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return true;
    return nCells <= RK_PDL_MAX_CELLS;
}
int fail(const std::string& m) { g_err = m; return 1; }

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"; \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)

inline int cdiv(long a, int b) { return (int)((a + b - 1) / b); }
inline int round_up(long a, int b) { return (int)(((a + b - 1) / b) * b); }

// ---------------------------------------------------------------- NCCL through dlopen (no link-time dependency)
typedef struct { char internal[128]; } NcclId;
struct Nccl {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) return false;
#define LD(f) f = (decltype(f))dlsym(lib, "nccl" #f); if (!f) return false;
        LD(GetUniqueId) LD(CommInitRank) LD(CommDestroy) LD(Send) LD(Recv) LD(AllReduce) LD(GroupStart) LD(GroupEnd) LD(GetErrorString)
#undef LD
        return true;
    }
} g_nccl;
constexpr int NCCL_FLOAT64 = 8, NCCL_UINT8 = 1, NCCL_SUM = 0;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int alloc(size_t b) {
        bytes = b;
        if (b == 0) { p = nullptr; return 0; }
        CK(cudaMalloc(&p, b));
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; }
    template <class T> T* as() const { return (T*)p; }
};

struct ModeDev {
    RheoModelDesc desc;
    ModelParams mp;
    DevBuf theta, thetaOld, thetaOldOld, tau, lam, R, fFene, bsrc, thetaB, tauB, gammaVals;   // thetaOldOld: backward and CrankNicolson ddt only
    DevBuf ddt0;              // CrankNicolson: the scheme's ddt0 field (6 planes)
    int ddt0TimeIndex = 0;    //                 time step at which ddt0 was last evaluated
    DevBuf corr;   // [nComp][K*NS] deferred face values received from the upwind neighbours (assembly.cuh)
    DevBuf lamCell, etaCell;   // thermo-dependent lambda / etaP per cell (rheo_gpu_upload_thermo); unallocated: the scalars of desc
};

struct HaloSeg { int nbrRank, h0, len; };

}  // namespace

struct RheoGpu {
    int device = 0;
    cudaStream_t stream = nullptr;
    RheoSchemeCtl ctl;
    Limiter lim;
    // sizes
    int N = 0, H = 0, NT = 0, NS = 0, NP = 0, K = 0, nInt = 0, nF = 0, nB = 0;
    int nComp = 6;                 // solved components
    int comps[6];                  // indices of solved components
    int nColours = 0;
    std::vector<int> colourStart;
    std::string orderingInfo;      // rheo_gpu_get_ordering
    bool blockMode = false;        // block ordering (host/ordering.hpp): colourStart holds the chunk colours' cell ranges
    DevBuf d_lev, d_chunkLev;
    long nGlobalCells = 0;
    // host-side maps
    std::vector<int> perm;         // perm[new] = old
    std::vector<int> faceOld;      // device face -> (old face+1), negative if flipped
    std::vector<int> h_nbr, h_fidx;
    std::vector<RheoPatchDesc> patches;
    std::vector<HaloSeg> segs;
    // boundary-face ranges (b0, len) whose host values are read at upload: U_b skips empty and processor patches,
    // phi skips empty patches (OpenFOAM's emptyFvPatchField has size 0: there is nothing to send)
    std::vector<std::pair<int, int>> ubRanges, phiBRanges;
    // device mesh
    DevBuf d_perm, d_faceOld, d_nbr, d_nbrA, d_fidx, d_Sf, d_w, d_C, d_V, d_rV, d_bcell, d_bkind, d_bthetaBC, d_btauBC, d_CfB;
    DevBuf d_haloCell, d_send, d_recv;
    DevBuf d_tileRec;              // per-tile mesh records streamed by k_flux_assemble (assembly.cuh) / k_flux3 (assembly3.cuh)
    int nTiles = 0;
    bool rec3 = false;             // d_tileRec holds version-3 records: k_flux3 + k_source_init run instead of k_flux_assemble + k_cell_source2 + k_krylov_init
    std::vector<int> sweepOrder;   // block ordering: chunks in geometric order (host/ordering.hpp)
    DevBuf d_tileOrder;            // rec3: the assembly's tile walk (sweepOrder padded to nTiles); RHEO_TILE_ORDER=0 walks in index order
    bool externalGradU = false;    // rheo_gpu_upload_grad_u: d_gradU holds the caller's gradient, the assembly must not overwrite it
    bool pdl = true;               // programmatic dependent launch on this handle's kernels (pdl_policy)
    int nHidden = 0;               // BMPLog: its fluidity equation is modes[0] (RHEO_MODEL_BMP_FLUIDITY), the caller's mode 0 is modes[1]
    DevBuf d_CfI;                  // [3][nInt] internal face centres, only when a patch uses linearExtrapolation with useRegression
    DevBuf d_gradUb;               // [9][nB] patch values of fvc::grad(U) (rheo_gpu_div_tau, allocated on first use)
    DevBuf d_rowsum, d_inflow;     // rec3: row sums of the matrix and inflow-slot masks (written by k_flux3 with the matrix)
    MeshView mv;
    // fields
    DevBuf d_U, d_Ub, d_phi, d_diag, d_rD, d_Fs, d_FsT, d_stage, d_tmpB;   // d_FsT: A^T coefficients, allocated when fvSolution selects PBiCG
    DevBuf d_Fell, d_gradU, d_sumPsi;   // d_sumPsi: sum of theta per (mode, solved component), written by k_cell_source2
    std::vector<ModeDev> modes;
    // Krylov
    DevBuf d_r, d_r0, d_p, d_y, d_v, d_s, d_z, d_t, d_ks, d_partials, d_red, d_counter, d_bcells;
    KrylovShared* h_ks = nullptr;  // pinned mirror of the device control block
    bool tauAssign = false;        // rheo_gpu_set_tau_assignment (alternative reading of `tau_ = ...`)
    int specIters = 1;             // Krylov iterations launched speculatively per batch (= last step's count)
    int nSms = 148;                // SM count (set at create)
    std::map<const void*, int> residentBlocks;   // kernel -> SM count x resident CTAs per SM
    std::map<std::pair<const void*, long>, int> fluxBlocks;
    int nBcells = 0;               // cells that own at least one ghost (processor) slot
    // comm
    void* comm = nullptr;
    int rank = 0, nRanks = 1;
    // NVLink peer-memory path (peer.cuh); NCCL is the bootstrap and the fallback
    bool p2p = false;
    PeerView pv{};
    DevBuf d_mailbox, d_peerSegs, d_segOfGhost, d_peerMisc;
    void* peerBase[MAX_RANKS] = {};
    unsigned long long haloSeq = 0, arSeq = 0;
    // stats
    // time levels (EXT-OF9 Time::deltaT0Value / GeometricField::nOldTimes), for the backward ddt scheme
    int nOldTimes = 0;
    double dtNow = 0, dt0 = 0;
    long launches = 0;
    std::string launchErr;         // first kernel whose launch left an error behind (diagnostics of rheo_gpu_step's message)
    long long h2dBytes = 0, d2hBytes = 0;   // host<->device bytes copied by the upload/download entry points
    int lastIters = 0;
    bool timing = false;
    bool ktiming = false;          // per-kernel CUDA-event timing (bench.py roofline pass; serialises launches)
    cudaEvent_t kev0 = nullptr, kev1 = nullptr;
    std::map<std::string, std::pair<double, long>> ktimes;
    cudaEvent_t ev[8];
    double phaseMs[7] = {0, 0, 0, 0, 0, 0, 0};
    size_t stageBytes = 0;
};

namespace {

// kernel launch with programmatic dependent launch enabled (kernels.cuh: pdl_sync)
template <class... KArgs, class... Args>
inline void launch_pdl(bool pdl, cudaStream_t stream, int grid, int block, size_t smem, void (*kern)(KArgs...), Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

#define LAUNCH(h, kern, grid, block, ...)                                      \
    do {                                                                       \
        if ((h)->ktiming) cudaEventRecord((h)->kev0, (h)->stream);             \
        launch_pdl((h)->pdl, (h)->stream, (grid), (block), 0, kern, __VA_ARGS__); \
        (h)->launches++;                                                       \
        if ((h)->launchErr.empty() && cudaPeekAtLastError() != cudaSuccess) (h)->launchErr = #kern; \
        if ((h)->ktiming) {                                                    \
            cudaEventRecord((h)->kev1, (h)->stream);                           \
            cudaEventSynchronize((h)->kev1);                                   \
            float ms_ = 0;                                                     \
            cudaEventElapsedTime(&ms_, (h)->kev0, (h)->kev1);                  \
            auto& kt_ = (h)->ktimes[#kern];                                    \
            kt_.first += ms_;                                                  \
            kt_.second += 1;                                                   \
        }                                                                      \
    } while (0)

#define LAUNCH_SM(h, kern, grid, block, smem, ...)                             \
    do {                                                                       \
        if ((h)->ktiming) cudaEventRecord((h)->kev0, (h)->stream);             \
        launch_pdl((h)->pdl, (h)->stream, (grid), (block), (smem), kern, __VA_ARGS__); \
        (h)->launches++;                                                       \
        if ((h)->launchErr.empty() && cudaPeekAtLastError() != cudaSuccess) (h)->launchErr = #kern; \
        if ((h)->ktiming) {                                                    \
            cudaEventRecord((h)->kev1, (h)->stream);                           \
            cudaEventSynchronize((h)->kev1);                                   \
            float ms_ = 0;                                                     \
            cudaEventElapsedTime(&ms_, (h)->kev0, (h)->kev1);                  \
            auto& kt_ = (h)->ktimes[#kern];                                    \
            kt_.first += ms_;                                                  \
            kt_.second += 1;                                                   \
        }                                                                      \
    } while (0)

Limiter make_limiter(int lim) {
    Limiter L{0, 1, 1, 1, 0, 0, 0, 1, 1};
    switch (lim) {   // gaussDefCmpwConvectionScheme/limiters.H:48-98
        case RHEO_LIMITER_CUBISTA: L = {1, 7. / 4., 3. / 4., 1. / 4., 0., 3. / 8., 3. / 4., 3. / 8., 3. / 4.}; break;
        case RHEO_LIMITER_MINMOD: L = {1, 1.5, .5, .5, 0., .5, .5, .5, 1.}; break;
        case RHEO_LIMITER_SMART: L = {1, 3., 3. / 4., 0., 0., 3. / 8., 1., 1. / 6., 5. / 6.}; break;
        case RHEO_LIMITER_WACEB: L = {1, 2., 3. / 4., 0., 0., 3. / 8., 1., 3. / 10., 5. / 6.}; break;
        case RHEO_LIMITER_SUPERBEE: L = {1, 0.5, 1.5, 0., 0.5, 0., 1., 1. / 2., 2. / 3.}; break;
        default: break;
    }
    return L;
}

// greedy colouring + colour-major permutation (the integer contract; restated independently in
// tests' numpy mesh reference and compared bit-exactly)
int colour_renumber(int n, int nInt, const int32_t* own, const int32_t* nei, std::vector<int>& perm, std::vector<int>& colourStart) {
    std::vector<int> start((size_t)n + 1, 0);
    for (int f = 0; f < nInt; ++f) { start[own[f] + 1]++; start[nei[f] + 1]++; }
    for (int c = 0; c < n; ++c) start[c + 1] += start[c];
    std::vector<int> adj((size_t)start[n]), fill(start.begin(), start.end() - 1);
    for (int f = 0; f < nInt; ++f) { adj[fill[own[f]]++] = nei[f]; adj[fill[nei[f]]++] = own[f]; }
    std::vector<int> colour(n);
    int nCol = 0;
    for (int c = 0; c < n; ++c) {
        uint64_t used = 0;
        for (int q = start[c]; q < start[c + 1]; ++q)
            if (adj[q] < c) used |= (uint64_t)1 << colour[adj[q]];
        int col = 0;
        while (used & ((uint64_t)1 << col)) ++col;
        if (col >= 63) return -1;
        colour[c] = col;
        nCol = std::max(nCol, col + 1);
    }
    colourStart.assign(nCol + 1, 0);
    for (int c = 0; c < n; ++c) colourStart[colour[c] + 1]++;
    for (int q = 0; q < nCol; ++q) colourStart[q + 1] += colourStart[q];
    std::vector<int> pos(colourStart.begin(), colourStart.end() - 1);
    perm.resize(n);
    for (int c = 0; c < n; ++c) perm[pos[colour[c]]++] = c;
    return nCol;
}

template <class T> int upload(DevBuf& b, const std::vector<T>& v) {
    if (b.alloc(v.size() * sizeof(T))) return 1;
    if (!v.empty()) CK(cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

int build_mesh(RheoGpu* h, const RheoMeshDesc* d, bool allowBlocks = true) {
    const int N = d->n_cells, nF = d->n_faces, nInt = d->n_internal_faces, nB = nF - nInt;
    h->segs.clear(); h->ubRanges.clear(); h->phiBRanges.clear(); h->blockMode = false;
    h->N = N; h->nF = nF; h->nInt = nInt; h->nB = nB;
    h->pdl = pdl_policy(N);
    h->patches.assign(d->patches, d->patches + d->n_patches);
    h->nComp = 0;
    for (int q = 0; q < 6; ++q) if (d->solved_components[q]) h->comps[h->nComp++] = q;
    if (h->nComp != 6 && h->nComp != 4) return fail("rheo_gpu_create: solved_components must select 6 (3-D) or 4 (2-D) components");

    // ---- renumbering: block ordering (natural order inside 256-cell blocks, blocks coloured: the reference's Krylov iteration
    // counts) for lattice meshes solved with PBiCGStab; greedy cell colouring otherwise (unstructured meshes; the device PBiCG,
    // pbicg.cuh, sweeps cell colours; RHEO_ORDERING=colour forces it)
    {
        const char* env = getenv("RHEO_ORDERING");
        const bool wantBlocks = allowBlocks && h->ctl.solver == RHEO_SOLVER_PBICGSTAB && !(env && std::string(env) == "colour");
        rk_host::BlockOrdering bo;
        if (wantBlocks && rk_host::block_renumber(N, nInt, d->owner, d->neighbour, d->C, bo)) {
            h->blockMode = true;
            h->perm = std::move(bo.perm);
            h->sweepOrder = std::move(bo.sweepOrder);
            h->colourStart = std::move(bo.colourStart);
            h->nColours = bo.nColours;
            h->orderingInfo = std::to_string(bo.tile[0]) + "x" + std::to_string(bo.tile[1]) + "x" + std::to_string(bo.tile[2]) + " blocks of a " +
                              std::to_string(bo.dims[0]) + "x" + std::to_string(bo.dims[1]) + "x" + std::to_string(bo.dims[2]) +
                              " lattice, natural order inside chunks of 256 cells, " + std::to_string(bo.nColours) + " chunk colours";
        } else {
            h->nColours = colour_renumber(N, nInt, d->owner, d->neighbour, h->perm, h->colourStart);
            if (h->nColours < 1) return fail("rheo_gpu_create: colouring failed (more than 63 colours)");
            h->orderingInfo = "cell colouring, " + std::to_string(h->nColours) + " colours";
        }
    }
    std::vector<int> iperm(N);
    for (int c = 0; c < N; ++c) iperm[h->perm[c]] = c;

    // ---- internal faces in upper-triangular order of the new numbering (bucket by new owner)
    std::vector<int> fo(nInt), fn(nInt);
    std::vector<char> flip(nInt);
    std::vector<int> cnt((size_t)N + 1, 0);
    for (int f = 0; f < nInt; ++f) {
        int o = iperm[d->owner[f]], n = iperm[d->neighbour[f]];
        flip[f] = o > n;
        if (o > n) std::swap(o, n);
        fo[f] = o; fn[f] = n;
        cnt[o + 1]++;
    }
    for (int c = 0; c < N; ++c) cnt[c + 1] += cnt[c];
    std::vector<int> order(nInt), pos(cnt.begin(), cnt.end() - 1);
    for (int f = 0; f < nInt; ++f) order[pos[fo[f]]++] = f;
    for (int c = 0; c < N; ++c) std::sort(order.begin() + cnt[c], order.begin() + cnt[c + 1], [&](int a, int b) { return fn[a] < fn[b]; });
    h->faceOld.resize(nF);
    std::vector<int> newOwn(nF), newNei(nInt);
    for (int q = 0; q < nInt; ++q) {
        const int f = order[q];
        h->faceOld[q] = flip[f] ? -(f + 1) : (f + 1);
        newOwn[q] = fo[f]; newNei[q] = fn[f];
    }
    for (int f = nInt; f < nF; ++f) { h->faceOld[f] = f + 1; newOwn[f] = iperm[d->owner[f]]; }

    // ---- ghosts: one per processor face, in patch order
    std::vector<int> ghostOfB(nB, -1), haloCell;
    std::vector<int> bkind(nB, RHEO_PATCH_EMPTY), bthetaBC(nB, RHEO_BC_EMPTY), btauBC(nB, RHEO_BC_EMPTY), bcell(nB);
    int H = 0;
    bool needCfI = false;
    for (const RheoPatchDesc& p : h->patches) {
        if (p.start < nInt || p.start + p.size > nF) return fail("rheo_gpu_create: patch range outside the boundary faces");
        if (p.type == RHEO_PATCH_PROCESSOR) h->segs.push_back({p.nbr_rank, H, p.size});
        for (int f = p.start; f < p.start + p.size; ++f) {
            const int b = f - nInt;
            bkind[b] = p.type; bthetaBC[b] = p.theta_bc; btauBC[b] = p.tau_bc;
            if (p.type == RHEO_PATCH_PROCESSOR) {
                ghostOfB[b] = H;
                haloCell.push_back(newOwn[f]);
                ++H;
            } else if (p.type != RHEO_PATCH_EMPTY) {
                if (p.theta_bc != RHEO_BC_FIXED_VALUE && p.theta_bc != RHEO_BC_ZERO_GRADIENT)
                    return fail("rheo_gpu_create: theta BC must be fixedValue or zeroGradient on physical patches");
                if (p.tau_bc != RHEO_BC_FIXED_VALUE && p.tau_bc != RHEO_BC_ZERO_GRADIENT && p.tau_bc != RHEO_BC_LINEAR_EXTRAPOLATION &&
                    p.tau_bc != RHEO_BC_LINEAR_EXTRAPOLATION_REG)
                    return fail("rheo_gpu_create: tau BC must be fixedValue, zeroGradient or linearExtrapolation on physical patches");
                if (p.tau_bc == RHEO_BC_LINEAR_EXTRAPOLATION_REG) needCfI = true;
            }
        }
    }
    for (int b = 0; b < nB; ++b) bcell[b] = newOwn[nInt + b];
    {   // merged upload ranges, in face order
        std::vector<RheoPatchDesc> ps(h->patches);
        std::sort(ps.begin(), ps.end(), [](const RheoPatchDesc& a, const RheoPatchDesc& b) { return a.start < b.start; });
        auto add = [](std::vector<std::pair<int, int>>& r, int b0, int len) {
            if (len <= 0) return;
            if (!r.empty() && r.back().first + r.back().second == b0) r.back().second += len;
            else r.push_back({b0, len});
        };
        for (const RheoPatchDesc& p : ps) {
            if (p.type == RHEO_PATCH_EMPTY) continue;
            add(h->phiBRanges, p.start - nInt, p.size);
            if (p.type != RHEO_PATCH_PROCESSOR) add(h->ubRanges, p.start - nInt, p.size);
        }
    }
    h->H = H; h->NT = N + H;
    h->NS = round_up(N, RT);   // row tiles of 256 cells (krylov.cuh); also a multiple of the assembly tile (32)
    h->NP = round_up(N + H, 32);

    // ---- ELL (slot order: internal faces by ascending new neighbour, then boundary faces in face order)
    std::vector<int> deg(N, 0);
    for (int q = 0; q < nInt; ++q) { deg[newOwn[q]]++; deg[newNei[q]]++; }
    for (int b = 0; b < nB; ++b) if (bkind[b] != RHEO_PATCH_EMPTY) deg[bcell[b]]++;
    int K = 0;
    for (int c = 0; c < N; ++c) K = std::max(K, deg[c]);
    h->K = K;
    if ((long)K * round_up(N, 32) >= (1L << 31)) return fail("rheo_gpu_create: K x cells exceeds 2^31 (32-bit slot indexing)");
    if (K > 32) return fail("rheo_gpu_create: a cell has more than 32 faces (tile kernels stage 40 B per slot and cell in shared memory)");
    const size_t ell = (size_t)K * h->NS;
    h->h_nbr.assign(ell, -1);
    h->h_fidx.assign(ell, 0);
    std::vector<int> nbrA(ell);
    std::fill(deg.begin(), deg.end(), 0);
    for (int q = 0; q < nInt; ++q) {
        const int o = newOwn[q], n = newNei[q];
        h->h_nbr[(size_t)deg[o] * h->NS + o] = n; h->h_fidx[(size_t)deg[o] * h->NS + o] = q; deg[o]++;
        h->h_nbr[(size_t)deg[n] * h->NS + n] = o; h->h_fidx[(size_t)deg[n] * h->NS + n] = ~q; deg[n]++;
    }
    for (int b = 0; b < nB; ++b) {
        if (bkind[b] == RHEO_PATCH_EMPTY) continue;
        const int c = bcell[b];
        h->h_nbr[(size_t)deg[c] * h->NS + c] = (bkind[b] == RHEO_PATCH_PROCESSOR) ? N + ghostOfB[b] : -(b + 2);
        h->h_fidx[(size_t)deg[c] * h->NS + c] = nInt + b;
        deg[c]++;
    }
    for (int s = 0; s < K; ++s)
        for (int c = 0; c < h->NS; ++c) {
            const int v = h->h_nbr[(size_t)s * h->NS + c];
            nbrA[ell_t(K, s, c)] = (v >= 0) ? v : std::min(c, N - 1);   // tile-major (kernels.cuh: ell_t)
        }

    if (h->blockMode) {   // levels of the in-chunk dependency graphs (blocksweep.cuh)
        std::vector<uint16_t> lev, chunkLev;
        if (!rk_host::chunk_levels(N, h->NS, K, h->h_nbr, lev, chunkLev)) return build_mesh(h, d, false);   // a chain longer than 255: cell colouring
        if (upload(h->d_lev, lev) || upload(h->d_chunkLev, chunkLev)) return 1;
    }
    {   // cells owning ghost slots (k_ghost)
        std::vector<int> bc;
        for (int c = 0; c < N; ++c) {
            bool g = false;
            for (int s = 0; s < K; ++s) if (h->h_nbr[(size_t)s * h->NS + c] >= N) g = true;
            if (g) bc.push_back(c);
        }
        h->nBcells = (int)bc.size();
        if (upload(h->d_bcells, bc)) return 1;
    }
    // ---- geometry in device order
    if (needCfI) {   // internal face centres: only the regression flavour of linearExtrapolation reads them (k_tau_bc_regress)
        std::vector<double> CfI(3 * (size_t)std::max(nInt, 1));
        for (int q = 0; q < nInt; ++q) {
            const int o = h->faceOld[q];
            const int f = (o > 0 ? o : -o) - 1;
            for (int e = 0; e < 3; ++e) CfI[(size_t)e * nInt + q] = d->Cf[3 * (size_t)f + e];
        }
        if (upload(h->d_CfI, CfI)) return 1;
    } else h->d_CfI.release();
    std::vector<double> Sf(3 * (size_t)nF), w(nF), C(3 * (size_t)h->NP, 0.0), V(N), rV(N), CfB(3 * (size_t)nB);
    for (int q = 0; q < nF; ++q) {
        const int o = h->faceOld[q];
        const int f = (o > 0 ? o : -o) - 1;
        const double sg = o > 0 ? 1.0 : -1.0;
        for (int e = 0; e < 3; ++e) Sf[(size_t)e * nF + q] = sg * d->Sf[3 * (size_t)f + e];
        w[q] = o > 0 ? d->weights[f] : 1.0 - d->weights[f];
    }
    for (int c = 0; c < N; ++c) {
        const int o = h->perm[c];
        for (int e = 0; e < 3; ++e) C[(size_t)e * h->NP + c] = d->C[3 * (size_t)o + e];
        V[c] = d->V[o];
        rV[c] = 1.0 / d->V[o];
    }
    for (int b = 0; b < nB; ++b) {
        for (int e = 0; e < 3; ++e) CfB[(size_t)e * nB + b] = d->Cf[3 * (size_t)(nInt + b) + e];
        if (ghostOfB[b] >= 0) {
            if (!d->nbr_C) return fail("rheo_gpu_create: processor patches need nbr_C");
            for (int e = 0; e < 3; ++e) C[(size_t)e * h->NP + N + ghostOfB[b]] = d->nbr_C[3 * (size_t)b + e];
        }
    }
    {   // tile records streamed by the assembly kernel
        // version 3 (assembly3.cuh: k_flux3, one thread per cell) for lattice-sized rows solved with PBiCGStab; version 1
        // (assembly.cuh: k_flux_assemble, run-time K) for unstructured meshes and for the device PBiCG.  RHEO_FLUX=1 forces version 1.
        const char* fenv = getenv("RHEO_FLUX");
        h->rec3 = h->ctl.solver == RHEO_SOLVER_PBICGSTAB && ((K == 6 && h->nComp == 6) || (K == 4 && h->nComp == 4)) && !(fenv && fenv[0] == '1');
        const int nTiles = h->NS / TILE;
        const size_t recBytes = h->rec3 ? tile_record3_bytes(K) : tile_record_bytes(K);
        std::vector<unsigned char> rec((size_t)nTiles * recBytes, 0);
        std::vector<int> slotOfOwner(nInt, 0);
        for (int s = 0; s < K; ++s)
            for (int c = 0; c < N; ++c) {
                const size_t e = (size_t)s * h->NS + c;
                const int fi = h->h_fidx[e];
                if (h->h_nbr[e] >= 0 && h->h_nbr[e] < N && fi >= 0) slotOfOwner[fi] = s;
            }
        for (int t = 0; t < nTiles; ++t) {
            unsigned char* base = rec.data() + (size_t)t * recBytes;
            int* rNb = (int*)base;
            int* rMeta = rNb + K * TILE;
            double* dbl = (double*)(base + (size_t)2 * K * TILE * sizeof(int));
            // version 1: S[3][K][T], W[K][T], D[3][K][T], rV[T], V[T];   version 3: G[3][K][T], D[3][K][T], G0[3][T], V[T]
            double* rS = dbl;
            double* rW = h->rec3 ? nullptr : rS + 3 * K * TILE;
            double* rD = h->rec3 ? rS + 3 * K * TILE : rW + K * TILE;
            double* rG0 = h->rec3 ? rD + 3 * K * TILE : nullptr;
            double* rRV = h->rec3 ? nullptr : rD + 3 * K * TILE;
            double* rVol = h->rec3 ? rG0 + 3 * TILE : rRV + TILE;
            for (int l = 0; l < TILE; ++l) {
                const int c = t * TILE + l;
                if (rRV) rRV[l] = c < N ? rV[c] : 0.0;
                rVol[l] = c < N ? V[c] : 0.0;
                double g0[3] = {0, 0, 0};
                for (int s = 0; s < K; ++s) {
                    const int i = s * TILE + l;
                    rNb[i] = -1; rMeta[i] = 0;
                    if (c >= N) continue;
                    const size_t e = (size_t)s * h->NS + c;
                    const int nb = h->h_nbr[e];
                    rNb[i] = nb;
                    if (nb == -1) continue;
                    const int fi = h->h_fidx[e];
                    const int f = fi >= 0 ? fi : ~fi;
                    const double sg = fi >= 0 ? 1.0 : -1.0;
                    const bool own = fi >= 0;   // == (nb > c) on internal faces: faces are in upper-triangular order
                    if (!h->rec3) {
                        for (int x = 0; x < 3; ++x) rS[x * K * TILE + i] = sg * Sf[(size_t)x * nF + f];
                        rW[i] = nb >= 0 ? w[f] : 0.0;   // patch slots: w = 0 makes the branch-free face value w (P - N) + N the patch value
                    } else {
                        // face value = a f_P + b f_N: owner / processor side  w, 1 - w;  neighbour side  1 - w, w;  patch  0, 1
                        const double wf = w[f];
                        const double bN = nb < 0 ? 1.0 : ((own || nb >= N) ? 1.0 - wf : wf);
                        const double aP = nb < 0 ? 0.0 : ((own || nb >= N) ? wf : 1.0 - wf);
                        for (int x = 0; x < 3; ++x) {
                            const double Sx = sg * Sf[(size_t)x * nF + f];
                            rS[x * K * TILE + i] = (bN * Sx) * rV[c];
                            g0[x] += aP * Sx;
                        }
                    }
                    if (nb >= 0) {
                        int rs = 0;
                        if (nb < N) {
                            if (own) {   // find our face in the neighbour's row
                                for (int s2 = 0; s2 < K; ++s2) if (h->h_nbr[(size_t)s2 * h->NS + nb] == c && h->h_fidx[(size_t)s2 * h->NS + nb] == ~f) rs = s2;
                            } else rs = slotOfOwner[f];
                        }
                        rMeta[i] = SLOT_CELL | (own ? SLOT_OWNER : 0) | (nb >= N ? SLOT_GHOST : 0) | (rs << 8);
                        for (int x = 0; x < 3; ++x) {   // d = C_N - C_P in the face's owner -> neighbour frame
                            const double Cc = C[(size_t)x * h->NP + c], Cn = C[(size_t)x * h->NP + nb];
                            rD[x * K * TILE + i] = own ? Cn - Cc : Cc - Cn;
                        }
                    } else {
                        rMeta[i] = SLOT_PATCH | (bthetaBC[-nb - 2] == RHEO_BC_ZERO_GRADIENT ? SLOT_PATCH_ZG : 0);
                    }
                }
                if (rG0 && c < N) for (int x = 0; x < 3; ++x) rG0[x * TILE + l] = g0[x] * rV[c];
            }
        }
        h->nTiles = nTiles;
        if (upload(h->d_tileRec, rec)) return 1;
        const char* oenv = getenv("RHEO_TILE_ORDER");
        if (h->rec3 && h->blockMode && !(oenv && oenv[0] == '0')) {
            std::vector<int> order(h->sweepOrder);
            for (int t = (int)order.size(); t < nTiles; ++t) order.push_back(t);   // padding tiles (no cells) last
            if ((int)order.size() != nTiles) return fail("rheo_gpu_create: tile order does not cover the tiles");
            if (upload(h->d_tileOrder, order)) return 1;
        } else h->d_tileOrder.release();
    }
    if (upload(h->d_perm, h->perm) || upload(h->d_faceOld, h->faceOld) || upload(h->d_nbr, h->h_nbr) || upload(h->d_nbrA, nbrA) ||
        upload(h->d_fidx, h->h_fidx) || upload(h->d_Sf, Sf) || upload(h->d_w, w) || upload(h->d_C, C) || upload(h->d_V, V) ||
        upload(h->d_rV, rV) || upload(h->d_bcell, bcell) || upload(h->d_bkind, bkind) || upload(h->d_bthetaBC, bthetaBC) ||
        upload(h->d_btauBC, btauBC) || upload(h->d_CfB, CfB) || upload(h->d_haloCell, haloCell))
        return 1;
    MeshView& m = h->mv;
    m.N = N; m.H = H; m.NT = h->NT; m.NS = h->NS; m.NP = h->NP; m.K = K; m.nInt = nInt; m.nF = nF; m.nB = nB;
    m.nbr = h->d_nbr.as<int>(); m.nbrA = h->d_nbrA.as<int>(); m.fidx = h->d_fidx.as<int>();
    m.Sf = h->d_Sf.as<double>(); m.w = h->d_w.as<double>(); m.C = h->d_C.as<double>(); m.V = h->d_V.as<double>(); m.rV = h->d_rV.as<double>();
    m.bcell = h->d_bcell.as<int>(); m.bkind = h->d_bkind.as<int>(); m.bthetaBC = h->d_bthetaBC.as<int>(); m.btauBC = h->d_btauBC.as<int>();
    m.CfB = h->d_CfB.as<double>();
    m.lev = h->d_lev.as<uint16_t>(); m.chunkLev = h->d_chunkLev.as<uint16_t>();
    h->nGlobalCells = N;
    return 0;
}

int zero(RheoGpu* h, DevBuf& b) {
    if (b.bytes) CK(cudaMemsetAsync(b.p, 0, b.bytes, h->stream));
    return 0;
}

int alloc_fields(RheoGpu* h, const RheoModelDesc* callerModes, int nCallerModes) {
    const size_t NP = h->NP, nB = std::max(h->nB, 1);
    const size_t d8 = sizeof(double);
    if (h->d_U.alloc(3 * NP * d8) || h->d_Ub.alloc(3 * nB * d8) || h->d_phi.alloc((size_t)std::max(h->nF, 1) * d8) ||
        h->d_diag.alloc(std::max<size_t>(NP, h->NS) * d8) || h->d_rD.alloc(std::max<size_t>(NP, h->NS) * d8) || h->d_Fs.alloc((size_t)h->K * h->NS * d8) ||
        h->d_Fell.alloc((size_t)h->K * h->NS * d8) ||
        h->d_gradU.alloc(9 * NP * d8) || h->d_tmpB.alloc(6 * nB * d8))
        return 1;
    zero(h, h->d_U); zero(h, h->d_Ub); zero(h, h->d_phi); zero(h, h->d_Fs); zero(h, h->d_diag); zero(h, h->d_rD);
    zero(h, h->d_Fell); zero(h, h->d_gradU);
    h->stageBytes = std::max<size_t>(std::max<size_t>(9 * (size_t)h->N, NP), std::max<size_t>(6 * nB, (size_t)h->nF)) * d8;   // NP: one padded plane (rheo_gpu_upload_grad_u)
    if (h->d_stage.alloc(h->stageBytes)) return 1;
    // BMPLog (BMPLog.C:142-201): the fluidity equation is solved before theta in every correct().  It runs through the same
    // assembly and solver as component xx of a padded symmTensor (the other components have a zero source and a zero initial
    // residual: no iterations), as a hidden mode IN FRONT of the caller's — do_step then meets it first.
    std::vector<RheoModelDesc> list(callerModes, callerModes + nCallerModes);
    h->nHidden = 0;
    for (int mi = 0; mi < nCallerModes; ++mi) {
        if (callerModes[mi].model == RHEO_MODEL_BMP_FLUIDITY) return fail("rheo_gpu_create: unknown constitutiveEq model");
        if (callerModes[mi].model != RHEO_MODEL_BMP_LOG) continue;
        if (nCallerModes != 1) return fail("rheo_gpu_create: BMPLog runs as a single-mode model (not inside multiMode)");
        const RheoModelDesc& q = callerModes[mi];
        if (!(q.bmp_G0 > 0 && q.bmp_Phi0 > 0 && q.bmp_PhiInf > 0)) return fail("rheo_gpu_create: BMPLog needs G0 > 0, Phi0 > 0 and PhiInf > 0");
        RheoModelDesc f = q;
        f.model = RHEO_MODEL_BMP_FLUIDITY;
        list.insert(list.begin(), f);
        h->nHidden = 1;
    }
    const RheoModelDesc* modes = list.data();
    const int nModes = (int)list.size();
    h->modes.resize(nModes);
    for (int mi = 0; mi < nModes; ++mi) {
        ModeDev& md = h->modes[mi];
        md.desc = modes[mi];
        const RheoModelDesc& q = modes[mi];
        if (q.model < RHEO_MODEL_OLDROYD_B_LOG || q.model > RHEO_MODEL_BMP_FLUIDITY) return fail("rheo_gpu_create: unknown constitutiveEq model");
        if (!(q.lambda > 0)) return fail("rheo_gpu_create: lambda must be positive");
        ModelParams& mp = md.mp;
        mp.model = q.model == RHEO_MODEL_BMP_LOG ? RHEO_MODEL_OLDROYD_B_LOG : q.model;   // theta of BMPLog: Oldroyd-BLog on per-cell rates (k_bmp_rates)
        mp.bmpK = q.bmp_k; mp.bmpPhi0 = q.bmp_Phi0; mp.bmpPhiInf = q.bmp_PhiInf;
        mp.ptt_function = q.ptt_function; mp.ml_max_iter = q.ml_max_iter;
        mp.etaP = q.etaP; mp.lambda = q.lambda; mp.alpha = q.alpha; mp.epsilon = q.epsilon; mp.zeta = q.zeta; mp.L2 = q.L2;
        mp.ml_rtol = q.ml_rtol; mp.gamma_beta = 1.0; mp.gamma_vals = nullptr;
        mp.wmK = q.wm_K; mp.wmN = q.wm_n; mp.wmA = q.wm_a;
        mp.rpLambdaR = q.rp_lambdaR; mp.rpBeta = q.rp_beta; mp.rpDelta = q.rp_delta; mp.rpChiMax = q.rp_chiMax;
        mp.xppLambdaS = q.xpp_lambdaS; mp.xppQ = q.xpp_q; mp.xppN = q.xpp_n;
        mp.sarTau0 = q.sar_tau0; mp.sarK = q.sar_k; mp.sarN = q.sar_n; mp.sarD0 = q.sar_dims[0]; mp.sarD1 = q.sar_dims[1]; mp.sarD2 = q.sar_dims[2];
        mp.sarPtt = q.sar_n == 1.0 ? q.sar_ptt : 0;   // SaramitoLog.C:161-165: no PTT function when n != 1
        if (q.model == RHEO_MODEL_SARAMITO_LOG) {
            if (!(q.sar_n > 0 && q.sar_k > 0 && q.sar_tau0 >= 0)) return fail("rheo_gpu_create: SaramitoLog needs n > 0, k > 0 and tau0 >= 0");
            if (!(q.sar_dims[0] + q.sar_dims[1] + q.sar_dims[2] > 0)) return fail("rheo_gpu_create: SaramitoLog needs at least one valid direction in dims");
            if (q.sar_ptt < 0 || q.sar_ptt > 2) return fail("The PTT function specified does not exist. Available PTT functions are: none linear exponential");
        }
        if (q.model == RHEO_MODEL_ROLIE_POLY_LOG && !(q.rp_lambdaR > 0)) return fail("rheo_gpu_create: Rolie-PolyLog needs lambdaR > 0");
        if (q.model == RHEO_MODEL_XPOMPOM_LOG && !(q.xpp_lambdaS > 0 && q.xpp_q > 0)) return fail("rheo_gpu_create: XPomPomLog needs lambdaS > 0 and q > 0");
        if (q.model == RHEO_MODEL_WM_CY_LOG && !(q.wm_a > 0)) return fail("rheo_gpu_create: WhiteMetznerCYLog needs a > 0");
        if (q.model == RHEO_MODEL_PTT_LOG && q.ptt_function == RHEO_PTT_GENERALIZED) {   // PTTLog.C:143-170
            if (q.ml_alpha <= 0 || q.ml_beta <= 0) return fail("Both alpha and beta should be positive values for the Mittag-Leffler function to converge.");
            std::vector<double> gv{std::tgamma(q.ml_beta)};
            int k = 0;
            while (k < q.ml_max_iter && gv.back() < 1e+100) { gv.push_back(std::tgamma(q.ml_alpha * k + q.ml_beta)); k++; }
            mp.ml_max_iter = k;
            mp.gamma_beta = gv[0];
            if (upload(md.gammaVals, gv)) return 1;
            mp.gamma_vals = md.gammaVals.as<double>();
        }
        if (md.theta.alloc(6 * NP * d8) || md.thetaOld.alloc(6 * NP * d8) || md.tau.alloc(6 * NP * d8) || md.lam.alloc(3 * NP * d8) ||
            md.R.alloc(9 * NP * d8) || md.fFene.alloc(NP * d8) || md.bsrc.alloc(6 * NP * d8) || md.thetaB.alloc(6 * nB * d8) || md.tauB.alloc(6 * nB * d8) ||
            md.corr.alloc((size_t)h->nComp * h->K * h->NS * d8) || md.thetaOldOld.alloc(h->ctl.ddt != RHEO_DDT_EULER ? 6 * NP * d8 : 0) ||
            md.ddt0.alloc(h->ctl.ddt == RHEO_DDT_CRANK_NICOLSON ? 6 * NP * d8 : 0))
            return 1;
        if (q.model == RHEO_MODEL_BMP_LOG) {
            if (md.lamCell.alloc(NP * d8) || md.etaCell.alloc(NP * d8)) return 1;
            LAUNCH(h, k_fill, cdiv(NP, BLOCK), BLOCK, NP, md.lamCell.as<double>(), 1.0);
            LAUNCH(h, k_fill, cdiv(NP, BLOCK), BLOCK, NP, md.etaCell.as<double>(), 1.0);
        }
        zero(h, md.corr); zero(h, md.thetaOldOld); zero(h, md.ddt0);
        zero(h, md.theta); zero(h, md.thetaOld); zero(h, md.tau); zero(h, md.fFene); zero(h, md.bsrc); zero(h, md.thetaB); zero(h, md.tauB); zero(h, md.R);
        // READ_IF_PRESENT defaults: eigVals = eigVecs = I (Oldroyd_BLog.C:76-113)
        LAUNCH(h, k_fill, cdiv(3 * NP, BLOCK), BLOCK, 3 * NP, md.lam.as<double>(), 1.0);
        for (int q9 : {0, 4, 8}) LAUNCH(h, k_fill, cdiv(NP, BLOCK), BLOCK, NP, md.R.as<double>() + (size_t)q9 * NP, 1.0);
    }
    const int nrhsMax = std::min(MAX_RHS, nModes * h->nComp);
    const size_t kv = (size_t)nrhsMax * NP * d8;
    if (h->d_r.alloc(kv) || h->d_r0.alloc(kv) || h->d_p.alloc(kv) || h->d_y.alloc(kv) || h->d_v.alloc(kv) || h->d_s.alloc(kv) ||
        h->d_z.alloc(kv) || h->d_t.alloc(kv))
        return 1;
    for (DevBuf* b : {&h->d_r, &h->d_r0, &h->d_p, &h->d_y, &h->d_v, &h->d_s, &h->d_z, &h->d_t}) zero(h, *b);
    // partial sums: one row of <= MAX_RED slots per CTA of the largest reducing grid (all of them persistent: one resident wave)
    const int nBlocks = std::max(cdiv(h->N, BLOCK), std::min(cdiv(h->N, TILE), 32 * h->nSms));
    if (h->rec3 && (h->d_rowsum.alloc(std::max<size_t>(NP, h->NS) * d8) || h->d_inflow.alloc(std::max<size_t>(NP, h->NS) * sizeof(unsigned)))) return 1;
    if (h->d_ks.alloc(sizeof(KrylovShared)) || h->d_partials.alloc((size_t)nBlocks * MAX_RED * d8) || h->d_red.alloc(4 * MAX_RED * d8) ||
        h->d_counter.alloc(sizeof(unsigned)) || h->d_sumPsi.alloc((size_t)nModes * 6 * d8))
        return 1;
    zero(h, h->d_counter); zero(h, h->d_red); zero(h, h->d_ks);
    CK(cudaHostAlloc((void**)&h->h_ks, sizeof(KrylovShared), cudaHostAllocDefault));
    const size_t hb = (size_t)std::max(h->H, 1) * MAX_RHS * d8;
    if (h->d_send.alloc(hb) || h->d_recv.alloc(hb)) return 1;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---------------------------------------------------------------- halo exchange
// records already laid out in d_send as [h][rec] (rec doubles per processor face, my ghost order) -> *recvOut points at the
// records of my ghosts (same layout): the peer-memory mailbox (peer.cuh) or d_recv (NCCL send/recv group)
int halo_sendrecv(RheoGpu* h, int rec, const double** recvOut) {
    *recvOut = h->d_recv.as<double>();
    if (h->H == 0) return 0;
    if (!h->comm) return fail("mesh has processor patches but rheo_gpu_comm_init was not called");
    if (h->p2p) {
        const unsigned long long seq = ++h->haloSeq;
        const int grid = std::max(1, std::min(cdiv((long)h->H * rec, BLOCK), h->nSms));
        LAUNCH(h, k_peer_halo, grid, BLOCK, h->pv, rec, seq, h->d_send.as<double>());
        *recvOut = h->pv.haloData + (seq & 1ull) * h->pv.haloCap;
        return 0;
    }
    g_nccl.GroupStart();
    for (const HaloSeg& s : h->segs) {
        const size_t off = (size_t)rec * s.h0, cnt = (size_t)rec * s.len;
        int rc = g_nccl.Send(h->d_send.as<double>() + off, cnt, NCCL_FLOAT64, s.nbrRank, h->comm, h->stream);
        if (!rc) rc = g_nccl.Recv(h->d_recv.as<double>() + off, cnt, NCCL_FLOAT64, s.nbrRank, h->comm, h->stream);
        if (rc) { g_nccl.GroupEnd(); return fail(std::string("ncclSend/Recv: ") + g_nccl.GetErrorString(rc)); }
    }
    int rc = g_nccl.GroupEnd();
    if (rc) return fail(std::string("ncclGroupEnd: ") + g_nccl.GetErrorString(rc));
    return 0;
}
// halo exchange of a set of planes (pack: buf[h * n + p])
int halo_exchange(RheoGpu* h, const PlaneList& pl) {
    if (h->H == 0) return 0;
    if (h->p2p) {   // pack + swap + unpack in one kernel over peer memory
        const int grid = std::max(1, std::min(cdiv(h->H, BLOCK), h->nSms));
        LAUNCH(h, k_peer_halo_planes, grid, BLOCK, h->pv, ++h->haloSeq, h->N, pl, h->d_haloCell.as<int>());
        return 0;
    }
    LAUNCH(h, k_halo_pack, cdiv(h->H, BLOCK), BLOCK, h->H, pl, h->d_haloCell.as<int>(), h->d_send.as<double>());
    const double* recv;
    if (halo_sendrecv(h, pl.n, &recv)) return 1;
    LAUNCH(h, k_halo_unpack, cdiv(h->H, BLOCK), BLOCK, h->H, h->N, pl, recv);
    return 0;
}
int halo_planes(RheoGpu* h, double* base, int nPlanes) {
    for (int p0 = 0; p0 < nPlanes; p0 += MAX_RHS) {
        PlaneList pl;
        pl.n = std::min(MAX_RHS, nPlanes - p0);
        for (int p = 0; p < pl.n; ++p) pl.p[p] = base + (size_t)(p0 + p) * h->NP;
        if (halo_exchange(h, pl)) return 1;
    }
    return 0;
}
// sum over the ranks of buf[0..n), then the scalar control step `what` (CTL_NONE: none).  One kernel on the peer-memory
// path; ncclAllReduce + k_ctl otherwise.  Single rank: only the control step (when the producing kernel did not run it).
int all_reduce_ctl(RheoGpu* h, double* buf, int n, int what, int nrhs, const SolveCtl& sc) {
    KrylovShared* ks = h->d_ks.as<KrylovShared>();
    if (h->nRanks <= 1) return 0;
    if (h->p2p) {
        LAUNCH(h, k_peer_allreduce_ctl, 1, 128, h->pv, buf, n, ++h->arSeq, what, ks, nrhs, sc);
        return 0;
    }
    int rc = g_nccl.AllReduce(buf, buf, (size_t)n, NCCL_FLOAT64, NCCL_SUM, h->comm, h->stream);
    if (rc) return fail(std::string("ncclAllReduce: ") + g_nccl.GetErrorString(rc));
    if (what != CTL_NONE) LAUNCH(h, k_ctl, 1, 32, what, ks, nrhs, buf, sc);
    return 0;
}
int all_reduce(RheoGpu* h, double* buf, int n) { return all_reduce_ctl(h, buf, n, CTL_NONE, 0, SolveCtl{}); }

// ---------------------------------------------------------------- peer-memory set-up (collective; called by rheo_gpu_comm_init)
struct PeerRec {   // what every rank publishes
    cudaIpcMemHandle_t handle;
    int ok, H;
    int h0For[MAX_RANKS];   // first ghost of my segment facing rank j, -1 if none
};
int setup_peer(RheoGpu* h) {
    const int R = h->nRanks, me = h->rank;
    const char* env = getenv("RHEO_P2P");
    const bool wanted = !(env && env[0] == '0') && R > 1 && R <= MAX_RANKS;
    // mailbox: flags | all-reduce slots | halo slots (2 parities each)
    const size_t flagBytes = (size_t)2 * R * sizeof(unsigned long long);
    const size_t offArFlag = 256 * ((flagBytes + 255) / 256), offArData = 2 * offArFlag;
    const size_t arBytes = (size_t)2 * R * AR_MAX * sizeof(double);
    const size_t offHalo = offArData + 256 * ((arBytes + 255) / 256);
    const long haloCap = (long)std::max(h->H, 1) * MAX_RHS;
    const size_t total = offHalo + (size_t)2 * haloCap * sizeof(double);
    PeerRec mine{};
    mine.ok = 0; mine.H = h->H;
    for (int j = 0; j < MAX_RANKS; ++j) mine.h0For[j] = -1;
    if (wanted) {
        bool dup = false;
        for (const HaloSeg& sg : h->segs) {
            if (sg.nbrRank < 0 || sg.nbrRank >= R || mine.h0For[sg.nbrRank] >= 0) dup = true;   // two patches towards the same rank: keep NCCL
            else mine.h0For[sg.nbrRank] = sg.h0;
        }
        if (!dup && h->d_mailbox.alloc(total) == 0 && cudaMemsetAsync(h->d_mailbox.p, 0, total, h->stream) == cudaSuccess &&
            cudaIpcGetMemHandle(&mine.handle, h->d_mailbox.p) == cudaSuccess)
            mine.ok = 1;
        cudaGetLastError();
    }
    // gather the records: sum of byte arrays in which only the owner's slot is non-zero
    std::vector<PeerRec> all(R);
    DevBuf g;
    if (g.alloc(sizeof(PeerRec) * R)) return 1;
    CK(cudaMemsetAsync(g.p, 0, g.bytes, h->stream));
    CK(cudaMemcpyAsync((char*)g.p + sizeof(PeerRec) * me, &mine, sizeof(PeerRec), cudaMemcpyHostToDevice, h->stream));
    int rc = g_nccl.AllReduce(g.p, g.p, g.bytes, NCCL_UINT8, NCCL_SUM, h->comm, h->stream);
    if (rc) return fail(std::string("ncclAllReduce (peer records): ") + g_nccl.GetErrorString(rc));
    CK(cudaMemcpyAsync(all.data(), g.p, g.bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    g.release();
    bool ok = wanted;
    for (int r = 0; r < R; ++r) ok = ok && all[r].ok;
    // map the peers
    double opened = 0.0;
    if (ok) {
        opened = 1.0;
        for (int r = 0; r < R && opened > 0; ++r) {
            if (r == me) { h->peerBase[r] = h->d_mailbox.p; continue; }
            if (cudaIpcOpenMemHandle(&h->peerBase[r], all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); h->peerBase[r] = nullptr; opened = 0.0; }
        }
    }
    // consensus: every rank must have mapped every peer
    double* tmp = h->d_red.as<double>();
    CK(cudaMemcpyAsync(tmp, &opened, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    rc = g_nccl.AllReduce(tmp, tmp, 1, NCCL_FLOAT64, NCCL_SUM, h->comm, h->stream);
    if (rc) return fail(std::string("ncclAllReduce (peer consensus): ") + g_nccl.GetErrorString(rc));
    double sum = 0;
    CK(cudaMemcpyAsync(&sum, tmp, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (!(sum > R - 0.5)) {
        for (int r = 0; r < R; ++r) if (r != me && h->peerBase[r]) { cudaIpcCloseMemHandle(h->peerBase[r]); h->peerBase[r] = nullptr; }
        h->p2p = false;
        return 0;
    }
    // device-side tables
    std::vector<PeerSeg> ps;
    std::vector<int> segOf(std::max(h->H, 1), 0);
    for (size_t q = 0; q < h->segs.size(); ++q) {
        const HaloSeg& sg = h->segs[q];
        const int nbrH0 = all[sg.nbrRank].h0For[me];
        if (nbrH0 < 0) return fail("rheo_gpu_comm_init: processor patch without a matching patch on the neighbour rank");
        ps.push_back({sg.nbrRank, sg.h0, sg.len, nbrH0});
        for (int i = 0; i < sg.len; ++i) segOf[sg.h0 + i] = (int)q;
    }
    if (ps.empty()) ps.push_back({me, 0, 0, 0});
    if (upload(h->d_peerSegs, ps) || upload(h->d_segOfGhost, segOf) || h->d_peerMisc.alloc(256)) return 1;
    CK(cudaMemsetAsync(h->d_peerMisc.p, 0, 256, h->stream));
    PeerView& pv = h->pv;
    pv.rank = me; pv.nRanks = R; pv.nSegs = (int)h->segs.size(); pv.H = h->H;
    auto view = [&](void* base, unsigned long long** hf, unsigned long long** af, double** ad, double** hd) {
        char* b = (char*)base;
        *hf = (unsigned long long*)b; *af = (unsigned long long*)(b + offArFlag); *ad = (double*)(b + offArData); *hd = (double*)(b + offHalo);
    };
    view(h->d_mailbox.p, &pv.haloFlag, &pv.arFlag, &pv.arData, &pv.haloData);
    pv.haloCap = haloCap;
    for (int r = 0; r < R; ++r) {
        view(h->peerBase[r], &pv.pHaloFlag[r], &pv.pArFlag[r], &pv.pArData[r], &pv.pHaloData[r]);
        pv.pHaloCap[r] = (long)std::max(all[r].H, 1) * MAX_RHS;
    }
    pv.segs = h->d_peerSegs.as<PeerSeg>(); pv.segOfGhost = h->d_segOfGhost.as<int>();
    pv.blockCounter = h->d_peerMisc.as<unsigned>(); pv.err = (int*)((char*)h->d_peerMisc.p + 128); pv.stat = (unsigned long long*)((char*)h->d_peerMisc.p + 192);
    CK(cudaStreamSynchronize(h->stream));
    h->p2p = true;
    return 0;
}

// A bounded peer-memory wait that expired leaves pv.err set (peer.cuh): halo data or reduced sums may then have been read
// without synchronisation.  Called after the stream has been synchronised, by every entry point that hands results to the host.
int check_peer_err(RheoGpu* h) {
    if (!h->p2p) return 0;
    int e = 0;
    CK(cudaMemcpy(&e, h->pv.err, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) return fail("peer-memory wait expired: a neighbour rank did not reach a halo swap / reduction within 20 s (peer.cuh); results of this handle are invalid");
    return 0;
}

// persistent-style grids: exactly one wave of resident CTAs (SM count x occupancy of that kernel), each
// striding over the cells — no partial second wave, and the block reductions are paid once per CTA.
template <class Kern> int resident_grid(RheoGpu* h, Kern kern, long n) {
    auto it = h->residentBlocks.find((const void*)kern);
    int blocks;
    if (it == h->residentBlocks.end()) {
        int perSm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, BLOCK, 0) != cudaSuccess || perSm < 1) { cudaGetLastError(); perSm = 2; }
        blocks = perSm * h->nSms;
        h->residentBlocks[(const void*)kern] = blocks;
    } else blocks = it->second;
    return std::max(1, std::min(cdiv(n, BLOCK), blocks));
}
#define GRID(h, kern, n) resident_grid(h, kern, n)
#define LAUNCH_K(h, kern, grid, block, ...)                                                  \
    do {                                                                                     \
        switch ((h)->K) {                                                                    \
            case 4: LAUNCH(h, (kern<4>), grid, block, __VA_ARGS__); break;                   \
            case 6: LAUNCH(h, (kern<6>), grid, block, __VA_ARGS__); break;                   \
            default: LAUNCH(h, (kern<0>), grid, block, __VA_ARGS__); break;                  \
        }                                                                                    \
    } while (0)

// persistent grid of the TMA-fed assembly kernel: one resident wave for this block size / dynamic shared memory
template <class Kern> int flux_grid(RheoGpu* h, Kern kern, int threads, size_t smem) {
    const long key = ((long)threads << 32) ^ (long)smem;
    auto it = h->fluxBlocks.find({(const void*)kern, key});
    int blocks;
    if (it == h->fluxBlocks.end()) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int perSm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, threads, smem) != cudaSuccess || perSm < 1) { cudaGetLastError(); perSm = 1; }
        blocks = perSm * h->nSms;
        h->fluxBlocks[{(const void*)kern, key}] = blocks;
    } else blocks = it->second;
    return std::max(1, std::min(h->nTiles, blocks));
}

#include "solve.inl"

int do_step(RheoGpu* h, double dt, RheoStepStats* statsOut) {
    if (!(dt > 0)) return fail("rheo_gpu_step: dt must be positive");
    if (h->ctl.ddt != RHEO_DDT_EULER && h->ctl.ddt != RHEO_DDT_BACKWARD && h->ctl.ddt != RHEO_DDT_CRANK_NICOLSON && h->ctl.ddt != RHEO_DDT_STEADY_STATE)
        return fail("rheo_gpu_step: the Euler, backward, CrankNicolson and steadyState ddt schemes are implemented");
    const bool steady = h->ctl.ddt == RHEO_DDT_STEADY_STATE;   // EXT-OF9 steadyStateDdtScheme::fvmDdt: no diagonal, no source
    if (h->ctl.ddt == RHEO_DDT_CRANK_NICOLSON && !(h->ctl.cn_psi >= 0 && h->ctl.cn_psi <= 1)) return fail("CrankNicolson coefficient should be >= 0 and <= 1");
    // EXT-OF9 backwardDdtScheme::fvmDdt: deltaT0 = great while the field has < 2 old times (first step = Euler)
    h->dtNow = dt;
    double ddtDiag = steady ? 0.0 : 1.0 / dt, c0 = 1.0, c00 = 0.0;
    const bool backward = h->ctl.ddt == RHEO_DDT_BACKWARD;
    if (backward) {
        const double deltaT0 = h->nOldTimes < 2 ? 1e15 : h->dt0;
        const double coefft = 1 + dt / (dt + deltaT0);
        c00 = dt * dt / (deltaT0 * (dt + deltaT0));
        c0 = coefft + c00;
        ddtDiag = coefft * (1.0 / dt);
    }
    // EXT-OF9 CrankNicolsonDdtScheme::fvmDdt, fresh start (see oracle.cpp): once per time index
    //   ddt0 = (coef0/deltaT0)(theta_old - theta_oldold) - offCentre(ddt0);   diag = (coef/dt) V;
    //   source = ((coef/dt) theta_old + offCentre(ddt0)) V  — the backward slots of the source kernel with c0 = coef,
    //   c00 = -dt*off and ddt0 in place of theta_oldold:  (1/dt) V (c0 theta_old - c00 ddt0)
    const bool crankNicolson = h->ctl.ddt == RHEO_DDT_CRANK_NICOLSON;
    if (crankNicolson) {
        const double psi = h->ctl.cn_psi, off = psi < 1 ? psi : 1.0;
        const int k = std::max(1, h->nOldTimes);
        for (ModeDev& md : h->modes) {
            if (k > md.ddt0TimeIndex) {
                if (k > 1) {
                    const double rDtCoef0 = (k > 2 ? 1 + psi : 1.0) / h->dt0;
                    const size_t n6 = (size_t)6 * h->NP;
                    LAUNCH(h, k_cn_ddt0, cdiv((long)n6, BLOCK), BLOCK, n6, rDtCoef0, off, md.thetaOld.as<double>(), md.thetaOldOld.as<double>(), md.ddt0.as<double>());
                }
                md.ddt0TimeIndex = k;
            }
        }
        const double coef = k > 1 ? 1 + psi : 1.0;
        c0 = coef; c00 = -dt * off;
        ddtDiag = coef / dt;
    }
    if (h->ctl.solver != RHEO_SOLVER_PBICGSTAB && h->ctl.solver != RHEO_SOLVER_PBICG) return fail("rheo_gpu_step: unknown solver (fvSolution solver PBiCGStab or PBiCG)");
    const bool pbicg = h->ctl.solver == RHEO_SOLVER_PBICG;
    if (pbicg) {
        if (!h->d_FsT.p) {
            if (h->d_FsT.alloc((size_t)h->K * h->NS * sizeof(double))) return 1;
            if (zero(h, h->d_FsT)) return 1;
        }
    }
    const int N = h->N, NP = h->NP, grid = cdiv(N, BLOCK);
    const int nModes = (int)h->modes.size();
    const double rDeltaT = steady ? 0.0 : 1.0 / dt;
    const int noConv = h->ctl.limiter == RHEO_LIMITER_NONE;
    CompList cl;
    cl.n = h->nComp;
    for (int j = 0; j < 6; ++j) cl.c[j] = j < h->nComp ? h->comps[j] : 0;
    const int tileGrid = cdiv(N, TILE);
    if (h->timing) cudaEventRecord(h->ev[0], h->stream);

    // ---- halo of U (grad U) and of theta (deferred correction / SpMV of the first residual)
    if (h->H) {   // one swap for the 3 planes of U and the 6 planes of theta of as many modes as fit one record
        PlaneList pl;
        pl.n = 0;
        for (int p = 0; p < 3; ++p) pl.p[pl.n++] = h->d_U.as<double>() + (size_t)p * NP;
        for (ModeDev& md : h->modes) {
            if (pl.n + 6 > MAX_RHS) { if (halo_exchange(h, pl)) return 1; pl.n = 0; }
            for (int p = 0; p < 6; ++p) pl.p[pl.n++] = md.theta.as<double>() + (size_t)p * NP;
        }
        if (pl.n && halo_exchange(h, pl)) return 1;
    }
    if (h->timing) cudaEventRecord(h->ev[1], h->stream);
    // ---- assembly + segregated solve, batch by batch (the matrix is shared by all modes: same phi, same dt; a batch = as many
    // modes as fit MAX_RHS right-hand sides — the Krylov vectors are sized for one batch)
    float msGrad = 0, msAsm = 0, msSolve = 0;
    h->lastIters = 0;
    const bool hrs = h->lim.hrs && !noConv;
    const int perBatch = h->nHidden ? 1 : std::max(1, MAX_RHS / h->nComp);   // BMPLog: the fluidity matrix has its own diagonal
    for (int m0 = 0; m0 < nModes; m0 += perBatch) {
        const int m1 = std::min(nModes, m0 + perBatch);
        const int nrhs = (m1 - m0) * h->nComp;
        cudaEvent_t e0 = h->ev[6], e1 = h->ev[7];
        const bool oneBatch = nModes <= perBatch;   // phase events are then only recorded, never waited for, inside the step
        if (h->timing) cudaEventRecord(e0, h->stream);
        // k_flux3 / k_flux_assemble: the upwind cell computes each deferred face value once (grad(U) and the matrix with the
        // first mode).  Processor faces: the face values of the batch travel in one message per neighbour, then k_ghost_corr
        // adds them on the receiving side.  Then the per-cell source (model term + ddt) and the first residual.
        for (int mi = m0; mi < m1; ++mi) {
            ModeDev& md = h->modes[mi];
            FluxArgs fa{};
            const bool fluidity = md.mp.model == RHEO_MODEL_BMP_FLUIDITY;
            fa.cl = cl; fa.nU = (mi == 0 && !h->externalGradU) ? 3 : 0; fa.lim = h->lim; fa.noConv = noConv;
            // BMPLog.C:158: `== - fvm::Sp(1/lambda, Phi)` puts V/lambda on the diagonal next to the ddt coefficient; PhiEqn.relax() has its own factor
            fa.rDeltaT = fluidity ? ddtDiag + 1.0 / md.mp.lambda : ddtDiag;
            fa.relax = fluidity ? md.desc.bmp_relax : h->ctl.relax;
            fa.writeMatrix = (mi == 0 || h->nHidden) ? 1 : 0;
            fa.bounded = h->ctl.bounded ? 1 : 0;
            fa.Fell = h->d_Fell.as<double>(); fa.theta = md.theta.as<double>(); fa.thetaB = md.thetaB.as<double>();
            fa.U = h->d_U.as<double>(); fa.Ub = h->d_Ub.as<double>(); fa.bsrc = md.bsrc.as<double>();
            fa.diag = h->d_diag.as<double>(); fa.rD = h->d_rD.as<double>(); fa.Fs = h->d_Fs.as<double>(); fa.FsT = pbicg ? h->d_FsT.as<double>() : nullptr;
            fa.corr = md.corr.as<double>(); fa.ghostCorr = h->d_send.as<double>(); fa.ghostStride = nrhs; fa.ghostOffset = (mi - m0) * h->nComp;
            fa.gradU = h->d_gradU.as<double>();
            const unsigned char* rec = h->d_tileRec.as<unsigned char>();
            if (h->rec3) {
                // A theta goes where the Krylov vector t will live (first written by the second product of the first iteration)
                fa.acc = h->d_t.as<double>() + (size_t)(mi - m0) * h->nComp * NP;
                fa.rowsum = h->d_rowsum.as<double>(); fa.inflow = h->d_inflow.as<unsigned>();
                fa.sumPartials = h->d_partials.as<double>(); fa.sumOut = h->d_sumPsi.as<double>() + (size_t)mi * h->nComp; fa.counter = h->d_counter.as<unsigned>();
                fa.tileOrder = h->d_tileOrder.as<int>();
                const size_t sm3 = tile_record3_bytes(h->K) + tile_flux_bytes(h->K);
                if (h->K == 6) LAUNCH_SM(h, (k_flux3<6, 6>), flux_grid(h, k_flux3<6, 6>, TILE, sm3), TILE, sm3, h->mv, fa, rec, h->nTiles);
                else LAUNCH_SM(h, (k_flux3<4, 4>), flux_grid(h, k_flux3<4, 4>, TILE, sm3), TILE, sm3, h->mv, fa, rec, h->nTiles);
                continue;
            }
            const size_t fluxSmem = 2 * (tile_record_bytes(h->K) + tile_flux_bytes(h->K));   // two stages
            const int threads = TILE * (cl.n + fa.nU);
            switch (h->K) {
                case 4: LAUNCH_SM(h, (k_flux_assemble<4>), flux_grid(h, k_flux_assemble<4>, threads, fluxSmem), threads, fluxSmem, h->mv, fa, rec, h->nTiles); break;
                case 6: LAUNCH_SM(h, (k_flux_assemble<6>), flux_grid(h, k_flux_assemble<6>, threads, fluxSmem), threads, fluxSmem, h->mv, fa, rec, h->nTiles); break;
                default: LAUNCH_SM(h, (k_flux_assemble<0>), flux_grid(h, k_flux_assemble<0>, threads, fluxSmem), threads, fluxSmem, h->mv, fa, rec, h->nTiles); break;
            }
        }
        if (h->timing && oneBatch) cudaEventRecord(h->ev[2], h->stream);   // end of the flux / matrix kernels
        if (h->H && hrs) {
            const double* recv;
            if (halo_sendrecv(h, nrhs, &recv)) return 1;
            for (int mi = m0; mi < m1; ++mi)
                if (h->nBcells) LAUNCH(h, k_ghost_corr, cdiv(h->nBcells, BLOCK), BLOCK, h->mv, h->nBcells, h->d_bcells.as<int>(), cl, h->d_Fs.as<double>(),
                                       recv, nrhs, (mi - m0) * h->nComp, h->modes[mi].bsrc.as<double>());
        }
        // gAverage(psi) of the solver's normFactor: the per-component sums of theta (k_flux3 / k_cell_source2), summed over the ranks
        double* sumPsi = h->d_sumPsi.as<double>() + (size_t)m0 * h->nComp;
        if (h->rec3 && all_reduce(h, sumPsi, nrhs)) return 1;
        const SolveCtl sc{h->ctl.tolerance, h->ctl.rel_tol, h->ctl.min_iter, h->ctl.max_iter};
        const int srcGrid = std::min(cdiv(N, SRC_BLOCK), 3 * h->nSms);
        for (int mi = m0; mi < m1; ++mi) {
            ModeDev& md = h->modes[mi];
            SourceArgs sa;
            sa.mp = md.mp; sa.rDeltaT = rDeltaT; sa.backward = (backward || crankNicolson) ? 1 : 0; sa.c0 = c0; sa.c00 = c00;
            sa.thetaOldOld = backward ? md.thetaOldOld.as<double>() : crankNicolson ? md.ddt0.as<double>() : md.thetaOld.as<double>();
            for (int q = 0; q < 6; ++q) sa.solvedIdx[q] = -1;
            for (int j = 0; j < h->nComp; ++j) sa.solvedIdx[h->comps[j]] = j;
            sa.gradU = h->d_gradU.as<double>(); sa.theta = md.theta.as<double>(); sa.thetaOld = md.thetaOld.as<double>();
            sa.lam = md.lam.as<double>(); sa.R = md.R.as<double>();
            sa.bsrc = md.bsrc.as<double>(); sa.fFene = md.fFene.as<double>();
            sa.tau = md.mp.model == RHEO_MODEL_BMP_FLUIDITY ? h->modes[mi + 1].tau.as<double>() : md.tau.as<double>();   // BMPLog.C:160: tau_ && symm(L)
            sa.lamCell = md.lamCell.as<double>(); sa.etaCell = md.etaCell.as<double>();
            sa.sumPartials = h->d_partials.as<double>(); sa.sumOut = h->d_sumPsi.as<double>() + (size_t)mi * h->nComp; sa.counter = h->d_counter.as<unsigned>();
            if (h->rec3) {
                SrcInitArgs si;
                si.s = sa;
                si.corr = hrs ? md.corr.as<double>() : nullptr;
                si.acc = h->d_t.as<double>() + (size_t)(mi - m0) * h->nComp * NP;
                si.rowsum = h->d_rowsum.as<double>(); si.inflow = h->d_inflow.as<unsigned>();
                si.sumPsi = h->d_sumPsi.as<double>() + (size_t)mi * h->nComp; si.nGlobal = (double)h->nGlobalCells;
                si.r = h->d_r.as<double>() + (size_t)(mi - m0) * NP * h->nComp; si.r0 = h->d_r0.as<double>() + (size_t)(mi - m0) * NP * h->nComp;
                si.partials = h->d_partials.as<double>(); si.out = h->d_red.as<double>() + MAX_RED; si.counter = h->d_counter.as<unsigned>();
                si.slotBase = 3 * (mi - m0) * h->nComp; si.nSlots = 3 * nrhs; si.totalBlocks = (m1 - m0) * srcGrid;
                si.ctlWhat = h->nRanks > 1 ? CTL_NONE : CTL_INIT; si.nrhs = nrhs;
                si.ks = h->d_ks.as<KrylovShared>(); si.sc = sc;
#define RK_SRC3(M)                                                                                            \
    do {                                                                                                      \
        if (h->K == 6) LAUNCH(h, (k_source_init<M, 6, 6>), srcGrid, SRC_BLOCK, h->mv, si);                    \
        else LAUNCH(h, (k_source_init<M, 4, 4>), srcGrid, SRC_BLOCK, h->mv, si);                              \
    } while (0)
                switch (md.mp.model) {
                    case RHEO_MODEL_OLDROYD_B_LOG: RK_SRC3(RHEO_MODEL_OLDROYD_B_LOG); break;
                    case RHEO_MODEL_GIESEKUS_LOG: RK_SRC3(RHEO_MODEL_GIESEKUS_LOG); break;
                    case RHEO_MODEL_PTT_LOG: RK_SRC3(RHEO_MODEL_PTT_LOG); break;
                    case RHEO_MODEL_FENE_P_LOG: RK_SRC3(RHEO_MODEL_FENE_P_LOG); break;
                    case RHEO_MODEL_FENE_CR_LOG: RK_SRC3(RHEO_MODEL_FENE_CR_LOG); break;
                    case RHEO_MODEL_WM_CY_LOG: RK_SRC3(RHEO_MODEL_WM_CY_LOG); break;
                    case RHEO_MODEL_ROLIE_POLY_LOG: RK_SRC3(RHEO_MODEL_ROLIE_POLY_LOG); break;
                    case RHEO_MODEL_SARAMITO_LOG: RK_SRC3(RHEO_MODEL_SARAMITO_LOG); break;
                    case RHEO_MODEL_BMP_FLUIDITY: RK_SRC3(RHEO_MODEL_BMP_FLUIDITY); break;
                    default: RK_SRC3(RHEO_MODEL_XPOMPOM_LOG); break;
                }
#undef RK_SRC3
                continue;
            }
            switch (md.mp.model) {
                case RHEO_MODEL_OLDROYD_B_LOG: LAUNCH(h, (k_cell_source2<RHEO_MODEL_OLDROYD_B_LOG>), srcGrid, SRC_BLOCK, h->mv, sa); break;
                case RHEO_MODEL_GIESEKUS_LOG: LAUNCH(h, (k_cell_source2<RHEO_MODEL_GIESEKUS_LOG>), srcGrid, SRC_BLOCK, h->mv, sa); break;
                case RHEO_MODEL_PTT_LOG: LAUNCH(h, (k_cell_source2<RHEO_MODEL_PTT_LOG>), srcGrid, SRC_BLOCK, h->mv, sa); break;
                case RHEO_MODEL_FENE_P_LOG: LAUNCH(h, (k_cell_source2<RHEO_MODEL_FENE_P_LOG>), srcGrid, SRC_BLOCK, h->mv, sa); break;
                case RHEO_MODEL_FENE_CR_LOG: LAUNCH(h, (k_cell_source2<RHEO_MODEL_FENE_CR_LOG>), srcGrid, SRC_BLOCK, h->mv, sa); break;
                case RHEO_MODEL_WM_CY_LOG: LAUNCH(h, (k_cell_source2<RHEO_MODEL_WM_CY_LOG>), srcGrid, SRC_BLOCK, h->mv, sa); break;
                case RHEO_MODEL_ROLIE_POLY_LOG: LAUNCH(h, (k_cell_source2<RHEO_MODEL_ROLIE_POLY_LOG>), srcGrid, SRC_BLOCK, h->mv, sa); break;
                case RHEO_MODEL_SARAMITO_LOG: LAUNCH(h, (k_cell_source2<RHEO_MODEL_SARAMITO_LOG>), srcGrid, SRC_BLOCK, h->mv, sa); break;
                case RHEO_MODEL_BMP_FLUIDITY: LAUNCH(h, (k_cell_source2<RHEO_MODEL_BMP_FLUIDITY>), srcGrid, SRC_BLOCK, h->mv, sa); break;
                default: LAUNCH(h, (k_cell_source2<RHEO_MODEL_XPOMPOM_LOG>), srcGrid, SRC_BLOCK, h->mv, sa); break;
            }
        }
        if (h->rec3 && h->nRanks > 1 && all_reduce_ctl(h, h->d_red.as<double>() + MAX_RED, 3 * nrhs, CTL_INIT, nrhs, sc)) return 1;
        if (h->timing) { cudaEventRecord(e1, h->stream); if (!oneBatch) { cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); msAsm += ms; } }

        // ---- segregated solve: all valid components of the batch's modes on the shared matrix
        RhsPtrs rp;
        rp.n = 0;
        for (int mi = m0; mi < m1; ++mi)
            for (int j = 0; j < h->nComp; ++j) {
                rp.psi[rp.n] = h->modes[mi].theta.as<double>() + (size_t)h->comps[j] * NP;
                rp.b[rp.n] = h->modes[mi].bsrc.as<double>() + (size_t)h->comps[j] * NP;
                rp.corr[rp.n] = hrs ? h->modes[mi].corr.as<double>() + (size_t)j * h->K * h->NS : nullptr;
                rp.n++;
            }
        int iters = 0;
        int rc;
        const bool initDone = h->rec3;   // k_source_init has formed r, r0, the norm factor and the initial residual
        if (pbicg) {
            if (h->nComp == 6) rc = (h->K == 6) ? solve_batch_pbicg<6, 6>(h, rp, m0, m1 - m0, &iters) : solve_batch_pbicg<6, 0>(h, rp, m0, m1 - m0, &iters);
            else rc = (h->K == 4) ? solve_batch_pbicg<4, 4>(h, rp, m0, m1 - m0, &iters) : solve_batch_pbicg<4, 0>(h, rp, m0, m1 - m0, &iters);
        } else if (h->nComp == 6) rc = (h->K == 6) ? solve_batch<6, 6>(h, rp, m0, m1 - m0, &iters, initDone) : solve_batch<6, 0>(h, rp, m0, m1 - m0, &iters, false);
        else rc = (h->K == 4) ? solve_batch<4, 4>(h, rp, m0, m1 - m0, &iters, initDone) : solve_batch<4, 0>(h, rp, m0, m1 - m0, &iters, false);
        if (rc) return rc;
        h->specIters = std::max(1, iters);
        int q = 0;
        for (int mi = m0; mi < m1; ++mi) {
            if (h->modes[mi].mp.model == RHEO_MODEL_BMP_FLUIDITY) {
                // theta of the BMPLog mode sees the fluidity of AFTER PhiEqn.solve() (BMPLog.C:177-196)
                ModeDev& T = h->modes[mi + 1];
                LAUNCH(h, k_bmp_rates, grid, BLOCK, N, h->modes[mi].theta.as<double>(), T.desc.bmp_G0, T.lamCell.as<double>(), T.etaCell.as<double>());
                h->lastIters = std::max(h->lastIters, h->h_ks->ctl[0].iters);
                q += h->nComp;
                continue;
            }
            RheoStepStats* stats = statsOut ? statsOut - h->nHidden : nullptr;   // the caller's array starts at its own mode 0
            if (stats) std::memset(&stats[mi], 0, sizeof(RheoStepStats));
            for (int j = 0; j < h->nComp; ++j, ++q) {
                const KrylovCtl& k = h->h_ks->ctl[q];
                h->lastIters = std::max(h->lastIters, k.iters);
                if (stats) {
                    const int cmp = h->comps[j];
                    stats[mi].initial_residual[cmp] = k.initRes;
                    stats[mi].final_residual[cmp] = k.finRes;
                    stats[mi].n_iterations[cmp] = k.iters;
                    stats[mi].converged[cmp] = (k.finRes < h->ctl.tolerance || (h->ctl.rel_tol > 1e-20 && k.finRes < h->ctl.rel_tol * k.initRes)) ? 1 : 0;
                }
            }
            if (stats) for (int cmp = 0; cmp < 6; ++cmp) if (std::find(h->comps, h->comps + h->nComp, cmp) == h->comps + h->nComp) stats[mi].converged[cmp] = 1;
        }
        if (h->timing && !oneBatch) { cudaEventRecord(e0, h->stream); cudaEventSynchronize(e0); float ms; cudaEventElapsedTime(&ms, e1, e0); msSolve += ms; }
    }
    if (h->timing) cudaEventRecord(h->ev[3], h->stream);

    // ---- eig + exp + tau
    for (ModeDev& md : h->modes)
    {
        if (md.mp.model == RHEO_MODEL_BMP_FLUIDITY) continue;   // the fluidity has no eigen-decomposition and no stress of its own
#define RK_EIG(M) LAUNCH(h, (k_eig_tau<M>), grid, BLOCK, N, NP, md.mp, md.theta.as<double>(), md.fFene.as<double>(), md.lam.as<double>(), md.R.as<double>(), md.tau.as<double>(), \
                         md.lamCell.as<double>(), md.etaCell.as<double>())
        switch (md.mp.model) {
            case RHEO_MODEL_OLDROYD_B_LOG: RK_EIG(RHEO_MODEL_OLDROYD_B_LOG); break;
            case RHEO_MODEL_GIESEKUS_LOG: RK_EIG(RHEO_MODEL_GIESEKUS_LOG); break;
            case RHEO_MODEL_PTT_LOG: RK_EIG(RHEO_MODEL_PTT_LOG); break;
            case RHEO_MODEL_FENE_P_LOG: RK_EIG(RHEO_MODEL_FENE_P_LOG); break;
            case RHEO_MODEL_FENE_CR_LOG: RK_EIG(RHEO_MODEL_FENE_CR_LOG); break;
            case RHEO_MODEL_WM_CY_LOG: RK_EIG(RHEO_MODEL_WM_CY_LOG); break;
            case RHEO_MODEL_ROLIE_POLY_LOG: RK_EIG(RHEO_MODEL_ROLIE_POLY_LOG); break;
            case RHEO_MODEL_SARAMITO_LOG: RK_EIG(RHEO_MODEL_SARAMITO_LOG); break;
            default: RK_EIG(RHEO_MODEL_XPOMPOM_LOG); break;
        }
#undef RK_EIG
    }
    if (h->timing) cudaEventRecord(h->ev[4], h->stream);
    // ---- theta BCs; tau.correctBoundaryConditions(): processor values first, then physical patches in order
    for (ModeDev& md : h->modes) {
        if (md.mp.model == RHEO_MODEL_BMP_FLUIDITY) {   // Phi.correctBoundaryConditions(): zeroGradient faces follow their cells
            if (h->nB) LAUNCH(h, k_bc_zero_gradient2, cdiv(h->nB, BLOCK), BLOCK, h->mv, (const double*)md.theta.as<double>(), md.thetaB.as<double>(), (const double*)nullptr, (double*)nullptr, 0, 0);
            continue;
        }
        if (h->H && halo_planes(h, md.tau.as<double>(), 6)) return 1;
        // patches in patch (= face) order, as EXT-OF9 GeometricBoundaryField::evaluate visits them: a linearExtrapolation
        // patch sees the values of the patches before it already updated and of those after it still old.  Consecutive
        // non-linearExtrapolation patches are one launch; the first launch also refreshes theta's zeroGradient faces.
        if (h->nB && h->tauAssign) {
            const double e = md.desc.model == RHEO_MODEL_OLDROYD_B_LOG ? -md.mp.etaP / md.mp.lambda : 0.0;   // Oldroyd_BLog.C:176 goes through innerP
            LAUNCH(h, k_tau_b_assign, cdiv(h->nB, BLOCK), BLOCK, h->mv, e, md.tauB.as<double>());
        }
        if (h->nB) {
            std::vector<const RheoPatchDesc*> ordered;
            for (const RheoPatchDesc& p : h->patches) ordered.push_back(&p);
            std::sort(ordered.begin(), ordered.end(), [](const RheoPatchDesc* a, const RheoPatchDesc* b) { return a->start < b->start; });
            bool thetaDone = false, pending = false;   // pending: a zeroGradient tau patch lies in [z0, current)
            int z0 = 0;   // first boundary face whose zeroGradient tau value has not been refreshed yet
            auto flush = [&](int z1) {
                if (pending || !thetaDone)
                    LAUNCH(h, k_bc_zero_gradient2, cdiv(h->nB, BLOCK), BLOCK, h->mv, thetaDone ? (const double*)nullptr : md.theta.as<double>(), md.thetaB.as<double>(),
                           md.tau.as<double>(), md.tauB.as<double>(), z0, z1);
                thetaDone = true; pending = false;
                z0 = z1;
            };
            for (const RheoPatchDesc* pp : ordered) {
                const RheoPatchDesc& p = *pp;
                if (p.type != RHEO_PATCH_EMPTY && p.type != RHEO_PATCH_PROCESSOR && p.tau_bc == RHEO_BC_ZERO_GRADIENT && p.size > 0) pending = true;
                if (p.type != RHEO_PATCH_EMPTY && p.type != RHEO_PATCH_PROCESSOR && p.tau_bc == RHEO_BC_LINEAR_EXTRAPOLATION_REG && p.size > 0) {
                    // useRegression true: cell values only, independent of the other patches' boundary values
                    LAUNCH(h, k_tau_bc_regress, cdiv(p.size, 128), 128, h->mv, h->d_CfI.as<double>(), p.start - h->nInt, p.size, md.tau.as<double>(), md.tauB.as<double>());
                    continue;
                }
                if (p.type == RHEO_PATCH_EMPTY || p.type == RHEO_PATCH_PROCESSOR || p.tau_bc != RHEO_BC_LINEAR_EXTRAPOLATION || p.size == 0) continue;
                const int b0 = p.start - h->nInt;
                flush(b0);
                LAUNCH(h, k_tau_bc_linext, cdiv(p.size, 128), 128, h->mv, b0, p.size, md.tau.as<double>(), md.tauB.as<double>(), h->d_tmpB.as<double>());
                LAUNCH(h, k_tau_bc_commit, cdiv(p.size, 128), 128, h->nB, b0, p.size, h->d_tmpB.as<double>(), md.tauB.as<double>());
                z0 = b0 + p.size;
            }
            flush(h->nB);
        }
    }
    if (h->timing) {
        cudaEventRecord(h->ev[5], h->stream);
        cudaEventSynchronize(h->ev[5]);
        float ms;
        cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->phaseMs[0] = ms;
        if (nModes <= perBatch) {   // one batch: free-running phases (events recorded in stream order, read here)
            cudaEventElapsedTime(&msGrad, h->ev[6], h->ev[2]);
            cudaEventElapsedTime(&msAsm, h->ev[6], h->ev[7]);
            cudaEventElapsedTime(&msSolve, h->ev[7], h->ev[3]);
        }
        h->phaseMs[1] = msGrad; h->phaseMs[2] = msAsm;
        h->phaseMs[3] = msSolve;
        cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]); h->phaseMs[4] = ms;
        cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]); h->phaseMs[5] = ms;
        cudaEventElapsedTime(&ms, h->ev[0], h->ev[5]); h->phaseMs[6] = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        const std::string where = h->launchErr.empty() ? std::string() : " (first failing launch: " + h->launchErr + ")";
        h->launchErr.clear();
        return fail(std::string("rheo_gpu_step: ") + cudaGetErrorString(e) + where);
    }
    return 0;
}

// host AoS -> device SoA through the staging buffer
int put_cells(RheoGpu* h, const double* src, int nc, double* dstPlanes) {
    CK(cudaMemcpyAsync(h->d_stage.p, src, (size_t)h->N * nc * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->h2dBytes += (long long)h->N * nc * sizeof(double);
    LAUNCH(h, k_aos_to_soa, cdiv(h->N, BLOCK), BLOCK, h->N, nc, h->d_perm.as<int>(), h->d_stage.as<double>(), dstPlanes, h->NP);
    return 0;
}
int put_bfaces(RheoGpu* h, const double* src, int nc, double* dstPlanes, const std::vector<std::pair<int, int>>* ranges = nullptr) {
    if (!h->nB) return 0;
    if (!ranges) {
        CK(cudaMemcpyAsync(h->d_stage.p, src, (size_t)h->nB * nc * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        h->h2dBytes += (long long)h->nB * nc * sizeof(double);
    } else
        for (const auto& r : *ranges) {   // faces outside the ranges keep stale staging values; no kernel reads them
            CK(cudaMemcpyAsync(h->d_stage.as<double>() + (size_t)r.first * nc, src + (size_t)r.first * nc, (size_t)r.second * nc * sizeof(double),
                               cudaMemcpyHostToDevice, h->stream));
            h->h2dBytes += (long long)r.second * nc * sizeof(double);
        }
    LAUNCH(h, k_aos_to_soa, cdiv(h->nB, BLOCK), BLOCK, h->nB, nc, (const int*)nullptr, h->d_stage.as<double>(), dstPlanes, h->nB);
    return 0;
}

}  // namespace

// ==================================================================== C-ABI
extern "C" {

const char* rheo_gpu_last_error(void) { return g_err.c_str(); }

int rheo_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int rheo_gpu_create(const RheoMeshDesc* mesh, const RheoModelDesc* modes, int32_t n_modes, const RheoSchemeCtl* ctl, int32_t device, RheoGpu** out) {
    if (!mesh || !modes || !ctl || !out || n_modes < 1) return fail("rheo_gpu_create: null/invalid argument");
    if (rheo_gpu_device_count() <= device) return fail("rheo_gpu_create: no CUDA device " + std::to_string(device) + " (this library has no CPU fallback)");
    CK(cudaSetDevice(device));
    RheoGpu* h = new RheoGpu();
    h->device = device;
    h->ctl = *ctl;
    h->lim = make_limiter(ctl->limiter);
    if (ctl->limiter < RHEO_LIMITER_UPWIND || ctl->limiter > RHEO_LIMITER_NONE) { delete h; return fail("The deferred limited scheme is not specified or does not exist. Valid schemes are: upwind cubista minmod smart waceb superbee none"); }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return fail("cudaStreamCreate failed"); }
    { int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device); h->nSms = sms; }
    for (auto& e : h->ev) cudaEventCreate(&e);
    cudaEventCreate(&h->kev0); cudaEventCreate(&h->kev1);
    if (build_mesh(h, mesh) || alloc_fields(h, modes, n_modes)) { rheo_gpu_destroy(h); return 1; }
    *out = h;
    return 0;
}

void rheo_gpu_destroy(RheoGpu* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int r = 0; r < MAX_RANKS; ++r) if (r != h->rank && h->peerBase[r]) cudaIpcCloseMemHandle(h->peerBase[r]);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    for (DevBuf* b : {&h->d_perm, &h->d_faceOld, &h->d_nbr, &h->d_nbrA, &h->d_fidx, &h->d_Sf, &h->d_w, &h->d_C, &h->d_V, &h->d_rV, &h->d_bcell,
                      &h->d_bkind, &h->d_bthetaBC, &h->d_btauBC, &h->d_CfB, &h->d_haloCell, &h->d_send, &h->d_recv,
                      &h->d_tileRec, &h->d_Fell, &h->d_gradU, &h->d_sumPsi, &h->d_mailbox, &h->d_peerSegs, &h->d_segOfGhost, &h->d_peerMisc,
                      &h->d_U, &h->d_Ub, &h->d_phi, &h->d_diag, &h->d_rD, &h->d_Fs, &h->d_FsT, &h->d_stage, &h->d_tmpB, &h->d_r, &h->d_r0, &h->d_p,
                      &h->d_y, &h->d_v, &h->d_s, &h->d_z, &h->d_t, &h->d_ks, &h->d_partials, &h->d_red, &h->d_counter, &h->d_bcells, &h->d_lev, &h->d_chunkLev, &h->d_rowsum, &h->d_inflow, &h->d_tileOrder, &h->d_gradUb, &h->d_CfI})
        b->release();
    for (ModeDev& md : h->modes)
        for (DevBuf* b : {&md.theta, &md.thetaOld, &md.tau, &md.lam, &md.R, &md.fFene, &md.bsrc, &md.thetaB, &md.tauB, &md.gammaVals, &md.corr, &md.thetaOldOld, &md.ddt0, &md.lamCell, &md.etaCell}) b->release();
    if (h->h_ks) cudaFreeHost(h->h_ks);
    for (auto& e : h->ev) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int rheo_gpu_nccl_unique_id(void* id128) {
    if (!g_nccl.load()) return fail("rheo_gpu_nccl_unique_id: cannot load libnccl.so.2");
    NcclId id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc) return fail(std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
    std::memcpy(id128, &id, 128);
    return 0;
}

int rheo_gpu_comm_init(RheoGpu* h, int32_t rank, int32_t n_ranks, const void* id128) {
    if (!h) return fail("rheo_gpu_comm_init: null handle");
    if (!g_nccl.load()) return fail("rheo_gpu_comm_init: cannot load libnccl.so.2");
    CK(cudaSetDevice(h->device));
    NcclId id;
    std::memcpy(&id, id128, 128);
    int rc = g_nccl.CommInitRank(&h->comm, n_ranks, id, rank);
    if (rc) return fail(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(rc));
    h->rank = rank; h->nRanks = n_ranks;
    // global cell count for gAverage
    double* tmp = h->d_red.as<double>();
    double n = (double)h->N;
    CK(cudaMemcpyAsync(tmp, &n, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (all_reduce(h, tmp, 1)) return 1;
    CK(cudaMemcpyAsync(&n, tmp, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->nGlobalCells = (long)(n + 0.5);
    return setup_peer(h);
}

int rheo_gpu_upload_state(RheoGpu* h, int32_t mode, const double* theta, const double* tau, const double* eigvals, const double* eigvecs,
                          const double* theta_b, const double* tau_b) {
    if (!h || mode < 0 || mode + h->nHidden >= (int)h->modes.size()) return fail("rheo_gpu_upload_state: bad handle/mode");
    CK(cudaSetDevice(h->device));
    ModeDev& md = h->modes[mode + h->nHidden];
    if (theta) {
        if (put_cells(h, theta, 6, md.theta.as<double>())) return 1;
        CK(cudaMemcpyAsync(md.thetaOld.p, md.theta.p, md.theta.bytes, cudaMemcpyDeviceToDevice, h->stream));
        if (md.thetaOldOld.p) CK(cudaMemcpyAsync(md.thetaOldOld.p, md.theta.p, md.theta.bytes, cudaMemcpyDeviceToDevice, h->stream));
        if (md.ddt0.p) { if (zero(h, md.ddt0)) return 1; }
        md.ddt0TimeIndex = 0;
        h->nOldTimes = 0;
    }
    if (tau && put_cells(h, tau, 6, md.tau.as<double>())) return 1;
    if (eigvals) {
        CK(cudaMemcpyAsync(h->d_stage.p, eigvals, (size_t)h->N * 9 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        LAUNCH(h, k_lam_from_tensor, cdiv(h->N, BLOCK), BLOCK, h->N, h->d_perm.as<int>(), h->d_stage.as<double>(), md.lam.as<double>(), h->NP);
    }
    if (eigvecs && put_cells(h, eigvecs, 9, md.R.as<double>())) return 1;
    if (theta_b) { if (put_bfaces(h, theta_b, 6, md.thetaB.as<double>())) return 1; }
    if (h->nB) LAUNCH(h, k_bc_zero_gradient, cdiv(h->nB, BLOCK), BLOCK, h->mv, h->d_bthetaBC.as<int>(), md.theta.as<double>(), md.thetaB.as<double>(), 6);
    if (tau_b) { if (put_bfaces(h, tau_b, 6, md.tauB.as<double>())) return 1; }
    else if (h->nB) LAUNCH(h, k_bc_zero_gradient, cdiv(h->nB, BLOCK), BLOCK, h->mv, h->d_btauBC.as<int>(), md.tau.as<double>(), md.tauB.as<double>(), 6);
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// correct(alpha, gradU) with gradU != nullptr (boilerLog.H:1): the caller's gradient, OpenFOAM tensor order L_ij = d_i U_j at 3i+j,
// into the planes the source kernels read (g[3k+d] = d_d U_k: the transposed component order)
int rheo_gpu_upload_grad_u(RheoGpu* h, const double* gradU9) {
    if (!h) return fail("rheo_gpu_upload_grad_u: null handle");
    CK(cudaSetDevice(h->device));
    if (!gradU9) { h->externalGradU = false; return 0; }
    if (put_cells(h, gradU9, 9, h->d_gradU.as<double>())) return 1;
    // plane 3i+j holds L_ij; the kernels want d_d U_k at 3k+d = L_dk: swap the off-diagonal pairs through the staging buffer
    const size_t pb = (size_t)h->NP * sizeof(double);
    double* g = h->d_gradU.as<double>();
    for (const auto& pr : {std::pair<int, int>{1, 3}, {2, 6}, {5, 7}}) {
        CK(cudaMemcpyAsync(h->d_stage.p, g + (size_t)pr.first * h->NP, pb, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(g + (size_t)pr.first * h->NP, g + (size_t)pr.second * h->NP, pb, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(g + (size_t)pr.second * h->NP, h->d_stage.p, pb, cudaMemcpyDeviceToDevice, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    h->externalGradU = true;
    return 0;
}

// BMPLog: the fluidity field (component xx of the hidden mode; the other components stay 0)
int rheo_gpu_upload_fluidity(RheoGpu* h, int32_t mode, const double* Phi, const double* Phi_b) {
    if (!h || !Phi) return fail("rheo_gpu_upload_fluidity: null argument");
    if (!h->nHidden || mode != 0) return fail("rheo_gpu_upload_fluidity: not a BMPLog mode");
    CK(cudaSetDevice(h->device));
    ModeDev& md = h->modes[0];
    if (zero(h, md.theta) || zero(h, md.thetaB)) return 1;
    if (put_cells(h, Phi, 1, md.theta.as<double>())) return 1;
    CK(cudaMemcpyAsync(md.thetaOld.p, md.theta.p, md.theta.bytes, cudaMemcpyDeviceToDevice, h->stream));
    if (md.thetaOldOld.p) CK(cudaMemcpyAsync(md.thetaOldOld.p, md.theta.p, md.theta.bytes, cudaMemcpyDeviceToDevice, h->stream));
    if (md.ddt0.p) { if (zero(h, md.ddt0)) return 1; }
    md.ddt0TimeIndex = 0;
    h->nOldTimes = 0;
    CK(cudaStreamSynchronize(h->stream));   // the staging buffer is reused by the next copy
    if (Phi_b && put_bfaces(h, Phi_b, 1, md.thetaB.as<double>())) return 1;
    if (h->nB) LAUNCH(h, k_bc_zero_gradient, cdiv(h->nB, BLOCK), BLOCK, h->mv, h->d_bthetaBC.as<int>(), md.theta.as<double>(), md.thetaB.as<double>(), 6);
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int rheo_gpu_upload_velocity(RheoGpu* h, const double* U, const double* U_b, const double* phi) {
    if (!h || !U || !phi) return fail("rheo_gpu_upload_velocity: null argument");
    CK(cudaSetDevice(h->device));
    if (put_cells(h, U, 3, h->d_U.as<double>())) return 1;
    if (h->nB && U_b && put_bfaces(h, U_b, 3, h->d_Ub.as<double>(), &h->ubRanges)) return 1;
    CK(cudaMemcpyAsync(h->d_stage.p, phi, (size_t)h->nInt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->h2dBytes += (long long)h->nInt * sizeof(double);
    for (const auto& r : h->phiBRanges) {
        CK(cudaMemcpyAsync(h->d_stage.as<double>() + h->nInt + r.first, phi + h->nInt + r.first, (size_t)r.second * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        h->h2dBytes += (long long)r.second * sizeof(double);
    }
    LAUNCH(h, k_phi_in, cdiv(h->nF, BLOCK), BLOCK, h->nF, h->d_faceOld.as<int>(), h->d_stage.as<double>(), h->d_phi.as<double>());
    LAUNCH(h, k_flux_ell, cdiv(h->N, BLOCK), BLOCK, h->mv, h->d_phi.as<double>(), h->d_Fell.as<double>());
    return 0;
}

int rheo_gpu_upload_thermo(RheoGpu* h, int32_t mode, const double* lambda_cell, const double* etaP_cell) {
    if (!h || mode < 0 || mode + h->nHidden >= (int)h->modes.size()) return fail("rheo_gpu_upload_thermo: bad handle/mode");
    if (h->nHidden) return fail("rheo_gpu_upload_thermo: BMPLog has no thermo-dependent parameters (the per-cell rates come from its fluidity)");
    if ((lambda_cell == nullptr) != (etaP_cell == nullptr)) return fail("rheo_gpu_upload_thermo: pass both lambda and etaP per cell, or neither");
    CK(cudaSetDevice(h->device));
    ModeDev& md = h->modes[mode];
    if (!lambda_cell) {
        CK(cudaStreamSynchronize(h->stream));
        md.lamCell.release(); md.etaCell.release();
        return 0;
    }
    if (!md.lamCell.p && (md.lamCell.alloc((size_t)h->NP * sizeof(double)) || md.etaCell.alloc((size_t)h->NP * sizeof(double)))) return 1;
    if (put_cells(h, lambda_cell, 1, md.lamCell.as<double>())) return 1;
    CK(cudaStreamSynchronize(h->stream));   // the staging buffer is reused by the next copy
    if (put_cells(h, etaP_cell, 1, md.etaCell.as<double>())) return 1;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int rheo_gpu_set_tau_assignment(RheoGpu* h, int32_t on) {
    if (!h) return fail("null handle");
    h->tauAssign = on != 0;
    return 0;
}

int rheo_gpu_store_old_time(RheoGpu* h) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    for (ModeDev& md : h->modes) {
        if (md.thetaOldOld.p) CK(cudaMemcpyAsync(md.thetaOldOld.p, md.thetaOld.p, md.theta.bytes, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(md.thetaOld.p, md.theta.p, md.theta.bytes, cudaMemcpyDeviceToDevice, h->stream));
    }
    h->nOldTimes++;
    h->dt0 = h->dtNow;   // Time::operator++: deltaT0_ = deltaT_
    return 0;
}

int rheo_gpu_step(RheoGpu* h, double dt, RheoStepStats* stats) {
    if (!h) return fail("rheo_gpu_step: null handle");
    CK(cudaSetDevice(h->device));
    return do_step(h, dt, stats);
}

int rheo_gpu_download(RheoGpu* h, int32_t mode, int32_t field, double* dst) {
    if (!h || !dst || mode < 0 || mode + h->nHidden >= (int)h->modes.size()) return fail("rheo_gpu_download: bad argument");
    CK(cudaSetDevice(h->device));
    ModeDev& md = h->modes[mode + h->nHidden];
    const int N = h->N, NP = h->NP, nB = h->nB;
    double* stage = h->d_stage.as<double>();
    size_t bytes = 0;
    switch (field) {
        case RHEO_FIELD_THETA: LAUNCH(h, k_soa_to_aos, cdiv(N, BLOCK), BLOCK, N, 6, h->d_perm.as<int>(), md.theta.as<double>(), stage, NP); bytes = (size_t)N * 6; break;
        case RHEO_FIELD_THETA_OLD: LAUNCH(h, k_soa_to_aos, cdiv(N, BLOCK), BLOCK, N, 6, h->d_perm.as<int>(), md.thetaOld.as<double>(), stage, NP); bytes = (size_t)N * 6; break;
        case RHEO_FIELD_TAU: LAUNCH(h, k_soa_to_aos, cdiv(N, BLOCK), BLOCK, N, 6, h->d_perm.as<int>(), md.tau.as<double>(), stage, NP); bytes = (size_t)N * 6; break;
        case RHEO_FIELD_TAU_TOTAL:
            for (size_t mi = h->nHidden; mi < h->modes.size(); ++mi) {
                if ((int)mi == h->nHidden) LAUNCH(h, k_soa_to_aos, cdiv(N, BLOCK), BLOCK, N, 6, h->d_perm.as<int>(), h->modes[mi].tau.as<double>(), stage, NP);
                else LAUNCH(h, k_soa_to_aos_acc, cdiv(N, BLOCK), BLOCK, N, 6, h->d_perm.as<int>(), h->modes[mi].tau.as<double>(), stage, NP);
            }
            bytes = (size_t)N * 6;
            break;
        case RHEO_FIELD_EIGVALS: LAUNCH(h, k_lam_to_tensor, cdiv(N, BLOCK), BLOCK, N, h->d_perm.as<int>(), md.lam.as<double>(), stage, NP); bytes = (size_t)N * 9; break;
        case RHEO_FIELD_EIGVECS: LAUNCH(h, k_soa_to_aos, cdiv(N, BLOCK), BLOCK, N, 9, h->d_perm.as<int>(), md.R.as<double>(), stage, NP); bytes = (size_t)N * 9; break;
        case RHEO_FIELD_THETA_B: if (nB) LAUNCH(h, k_soa_to_aos, cdiv(nB, BLOCK), BLOCK, nB, 6, (const int*)nullptr, md.thetaB.as<double>(), stage, nB); bytes = (size_t)nB * 6; break;
        case RHEO_FIELD_TAU_B: if (nB) LAUNCH(h, k_soa_to_aos, cdiv(nB, BLOCK), BLOCK, nB, 6, (const int*)nullptr, md.tauB.as<double>(), stage, nB); bytes = (size_t)nB * 6; break;
        case RHEO_FIELD_TAU_B_TOTAL:   // multiMode::divTau sums each mode's divTau, i.e. sees the sum of the modes' patch values
            for (size_t mi = h->nHidden; nB && mi < h->modes.size(); ++mi) {
                if ((int)mi == h->nHidden) LAUNCH(h, k_soa_to_aos, cdiv(nB, BLOCK), BLOCK, nB, 6, (const int*)nullptr, h->modes[mi].tauB.as<double>(), stage, nB);
                else LAUNCH(h, k_soa_to_aos_acc, cdiv(nB, BLOCK), BLOCK, nB, 6, (const int*)nullptr, h->modes[mi].tauB.as<double>(), stage, nB);
            }
            bytes = (size_t)nB * 6;
            break;
        case RHEO_FIELD_FLUIDITY:
            if (!h->nHidden) return fail("rheo_gpu_download: not a BMPLog model");
            LAUNCH(h, k_soa_to_aos, cdiv(N, BLOCK), BLOCK, N, 1, h->d_perm.as<int>(), h->modes[0].theta.as<double>(), stage, NP); bytes = (size_t)N; break;
        case RHEO_FIELD_FLUIDITY_B:
            if (!h->nHidden) return fail("rheo_gpu_download: not a BMPLog model");
            if (nB) LAUNCH(h, k_soa_to_aos, cdiv(nB, BLOCK), BLOCK, nB, 1, (const int*)nullptr, h->modes[0].thetaB.as<double>(), stage, nB); bytes = (size_t)nB; break;
        default: return fail("rheo_gpu_download: unknown field");
    }
    if (bytes) CK(cudaMemcpyAsync(dst, stage, bytes * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    h->d2hBytes += (long long)bytes * sizeof(double);
    CK(cudaStreamSynchronize(h->stream));
    return check_peer_err(h);
}

// constitutiveEq::divTau, explicit part (momentum.cuh)
int rheo_gpu_div_tau(RheoGpu* h, int32_t stabilization, double* div_out) {
    if (!h || !div_out) return fail("rheo_gpu_div_tau: null argument");
    if (stabilization != RHEO_STAB_NONE && stabilization != RHEO_STAB_BSD && stabilization != RHEO_STAB_COUPLING) return fail("rheo_gpu_div_tau: unknown stabilization");
    if ((int)h->modes.size() - h->nHidden > MAX_MODES_DIV) return fail("rheo_gpu_div_tau: more than 8 modes");
    CK(cudaSetDevice(h->device));
    const bool coupling = stabilization == RHEO_STAB_COUPLING;
    if (coupling && h->externalGradU) return fail("rheo_gpu_div_tau: stabilization coupling needs fvc::grad(U), but a caller-supplied gradU is in place (rheo_gpu_upload_grad_u(h, NULL) first)");
    DivTauArgs a{};
    a.nModes = (int)h->modes.size() - h->nHidden;
    for (int mi = 0; mi < a.nModes; ++mi) {
        ModeDev& md = h->modes[mi + h->nHidden];
        if (coupling && md.etaCell.p && !h->nHidden) return fail("rheo_gpu_div_tau: stabilization coupling with a temperature-dependent etaP is not implemented (download tau instead)");
        a.tau[mi] = md.tau.as<double>(); a.tauB[mi] = md.tauB.as<double>();
        a.rRho[mi] = 1.0 / md.desc.rho;
        if (coupling) a.coefGrad += md.desc.etaP / md.desc.rho;
    }
    a.gradU = h->d_gradU.as<double>(); a.U = h->d_U.as<double>(); a.Ub = h->d_Ub.as<double>();
    a.perm = h->d_perm.as<int>(); a.out = h->d_stage.as<double>();
    if (coupling) {
        if (h->H && halo_planes(h, h->d_gradU.as<double>(), 9)) return 1;
        if (h->nB) {
            if (!h->d_gradUb.p && h->d_gradUb.alloc((size_t)9 * h->nB * sizeof(double))) return 1;
            LAUNCH(h, k_gradU_patch, cdiv(h->nB, BLOCK), BLOCK, h->mv, a.gradU, a.U, a.Ub, h->d_gradUb.as<double>());
        }
        a.gradUb = h->d_gradUb.as<double>();
    }
    a.tileOrder = h->d_tileOrder.as<int>(); a.nTiles = h->nTiles;
    LAUNCH(h, k_div_tau, cdiv(a.tileOrder ? h->nTiles * TILE : h->N, BLOCK), BLOCK, h->mv, a);
    CK(cudaMemcpyAsync(div_out, h->d_stage.p, (size_t)h->N * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    h->d2hBytes += (long long)h->N * 3 * sizeof(double);
    CK(cudaStreamSynchronize(h->stream));
    return check_peer_err(h);
}

int rheo_gpu_correct(RheoGpu* h, const double* U, const double* U_b, const double* phi, double dt, int32_t new_time_step, double* tau_out,
                     double* tau_b_out, RheoStepStats* stats) {
    if (rheo_gpu_upload_velocity(h, U, U_b, phi)) return 1;
    if (new_time_step && rheo_gpu_store_old_time(h)) return 1;
    if (rheo_gpu_step(h, dt, stats)) return 1;
    if (tau_out && rheo_gpu_download(h, 0, RHEO_FIELD_TAU_TOTAL, tau_out)) return 1;
    if (tau_b_out && rheo_gpu_download(h, 0, RHEO_FIELD_TAU_B_TOTAL, tau_b_out)) return 1;
    return 0;
}

int rheo_gpu_get_renumbering(RheoGpu* h, int32_t* perm, int32_t* n_colours, int32_t* colour_start) {
    if (!h) return fail("null handle");
    if (perm) std::copy(h->perm.begin(), h->perm.end(), perm);
    if (n_colours) *n_colours = h->nColours;
    if (colour_start) std::copy(h->colourStart.begin(), h->colourStart.end(), colour_start);
    return 0;
}

int rheo_gpu_get_ordering(RheoGpu* h, char* buf, int32_t buflen) {
    if (!h || !buf || buflen < 1) return fail("rheo_gpu_get_ordering: bad argument");
    snprintf(buf, (size_t)buflen, "%s", h->orderingInfo.c_str());
    return 0;
}

int rheo_gpu_get_levels(RheoGpu* h, int32_t* fwd, int32_t* bwd) {
    if (!h || !fwd || !bwd) return fail("rheo_gpu_get_levels: bad argument");
    if (!h->blockMode) return fail("rheo_gpu_get_levels: this handle uses the cell colouring (no in-chunk levels)");
    CK(cudaSetDevice(h->device));
    std::vector<uint16_t> lev(h->NS);
    CK(cudaMemcpy(lev.data(), h->d_lev.p, (size_t)h->NS * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    for (int c = 0; c < h->N; ++c) { fwd[c] = lev[c] & 255; bwd[c] = lev[c] >> 8; }
    return 0;
}

int rheo_gpu_get_ell(RheoGpu* h, int32_t* K, int32_t* nbr, int32_t* face) {
    if (!h) return fail("null handle");
    if (K) *K = h->K;
    for (int s = 0; s < h->K; ++s)
        for (int c = 0; c < h->N; ++c) {
            if (nbr) nbr[(size_t)s * h->N + c] = h->h_nbr[(size_t)s * h->NS + c];
            if (face) face[(size_t)s * h->N + c] = h->h_fidx[(size_t)s * h->NS + c];
        }
    return 0;
}

int64_t rheo_gpu_launch_count(const RheoGpu* h) { return h ? h->launches : 0; }
int rheo_gpu_last_iterations(const RheoGpu* h) { return h ? h->lastIters : -1; }
int rheo_gpu_comm_stats(RheoGpu* h, int32_t* mode, double* wait_ms2, int64_t* waits2) {
    if (!h) return 1;
    if (mode) *mode = h->nRanks <= 1 ? 0 : (h->p2p ? 2 : 1);
    unsigned long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (h->p2p) {
        CK(cudaSetDevice(h->device));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpy(st, h->pv.stat, sizeof st, cudaMemcpyDeviceToHost));
    }
    if (wait_ms2) { wait_ms2[0] = st[0] * 1e-6; wait_ms2[1] = st[1] * 1e-6; }
    if (waits2) { waits2[0] = (int64_t)st[2]; waits2[1] = (int64_t)st[3]; }
    if (getenv("RHEO_PEER_TRACE")) fprintf(stderr, "peer halo_planes phases (ms total): store %.3f fence %.3f publish+wait %.3f unpack %.3f\n", st[4] * 1e-6, st[5] * 1e-6, st[6] * 1e-6, st[7] * 1e-6);
    return 0;
}
int rheo_gpu_transfer_bytes(const RheoGpu* h, int64_t* h2d, int64_t* d2h) {
    if (!h) return 1;
    if (h2d) *h2d = h->h2dBytes;
    if (d2h) *d2h = h->d2hBytes;
    return 0;
}
int rheo_gpu_set_phase_timing(RheoGpu* h, int32_t enabled) { if (!h) return 1; h->timing = enabled != 0; return 0; }
int rheo_gpu_get_phase_times(RheoGpu* h, double* ms7) { if (!h || !ms7) return 1; std::copy(h->phaseMs, h->phaseMs + 7, ms7); return 0; }
int rheo_gpu_set_kernel_timing(RheoGpu* h, int32_t enabled) {
    if (!h) return 1;
    h->ktiming = enabled != 0;
    h->ktimes.clear();
    return 0;
}
int rheo_gpu_get_kernel_times(RheoGpu* h, char* buf, int32_t buflen) {
    if (!h || !buf || buflen < 1) return 1;
    std::string out;
    for (auto& kv : h->ktimes) {
        char line[256];
        snprintf(line, sizeof line, "%s %ld %.6f\n", kv.first.c_str(), kv.second.second, kv.second.first);
        out += line;
    }
    snprintf(buf, (size_t)buflen, "%s", out.c_str());
    return 0;
}
int rheo_gpu_stream(RheoGpu* h, void** s) { if (!h || !s) return 1; *s = (void*)h->stream; return 0; }
int rheo_gpu_synchronize(RheoGpu* h) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return check_peer_err(h);
}

int rheo_gpu_eig_exp(int32_t device, int32_t n, const double* theta6, double* eigvals9, double* eigvecs9) {
    if (rheo_gpu_device_count() <= device) return fail("rheo_gpu_eig_exp: no CUDA device (no CPU fallback)");
    CK(cudaSetDevice(device));
    double *a = nullptr, *b = nullptr, *c = nullptr;
    CK(cudaMalloc(&a, (size_t)n * 6 * 8)); CK(cudaMalloc(&b, (size_t)n * 9 * 8)); CK(cudaMalloc(&c, (size_t)n * 9 * 8));
    CK(cudaMemcpy(a, theta6, (size_t)n * 6 * 8, cudaMemcpyHostToDevice));
    k_eig_exp_aos<<<cdiv(n, BLOCK), BLOCK>>>(n, a, b, c);
    CK(cudaMemcpy(eigvals9, b, (size_t)n * 9 * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(eigvecs9, c, (size_t)n * 9 * 8, cudaMemcpyDeviceToHost));
    cudaFree(a); cudaFree(b); cudaFree(c);
    return 0;
}

int rheo_gpu_abi_sizes(int32_t* out8) {
    if (!out8) return fail("rheo_gpu_abi_sizes: null argument");
    const size_t sz[8] = {sizeof(RheoPatchDesc), sizeof(RheoMeshDesc), sizeof(RheoModelDesc), sizeof(RheoSchemeCtl),
                          sizeof(RheoStepStats), sizeof(RheoSynthSpec), sizeof(RheoPatchRule), sizeof(RheoPatchSpec)};
    for (int i = 0; i < 8; ++i) out8[i] = (int32_t)sz[i];
    return 0;
}

}  // extern "C"
