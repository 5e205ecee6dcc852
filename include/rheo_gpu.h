/* rheo_gpu.h — C-ABI of the B200-native viscoelastic stress step (log-conformation models).
 *
 * One handle == one `Foam::constitutiveEq` object of the reference (or one `multiMode` wrapper with
 * n_modes sub-models) living on one GPU / one MPI rank.  Each entry point cites the reference
 * interface it replaces; paths are relative to of90/src/libs/constitutiveEquations/constitutiveEqs/.
 *
 *   rheo_gpu_create            <- XxxLog constructors (Oldroyd-B/Oldroyd-BLog/Oldroyd_BLog.C:43-122,
 *                                 Giesekus/GiesekusLog/GiesekusLog.C:43-123, PTT/PTTLog/PTTLog.C:58-171,
 *                                 FENE-P/FENE-PLog/FENE_PLog.C:43-123, multiMode/multiMode.C:41-92) and
 *                                 the run-time selection call constitutiveEq/newConstitutiveEq.C:32-61
 *   rheo_gpu_upload_state      <- MUST_READ tau/theta and READ_IF_PRESENT eigVals/eigVecs
 *                                 (Oldroyd_BLog.C:52-113; defaults = identity)
 *   rheo_gpu_upload_velocity   <- the const references U(), phi() the model holds
 *                                 (constitutiveEq/constitutiveEq.H:72-75,285-295)
 *   rheo_gpu_store_old_time    <- theta_.oldTime() bookkeeping done by OpenFOAM when runTime++
 *                                 (fvm::ddt(theta_), Oldroyd_BLog.C:143)
 *   rheo_gpu_step              <- constitutiveEq::correct() (constitutiveEq.H:346-350; bodies:
 *                                 Oldroyd_BLog.C:127-179, GiesekusLog.C:128-176, PTTLog.C:176-268,
 *                                 FENE_PLog.C:128-182, multiMode.C:247-260)
 *   rheo_gpu_download          <- tau() (constitutiveEq.H:314; multiMode.C:216-226) and the AUTO_WRITE
 *                                 of tau/theta/eigVals/eigVecs at write time (Oldroyd_BLog.C:52-113)
 *   rheo_gpu_correct           <- correct() + tau() as a CPU momentum predictor uses them
 *                                 (of90/src/solvers/rheoFoam/rheoFoam.C:147-153, UEqn.H:12)
 *   rheo_gpu_last_error        <- FatalErrorInFunction text (newConstitutiveEq.C:47-58); every call
 *                                 returns 0 on success, non-zero on error (the C++ shim turns that
 *                                 into FatalError).
 *
 * Host arrays use OpenFOAM's AoS layouts (vector 3, symmTensor 6 = xx,xy,xz,yy,yz,zz, tensor 9
 * row-major); on the device everything is FP64 structure-of-arrays in a colour-sorted numbering
 * (see DESIGN.md).  There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef RHEO_GPU_H
#define RHEO_GPU_H

#include <stdint.h>
#include "rheo_mesh.h"

#ifdef __cplusplus
extern "C" {
#endif

/* constitutive models (`type` in constitutiveProperties) */
#define RHEO_MODEL_OLDROYD_B_LOG 0
#define RHEO_MODEL_GIESEKUS_LOG  1
#define RHEO_MODEL_PTT_LOG       2
#define RHEO_MODEL_FENE_P_LOG    3
#define RHEO_MODEL_FENE_CR_LOG   4   /* FENE-CR/FENE-CRLog/FENE_CRLog.C:128-182 (further Log models) */
#define RHEO_MODEL_WM_CY_LOG     5   /* WhiteMetznerCY/WhiteMetznerCYLog/WhiteMetznerCYLog.C:145-211 */
#define RHEO_MODEL_ROLIE_POLY_LOG 6  /* Rolie-Poly/Rolie-PolyLog/RoliePolyLog.C:130-215 (lambda = lambdaD) */
#define RHEO_MODEL_XPOMPOM_LOG   7   /* XPomPom/XPomPomLog/XPomPomLog.C:130-198 (lambda = lambdaB, alpha = anisotropy) */

#define RHEO_MODEL_BMP_LOG       9   /* otherModels/BMP/BMPLog/BMPLog.C:142-201 (thixotropic: the fluidity Phi obeys its own transport
                                       equation, solved before theta in every correct(); single-mode only) */
#define RHEO_MODEL_BMP_FLUIDITY  10  /* internal: the fluidity equation of a BMPLog mode (BMPLog.C:151-163), carried as component xx of
                                       a padded symmTensor through the same assembly and solver; not selectable by the caller */
#define RHEO_MODEL_SARAMITO_LOG  8   /* otherModels/Saramito/SaramitoLog/SaramitoLog.C:143-245 (elasto-viscoplastic: the relaxation
                                        term is switched by max(0, (|tau_d| - tau0)/(k |tau_d|^n))^(1/n) of the CURRENT tau) */
/* PTTLog destructionFunctionType (PTT/PTTLog/PTTLog.C:41-50,190-237) */
#define RHEO_PTT_LINEAR      0
#define RHEO_PTT_EXPONENTIAL 1
#define RHEO_PTT_GENERALIZED 2

/* GaussDefCmpw limiter (gaussDefCmpwConvectionScheme/limiters.H:48-98) */
#define RHEO_LIMITER_UPWIND   0
#define RHEO_LIMITER_CUBISTA  1
#define RHEO_LIMITER_MINMOD   2
#define RHEO_LIMITER_SMART    3
#define RHEO_LIMITER_WACEB    4
#define RHEO_LIMITER_SUPERBEE 5
#define RHEO_LIMITER_NONE     6   /* no convection */

#define RHEO_DDT_EULER    0
#define RHEO_DDT_BACKWARD 1   /* EXT-OF9 backwardDdtScheme: Euler until the field has two old times, variable-step coefficients after */

#define RHEO_DDT_CRANK_NICOLSON 2   /* EXT-OF9 CrankNicolsonDdtScheme `CrankNicolson <psi>` (tutorial Cavity/Oldroyd-BLog/system/fvSchemes:
                                      `CrankNicolson 1`): ddt0-field formulation on a fresh start (first step Euler), cn_psi = off-centring */
#define RHEO_DDT_STEADY_STATE 3      /* EXT-OF9 steadyStateDdtScheme: fvm::ddt contributes nothing (tutorials rheoFilmFoam/UCM/system/fvSchemes:
                                      `default steadyState`); used with relaxationFactors.equations.theta < 1 and `bounded` convection */
#define RHEO_SOLVER_PBICGSTAB 0
#define RHEO_SOLVER_PBICG     1

typedef struct RheoModelDesc {
    int32_t model;          /* RHEO_MODEL_*                                   */
    double  rho, etaS, etaP, lambda;
    double  alpha;          /* GiesekusLog mobility                            */
    double  epsilon, zeta;  /* PTTLog                                          */
    int32_t ptt_function;   /* RHEO_PTT_*                                      */
    double  ml_alpha, ml_beta, ml_rtol;  /* PTTLog generalized (Mittag-Leffler) */
    int32_t ml_max_iter;
    double  L2;             /* FENE-PLog / FENE-CRLog extensibility            */
    double  wm_K, wm_n, wm_a; /* WhiteMetznerCYLog: eta, lambda *= (1 + (K gdot)^a)^((n-1)/a); the Log version requires
                               m = n, L = K, b = a (WhiteMetznerCYLog.C:132-140: the caller checks and passes one set) */
    double  rp_lambdaR, rp_beta, rp_delta, rp_chiMax;   /* Rolie-PolyLog (RoliePolyLog.C:114-121)            */
    double  xpp_lambdaS, xpp_q, xpp_n;                  /* XPomPomLog   (XPomPomLog.C:115-122)               */
    double  sar_tau0, sar_k, sar_n;   /* SaramitoLog yield stress, consistency (viscosity units; = etaP when n == 1, SaramitoLog.C:116),
                                         index; epsilon / zeta / ptt_function (linear | exponential, only with n == 1) as for PTTLog */
    double  sar_dims[3];              /* SaramitoLog `dims` (1 = valid geometric direction, SaramitoLog.C:117-119,128-130)           */
    int32_t sar_ptt;                  /* SaramitoLog PTTfunction: 0 none, 1 linear, 2 exponential (SaramitoLog.C:133-165)            */
    double  bmp_G0, bmp_k, bmp_Phi0, bmp_PhiInf;   /* BMPLog (BMPLog.C:129-136): elastic modulus, structure break-down constant, zero- and
                                                      infinite-shear fluidities; lambda = structure build-up time; etaP only enters divTau */
    double  bmp_relax;                /* relaxationFactors.equations.Phi (PhiEqn.relax(), BMPLog.C:165); <= 0 : no-op                */
} RheoModelDesc;

typedef struct RheoSchemeCtl {
    int32_t limiter;        /* divSchemes div(phi,theta) GaussDefCmpw <limiter>          */
    int32_t ddt;            /* ddtSchemes: RHEO_DDT_EULER | RHEO_DDT_BACKWARD            */
    int32_t solver;         /* fvSolution solvers.theta.solver                           */
    double  tolerance;      /* fvSolution tolerance                                      */
    double  rel_tol;        /* fvSolution relTol                                         */
    int32_t min_iter;
    int32_t max_iter;
    double  relax;          /* relaxationFactors.equations.theta; <= 0 : relax() is a no-op */
    double  cn_psi;         /* CrankNicolson off-centring coefficient psi in [0,1] (1 = Crank-Nicolson, 0 = Euler); other ddt: unused */
    int32_t bounded;        /* divSchemes `bounded GaussDefCmpw <limiter>` (EXT-OF9 boundedConvectionScheme: fvmDiv - fvm::Sp(div(phi)),
                               rheoFilmFoam/UCM/system/fvSchemes:35): the net outflow of the cell is taken off the diagonal */
    int32_t pad_;
} RheoSchemeCtl;

/* mirrors OpenFOAM SolverPerformance<symmTensor> per mode */
typedef struct RheoStepStats {
    double  initial_residual[6];
    double  final_residual[6];
    int32_t n_iterations[6];
    int32_t converged[6];
} RheoStepStats;

typedef struct RheoGpu RheoGpu;

/* what rheo_gpu_download / rheo_gpu_upload_field move */
#define RHEO_FIELD_THETA      0  /* symmTensor, 6/cell */
#define RHEO_FIELD_TAU        1  /* symmTensor, 6/cell */
#define RHEO_FIELD_EIGVALS    2  /* tensor, 9/cell (diagonal = exp(eig)) */
#define RHEO_FIELD_EIGVECS    3  /* tensor, 9/cell (columns = eigenvectors) */
#define RHEO_FIELD_THETA_B    4  /* symmTensor, 6/boundary face */
#define RHEO_FIELD_TAU_B      5  /* symmTensor, 6/boundary face */
#define RHEO_FIELD_TAU_TOTAL  6  /* sum over modes of tau, 6/cell (multiMode::tau) */
#define RHEO_FIELD_THETA_OLD  7
#define RHEO_FIELD_FLUIDITY    9  /* BMPLog: Phi, 1/cell      */
#define RHEO_FIELD_FLUIDITY_B  10 /* BMPLog: Phi, 1/boundary face */
#define RHEO_FIELD_TAU_B_TOTAL 8 /* sum over modes of the boundary stress, 6/boundary face: what multiMode::divTau sees on the
                                    patches (each mode's own linearExtrapolation / zeroGradient / fixedValue values, summed) */

/* constitutiveProperties `stabilization` (constitutiveEq/constitutiveEq.H: soNone, soBSD, soCoupling) */
#define RHEO_STAB_NONE      0
#define RHEO_STAB_BSD       1
#define RHEO_STAB_COUPLING  2

int rheo_gpu_device_count(void);

/* Build the device-resident model.  The mesh arrays are only read during the call. */
int rheo_gpu_create(const RheoMeshDesc* mesh, const RheoModelDesc* modes, int32_t n_modes,
                    const RheoSchemeCtl* ctl, int32_t device, RheoGpu** out);
void rheo_gpu_destroy(RheoGpu* h);

/* Multi-GPU: one rank per GPU.  Rank 0 obtains an id (128 bytes), the host broadcasts it by any
 * means (MPI_Bcast in the OpenFOAM shim, torch.distributed in bench.py), every rank then joins.
 * Halo swaps of processor patches and Krylov reductions then go through NCCL on the model's stream. */
int rheo_gpu_nccl_unique_id(void* id128);
int rheo_gpu_comm_init(RheoGpu* h, int32_t rank, int32_t n_ranks, const void* id128);

/* State.  NULL eigvals/eigvecs => identity (READ_IF_PRESENT default); NULL theta_b/tau_b => boundary
 * values derived from the BC (fixedValue patches then hold 0). */
int rheo_gpu_upload_state(RheoGpu* h, int32_t mode, const double* theta, const double* tau,
                          const double* eigvals, const double* eigvecs,
                          const double* theta_b, const double* tau_b);

/* BMPLog: the fluidity field Phi (MUST_READ in BMPLog.C:112-122) of `mode`: Phi[n_cells], Phi_b[n_boundary_faces] or NULL (boundary
 * values derived from the BC kinds, which are those of theta: fixedValue patches then hold 0). */
int rheo_gpu_upload_fluidity(RheoGpu* h, int32_t mode, const double* Phi, const double* Phi_b);

/* U [3*n_cells], U_b [3*n_boundary_faces] (faces of processor/empty patches are neither read nor copied),
 * phi [n_faces] (internal then boundary; faces of empty patches are neither read nor copied — OpenFOAM's
 * emptyFvPatchField has size 0, so the shim has nothing to put there).  Pageable or pinned host memory. */
int rheo_gpu_upload_velocity(RheoGpu* h, const double* U, const double* U_b, const double* phi);

/* Caller-supplied velocity gradient: correct(alpha, gradU) with gradU != nullptr (constitutiveEq.H:346-350; utils/boilerLog.H:1
 * `L(gradU == nullptr ? fvc::grad(U)() : *gradU)`, as filmModel.C:408 calls it).  gradU9[9*n_cells] in OpenFOAM's tensor order
 * (L_ij = d_i U_j at 3i+j) replaces the device's own Gauss-linear grad(U) in every following step; NULL returns to fvc::grad(U).
 * (`alpha` needs no entry point: no *Log model reads it inside correct().) */
int rheo_gpu_upload_grad_u(RheoGpu* h, const double* gradU9);

/* Temperature-dependent relaxation time and polymer viscosity (Oldroyd_BLog.C:133-135 and the same lines of GiesekusLog,
 * PTTLog, FENE-PLog, FENE-CRLog, WhiteMetznerCYLog: `lambda = thermoLambdaPtr_->createField(lambda_)`): per-cell values
 * lambda_cell[n_cells], etaP_cell[n_cells] in the caller's numbering replace the scalars of RheoModelDesc for `mode`; the caller
 * evaluates the thermoFunction (include/rheo_mesh.h: rheo_thermo_factor restates Arrhenius / ArrheniusModified / WLF / VFT).
 * NULL, NULL restores the scalar parameters.  Call again whenever T has changed. */
int rheo_gpu_upload_thermo(RheoGpu* h, int32_t mode, const double* lambda_cell, const double* etaP_cell);

int rheo_gpu_store_old_time(RheoGpu* h);
/* Off by default.  Alternative reading of `tau_ = ...` (Oldroyd_BLog.C:176 and the same line of the other models) before
 * tau_.correctBoundaryConditions(): GeometricField::operator= leaves the boundary value of the right-hand expression on every
 * non-fixed tau patch (0, or -etaP/lambda I for Oldroyd-BLog), which a linearExtrapolation patch listed EARLIER then sees
 * (DESIGN.md section 6).  Switch on to reproduce oracle/_ref on meshes that list wall patches before zeroGradient patches. */
int rheo_gpu_set_tau_assignment(RheoGpu* h, int32_t on);

/* One constitutiveEq::correct() for all modes, inputs already resident in HBM.
 * stats: array of n_modes entries or NULL (NULL avoids the final device->host read). */
int rheo_gpu_step(RheoGpu* h, double dt, RheoStepStats* stats);

int rheo_gpu_download(RheoGpu* h, int32_t mode, int32_t field, double* dst);

/* Host-buffer convenience = upload_velocity + store_old_time (if new_time_step) + step +
 * download(TAU_TOTAL) (+ TAU_B_TOTAL, the boundary stress summed over the modes, when tau_b != NULL): what the reference plugin call costs
 * when the momentum predictor stays on the CPU. */
int rheo_gpu_correct(RheoGpu* h, const double* U, const double* U_b, const double* phi, double dt,
                     int32_t new_time_step, double* tau_out, double* tau_b_out, RheoStepStats* stats);

/* ---- introspection used by the parity tests and bench.py ---- */
/* renumbering actually used on the device: perm[new] = old cell, n_colours, colour_start[n_colours+1] */
/* Explicit part of constitutiveEq::divTau(U) (constitutiveEq/constitutiveEq.C:72-132, summed over the modes as
 * multiMode/multiMode.C:143-157 does), evaluated on the device from the stress of the last step and the velocity last uploaded:
 *   div_out[3*n_cells] = sum_modes fvc::div(tau_m/rho_m)  -  (RHEO_STAB_COUPLING only) fvc::div((etaP_m/rho_m) fvc::grad(U)),
 * both `Gauss linear` (csrc/gpu/momentum.cuh).  The caller adds fvm::laplacian((etaP+etaS)/rho, U) (and, for RHEO_STAB_BSD,
 * subtracts its own fvc::laplacian(etaP/rho, U)): with this call tau itself only leaves the device at write time.
 * Collective over the ranks of a decomposed case (the velocity gradient of the ghost cells is swapped). */
int rheo_gpu_div_tau(RheoGpu* h, int32_t stabilization, double* div_out);

int rheo_gpu_get_renumbering(RheoGpu* h, int32_t* perm, int32_t* n_colours, int32_t* colour_start);
/* one line of text naming the cell ordering the DILU substitutions run in on this handle, e.g.
 * "8x8x4 blocks (256 cells), natural order inside, 2 block colours" or "cell colouring, 2 colours" */
int rheo_gpu_get_ordering(RheoGpu* h, char* buf, int32_t buflen);
/* block ordering only: levels of every cell (new numbering) in the dependency graph of its 256-cell chunk — longest chain of
 * lower- (fwd) / higher-numbered (bwd) neighbours inside the chunk; the level-scheduled substitutions run in this order */
int rheo_gpu_get_levels(RheoGpu* h, int32_t* fwd, int32_t* bwd);
/* ELL width K and the neighbour table nbr[K*n_cells] (slot-major, in NEW numbering: >=0 cell,
 * >= n_cells ghost, -1 empty, <=-2 boundary face -(b+2)) and face table (face index, ~face when
 * the cell is the face's neighbour) */
int rheo_gpu_get_ell(RheoGpu* h, int32_t* K, int32_t* nbr, int32_t* face);
/* kernels launched by this handle so far (all of them are this library's own) */
int64_t rheo_gpu_launch_count(const RheoGpu* h);
/* Krylov iterations (max over components and modes) of the last step */
int rheo_gpu_last_iterations(const RheoGpu* h);
/* multi-GPU path in use: mode 0 single rank, 1 NCCL send/recv + all-reduce, 2 NVLink peer-memory mailboxes (peer.cuh);
 * for mode 2 the device time rank-local thread 0 has spent waiting for [0] neighbours' halo records and [1] the other
 * ranks' partial sums since comm_init (ms), and the number of such waits — the synchronisation cost of the decomposition */
int rheo_gpu_comm_stats(RheoGpu* h, int32_t* mode, double* wait_ms2, int64_t* waits2);
/* bytes copied host->device / device->host by the upload/download/correct entry points of this handle so far */
int rheo_gpu_transfer_bytes(const RheoGpu* h, int64_t* h2d, int64_t* d2h);
/* per-phase device times of the last step measured with CUDA events when enabled (ms):
 * [0] halo+bc  [1] grad(theta)  [2] assemble  [3] solve  [4] eig+tau  [5] tau bc  [6] total */
int rheo_gpu_set_phase_timing(RheoGpu* h, int32_t enabled);
int rheo_gpu_get_phase_times(RheoGpu* h, double* ms7);
/* per-kernel CUDA-event timing (serialises the launches; used by bench.py's roofline pass only).
 * get: text lines "<kernel> <launches> <total ms>" */
int rheo_gpu_set_kernel_timing(RheoGpu* h, int32_t enabled);
int rheo_gpu_get_kernel_times(RheoGpu* h, char* buf, int32_t buflen);
/* device buffers for callers that keep U/phi on the GPU (SoA, renumbered) */
int rheo_gpu_stream(RheoGpu* h, void** cuda_stream);
int rheo_gpu_synchronize(RheoGpu* h);

/* stand-alone per-cell kernels on host arrays (n cells, AoS) — used by the unit parity tests */
int rheo_gpu_eig_exp(int32_t device, int32_t n, const double* theta6, double* eigvals9, double* eigvecs9);

const char* rheo_gpu_last_error(void);

/* sizeof() of the public structs as compiled into the library, in the order RheoPatchDesc, RheoMeshDesc,
 * RheoModelDesc, RheoSchemeCtl, RheoStepStats, RheoSynthSpec, RheoPatchRule, RheoPatchSpec — lets a binding
 * (ctypes here, cgo/JNI elsewhere) verify its mirror of the layouts before the first call. */
int rheo_gpu_abi_sizes(int32_t* out8);

#ifdef __cplusplus
}
#endif
#endif /* RHEO_GPU_H */
