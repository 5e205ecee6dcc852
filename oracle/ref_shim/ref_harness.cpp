// ref_harness.cpp — C entry points of oracle/_ref/libref_stress.so (TEST INFRASTRUCTURE ONLY).
//
// Runs the reference's OWN text for the stress step (see ref_ce.H for the list of files and line ranges)
// on a mesh given as a RheoMeshDesc (ref_correct), or on R decomposed meshes with processor patches, one thread per rank
// (ref_correct_multi), so that tests/ can pin oracle/oracle.cpp against it.  What is NOT the
// reference's here, and therefore not pinned by it: the OpenFOAM-9 layer (minifoam.H), the linear solver
// (a Jacobi iteration to round-off — the pinned quantity is the solution of the assembled system, not an
// iteration history) and the eigen-solver call (Eigen 3.2.9 is absent; the reference's own jacobi.H is
// called instead, its documented alternative at constitutiveEq.C:418-426).
#define NoRepository
#include "ref_ce.H"
#include "gaussDefCmpwConvectionScheme.H"   // the reference's (with its .C and limiters.H through NoRepository)

#include "../../include/rheo_gpu.h"

#include <cstring>
#include <thread>

namespace Foam
{

std::string refHarness::limiterName = "cubista";

// ---- reference text: constitutiveEq::decomposeGradU, constitutiveEq::innerP (both overloads) --------
#include "constitutiveEq_decomposeGradU.inc"
#include "constitutiveEq_innerP.inc"

void constitutiveEq::calcEig(const volSymmTensorField& theta, volTensorField& vals, volTensorField& vecs)
{
    forAll(theta, cellI)
    {
        int N = 3;
        int NROT = 0;
        vals[cellI] *= 0.;
        jacobi(theta[cellI], N, vals[cellI], vecs[cellI], NROT);
    }
}

tmp<fvSymmTensorMatrix> fvm::div(const surfaceScalarField& phi, const volSymmTensorField& vf)
{
    Istream is(refHarness::limiterName);
    fv::gaussDefCmpwConvectionScheme<symmTensor> scheme(vf.mesh(), phi, is);
    return scheme.fvmDiv(phi, vf);
}

tmp<fvScalarMatrix> fvm::div(const surfaceScalarField& phi, const volScalarField& vf)
{
    Istream is(refHarness::limiterName);
    fv::gaussDefCmpwConvectionScheme<scalar> scheme(vf.mesh(), phi, is);
    return scheme.fvmDiv(phi, vf);
}

// ---- linearExtrapolation: class shells + the reference's updateCoeffs() text ------------------------
template<class Type>
class fixedValueFvPatchField
{
public:
    typedef GeometricField<Type, fvPatchField, volMesh> VolField;
    VolField& gf_;
    label patchi_;
    bool updated_;
    fixedValueFvPatchField(VolField& gf, label patchi) : gf_(gf), patchi_(patchi), updated_(false) {}
    bool updated() const { return updated_; }
    const VolField& internalField() const { return gf_; }
    const fvPatch& patch() const { return gf_.boundaryField()[patchi_].patch(); }
    const objectRegistry& db() const { return gf_.mesh(); }
    void operator==(const Field<Type>& v) { static_cast<Field<Type>&>(gf_.boundaryFieldRef()[patchi_]) = v; }
    void updateCoeffs() { updated_ = true; }
};

template<class Type>
class linearExtrapolationFvPatchField : public fixedValueFvPatchField<Type>
{
public:
    bool useReg_;
    linearExtrapolationFvPatchField(GeometricField<Type, fvPatchField, volMesh>& gf, label patchi, bool useReg)
    : fixedValueFvPatchField<Type>(gf, patchi), useReg_(useReg) {}
    void updateCoeffs();
};

#include "linearExtrapolation_updateCoeffs.inc"

namespace refHarness
{
bool useRegression = false;
// BMPLog: fluidity per cell (+ per boundary face).  fluiditySolve == 0: Phi AFTER PhiEqn.solve() is handed in and the text without
// the fluidity equation runs (BMPLog); != 0: Phi BEFORE the call is handed in, the whole correct() runs (BMPLogFull) and the
// new fluidity is written back into the same arrays
thread_local double* fluidity = nullptr;
thread_local double* fluidityB = nullptr;
thread_local int fluiditySolve = 0;
}

// GeometricField::Boundary::evaluate (EXT-OF9): patches in order, each patch field's evaluate()
template<class Type, template<class> class PatchField, class GeoMesh>
void GeometricField<Type, PatchField, GeoMesh>::correctBoundaryConditions()
{
    if constexpr (std::is_same<GeoMesh, volMesh>::value)
    {
        // non-blocking processor patches: every send and receive completes before the first evaluate()
        evaluateCoupled(*this);
        forAll(boundary_, patchi)
        {
            auto& pf = boundary_[patchi];
            if (pf.kind_ == pfZeroGradient)
            {
                static_cast<Field<Type>&>(pf) = pf.patchInternalField();
            }
            else if (pf.kind_ == pfLinearExtrapolation || pf.kind_ == pfLinearExtrapolationReg)
            {
                if constexpr (std::is_same<Type, symmTensor>::value)
                {
                    // useRegression (linearExtrapolationFvPatchField.C:72): per patch (RHEO_BC_LINEAR_EXTRAPOLATION_REG) or for the whole call
                    linearExtrapolationFvPatchField<Type> bc(*this, patchi, refHarness::useRegression || pf.kind_ == pfLinearExtrapolationReg);
                    bc.updateCoeffs();
                }
            }
        }
    }
}

template void GeometricField<symmTensor, fvPatchField, volMesh>::correctBoundaryConditions();
template void GeometricField<scalar, fvPatchField, volMesh>::correctBoundaryConditions();

// ---- RheoMeshDesc -> fvMesh ---------------------------------------------------------------------------
static inline vector v3(const double* p, label i) { return vector(p[3*i], p[3*i + 1], p[3*i + 2]); }

static void buildMesh(fvMesh& m, const RheoMeshDesc* d)
{
    m.nCells_ = d->n_cells;
    m.nFaces_ = d->n_faces;
    m.nInternalFaces_ = d->n_internal_faces;
    const label nif = d->n_internal_faces;
    m.owner_.resize(nif); m.neighbour_.resize(nif); m.Sf_.resize(nif); m.Cf_.resize(nif); m.weights_.resize(nif);
    m.faceOwner_.resize(d->n_faces); m.allSf_.resize(d->n_faces); m.allCf_.resize(d->n_faces);
    for (label f = 0; f < d->n_faces; f++)
    {
        m.faceOwner_[f] = d->owner[f];
        m.allSf_[f] = v3(d->Sf, f);
        m.allCf_[f] = v3(d->Cf, f);
    }
    for (label f = 0; f < nif; f++)
    {
        m.owner_[f] = d->owner[f]; m.neighbour_[f] = d->neighbour[f];
        m.Sf_[f] = m.allSf_[f]; m.Cf_[f] = m.allCf_[f]; m.weights_[f] = d->weights[f];
    }
    m.C_.resize(d->n_cells); m.V_.resize(d->n_cells);
    for (label c = 0; c < d->n_cells; c++) { m.C_[c] = v3(d->C, c); m.V_[c] = d->V[c]; }
    for (int k = 0; k < 6; k++) m.valid_[k] = d->solved_components[k] != 0;
    m.boundary_.mesh_ = &m;
    m.boundary_.resize(d->n_patches);
    for (label p = 0; p < d->n_patches; p++)
    {
        const RheoPatchDesc& pd = d->patches[p];
        fvPatch& fp = m.boundary_[p];
        fp.bm_ = &m.boundary_; fp.index_ = p; fp.start_ = pd.start; fp.kind_ = pd.type;
        fp.coupled_ = pd.type == RHEO_PATCH_PROCESSOR;
        fp.nbrRank_ = pd.nbr_rank;
        if (fp.coupled_ && !refHarness::world())
            FatalError << "processor patches need ref_correct_multi (one thread per rank)" << abort(FatalError);
        if (pd.type == RHEO_PATCH_EMPTY) continue;   // emptyFvPatch::size() == 0
        for (label i = 0; i < pd.size; i++)
        {
            const label f = pd.start + i;
            fp.faceCells_.append(d->owner[f]);
            fp.Cf_.append(m.allCf_[f]);
            fp.Sf_.append(m.allSf_[f]);
            if (fp.coupled_)   // processorFvPatch::delta(): neighbour cell centre - own cell centre; linear weight of the owner side
            {
                fp.delta_.append(v3(d->nbr_C, f - d->n_internal_faces) - m.C_[d->owner[f]]);
                fp.weights_.append(d->weights[f]);
            }
            else fp.delta_.append(m.allCf_[f] - m.C_[d->owner[f]]);
        }
    }
}

static int bcKind(int bc)
{
    switch (bc)
    {
        case RHEO_BC_FIXED_VALUE: return pfFixedValue;
        case RHEO_BC_ZERO_GRADIENT: return pfZeroGradient;
        case RHEO_BC_LINEAR_EXTRAPOLATION: return pfLinearExtrapolation;
        case RHEO_BC_LINEAR_EXTRAPOLATION_REG: return pfLinearExtrapolationReg;
        case RHEO_BC_EMPTY: return pfEmpty;
        case RHEO_BC_PROCESSOR: return pfProcessor;
    }
    FatalError << "patch field kind not available in the reference harness" << abort(FatalError);
    return pfCalculated;
}

template<class Type, template<class> class PF, class GM>
static void loadField(GeometricField<Type, PF, GM>& gf, const double* internal, const double* boundary,
                      const RheoMeshDesc* d)
{
    const int nc = pTraits<Type>::nComponents;
    if (internal) forAll(gf, i) for (int k = 0; k < nc; k++) setCmpt(gf[i], k, internal[nc*i + k]);
    if (boundary) forAll(gf.boundaryField(), p)
    {
        auto& pf = gf.boundaryFieldRef()[p];
        forAll(pf, i)
        {
            const label b = d->patches[p].start + i - d->n_internal_faces;
            for (int k = 0; k < nc; k++) setCmpt(pf[i], k, boundary[nc*b + k]);
        }
    }
}
template<class Type, template<class> class PF, class GM>
static void storeField(const GeometricField<Type, PF, GM>& gf, double* internal, double* boundary,
                       const RheoMeshDesc* d)
{
    const int nc = pTraits<Type>::nComponents;
    if (internal) forAll(gf, i) for (int k = 0; k < nc; k++) internal[nc*i + k] = cmptOf(gf[i], k);
    if (boundary) forAll(gf.boundaryField(), p)
    {
        const auto& pf = gf.boundaryField()[p];
        forAll(pf, i)
        {
            const label b = d->patches[p].start + i - d->n_internal_faces;
            for (int k = 0; k < nc; k++) boundary[nc*b + k] = cmptOf(pf[i], k);
        }
    }
}

static const char* limiterWord(int l)
{
    static const char* names[] = {"upwind", "cubista", "minmod", "smart", "waceb", "superbee", "none"};
    return (l >= 0 && l < 7) ? names[l] : "?";
}

}  // namespace Foam

using namespace Foam;

extern "C" {

// utils/jacobi.H on n symmetric tensors: expD3 = exp(eigenvalue) in the order jacobi leaves them, V9 = its
// eigenvector tensor (row-major; column k belongs to expD3[k]), nrot = Jacobi rotations used
void ref_jacobi(int n, const double* theta6, double* expD3, double* V9, int* nrot)
{
    for (int c = 0; c < n; c++)
    {
        symmTensor t;
        for (int k = 0; k < 6; k++) t[k] = theta6[6*c + k];
        tensor D, V;
        int N = 3, NROT = 0;
        jacobi(t, N, D, V, NROT);
        expD3[3*c] = D.xx(); expD3[3*c + 1] = D.yy(); expD3[3*c + 2] = D.zz();
        for (int k = 0; k < 9; k++) V9[9*c + k] = V[k];
        if (nrot) nrot[c] = NROT;
    }
}

// limiters.H: the (alpha, beta, bounds) rows of a limiter; returns the number of alpha entries
int ref_lims(int limiter, double* alpha3, double* beta3, double* bounds2)
{
    fvMesh mesh;
    mesh.nCells_ = mesh.nFaces_ = mesh.nInternalFaces_ = 0;
    mesh.boundary_.mesh_ = &mesh;
    surfaceScalarField phi(IOobject("phi"), mesh, dimensionSet());
    volSymmTensorField vf(IOobject("theta"), mesh, dimensionSet());
    Istream is(limiterWord(limiter));
    fv::gaussDefCmpwConvectionScheme<symmTensor> scheme(mesh, phi, is);
    scalarList a, b, bo;
    scheme.lims(a, b, bo, phi, vf);
    for (size_t i = 0; i < a.size() && i < 3; i++) alpha3[i] = a[i];
    for (size_t i = 0; i < b.size() && i < 3; i++) beta3[i] = b[i];
    for (size_t i = 0; i < bo.size() && i < 2; i++) bounds2[i] = bo[i];
    return int(a.size());
}

// constitutiveEq::decomposeGradU (+ innerP) on n cells: M, eigVals, eigVecs -> omega, B (tensors, row-major)
void ref_decompose_gradU(int n, const double* M9, const double* eigVals9, const double* eigVecs9,
                         double* omega9, double* B9)
{
    fvMesh mesh;
    mesh.nCells_ = n; mesh.nFaces_ = mesh.nInternalFaces_ = 0;
    mesh.boundary_.mesh_ = &mesh;
    volVectorField U(IOobject("U"), mesh, dimensionSet());
    surfaceScalarField phi(IOobject("phi"), mesh, dimensionSet());
    volTensorField M(IOobject("M"), mesh, dimensionSet()), vals(M), vecs(M), omega(M), B(M);
    for (int c = 0; c < n; c++) for (int k = 0; k < 9; k++)
    {
        M[c][k] = M9[9*c + k]; vals[c][k] = eigVals9[9*c + k]; vecs[c][k] = eigVecs9[9*c + k];
    }
    constitutiveEq ce(U, phi);
    ce.decomposeGradU(M, vals, vecs, omega, B);
    for (int c = 0; c < n; c++) for (int k = 0; k < 9; k++) { omega9[9*c + k] = omega[c][k]; B9[9*c + k] = B[c][k]; }
}

// constitutiveEq::innerP on n cells
void ref_innerP(int n, const double* t1, const double* t2, int isFirstT, double* out9)
{
    fvMesh mesh;
    mesh.nCells_ = n; mesh.nFaces_ = mesh.nInternalFaces_ = 0;
    mesh.boundary_.mesh_ = &mesh;
    volVectorField U(IOobject("U"), mesh, dimensionSet());
    surfaceScalarField phi(IOobject("phi"), mesh, dimensionSet());
    volTensorField A(IOobject("A"), mesh, dimensionSet()), Bf(A);
    for (int c = 0; c < n; c++) for (int k = 0; k < 9; k++) { A[c][k] = t1[9*c + k]; Bf[c][k] = t2[9*c + k]; }
    constitutiveEq ce(U, phi);
    volTensorField r(ce.innerP(A, Bf, isFirstT != 0));
    for (int c = 0; c < n; c++) for (int k = 0; k < 9; k++) out9[9*c + k] = r[c][k];
}

// One XxxLog::correct() of the reference on a single-rank mesh.
//   theta, eigvals, eigvecs, tau (+ boundary values theta_b, tau_b; boundary arrays are indexed by
//   face - n_internal_faces like everywhere in this repo): state before the call in, state after the call out.
//   theta is also the old-time level (the call is the first correct() of a time step).
//   asm_lower/upper [n_internal_faces], asm_diag [n_cells], asm_source [6*n_cells], asm_iC/asm_bC [6*n_bfaces]:
//   optional (NULL) — the thetaEqn the reference assembled, as handed to solve().
// Returns 0, or -1 for a model the harness has no reference text for.
struct RankArrays
{
    const double *U, *U_b, *phi;
    double *theta, *theta_b, *tau, *tau_b, *eigvals, *eigvecs;
    double *asm_lower, *asm_upper, *asm_diag, *asm_source, *asm_iC, *asm_bC;
};

static int correctOnMesh(fvMesh& mesh, const RheoMeshDesc* md, const RheoModelDesc* mm, double dt, const RankArrays& a)
{
    const double *U = a.U, *U_b = a.U_b, *phi = a.phi;
    double *theta = a.theta, *theta_b = a.theta_b, *tau = a.tau, *tau_b = a.tau_b, *eigvals = a.eigvals, *eigvecs = a.eigvecs;
    double *asm_lower = a.asm_lower, *asm_upper = a.asm_upper, *asm_diag = a.asm_diag, *asm_source = a.asm_source, *asm_iC = a.asm_iC,
           *asm_bC = a.asm_bC;
    mesh.time_.deltaT_ = dt;

    volVectorField Uf(IOobject("U"), mesh, dimensionSet());
    surfaceScalarField phif(IOobject("phi"), mesh, dimensionSet());
    volSymmTensorField thetaf(IOobject("theta"), mesh, dimensionSet());
    volSymmTensorField tauf(IOobject("tau"), mesh, dimensionSet());
    volTensorField valsf(IOobject("eigVals"), mesh, dimensionedTensor("I", dimless, tensor::I),
                         extrapolatedCalculatedFvPatchField<tensor>::typeName);
    volTensorField vecsf(valsf);
    for (label p = 0; p < md->n_patches; p++)
    {
        const bool empty = md->patches[p].type == RHEO_PATCH_EMPTY;
        const bool proc = md->patches[p].type == RHEO_PATCH_PROCESSOR;
        Uf.setPatchKind(p, empty ? pfEmpty : proc ? pfProcessor : pfFixedValue);
        thetaf.setPatchKind(p, empty ? pfEmpty : bcKind(md->patches[p].theta_bc));
        tauf.setPatchKind(p, empty ? pfEmpty : bcKind(md->patches[p].tau_bc));
    }
    loadField(Uf, U, U_b, md);
    loadField(thetaf, theta, theta_b, md);
    loadField(tauf, tau, tau_b, md);
    loadField(valsf, eigvals, (const double*) nullptr, md);
    loadField(vecsf, eigvecs, (const double*) nullptr, md);
    forAll(phif, f) phif[f] = phi[f];
    forAll(phif.boundaryField(), p) forAll(phif.boundaryField()[p], i)
        phif.boundaryFieldRef()[p][i] = phi[md->patches[p].start + i];
    // processor patches hold the neighbour cells' values (processorFvPatchField after evaluate())
    evaluateCoupled(Uf); evaluateCoupled(thetaf); evaluateCoupled(tauf);
    thetaf.storeOldTime();
    tauf.store();      // linearExtrapolation looks tau up by name in the registry
    thetaf.store();

    auto run = [&](auto& model)
    {
        model.rho_ = dimensionedScalar("rho", mm->rho);
        model.etaS_ = dimensionedScalar("etaS", mm->etaS);
        model.etaP_ = dimensionedScalar("etaP", mm->etaP);
        model.lambda_ = dimensionedScalar("lambda", mm->lambda);
        model.alpha_ = dimensionedScalar("alpha", mm->model == RHEO_MODEL_PTT_LOG ? mm->ml_alpha : mm->alpha);
        model.beta_ = dimensionedScalar("beta", mm->ml_beta);
        model.epsilon_ = dimensionedScalar("epsilon", mm->epsilon);
        model.zeta_ = dimensionedScalar("zeta", mm->zeta);
        model.L2_ = dimensionedScalar("L2", mm->L2);
        // WhiteMetznerCYLog requires m = n, L = K, b = a (WhiteMetznerCYLog.C:132-140); the descriptor carries one set
        model.K_ = model.L_ = dimensionedScalar("K", mm->wm_K);
        model.n_ = model.m_ = dimensionedScalar("n", mm->model == RHEO_MODEL_XPOMPOM_LOG ? mm->xpp_n : mm->wm_n);
        model.a_ = model.b_ = dimensionedScalar("a", mm->wm_a);
        model.lambdaR_ = dimensionedScalar("lambdaR", mm->rp_lambdaR);
        model.lambdaD_ = dimensionedScalar("lambdaD", mm->lambda);
        model.chiMax_ = dimensionedScalar("chiMax", mm->rp_chiMax);
        model.delta_ = dimensionedScalar("delta", mm->rp_delta);
        if (mm->model == RHEO_MODEL_ROLIE_POLY_LOG) model.beta_ = dimensionedScalar("beta", mm->rp_beta);
        model.lambdaS_ = dimensionedScalar("lambdaS", mm->xpp_lambdaS);
        model.lambdaB_ = dimensionedScalar("lambdaB", mm->lambda);
        model.q_ = dimensionedScalar("q", mm->xpp_q);
        if (mm->model == RHEO_MODEL_SARAMITO_LOG)
        {
            // SaramitoLog.C:113-130 (constructor)
            model.tau0_ = dimensionedScalar("tau0", mm->sar_tau0);
            model.n_ = dimensionedScalar("n", mm->sar_n);
            model.k_ = dimensionedScalar("k", mm->sar_k);
            model.nDims = scalar(mm->sar_dims[0] + mm->sar_dims[1] + mm->sar_dims[2]);
            model.ItensorCorr = dimensionedSymmTensor("Identity", symm(tensor::I));
            model.ItensorCorr.value().xx() = mm->sar_dims[0];
            model.ItensorCorr.value().yy() = mm->sar_dims[1];
            model.ItensorCorr.value().zz() = mm->sar_dims[2];
            model.funcPTT = mm->sar_ptt;
        }
        model.correct();
    };
    switch (mm->model)
    {
        case RHEO_MODEL_OLDROYD_B_LOG: { constitutiveEqs::Oldroyd_BLog m(Uf, phif, tauf, thetaf, valsf, vecsf); run(m); break; }
        case RHEO_MODEL_GIESEKUS_LOG: { constitutiveEqs::GiesekusLog m(Uf, phif, tauf, thetaf, valsf, vecsf); run(m); break; }
        case RHEO_MODEL_FENE_P_LOG: { constitutiveEqs::FENE_PLog m(Uf, phif, tauf, thetaf, valsf, vecsf); run(m); break; }
        case RHEO_MODEL_FENE_CR_LOG: { constitutiveEqs::FENE_CRLog m(Uf, phif, tauf, thetaf, valsf, vecsf); run(m); break; }
        case RHEO_MODEL_WM_CY_LOG: { constitutiveEqs::WhiteMetznerCYLog m(Uf, phif, tauf, thetaf, valsf, vecsf); run(m); break; }
        case RHEO_MODEL_ROLIE_POLY_LOG: { constitutiveEqs::RoliePolyLog m(Uf, phif, tauf, thetaf, valsf, vecsf); run(m); break; }
        case RHEO_MODEL_XPOMPOM_LOG: { constitutiveEqs::XPomPomLog m(Uf, phif, tauf, thetaf, valsf, vecsf); run(m); break; }
        case RHEO_MODEL_SARAMITO_LOG: { constitutiveEqs::SaramitoLog m(Uf, phif, tauf, thetaf, valsf, vecsf); run(m); break; }
        case RHEO_MODEL_BMP_LOG:
        {
            // the theta equation and theta -> tau of BMPLog::correct (BMPLog.C:168-199) with the fluidity handed in (ref_set_fluidity)
            if (!refHarness::fluidity) return -1;
            volScalarField Phif(IOobject("Phi"), mesh, dimensionSet());
            if (!refHarness::fluiditySolve)
            {
                forAll(Phif, c) Phif[c] = refHarness::fluidity[c];
                constitutiveEqs::BMPLog m(Uf, phif, tauf, thetaf, valsf, vecsf);
                m.PhiPtr_ = &Phif;
                m.G0_ = dimensionedScalar("G0", mm->bmp_G0);
                run(m);
                break;
            }
            // the whole BMPLog::correct (BMPLog.C:142-201): Phi carries theta's patch kinds (fixedValue inlet, zeroGradient elsewhere)
            for (label p = 0; p < md->n_patches; p++)
                Phif.setPatchKind(p, md->patches[p].type == RHEO_PATCH_EMPTY ? pfEmpty : bcKind(md->patches[p].theta_bc));
            loadField(Phif, refHarness::fluidity, (const double*) refHarness::fluidityB, md);
            evaluateCoupled(Phif);
            Phif.storeOldTime();
            constitutiveEqs::BMPLogFull m(Uf, phif, tauf, thetaf, valsf, vecsf);
            m.PhiPtr_ = &Phif;
            m.G0_ = dimensionedScalar("G0", mm->bmp_G0);
            m.Phi0_ = dimensionedScalar("Phi0", mm->bmp_Phi0);
            m.PhiInf_ = dimensionedScalar("PhiInf", mm->bmp_PhiInf);
            m.k_ = dimensionedScalar("k", mm->bmp_k);
            run(m);
            storeField(Phif, refHarness::fluidity, refHarness::fluidityB, md);
            break;
        }
        case RHEO_MODEL_PTT_LOG:
        {
            constitutiveEqs::PTTLog m(Uf, phif, tauf, thetaf, valsf, vecsf);
            m.PTTFunction_ = mm->ptt_function == RHEO_PTT_LINEAR ? LogModelShell::pfLinear
                           : mm->ptt_function == RHEO_PTT_EXPONENTIAL ? LogModelShell::pfExpt : LogModelShell::pfGen;
            if (m.PTTFunction_ == LogModelShell::pfGen)
            {
                // PTTLog.C:144-170 (constructor): table of Gamma values for the Mittag-Leffler series
                m.MLrtol_ = mm->ml_rtol;
                m.MLmaxIter_ = mm->ml_max_iter;
                m.gammaFunValues_.append(tgamma(mm->ml_beta));
                int k(0);
                while (k < m.MLmaxIter_ && m.gammaFunValues_.last() < 1e+100)
                {
                    m.gammaFunValues_.append(tgamma(mm->ml_alpha*k + mm->ml_beta));
                    k++;
                }
                m.MLmaxIter_ = k;
            }
            run(m);
            break;
        }
        default: return -1;
    }

    storeField(thetaf, theta, theta_b, md);
    storeField(tauf, tau, tau_b, md);
    storeField(valsf, eigvals, (double*) nullptr, md);
    storeField(vecsf, eigvecs, (double*) nullptr, md);
    auto& last = fvMatrix<symmTensor>::lastSolved();
    if (last)
    {
        if (asm_lower) forAll(last->lower(), f) asm_lower[f] = last->lower()[f];
        if (asm_upper) forAll(last->upper(), f) asm_upper[f] = last->upper()[f];
        if (asm_diag) forAll(last->diag(), c) asm_diag[c] = last->diag()[c];
        if (asm_source) forAll(last->source(), c) for (int k = 0; k < 6; k++) asm_source[6*c + k] = last->source()[c][k];
        forAll(last->internalCoeffs(), p) forAll(last->internalCoeffs()[p], i)
        {
            const label b = md->patches[p].start + i - md->n_internal_faces;
            for (int k = 0; k < 6; k++)
            {
                if (asm_iC) asm_iC[6*b + k] = last->internalCoeffs()[p][i][k];
                if (asm_bC) asm_bC[6*b + k] = last->boundaryCoeffs()[p][i][k];
            }
        }
        last.reset();
    }
    return 0;
}


// BMPLog: the fluidity field (after PhiEqn.solve()) the next ref_correct of this thread uses; NULL clears it
void ref_set_fluidity(double* Phi, double* Phi_b, int solve) { refHarness::fluidity = Phi; refHarness::fluidityB = Phi_b; refHarness::fluiditySolve = solve; }

int ref_correct(const RheoMeshDesc* md, const RheoModelDesc* mm, int limiter, double dt, int use_regression,
                const double* U, const double* U_b, const double* phi,
                double* theta, double* theta_b, double* tau, double* tau_b, double* eigvals, double* eigvecs,
                double* asm_lower, double* asm_upper, double* asm_diag, double* asm_source,
                double* asm_iC, double* asm_bC)
{
    fvMesh mesh;
    buildMesh(mesh, md);
    refHarness::limiterName = limiterWord(limiter);
    refHarness::useRegression = use_regression != 0;
    const RankArrays a{U, U_b, phi, theta, theta_b, tau, tau_b, eigvals, eigvecs, asm_lower, asm_upper, asm_diag, asm_source, asm_iC, asm_bC};
    return correctOnMesh(mesh, md, mm, dt, a);
}

// The same on R sub-domain meshes with processor patches (decomposePar layout): one thread per rank, lock-step collectives in
// place of Pstream (minifoam.H: refHarness::World).  Arrays are given per rank: U[r], U_b[r], ... (boundary arrays indexed by
// face - n_internal_faces of that rank's mesh; values on processor faces are ignored on input).
int ref_correct_multi(int R, const RheoMeshDesc* const* md, const RheoModelDesc* mm, int limiter, double dt, int use_regression,
                      const double* const* U, const double* const* U_b, const double* const* phi,
                      double* const* theta, double* const* theta_b, double* const* tau, double* const* tau_b,
                      double* const* eigvals, double* const* eigvecs)
{
    refHarness::World world;
    world.R = R;
    world.mail.resize(R);
    world.red.assign(R, 0.0);
    refHarness::world() = &world;
    refHarness::limiterName = limiterWord(limiter);
    refHarness::useRegression = use_regression != 0;
    std::vector<fvMesh> meshes(R);
    for (int r = 0; r < R; r++)
    {
        buildMesh(meshes[r], md[r]);
        world.mail[r].resize(md[r]->n_patches);
    }
    for (int r = 0; r < R; r++)   // the patch of the neighbour rank that faces this one
        for (fvPatch& p : meshes[r].boundary_)
            if (p.coupled())
            {
                const fvBoundaryMesh& nb = meshes[p.nbrRank_].boundary_;
                for (size_t q = 0; q < nb.size(); q++) if (nb[q].coupled() && nb[q].nbrRank_ == r) p.nbrPatch_ = label(q);
                if (p.nbrPatch_ < 0 || nb[p.nbrPatch_].size() != p.size())
                    FatalError << "processor patches of ranks " << r << " and " << p.nbrRank_ << " do not match" << abort(FatalError);
            }
    std::vector<int> rc(R, 0);
    std::vector<std::thread> threads;
    for (int r = 0; r < R; r++)
        threads.emplace_back([&, r]()
        {
            refHarness::myRank() = r;
            const RankArrays a{U[r], U_b[r], phi[r], theta[r], theta_b[r], tau[r], tau_b[r], eigvals[r], eigvecs[r],
                               nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
            rc[r] = correctOnMesh(meshes[r], md[r], mm, dt, a);
        });
    for (auto& t : threads) t.join();
    refHarness::world() = nullptr;
    for (int r = 0; r < R; r++) if (rc[r]) return rc[r];
    return 0;
}

}  // extern "C"
