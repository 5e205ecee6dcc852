/*---------------------------------------------------------------------------*\
  of90_LogConformationGPU.C — see the header.  OpenFOAM-9 + rheoTool only.

  Registered type names (constant/constitutiveProperties -> parameters -> type):
      Oldroyd-BLogGPU  GiesekusLogGPU  PTTLogGPU  FENE-PLogGPU  FENE-CRLogGPU
      WhiteMetznerCYLogGPU  Rolie-PolyLogGPU  XPomPomLogGPU  SaramitoLogGPU  BMPLogGPU  multiModeLogGPU
  Dictionary keys are those of the CPU models (Oldroyd_BLog.C:114-119,
  GiesekusLog.C:117, PTTLog.C:129-139, FENE_PLog.C:114-119, multiMode.C:73-92);
  fvSchemes div(phi,theta<name>) must be `GaussDefCmpw <limiter>`, ddtSchemes Euler or backward,
  fvSolution solvers.theta<name> supplies tolerance/relTol/minIter/maxIter.
\*---------------------------------------------------------------------------*/
#include "of90_LogConformationGPU.H"
#include "addToRunTimeSelectionTable.H"
#include "processorFvPatch.H"
#include "emptyFvPatch.H"
#include "wallFvPatch.H"
#include "fixedValueFvPatchFields.H"
#include "zeroGradientFvPatchFields.H"
#include "linearExtrapolationFvPatchField.H"
#include "extrapolatedCalculatedFvPatchFields.H"
#include "Pstream.H"

namespace Foam
{
namespace constitutiveEqs
{
    defineTypeNameAndDebug(LogConformationGPU, 0);

    // the same class under the five user-facing names
    #define RHEO_GPU_REGISTER(ClassAlias, UserName)                                              \
        class ClassAlias : public LogConformationGPU                                             \
        {                                                                                        \
        public:                                                                                  \
            TypeName(UserName);                                                                  \
            ClassAlias(const word& n, const volVectorField& U, const surfaceScalarField& phi,    \
                       const dictionary& d) : LogConformationGPU(n, U, phi, d) {}                \
        };                                                                                       \
        defineTypeNameAndDebug(ClassAlias, 0);                                                   \
        addToRunTimeSelectionTable(constitutiveEq, ClassAlias, dictionary);

    RHEO_GPU_REGISTER(Oldroyd_BLogGPU, "Oldroyd-BLogGPU")
    RHEO_GPU_REGISTER(GiesekusLogGPU, "GiesekusLogGPU")
    RHEO_GPU_REGISTER(PTTLogGPU, "PTTLogGPU")
    RHEO_GPU_REGISTER(FENE_PLogGPU, "FENE-PLogGPU")
    RHEO_GPU_REGISTER(FENE_CRLogGPU, "FENE-CRLogGPU")
    RHEO_GPU_REGISTER(WhiteMetznerCYLogGPU, "WhiteMetznerCYLogGPU")
    RHEO_GPU_REGISTER(RoliePolyLogGPU, "Rolie-PolyLogGPU")
    RHEO_GPU_REGISTER(XPomPomLogGPU, "XPomPomLogGPU")
    RHEO_GPU_REGISTER(SaramitoLogGPU, "SaramitoLogGPU")
    RHEO_GPU_REGISTER(BMPLogGPU, "BMPLogGPU")
    RHEO_GPU_REGISTER(multiModeLogGPU, "multiModeLogGPU")
}
}

using namespace Foam;
using namespace Foam::constitutiveEqs;

// * * * * * * * * * * * * * * * helpers * * * * * * * * * * * * * * * * * * //

void LogConformationGPU::check(int rc, const char* where) const
{
    if (rc)
    {
        // newConstitutiveEq.C:47-58 convention: fatal, no return codes above the C layer
        FatalErrorInFunction << where << ": " << rheo_gpu_last_error() << exit(FatalError);
    }
}

void LogConformationGPU::readMode(const word& type, const dictionary& dict, RheoModelDesc& m) const
{
    std::memset(&m, 0, sizeof(m));
    m.rho    = dimensionedScalar(dict.lookup("rho")).value();
    m.etaS   = dimensionedScalar(dict.lookup("etaS")).value();
    m.etaP   = dimensionedScalar(dict.lookup("etaP")).value();
    m.lambda = dict.found("lambda") ? dimensionedScalar(dict.lookup("lambda")).value() : 1;   // Rolie-Poly / XPomPom name theirs lambdaD / lambdaB
    m.L2 = 100; m.ml_alpha = 1; m.ml_beta = 1; m.ml_rtol = 1e-12; m.ml_max_iter = 200;
    if (type == "Oldroyd-BLogGPU" || type == "Oldroyd-BLog") m.model = RHEO_MODEL_OLDROYD_B_LOG;
    else if (type == "GiesekusLogGPU" || type == "GiesekusLog")
    {
        m.model = RHEO_MODEL_GIESEKUS_LOG;
        m.alpha = dimensionedScalar(dict.lookup("alpha")).value();
    }
    else if (type == "PTTLogGPU" || type == "PTTLog")
    {
        m.model   = RHEO_MODEL_PTT_LOG;
        m.epsilon = dimensionedScalar(dict.lookup("epsilon")).value();
        m.zeta    = dimensionedScalar(dict.lookup("zeta")).value();
        const word f(dict.lookup("destructionFunctionType"));      // PTTLog.C:41-50
        m.ptt_function = (f == "linear") ? RHEO_PTT_LINEAR : (f == "exponential") ? RHEO_PTT_EXPONENTIAL : RHEO_PTT_GENERALIZED;
        if (m.ptt_function == RHEO_PTT_GENERALIZED)                 // PTTLog.C:143-170
        {
            m.ml_alpha = dimensionedScalar(dict.lookup("alpha")).value();
            m.ml_beta  = dimensionedScalar(dict.lookup("beta")).value();
        }
    }
    else if (type == "FENE-PLogGPU" || type == "FENE-PLog")
    {
        m.model = RHEO_MODEL_FENE_P_LOG;
        m.L2 = dimensionedScalar(dict.lookup("L2")).value();
    }
    else if (type == "FENE-CRLogGPU" || type == "FENE-CRLog")      // FENE_CRLog.C:114-119
    {
        m.model = RHEO_MODEL_FENE_CR_LOG;
        m.L2 = dimensionedScalar(dict.lookup("L2")).value();
    }
    else if (type == "WhiteMetznerCYLogGPU" || type == "WhiteMetznerCYLog")
    {
        m.model = RHEO_MODEL_WM_CY_LOG;
        const scalar K = dimensionedScalar(dict.lookup("K")).value(), Lp = dimensionedScalar(dict.lookup("L")).value();
        const scalar n = dimensionedScalar(dict.lookup("n")).value(), mm = dimensionedScalar(dict.lookup("m")).value();
        const scalar a = dimensionedScalar(dict.lookup("a")).value(), b = dimensionedScalar(dict.lookup("b")).value();
        if (mm != n || K != Lp || a != b)                         // WhiteMetznerCYLog.C:132-140
        {
            FatalErrorInFunction << "The Log version of the WhiteMetznerCY model can only be used if:\n"
                << "\n   m=n   and   K=L   and   a=b\n" << abort(FatalError);
        }
        m.wm_K = K; m.wm_n = n; m.wm_a = a;
    }
    else if (type == "Rolie-PolyLogGPU" || type == "Rolie-PolyLog")      // RoliePolyLog.C:114-121
    {
        m.model = RHEO_MODEL_ROLIE_POLY_LOG;
        m.lambda     = dimensionedScalar(dict.lookup("lambdaD")).value();
        m.rp_lambdaR = dimensionedScalar(dict.lookup("lambdaR")).value();
        m.rp_beta    = dimensionedScalar(dict.lookup("beta")).value();
        m.rp_delta   = dimensionedScalar(dict.lookup("delta")).value();
        m.rp_chiMax  = dimensionedScalar(dict.lookup("chiMax")).value();
    }
    else if (type == "XPomPomLogGPU" || type == "XPomPomLog")            // XPomPomLog.C:115-122
    {
        m.model = RHEO_MODEL_XPOMPOM_LOG;
        m.lambda      = dimensionedScalar(dict.lookup("lambdaB")).value();
        m.xpp_lambdaS = dimensionedScalar(dict.lookup("lambdaS")).value();
        m.alpha       = dimensionedScalar(dict.lookup("alpha")).value();
        m.xpp_q       = dimensionedScalar(dict.lookup("q")).value();
        m.xpp_n       = dimensionedScalar(dict.lookup("n")).value();
    }
    else if (type == "BMPLogGPU" || type == "BMPLog")                    // BMPLog.C:129-136
    {
        m.model      = RHEO_MODEL_BMP_LOG;
        m.bmp_G0     = dimensionedScalar(dict.lookup("G0")).value();
        m.bmp_k      = dimensionedScalar(dict.lookup("k")).value();
        m.bmp_Phi0   = dimensionedScalar(dict.lookup("Phi0")).value();
        m.bmp_PhiInf = dimensionedScalar(dict.lookup("PhiInf")).value();
        // PhiEqn.relax() (BMPLog.C:165): relaxationFactors.equations.Phi<name>; its solver entry must equal theta's (the tutorial's
        // fvSolution groups "(theta|tau|C|A|Phi)"), and so must its convection scheme (div(phi,Phi): readSchemes checks both)
        m.bmp_relax  = 0;
    }
    else if (type == "SaramitoLogGPU" || type == "SaramitoLog")          // SaramitoLog.C:108-165
    {
        m.model    = RHEO_MODEL_SARAMITO_LOG;
        m.epsilon  = dimensionedScalar(dict.lookup("epsilon")).value();
        m.zeta     = dimensionedScalar(dict.lookup("zeta")).value();
        m.sar_tau0 = dimensionedScalar(dict.lookup("tau0")).value();
        m.sar_n    = dimensionedScalar(dict.lookup("n")).value();
        m.sar_k    = (m.sar_n == 1) ? m.etaP : dimensionedScalar(dict.lookup("k")).value();   // :116
        const vector dims(dict.lookup("dims"));
        m.sar_dims[0] = dims.x(); m.sar_dims[1] = dims.y(); m.sar_dims[2] = dims.z();
        m.sar_ptt = 0;
        if (m.sar_n == 1)                                                 // :133-165 (no PTT function when n != 1)
        {
            const word f(dict.lookup("PTTfunction"));
            if (f == "none") m.sar_ptt = 0;
            else if (f == "linear") m.sar_ptt = 1;
            else if (f == "exponential") m.sar_ptt = 2;
            else
            {
                FatalErrorInFunction << "\nThe PTT function specified does not exist.\n" << "\nAvailable PTT functions are:\n"
                    << "\n. none" << "\n. linear" << "\n. exponential" << abort(FatalError);
            }
        }
    }
    else
    {
        FatalErrorInFunction << "Unknown GPU constitutiveEq type " << type << exit(FatalError);
    }
    // thermoLambda / thermoEta sub-dictionaries (Oldroyd_BLog.C:118-119): the thermoFunction objects stay on the host — they
    // need the T field — and correct() uploads lambda(T), etaP(T) per cell (rheo_gpu_upload_thermo)
    thermoLambda_.append(thermoFunction::New("thermoLambda", U().mesh(), dict).ptr());
    thermoEta_.append(thermoFunction::New("thermoEta", U().mesh(), dict).ptr());
    if (dict.found("thermoLambda") || dict.found("thermoEta")) hasThermo_ = true;
}

void LogConformationGPU::readSchemes(const fvMesh& mesh, const word& thetaName)
{
    std::memset(&ctl_, 0, sizeof(ctl_));
    ITstream& div = mesh.divScheme("div(phi," + thetaName + ")");
    word w(div);
    if (w == "bounded") { ctl_.bounded = 1; w = word(div); }   // EXT-OF9 boundedConvectionScheme: fvmDiv - fvm::Sp(div(phi))
    if (w != "GaussDefCmpw") FatalErrorInFunction << "div(phi," << thetaName << ") must be GaussDefCmpw" << exit(FatalError);
    const word lim(div);
    const char* names[] = {"upwind", "cubista", "minmod", "smart", "waceb", "superbee", "none"};   // limiters.H:48-98
    ctl_.limiter = -1;
    for (int i = 0; i < 7; ++i) if (lim == names[i]) ctl_.limiter = i;
    const word ddtName(mesh.ddtScheme("ddt(" + thetaName + ")"));
    if (ddtName == "Euler") ctl_.ddt = RHEO_DDT_EULER;
    else if (ddtName == "backward") ctl_.ddt = RHEO_DDT_BACKWARD;
    else if (ddtName == "steadyState") ctl_.ddt = RHEO_DDT_STEADY_STATE;
    else if (ddtName == "CrankNicolson")
    {
        // `CrankNicolson <psi>`: the stream returned by ddtScheme() still holds the off-centring coefficient
        ITstream& is = mesh.ddtScheme("ddt(" + thetaName + ")");
        const word again(is);
        ctl_.ddt = RHEO_DDT_CRANK_NICOLSON;
        ctl_.cn_psi = is.eof() ? 1 : readScalar(is);
        if (ctl_.cn_psi < 0 || ctl_.cn_psi > 1) FatalErrorInFunction << "CrankNicolson coefficient = " << ctl_.cn_psi << " should be >= 0 and <= 1" << exit(FatalError);
    }
    else FatalErrorInFunction << "ddtSchemes Euler, backward, CrankNicolson and steadyState are available on the GPU path, not " << ddtName << exit(FatalError);
    // the device hard-codes Gauss linear for grad(U) (boilerLog.H:1), for the per-component grad(theta) of phifDefC
    // (gaussDefCmpwConvectionScheme.C:254) and for `linExtrapGrad` (linearExtrapolationFvPatchField.C:128)
    {
        const char* grads[] = {"grad(U)", "linExtrapGrad", "grad(theta)"};
        for (const char* gname : grads)
        {
            ITstream& gs = mesh.gradScheme(gname);
            const word g0(gs);
            const word g1(gs.eof() ? word("") : word(gs));
            if (g0 != "Gauss" || g1 != "linear")
            {
                FatalErrorInFunction << "gradSchemes entry for " << gname << " is `" << g0 << ' ' << g1
                    << "`; the GPU stress step implements `Gauss linear` gradients only" << exit(FatalError);
            }
        }
    }
    // deviceDivTau: div(tau) and div(grad(U)) (constitutiveEq.C:100-126) are evaluated on the device as `Gauss linear`
    if (deviceDivTau_)
    {
        const char* divs[] = {"div(tau)", "div(grad(U))"};
        for (const char* dname : divs)
        {
            if (word(dname) == "div(grad(U))" && stabOption_ != soCoupling) continue;
            ITstream& ds = mesh.divScheme(dname);
            const word d0(ds);
            const word d1(ds.eof() ? word("") : word(ds));
            if (d0 != "Gauss" || d1 != "linear")
            {
                FatalErrorInFunction << "divSchemes entry for " << dname << " is `" << d0 << ' ' << d1
                    << "`; the device divTau implements `Gauss linear` only (set deviceDivTau false)" << exit(FatalError);
            }
        }
    }
    const dictionary& sol = mesh.solverDict(thetaName);
    const word solver(sol.lookup("solver"));
    // the solver fvSolution names is the solver that runs: PBiCG (what every Log tutorial selects; pbicg.cuh) or PBiCGStab (the
    // tuned path: block ordering, fused kernels), each on any number of ranks
    if (solver == "PBiCGStab") ctl_.solver = RHEO_SOLVER_PBICGSTAB;
    else if (solver == "PBiCG") ctl_.solver = RHEO_SOLVER_PBICG;   // one rank or several (solve.inl: solve_batch_pbicg)
    else
    {
        FatalErrorInFunction << "fvSolution selects " << solver << " for " << thetaName
            << "; the GPU path implements PBiCG and PBiCGStab (+ DILU)" << exit(FatalError);
    }
    ctl_.tolerance = sol.lookupOrDefault<scalar>("tolerance", 1e-6);
    ctl_.rel_tol   = sol.lookupOrDefault<scalar>("relTol", 0);
    ctl_.min_iter  = sol.lookupOrDefault<label>("minIter", 0);
    ctl_.max_iter  = sol.lookupOrDefault<label>("maxIter", 1000);
    ctl_.relax     = mesh.relaxEquation(thetaName) ? mesh.equationRelaxationFactor(thetaName) : 0;
}

void LogConformationGPU::buildMeshDesc
(
    const fvMesh& mesh,
    List<int32_t>& owner, List<int32_t>& neighbour,
    List<RheoPatchDesc>& patches, vectorField& nbrC, scalarField& weights,
    RheoMeshDesc& d
) const
{
    const label nF = mesh.nFaces(), nI = mesh.nInternalFaces();
    owner.setSize(nF); neighbour.setSize(nI);
    forAll(owner, f) owner[f] = mesh.faceOwner()[f];
    forAll(neighbour, f) neighbour[f] = mesh.faceNeighbour()[f];
    weights.setSize(nF, 1.0);
    nbrC.setSize(nF - nI, vector::zero);
    const surfaceScalarField& w = mesh.weights();
    forAll(w, f) weights[f] = w[f];
    const volSymmTensorField& th = theta_[0];
    const volSymmTensorField& ta = tau_[0];
    patches.setSize(mesh.boundary().size());
    forAll(mesh.boundary(), pI)
    {
        const fvPatch& p = mesh.boundary()[pI];
        RheoPatchDesc& q = patches[pI];
        q.start = p.start(); q.size = p.size(); q.nbr_rank = -1;
        if (isA<processorFvPatch>(p))
        {
            q.type = RHEO_PATCH_PROCESSOR; q.theta_bc = q.tau_bc = RHEO_BC_PROCESSOR;
            q.nbr_rank = refCast<const processorFvPatch>(p).neighbProcNo();
            const vectorField nc(mesh.C().boundaryField()[pI].patchNeighbourField());
            forAll(nc, i) { nbrC[p.start() - nI + i] = nc[i]; weights[p.start() + i] = w.boundaryField()[pI][i]; }
        }
        else if (isA<emptyFvPatch>(p)) { q.type = RHEO_PATCH_EMPTY; q.theta_bc = q.tau_bc = RHEO_BC_EMPTY; }
        else
        {
            q.type = isA<wallFvPatch>(p) ? RHEO_PATCH_WALL : RHEO_PATCH_PATCH;
            const fvPatchSymmTensorField& tb = th.boundaryField()[pI];
            const fvPatchSymmTensorField& ub = ta.boundaryField()[pI];
            q.theta_bc = isA<zeroGradientFvPatchSymmTensorField>(tb) ? RHEO_BC_ZERO_GRADIENT : RHEO_BC_FIXED_VALUE;
            q.tau_bc = isA<linearExtrapolationFvPatchField<symmTensor>>(ub) ? RHEO_BC_LINEAR_EXTRAPOLATION
                     : isA<zeroGradientFvPatchSymmTensorField>(ub) ? RHEO_BC_ZERO_GRADIENT : RHEO_BC_FIXED_VALUE;
            if (q.tau_bc == RHEO_BC_LINEAR_EXTRAPOLATION)
            {
                // useReg_ is private (linearExtrapolationFvPatchField.H:82) but write() prints it (.C:230): `useRegression true`
                // selects the least-squares branch (.C:152-219) instead of the gradient branch (.C:101-151)
                OStringStream os;
                ub.write(os);
                const string txt(os.str());
                const auto at = txt.find("useRegression");
                if (at != string::npos)
                {
                    IStringStream is(txt.substr(at + 13));
                    const Switch sw(is);
                    if (sw) q.tau_bc = RHEO_BC_LINEAR_EXTRAPOLATION_REG;
                }
            }
        }
    }
    d.n_cells = mesh.nCells(); d.n_faces = nF; d.n_internal_faces = nI; d.n_patches = patches.size();
    d.owner = owner.begin(); d.neighbour = neighbour.begin();
    d.Sf = reinterpret_cast<const double*>(mesh.faceAreas().begin());      // vector == 3 contiguous doubles
    d.Cf = reinterpret_cast<const double*>(mesh.faceCentres().begin());
    d.C  = reinterpret_cast<const double*>(mesh.cellCentres().begin());
    d.V  = mesh.cellVolumes().begin();
    d.weights = weights.begin();
    d.nbr_C = reinterpret_cast<const double*>(nbrC.begin());
    d.patches = patches.begin();
    const Vector<label>& sd = mesh.solutionD();                            // validComponents<symmTensor>
    const int valid[6] = {sd.x() > 0, sd.x() > 0 && sd.y() > 0, sd.x() > 0 && sd.z() > 0, sd.y() > 0, sd.y() > 0 && sd.z() > 0, 1};
    for (int c = 0; c < 6; ++c) d.solved_components[c] = valid[c];
    // 2-D in (x,y): XX XY YY ZZ solved
    if (sd.z() < 0) { d.solved_components[2] = 0; d.solved_components[4] = 0; d.solved_components[5] = 1; }
}

// * * * * * * * * * * * * * * * * Constructor * * * * * * * * * * * * * * * * //

LogConformationGPU::LogConformationGPU
(
    const word& name,
    const volVectorField& U,
    const surfaceScalarField& phi,
    const dictionary& dict
)
:
    constitutiveEq(name, U, phi),
    tauTotal_
    (
        IOobject("tauGPUTotal" + name, U.time().timeName(), U.mesh(), IOobject::NO_READ, IOobject::NO_WRITE),
        U.mesh(),
        dimensionedSymmTensor("zero", dimensionSet(1, -1, -2, 0, 0, 0, 0), symmTensor::zero),
        extrapolatedCalculatedFvPatchField<symmTensor>::typeName
    ),
    rho_("rho", dimDensity, 0), etaS_("etaS", dimPressure*dimTime, 0), etaP_("etaP", dimPressure*dimTime, 0),
    gpu_(nullptr),
    lastTimeIndex_(-1),
    deviceDivTau_(dict.lookupOrDefault<Switch>("deviceDivTau", true))
{
    const fvMesh& mesh = U.mesh();
    const word type(dict.lookup("type"));
    wordList suffix;
    List<const dictionary*> dicts;
    PtrList<entry> modelEntries;
    if (type == "multiModeLogGPU")                                           // multiMode.C:73-92
    {
        modelEntries.transfer(PtrList<entry>(dict.lookup("models"))());
        forAll(modelEntries, i) { suffix.append(name + modelEntries[i].keyword()); dicts.append(&modelEntries[i].dict()); }
    }
    else { suffix.append(name); dicts.append(&dict); }

    modes_.setSize(suffix.size());
    forAll(suffix, i)
    {
        const word sub = (type == "multiModeLogGPU") ? word(dicts[i]->lookup("type")) : type;
        readMode(sub, *dicts[i], modes_[i]);
        etaS_.value() += modes_[i].etaS; etaP_.value() += modes_[i].etaP; rho_.value() += modes_[i].rho/suffix.size();
        // same IOobjects as the CPU models (Oldroyd_BLog.C:52-113)
        tau_.append(new volSymmTensorField(IOobject("tau" + suffix[i], U.time().timeName(), mesh, IOobject::MUST_READ, IOobject::AUTO_WRITE), mesh));
        theta_.append(new volSymmTensorField(IOobject("theta" + suffix[i], U.time().timeName(), mesh, IOobject::MUST_READ, IOobject::AUTO_WRITE), mesh));
        eigVals_.append(new volTensorField(IOobject("eigVals" + suffix[i], U.time().timeName(), mesh, IOobject::READ_IF_PRESENT, IOobject::AUTO_WRITE),
                                           mesh, dimensionedTensor("I", dimless, pTraits<tensor>::I), extrapolatedCalculatedFvPatchField<tensor>::typeName));
        eigVecs_.append(new volTensorField(IOobject("eigVecs" + suffix[i], U.time().timeName(), mesh, IOobject::READ_IF_PRESENT, IOobject::AUTO_WRITE),
                                           mesh, dimensionedTensor("I", dimless, pTraits<tensor>::I), extrapolatedCalculatedFvPatchField<tensor>::typeName));
    }
    checkForStab(dict);                                                     // constitutiveEq.C:430-439
    // the modes are batched on ONE matrix with ONE set of controls: every mode's schemes and solver entry must be identical
    // (checked, not assumed: the reference lets thetaM1, thetaM2, ... have their own fvSchemes / fvSolution entries)
    readSchemes(mesh, "theta" + suffix[0]);
    {
        const RheoSchemeCtl first = ctl_;
        for (label i = 1; i < suffix.size(); ++i)
        {
            readSchemes(mesh, "theta" + suffix[i]);
            if (std::memcmp(&first, &ctl_, sizeof(RheoSchemeCtl)) != 0)
            {
                FatalErrorInFunction << "schemes / solver controls of theta" << suffix[i] << " differ from those of theta" << suffix[0]
                    << "; the batched multiMode solve of the GPU path needs identical controls" << exit(FatalError);
            }
        }
    }

    if (modes_[0].model == RHEO_MODEL_BMP_LOG)                              // BMPLog.C:112-136
    {
        if (modes_.size() != 1) FatalErrorInFunction << "BMPLog runs as a single-mode model on the GPU path" << exit(FatalError);
        Phi_.reset(new volScalarField(IOobject("Phi" + name, U.time().timeName(), mesh, IOobject::MUST_READ, IOobject::AUTO_WRITE), mesh));
        // the fluidity equation shares theta's assembly and solver: same convection scheme, same solver entry, same patch kinds
        {
            const RheoSchemeCtl first = ctl_;
            readSchemes(mesh, "Phi" + name);   // reads div(phi,Phi<name>), solvers.Phi<name>, relaxationFactors.equations.Phi<name>
            modes_[0].bmp_relax = ctl_.relax;
            ctl_.relax = first.relax;
            if (std::memcmp(&first, &ctl_, sizeof(RheoSchemeCtl)) != 0)
            {
                FatalErrorInFunction << "schemes / solver controls of Phi" << name << " differ from those of theta" << name
                    << "; the GPU path solves both with one set of controls" << exit(FatalError);
            }
        }
        forAll(Phi_().boundaryField(), pI)
        {
            const fvPatchScalarField& pf = Phi_().boundaryField()[pI];
            if (pf.patch().coupled() || isA<emptyFvPatch>(pf.patch())) continue;
            const bool fixedPhi = isA<fixedValueFvPatchScalarField>(pf), fixedTheta = isA<fixedValueFvPatchSymmTensorField>(theta_[0].boundaryField()[pI]);
            if (fixedPhi != fixedTheta || (!fixedPhi && !isA<zeroGradientFvPatchScalarField>(pf)))
            {
                FatalErrorInFunction << "patch " << pf.patch().name() << ": Phi" << name << " must be fixedValue where theta" << name
                    << " is fixedValue and zeroGradient elsewhere (the fluidity is carried through theta's assembly)" << exit(FatalError);
            }
        }
    }

    List<int32_t> owner, neighbour; List<RheoPatchDesc> patches; vectorField nbrC; scalarField weights;
    RheoMeshDesc d;
    buildMeshDesc(mesh, owner, neighbour, patches, nbrC, weights, d);
    // one rank <-> one GPU: ranks of a node take the node's devices round-robin
    const int nDev = rheo_gpu_device_count();
    if (nDev < 1) FatalErrorInFunction << "no CUDA device: the GPU stress step has no CPU fallback" << exit(FatalError);
    check(rheo_gpu_create(&d, modes_.begin(), modes_.size(), &ctl_, Pstream::myProcNo() % nDev, &gpu_), "rheo_gpu_create");
    if (Pstream::parRun())
    {
        List<char> id(128, 0);
        if (Pstream::master()) check(rheo_gpu_nccl_unique_id(id.begin()), "rheo_gpu_nccl_unique_id");
        Pstream::scatter(id);                                               // MPI_Bcast of the NCCL unique id
        check(rheo_gpu_comm_init(gpu_, Pstream::myProcNo(), Pstream::nProcs(), id.begin()), "rheo_gpu_comm_init");
    }
    uploadState();
}

LogConformationGPU::~LogConformationGPU()
{
    rheo_gpu_destroy(gpu_);
}

// * * * * * * * * * * * * * * * Member Functions  * * * * * * * * * * * * * //

static void gatherBoundary(const GeometricField<symmTensor, fvPatchField, volMesh>& f, symmTensorField& out, label nI)
{
    out.setSize(f.mesh().nFaces() - nI, symmTensor::zero);
    forAll(f.boundaryField(), pI)
    {
        const fvPatchSymmTensorField& pf = f.boundaryField()[pI];
        if (!pf.size() || pf.patch().coupled() || isA<emptyFvPatch>(pf.patch())) continue;
        forAll(pf, i) out[pf.patch().start() - nI + i] = pf[i];
    }
}

void LogConformationGPU::uploadState()
{
    const label nI = U().mesh().nInternalFaces();
    forAll(theta_, i)
    {
        symmTensorField thB, tauB;
        gatherBoundary(theta_[i], thB, nI);
        gatherBoundary(tau_[i], tauB, nI);
        // symmTensor / tensor are 6 / 9 contiguous doubles in OpenFOAM's component order = the C-ABI's AoS order
        check(rheo_gpu_upload_state(gpu_, i,
              reinterpret_cast<const double*>(theta_[i].primitiveField().begin()),
              reinterpret_cast<const double*>(tau_[i].primitiveField().begin()),
              reinterpret_cast<const double*>(eigVals_[i].primitiveField().begin()),
              reinterpret_cast<const double*>(eigVecs_[i].primitiveField().begin()),
              reinterpret_cast<const double*>(thB.begin()), reinterpret_cast<const double*>(tauB.begin())),
              "rheo_gpu_upload_state");
    }
    if (Phi_.valid())   // BMPLog's fluidity (BMPLog.C:112-122)
    {
        scalarField PhiB(U().mesh().nFaces() - nI, 0.0);
        forAll(Phi_().boundaryField(), pI)
        {
            const fvPatchScalarField& pf = Phi_().boundaryField()[pI];
            if (!pf.size() || pf.patch().coupled() || isA<emptyFvPatch>(pf.patch())) continue;
            forAll(pf, i) PhiB[pf.patch().start() - nI + i] = pf[i];
        }
        check(rheo_gpu_upload_fluidity(gpu_, 0, Phi_().primitiveField().begin(), PhiB.begin()), "rheo_gpu_upload_fluidity");
    }
}

void LogConformationGPU::downloadAll()
{
    forAll(theta_, i)
    {
        check(rheo_gpu_download(gpu_, i, RHEO_FIELD_THETA,   reinterpret_cast<double*>(theta_[i].primitiveFieldRef().begin())), "download theta");
        check(rheo_gpu_download(gpu_, i, RHEO_FIELD_EIGVALS, reinterpret_cast<double*>(eigVals_[i].primitiveFieldRef().begin())), "download eigVals");
        check(rheo_gpu_download(gpu_, i, RHEO_FIELD_EIGVECS, reinterpret_cast<double*>(eigVecs_[i].primitiveFieldRef().begin())), "download eigVecs");
        theta_[i].correctBoundaryConditions();
        eigVals_[i].correctBoundaryConditions();
        eigVecs_[i].correctBoundaryConditions();
    }
    if (Phi_.valid())
    {
        check(rheo_gpu_download(gpu_, 0, RHEO_FIELD_FLUIDITY, Phi_().primitiveFieldRef().begin()), "download Phi");
        Phi_().correctBoundaryConditions();
    }
}

void LogConformationGPU::correct(const volScalarField* alpha, const volTensorField* gradU)
{
    // `alpha` (constitutiveTwoPhaseMixture.H:130-131) is part of the signature only: no *Log model reads it inside correct()
    // (Oldroyd_BLog.C:127-179 and the same bodies of the other models), so it is ignored here as it is there.
    // `gradU` (filmModel.C:408) replaces fvc::grad(U) (utils/boilerLog.H:1): hand it to the device, or return to its own gradient.
    check(rheo_gpu_upload_grad_u(gpu_, gradU ? reinterpret_cast<const double*>(gradU->primitiveField().begin()) : nullptr), "rheo_gpu_upload_grad_u");
    if (gradU && deviceDivTau_ && stabOption_ == soCoupling)
    {
        FatalErrorInFunction << "a caller-supplied gradU with stabilization coupling needs `deviceDivTau false` (divTau uses fvc::grad(U))"
                             << exit(FatalError);
    }
    const fvMesh& mesh = U().mesh();
    const label nI = mesh.nInternalFaces();
    // U boundary values and phi (internal + boundary) in face order
    vectorField Ub(mesh.nFaces() - nI, vector::zero);
    scalarField ph(mesh.nFaces(), 0.0);
    forAll(phi(), f) ph[f] = phi()[f];
    forAll(U().boundaryField(), pI)
    {
        const fvPatch& p = mesh.boundary()[pI];
        if (!p.size() || isA<emptyFvPatch>(p)) continue;
        forAll(p, i)
        {
            if (!p.coupled()) Ub[p.start() - nI + i] = U().boundaryField()[pI][i];
            ph[p.start() + i] = phi().boundaryField()[pI][i];
        }
    }
    // theta_.oldTime() bookkeeping of fvm::ddt (Oldroyd_BLog.C:143): once per time step, not per inner iteration
    const bool newStep = U().time().timeIndex() != lastTimeIndex_;
    lastTimeIndex_ = U().time().timeIndex();

    // temperature-dependent lambda / etaP (Oldroyd_BLog.C:133-135): evaluated here with the reference's own thermoFunction
    // objects, one scalar per cell and mode to the device
    if (hasThermo_)
    {
        forAll(modes_, i)
        {
            const volScalarField lam(thermoLambda_[i].createField(dimensionedScalar("lambda", dimTime, modes_[i].lambda)));
            const volScalarField eta(thermoEta_[i].createField(dimensionedScalar("etaP", dimMass/(dimLength*dimTime), modes_[i].etaP)));
            check(rheo_gpu_upload_thermo(gpu_, i, lam.primitiveField().begin(), eta.primitiveField().begin()), "rheo_gpu_upload_thermo");
        }
    }

    symmTensorField tauB(mesh.nFaces() - nI);
    List<RheoStepStats> stats(modes_.size());
    // deviceDivTau: the momentum predictor takes div(tau) from the device (divTau below), so tau only comes back at write time
    const bool wantTau = !deviceDivTau_ || U().time().writeTime();
    check(rheo_gpu_correct(gpu_,
          reinterpret_cast<const double*>(U().primitiveField().begin()), reinterpret_cast<const double*>(Ub.begin()), ph.begin(),
          U().time().deltaTValue(), newStep ? 1 : 0,
          wantTau ? reinterpret_cast<double*>(tauTotal_.primitiveFieldRef().begin()) : nullptr,
          wantTau ? reinterpret_cast<double*>(tauB.begin()) : nullptr, stats.begin()),
          "rheo_gpu_correct");
    tauOnHost_ = wantTau;

    // OpenFOAM-style solver report (SolverPerformance<symmTensor>)
    static const char* cmpt[6] = {"XX", "XY", "XZ", "YY", "YZ", "ZZ"};
    forAll(stats, i) for (int c = 0; c < 6; ++c) if (stats[i].n_iterations[c] || stats[i].initial_residual[c] > 0)
        Info<< (ctl_.solver == RHEO_SOLVER_PBICG ? "B200-PBiCG:  Solving for " : "B200-PBiCGStab:  Solving for ") << theta_[i].name() << cmpt[c] << ", Initial residual = " << stats[i].initial_residual[c]
            << ", Final residual = " << stats[i].final_residual[c] << ", No Iterations " << stats[i].n_iterations[c] << endl;

    // boundary values of tau for the momentum predictor: the device returns the patch values SUMMED OVER THE MODES
    // (RHEO_FIELD_TAU_B_TOTAL) — multiMode::divTau (multiMode.C) sums each mode's divTau, i.e. uses each mode's own
    // linearExtrapolation / zeroGradient / fixedValue patch values
    if (wantTau) forAll(tauTotal_.boundaryField(), pI)
    {
        fvPatchSymmTensorField& pf = tauTotal_.boundaryFieldRef()[pI];
        if (!pf.size() || pf.patch().coupled() || isA<emptyFvPatch>(pf.patch())) continue;
        forAll(pf, i) pf[i] = tauB[pf.patch().start() - nI + i];
    }
    if (wantTau && modes_.size() == 1)
    {
        tau_[0].primitiveFieldRef() = tauTotal_.primitiveField();
        tau_[0].boundaryFieldRef() = tauTotal_.boundaryField();
    }

    if (U().time().writeTime())
    {
        downloadAll();
        if (modes_.size() > 1) forAll(tau_, i)
            check(rheo_gpu_download(gpu_, i, RHEO_FIELD_TAU, reinterpret_cast<double*>(tau_[i].primitiveFieldRef().begin())), "download tau");
    }
}


// tau() (constitutiveEq.H:314; multiMode.C:216-226).  With deviceDivTau the host copy is stale between write times: a caller
// that still asks for it (post-processing function objects) gets a fresh download.
Foam::tmp<Foam::volSymmTensorField> Foam::constitutiveEqs::LogConformationGPU::tau() const
{
    if (!tauOnHost_)
    {
        const label nI = U().mesh().nInternalFaces();
        symmTensorField tauB(U().mesh().nFaces() - nI);
        check(rheo_gpu_download(gpu_, 0, RHEO_FIELD_TAU_TOTAL, reinterpret_cast<double*>(tauTotal_.primitiveFieldRef().begin())), "download tau");
        check(rheo_gpu_download(gpu_, 0, RHEO_FIELD_TAU_B_TOTAL, reinterpret_cast<double*>(tauB.begin())), "download tau_b");
        forAll(tauTotal_.boundaryField(), pI)
        {
            fvPatchSymmTensorField& pf = tauTotal_.boundaryFieldRef()[pI];
            if (!pf.size() || pf.patch().coupled() || isA<emptyFvPatch>(pf.patch())) continue;
            forAll(pf, i) pf[i] = tauB[pf.patch().start() - nI + i];
        }
        tauOnHost_ = true;
    }
    return tauTotal_;
}

// constitutiveEq::divTau (constitutiveEq.C:72-132): the fvc::div terms come from the device (rheo_gpu_div_tau: Gauss linear,
// which readSchemes has checked for div(tau) and div(grad(U))), the implicit laplacian and BSD's explicit laplacian are
// OpenFOAM's.  multiMode (multiMode.C:143-157) sums the modes' matrices: the device sums tau_m/rho_m and etaP_m/rho_m, the
// implicit part uses the summed etaP_ (constructor) like the single-mode classes.
Foam::tmp<Foam::fvVectorMatrix> Foam::constitutiveEqs::LogConformationGPU::divTau(const volVectorField& U) const
{
    if (!deviceDivTau_ || solveCoupled_) return constitutiveEq::divTau(U);
    const int32_t stab = stabOption_ == soNone ? RHEO_STAB_NONE : (stabOption_ == soBSD ? RHEO_STAB_BSD : RHEO_STAB_COUPLING);
    volVectorField divExplicit
    (
        IOobject("divTauExplicit", U.time().timeName(), U.mesh(), IOobject::NO_READ, IOobject::NO_WRITE),
        U.mesh(),
        dimensionedVector("zero", dimensionSet(0, 1, -2, 0, 0, 0, 0), vector::zero)   // div(tau/rho)
    );
    check(rheo_gpu_div_tau(gpu_, stab, reinterpret_cast<double*>(divExplicit.primitiveFieldRef().begin())), "rheo_gpu_div_tau");
    switch (stabOption_)
    {
        case soNone:
            return divExplicit + fvm::laplacian(etaS()/rho(), U, "laplacian(eta,U)");
        case soBSD:
            return divExplicit - fvc::laplacian(etaP()/rho(), U, "laplacian(eta,U)")
                 + fvm::laplacian((etaP() + etaS())/rho(), U, "laplacian(eta,U)");
        default:
            return divExplicit + fvm::laplacian((etaP() + etaS())/rho(), U, "laplacian(eta,U)");
    }
}
