// ref_model.cpp — one translation unit per Log model (TEST INFRASTRUCTURE ONLY, see ref_ce.H).
// Compiled with -DREF_MODEL_<name>; the function body is the reference's text, cut by oracle/Makefile from
// constitutiveEqs/<family>/<name>/<name>.C ("void Foam::constitutiveEqs::<name>::correct" to end of file)
// into oracle/_ref/gen/<name>_correct.inc.  boilerLog.H (included by that text) is the reference's file and
// switches on PTTLog_H exactly as it does when PTTLog.C includes PTTLog.H.
#if defined(REF_MODEL_PTTLog)
#define PTTLog_H
#endif
#if defined(REF_MODEL_SaramitoLog)
#define SaramitoLog_H
#endif
#include "ref_ce.H"

#if defined(REF_MODEL_Oldroyd_BLog)
#include "Oldroyd_BLog_correct.inc"
#elif defined(REF_MODEL_GiesekusLog)
#include "GiesekusLog_correct.inc"
#elif defined(REF_MODEL_PTTLog)
#include "PTTLog_correct.inc"
#elif defined(REF_MODEL_FENE_PLog)
#include "FENE_PLog_correct.inc"
#elif defined(REF_MODEL_FENE_CRLog)
#include "FENE_CRLog_correct.inc"
#elif defined(REF_MODEL_WhiteMetznerCYLog)
#include "WhiteMetznerCYLog_correct.inc"
#elif defined(REF_MODEL_RoliePolyLog)
#include "RoliePolyLog_correct.inc"
#elif defined(REF_MODEL_XPomPomLog)
#include "XPomPomLog_correct.inc"
#elif defined(REF_MODEL_SaramitoLog)
#include "SaramitoLog_correct.inc"
#elif defined(REF_MODEL_BMPLogFull)
#define Phi_ (*PhiPtr_)
#include "BMPLogFull_correct.inc"
#elif defined(REF_MODEL_BMPLog)
#define Phi_ (*PhiPtr_)   // BMPLog.H: volScalarField Phi_; here a field the harness fills with the fluidity of after PhiEqn.solve()
#include "BMPLog_correct.inc"
#else
#error "define REF_MODEL_<name>"
#endif
