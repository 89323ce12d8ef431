/*---------------------------------------------------------------------------*\
  See gpuNonLinGeomTotalLagSolid.H.  Source only: needs OpenFOAM + solids4foam to compile.
\*---------------------------------------------------------------------------*/
#include "gpuNonLinGeomTotalLagSolid.H"
#include "addToRunTimeSelectionTable.H"
#include "fvm.H"
#include "fvc.H"
#include "solidTractionFvPatchVectorField.H"

namespace Foam
{
namespace solidModels
{

defineTypeNameAndDebug(gpuNonLinGeomTotalLagSolid, 0);
addToRunTimeSelectionTable(solidModel, gpuNonLinGeomTotalLagSolid, dictionary);      // as nonLinGeomTotalLagSolid.C:40-44


void gpuNonLinGeomTotalLagSolid::downloadState()
{
    gpu_.downloadVector(D(), S4F_FIELD_D, S4F_FIELD_D_B);
    gpu_.downloadTensor(gradD(), S4F_FIELD_GRAD_D, S4F_FIELD_GRAD_D_B);
    gpu_.downloadVector(DD(), S4F_FIELD_DD, S4F_FIELD_DD_B);
    gpu_.downloadTensor(gradDD(), S4F_FIELD_GRAD_DD, -1);
    gpu_.downloadSymmTensor(sigma(), S4F_FIELD_SIGMA, S4F_FIELD_SIGMA_B);
    gpu_.downloadTensor(F_, S4F_FIELD_F, -1);
    gpu_.download(S4F_FIELD_J, J_.primitiveFieldRef().data(), "downloadState()");
    // the kinematic fields of the solid model (nonLinGeomTotalLagSolid.C:190-201) from the device's grad(D)
    F_.correctBoundaryConditions();
    Finv_ = inv(F_);
    J_.correctBoundaryConditions();
}


gpuNonLinGeomTotalLagSolid::gpuNonLinGeomTotalLagSolid(Time& runTime, const word& region)
:
    solidModel(typeName, runTime, region),
    F_
    (
        IOobject("F", runTime.timeName(), mesh(), IOobject::READ_IF_PRESENT, IOobject::AUTO_WRITE),
        mesh(),
        dimensionedTensor("I", dimless, I)
    ),
    Finv_(IOobject("Finv", runTime.timeName(), mesh(), IOobject::READ_IF_PRESENT, IOobject::NO_WRITE), inv(F_)),
    J_(IOobject("J", runTime.timeName(), mesh(), IOobject::READ_IF_PRESENT, IOobject::NO_WRITE), det(F_)),
    impK_(mechanical().impK()),
    rImpK_(1.0/impK_),
    gpu_(mesh(), solidModelDict().subOrEmptyDict("gpu"))
{
    DDisRequired();

    // old-time fields on the host as the CPU model creates them (nonLinGeomTotalLagSolid.C:107)
    fvm::d2dt2(DD());

    gpu_.mirrorMesh();
    const bool pointStencil =
        word(mesh().gradSchemes().lookupOrDefault<word>("default", "leastSquares")) == "pointCellsLeastSquares";
    gpu_.mirrorGeometry(pointStencil);
    gpu_.mirrorLaw(mechanical());                      // finite-strain law block from the gpu* law shell
    gpuSolidBridge::loopControls lc = {nCorr(), solutionTol(), alternativeTol(), materialTol()};
    gpu_.mirrorControls(S4F_MODEL_NONLIN_TL, "DD", solidModelDict(), lc, g().value());
    gpu_.mirrorBoundaryConditions(DD());

    gpu_.upload(S4F_FIELD_D, reinterpret_cast<const double*>(D().internalField().cdata()), "ctor");
    gpu_.upload(S4F_FIELD_D_OLD, reinterpret_cast<const double*>(D().oldTime().internalField().cdata()), "ctor");
    gpu_.upload(S4F_FIELD_F, reinterpret_cast<const double*>(F_.internalField().cdata()), "ctor");

    // consistent start incl. the restart branch (nonLinGeomTotalLagSolid.C:109-120): grad(D), F, Finv, J on the device
    gpu_.check(s4fgpu_initialise(gpu_.handle()), "gpuNonLinGeomTotalLagSolid::gpuNonLinGeomTotalLagSolid(...)");
}


gpuNonLinGeomTotalLagSolid::~gpuNonLinGeomTotalLagSolid()
{}


bool gpuNonLinGeomTotalLagSolid::evolve()
{
    Info<< "Evolving solid solver on the GPU" << endl;

    gpu_.newTimeStepIfNeeded();
    gpu_.mirrorBoundaryConditions(DD());

    s4fgpu_stats st;
    gpu_.check(s4fgpu_evolve(gpu_.handle(), &st), "evolve()");      // the do-while loop nonLinGeomTotalLagSolid.C:147-223

    Info<< "    Corr, res, relRes, matRes, iters" << nl
        << "    " << st.nCorr << ", " << st.solverPerfInitRes << ", " << st.relResidual << ", "
        << st.materialResidual << ", " << st.nIterations[0] + st.nIterations[1] + st.nIterations[2]
        << nl << endl;

    downloadState();

    // post-loop host work as in the reference model
    mechanical().interpolate(D(), gradD(), pointD());
    mechanical().interpolate(DD(), gradDD(), pointDD());
    U() = fvc::ddt(D());

    return st.converged;
}


tmp<vectorField> gpuNonLinGeomTotalLagSolid::tractionBoundarySnGrad
(
    const vectorField& traction,
    const scalarField& pressure,
    const fvPatch& patch
) const
{
    // host version of nonLinGeomTotalLagSolid.C:263-325 (deformed normal from Finv, Nanson's formula) for boundary conditions evaluated
    // on the host; the device evaluates its traction patches itself (k_bc_update)
    const label patchID = patch.index();
    const scalarField& pImpK = impK_.boundaryField()[patchID];
    const scalarField& pRImpK = rImpK_.boundaryField()[patchID];
    const tensorField& pGrad = gradDD().boundaryField()[patchID];
    const symmTensorField& pSigma = sigma().boundaryField()[patchID];
    const tensorField& pFinv = Finv_.boundaryField()[patchID];
    const vectorField n(patch.nf());
    vectorField nCurrent(pFinv.T() & n);
    nCurrent /= mag(nCurrent);

    return tmp<vectorField>
    (
        new vectorField(((traction - nCurrent*pressure) - (nCurrent & pSigma) + pImpK*(n & pGrad))*pRImpK)
    );
}


void gpuNonLinGeomTotalLagSolid::setTraction(const label interfaceI, const label patchID, const vectorField& faceZoneTraction)
{
    solidModel::setTraction(interfaceI, patchID, faceZoneTraction);
    const solidTractionFvPatchVectorField& t =
        refCast<const solidTractionFvPatchVectorField>(DD().boundaryField()[patchID]);
    gpu_.check
    (
        s4fgpu_set_bc
        (
            gpu_.handle(), patchID, S4F_BC_SOLID_TRACTION,
            reinterpret_cast<const double*>(t.traction().cdata()), t.pressure().cdata()
        ),
        "setTraction()"
    );
}


void gpuNonLinGeomTotalLagSolid::updateTotalFields()
{
    // law history commit (neoHookeanElasticMisesPlastic.C:1526-1601) on the device
    gpu_.check(s4fgpu_update_total_fields(gpu_.handle()), "updateTotalFields()");
    solidModel::updateTotalFields();
}

} // End namespace solidModels
} // End namespace Foam
