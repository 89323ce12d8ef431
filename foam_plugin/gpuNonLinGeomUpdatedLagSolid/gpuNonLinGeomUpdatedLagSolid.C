/*---------------------------------------------------------------------------*\
  See gpuNonLinGeomUpdatedLagSolid.H.  Source only: needs OpenFOAM + solids4foam to compile.
\*---------------------------------------------------------------------------*/
#include "gpuNonLinGeomUpdatedLagSolid.H"
#include "addToRunTimeSelectionTable.H"
#include "fvc.H"
#include "fvm.H"

namespace Foam
{
namespace solidModels
{

defineTypeNameAndDebug(gpuNonLinGeomUpdatedLagSolid, 0);
addToRunTimeSelectionTable(solidModel, gpuNonLinGeomUpdatedLagSolid, dictionary);   // nonLinGeomUpdatedLagSolid.C:40-43


void gpuNonLinGeomUpdatedLagSolid::downloadState()
{
    gpu_.downloadVector(D(), S4F_FIELD_D, S4F_FIELD_D_B);
    gpu_.downloadVector(DD(), S4F_FIELD_DD, S4F_FIELD_DD_B);
    gpu_.downloadTensor(gradD(), S4F_FIELD_GRAD_D, S4F_FIELD_GRAD_D_B);
    gpu_.downloadTensor(gradDD(), S4F_FIELD_GRAD_DD, -1);
    gpu_.downloadSymmTensor(sigma(), S4F_FIELD_SIGMA, S4F_FIELD_SIGMA_B);
    gpu_.downloadTensor(F_, S4F_FIELD_F, -1);
    gpu_.download(S4F_FIELD_J, J_.primitiveFieldRef().data(), "downloadState()");
    gpu_.download(S4F_FIELD_RHO, rho_.primitiveFieldRef().data(), "downloadState()");
}


gpuNonLinGeomUpdatedLagSolid::gpuNonLinGeomUpdatedLagSolid(Time& runTime, const word& region)
:
    solidModel(typeName, runTime, region),
    F_(IOobject("F", runTime.timeName(), mesh(), IOobject::READ_IF_PRESENT, IOobject::AUTO_WRITE), mesh(), dimensionedTensor("I", dimless, I)),
    J_(IOobject("J", runTime.timeName(), mesh(), IOobject::READ_IF_PRESENT, IOobject::NO_WRITE), det(F_)),
    rho_(IOobject("rho", runTime.timeName(), mesh(), IOobject::READ_IF_PRESENT, IOobject::AUTO_WRITE), mechanical().rho()),
    impK_(mechanical().impK()),
    rImpK_(1.0/impK_),
    gpu_(mesh(), solidModelDict().subOrEmptyDict("gpu"))
{
    DDisRequired();
    fvm::d2dt2(rho_, DD());                     // old-time levels on the host, as the CPU model (:143-145)
    fvc::d2dt2(rho_, D().oldTime());

    gpu_.mirrorMesh();
    gpu_.mirrorGeometry(true);                  // with points()/faces(): vol->point interpolation runs on the device
    gpu_.mirrorLaw(mechanical());               // gpuNeoHookeanElastic | gpuNeoHookeanElasticMisesPlastic
    gpuSolidBridge::loopControls lc = {nCorr(), solutionTol(), alternativeTol(), materialTol()};
    gpu_.mirrorControls(S4F_MODEL_NONLIN_UL, "DD", solidModelDict(), lc, g().value());
    // the fixedDisplacement patches hand over the TOTAL displacement, the device subtracts D.oldTime()
    // (fixedDisplacementFvPatchVectorField.C:279-287)
    gpu_.mirrorBoundaryConditions(DD());

    gpu_.upload(S4F_FIELD_D, reinterpret_cast<const double*>(D().internalField().cdata()), "ctor");
    gpu_.upload(S4F_FIELD_D_OLD, reinterpret_cast<const double*>(D().oldTime().internalField().cdata()), "ctor");
    gpu_.upload(S4F_FIELD_F, reinterpret_cast<const double*>(F_.internalField().cdata()), "ctor");
    gpu_.check(s4fgpu_initialise(gpu_.handle()), "gpuNonLinGeomUpdatedLagSolid::gpuNonLinGeomUpdatedLagSolid(...)");
}


gpuNonLinGeomUpdatedLagSolid::~gpuNonLinGeomUpdatedLagSolid()
{}


bool gpuNonLinGeomUpdatedLagSolid::evolve()
{
    Info<< "Evolving solid solver on the GPU" << nl
        << "Solving the updated Lagrangian form of the momentum equation for DD" << endl;

    gpu_.newTimeStepIfNeeded();                 // once per time index, not once per evolve()
    gpu_.mirrorBoundaryConditions(DD());

    s4fgpu_stats st;
    gpu_.check(s4fgpu_evolve(gpu_.handle(), &st), "evolve()");      // the do-while loop nonLinGeomUpdatedLagSolid.C:166-240

    Info<< "    Corr, res, relRes, matRes, iters" << nl
        << "    " << st.nCorr << ", " << st.solverPerfInitRes << ", " << st.relResidual << ", "
        << st.materialResidual << ", " << st.nIterations[0] + st.nIterations[1] + st.nIterations[2]
        << nl << endl;

    downloadState();

    // mechanical().interpolate(DD(), gradDD(), pointDD()) (:249), on the device
    gpu_.check
    (
        s4fgpu_interpolate_to_points
        (
            gpu_.handle(), S4F_FIELD_DD, S4F_POINT_INTERP_GRAD, reinterpret_cast<double*>(pointDD().primitiveFieldRef().data())
        ),
        "evolve()"
    );
    pointD() = pointD().oldTime() + pointDD();
    U() = fvc::ddt(D());
    return st.converged;
}


tmp<vectorField> gpuNonLinGeomUpdatedLagSolid::tractionBoundarySnGrad
(
    const vectorField& traction, const scalarField& pressure, const fvPatch& patch
) const
{
    // host version of nonLinGeomUpdatedLagSolid.C:263-357 for boundary conditions evaluated on the host; the device
    // evaluates its traction patches itself (k_bc_update, deformed normal from relFinv)
    const label patchID = patch.index();
    const vectorField n(patch.nf());
    return tmp<vectorField>
    (
        new vectorField
        (
            ((traction - n*pressure) - (n & sigma().boundaryField()[patchID])
          + impK_.boundaryField()[patchID]*(n & gradDD().boundaryField()[patchID]))*rImpK_.boundaryField()[patchID]
        )
    );
}


void gpuNonLinGeomUpdatedLagSolid::updateTotalFields()
{
    // moveMesh(oldPoints, DD(), pointDD()) (solidModel.C:2008-2148): the interpolation it starts with runs on the device
    pointVectorField& pDD = pointDD();
    gpu_.check
    (
        s4fgpu_interpolate_to_points
        (
            gpu_.handle(), S4F_FIELD_DD, S4F_POINT_INTERP_PATCH, reinterpret_cast<double*>(pDD.primitiveFieldRef().data())
        ),
        "updateTotalFields()"
    );

    // rho_ = rho_.oldTime()/relJ_, gradD = fvc::grad(D.oldTime() + DD), law history: on the device
    gpu_.check(s4fgpu_update_total_fields(gpu_.handle()), "updateTotalFields()");

    // symmetry-plane / empty corrections and mesh().movePoints(newPoints): the reference's own host code.  Note that
    // solidModel::moveMesh starts by interpolating DD to the points again on the host (solidModel.C:2020); the device
    // values above are the same field (tests/test_gpu_parity.py: vol->point parity), so either may be kept.
    const vectorField oldPoints(mesh().points());
    moveMesh(oldPoints, DD(), pDD);

    // the moved geometry goes back to the device; fields, boundary data and history are kept
    gpu_.mirrorGeometry(true);

    solidModel::updateTotalFields();
}

} // End namespace solidModels
} // End namespace Foam
