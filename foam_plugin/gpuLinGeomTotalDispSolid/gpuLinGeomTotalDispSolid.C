/*---------------------------------------------------------------------------*\
  See gpuLinGeomTotalDispSolid.H.  Source only: needs OpenFOAM + solids4foam to compile.
\*---------------------------------------------------------------------------*/
#include "gpuLinGeomTotalDispSolid.H"
#include "addToRunTimeSelectionTable.H"
#include "fvm.H"
#include "fvc.H"
#include "solidTractionFvPatchVectorField.H"

namespace Foam
{
namespace solidModels
{

defineTypeNameAndDebug(gpuLinGeomTotalDispSolid, 0);
addToRunTimeSelectionTable(solidModel, gpuLinGeomTotalDispSolid, dictionary);      // as linGeomTotalDispSolid.C:40-41


void gpuLinGeomTotalDispSolid::downloadState()
{
    gpu_.downloadVector(D(), S4F_FIELD_D, S4F_FIELD_D_B);
    gpu_.downloadTensor(gradD(), S4F_FIELD_GRAD_D, S4F_FIELD_GRAD_D_B);
    gpu_.downloadSymmTensor(sigma(), S4F_FIELD_SIGMA, S4F_FIELD_SIGMA_B);
}


gpuLinGeomTotalDispSolid::gpuLinGeomTotalDispSolid
(
    Time& runTime,
    const word& region
)
:
    solidModel(typeName, runTime, region),
    impK_(mechanical().impK()),
    rImpK_(1.0/impK_),
    gpu_(mesh(), solidModelDict().subOrEmptyDict("gpu"))
{
    DisRequired();

    // old-time fields exist on the host exactly as for the CPU model (linGeomTotalDispSolid.C:79)
    fvm::d2dt2(D());

    gpu_.mirrorMesh();
    const bool pointStencil =
        word(mesh().gradSchemes().lookupOrDefault<word>("default", "leastSquares")) == "pointCellsLeastSquares";
    gpu_.mirrorGeometry(pointStencil);
    gpu_.mirrorLaw(mechanical());                      // the block parsed by the gpu* law shell
    gpuSolidBridge::loopControls lc = {nCorr(), solutionTol(), alternativeTol(), materialTol()};
    gpu_.mirrorControls(S4F_MODEL_LIN_GEOM_TOTAL_DISP, "D", solidModelDict(), lc, g().value());
    gpu_.mirrorBoundaryConditions(D());

    gpu_.upload(S4F_FIELD_D, reinterpret_cast<const double*>(D().internalField().cdata()), "ctor");
    gpu_.upload(S4F_FIELD_D_OLD, reinterpret_cast<const double*>(D().oldTime().internalField().cdata()), "ctor");
    gpu_.upload(S4F_FIELD_D_OLDOLD, reinterpret_cast<const double*>(D().oldTime().oldTime().internalField().cdata()), "ctor");

    // D.correctBoundaryConditions(); D.storePrevIter(); mechanical().grad(D, gradD)  (:82-84)
    gpu_.check(s4fgpu_initialise(gpu_.handle()), "gpuLinGeomTotalDispSolid::gpuLinGeomTotalDispSolid(...)");
}


gpuLinGeomTotalDispSolid::~gpuLinGeomTotalDispSolid()
{}


bool gpuLinGeomTotalDispSolid::evolve()
{
    Info<< "Evolving solid solver on the GPU" << endl;

    gpu_.newTimeStepIfNeeded();                 // once per time index, not once per evolve()
    gpu_.mirrorBoundaryConditions(D());         // time-varying tractions / displacements

    s4fgpu_stats st;
    gpu_.check(s4fgpu_evolve(gpu_.handle(), &st), "evolve()");

    // the reference's log line (solidModelTemplates.C:153-163), so log scrapers keep working
    Info<< "    Corr, res, relRes, matRes, iters" << nl
        << "    " << st.nCorr << ", " << st.solverPerfInitRes << ", " << st.relResidual << ", "
        << st.materialResidual << ", " << st.nIterations[0] + st.nIterations[1] + st.nIterations[2]
        << nl << endl;

    downloadState();

    // post-loop host work exactly as linGeomTotalDispSolid.C:212-223
    mechanical().interpolate(D(), gradD(), pointD());
    U() = fvc::ddt(D());

    return st.converged;
}


tmp<vectorField> gpuLinGeomTotalDispSolid::tractionBoundarySnGrad
(
    const vectorField& traction,
    const scalarField& pressure,
    const fvPatch& patch
) const
{
    // same expression as linGeomTotalDispSolid.C:235-271, on the host copies
    const label patchID = patch.index();
    const scalarField& pImpK = impK_.boundaryField()[patchID];
    const scalarField& pRImpK = rImpK_.boundaryField()[patchID];
    const tensorField& pGradD = gradD().boundaryField()[patchID];
    const symmTensorField& pSigma = sigma().boundaryField()[patchID];
    const vectorField n(patch.nf());

    return tmp<vectorField>
    (
        new vectorField(((traction - n*pressure) - (n & (pSigma - pImpK*pGradD)))*pRImpK)
    );
}


void gpuLinGeomTotalDispSolid::setTraction
(
    const label interfaceI,
    const label patchID,
    const vectorField& faceZoneTraction
)
{
    solidModel::setTraction(interfaceI, patchID, faceZoneTraction);   // fills the host patch field
    const solidTractionFvPatchVectorField& t =
        refCast<const solidTractionFvPatchVectorField>(D().boundaryField()[patchID]);
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


void gpuLinGeomTotalDispSolid::updateTotalFields()
{
    gpu_.check(s4fgpu_update_total_fields(gpu_.handle()), "updateTotalFields()");
    solidModel::updateTotalFields();
}


void gpuLinGeomTotalDispSolid::writeFields(const Time& runTime)
{
    solidModel::writeFields(runTime);
}

} // End namespace solidModels
} // End namespace Foam
