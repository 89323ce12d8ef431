/*---------------------------------------------------------------------------*\
  See gpuUnsSolids.H.  Source only: needs OpenFOAM + solids4foam to compile.
\*---------------------------------------------------------------------------*/
#include "gpuUnsSolids.H"
#include "addToRunTimeSelectionTable.H"
#include "fvm.H"
#include "fvc.H"
#include "solidTractionFvPatchVectorField.H"

namespace Foam
{
namespace solidModels
{

defineTypeNameAndDebug(gpuUnsLinGeomSolid, 0);
addToRunTimeSelectionTable(solidModel, gpuUnsLinGeomSolid, dictionary);                 // as unsLinGeomSolid.C:39-40
defineTypeNameAndDebug(gpuUnsNonLinGeomTotalLagSolid, 0);
addToRunTimeSelectionTable(solidModel, gpuUnsNonLinGeomTotalLagSolid, dictionary);      // as unsNonLinGeomTotalLagSolid.C:38-43
defineTypeNameAndDebug(gpuUnsNonLinGeomUpdatedLagSolid, 0);
addToRunTimeSelectionTable(solidModel, gpuUnsNonLinGeomUpdatedLagSolid, dictionary);    // as unsNonLinGeomUpdatedLagSolid.C:38-43


void gpuUnsSolidBase::downloadState()
{
    gpu_.downloadVector(D(), S4F_FIELD_D, S4F_FIELD_D_B);
    gpu_.downloadTensor(gradD(), S4F_FIELD_GRAD_D, S4F_FIELD_GRAD_D_B);
    gpu_.downloadSymmTensor(sigma(), S4F_FIELD_SIGMA, S4F_FIELD_SIGMA_B);
    if (incrementalModel())
    {
        gpu_.downloadVector(DD(), S4F_FIELD_DD, S4F_FIELD_DD_B);
        gpu_.downloadTensor(gradDD(), S4F_FIELD_GRAD_DD, -1);
    }

    // face fields: [F + B] AoS in the order internal faces, then the boundary faces in patch order (the fv faces of the
    // mirror: empty patches are not part of it)
    const label nI = mesh().nInternalFaces();
    label nFv = nI;
    forAll(mesh().boundary(), patchI)
    {
        if (!isA<emptyPolyPatch>(mesh().boundaryMesh()[patchI])) nFv += mesh().boundary()[patchI].size();
    }
    List<symmTensor> sf(nFv);
    List<tensor> gf(nFv);
    gpu_.download(S4F_FIELD_SIGMA_F, reinterpret_cast<double*>(sf.begin()), "downloadState()");
    gpu_.download(S4F_FIELD_GRAD_D_F, reinterpret_cast<double*>(gf.begin()), "downloadState()");
    SubList<symmTensor>(sigmaf_.primitiveFieldRef(), nI) = SubList<symmTensor>(sf, nI);
    SubList<tensor>(gradDf_.primitiveFieldRef(), nI) = SubList<tensor>(gf, nI);
    label at = nI;
    forAll(mesh().boundary(), patchI)
    {
        if (isA<emptyPolyPatch>(mesh().boundaryMesh()[patchI])) continue;
        const label n = mesh().boundary()[patchI].size();
        sigmaf_.boundaryFieldRef()[patchI] = SubList<symmTensor>(sf, n, at);
        gradDf_.boundaryFieldRef()[patchI] = SubList<tensor>(gf, n, at);
        at += n;
    }
}


gpuUnsSolidBase::gpuUnsSolidBase(const word& type, const int modelEnum, Time& runTime, const word& region)
:
    solidModel(type, runTime, region),
    modelEnum_(modelEnum),
    sigmaf_
    (
        IOobject("sigmaf", runTime.timeName(), mesh(), IOobject::READ_IF_PRESENT, IOobject::AUTO_WRITE),
        mesh(),
        dimensionedSymmTensor("zero", dimForce/dimArea, symmTensor::zero)
    ),
    gradDf_
    (
        IOobject("grad(" + word(modelEnum == S4F_MODEL_UNS_NONLIN_UL ? "DD" : "D") + ")f", runTime.timeName(), mesh(),
                 IOobject::READ_IF_PRESENT, IOobject::NO_WRITE),
        mesh(),
        dimensionedTensor("0", dimless, tensor::zero)
    ),
    impK_(mechanical().impK()),
    rImpK_(1.0/impK_),
    gpu_(mesh(), solidModelDict().subOrEmptyDict("gpu"))
{
    if (incrementalModel())
    {
        DDisRequired();
        fvm::d2dt2(DD());            // old-time fields on the host as the CPU model creates them
    }
    else
    {
        DisRequired();
        fvm::d2dt2(D());
    }

    gpu_.mirrorMesh();
    gpu_.mirrorGeometry(true);       // the vertex-based gradients need points() and faces()
    gpu_.mirrorLaw(mechanical());    // gpuLinearElastic / gpuNeoHookeanElastic: the library refuses any other law on the faces
    gpuSolidBridge::loopControls lc = {nCorr(), solutionTol(), alternativeTol(), materialTol()};
    gpu_.mirrorControls(modelEnum_, incrementalModel() ? "DD" : "D", solidModelDict(), lc, g().value());
    gpu_.mirrorBoundaryConditions(solved());

    gpu_.upload(S4F_FIELD_D, reinterpret_cast<const double*>(D().internalField().cdata()), "ctor");
    gpu_.upload(S4F_FIELD_D_OLD, reinterpret_cast<const double*>(D().oldTime().internalField().cdata()), "ctor");

    // consistent start as the CPU constructors do it (unsLinGeomSolid.C:86-89): boundary conditions, point interpolation and the
    // gradients; sigmaf keeps its initial value
    gpu_.check(s4fgpu_initialise(gpu_.handle()), "gpuUnsSolidBase::gpuUnsSolidBase(...)");
}


gpuUnsSolidBase::~gpuUnsSolidBase()
{}


bool gpuUnsSolidBase::evolve()
{
    Info<< "Evolving solid solver on the GPU" << endl;

    gpu_.newTimeStepIfNeeded();
    gpu_.mirrorBoundaryConditions(solved());

    s4fgpu_stats st;
    // the do-while loops unsLinGeomSolid.C:108-175, unsNonLinGeomTotalLagSolid.C:247-380 (its own convergence test: the
    // relative change of D against the increment of the step, never on the first iteration), unsNonLinGeomUpdatedLagSolid.C:264-316
    gpu_.check(s4fgpu_evolve(gpu_.handle(), &st), "evolve()");

    Info<< "    Corr, res, relRes, iters" << nl
        << "    " << st.nCorr << ", " << st.solverPerfInitRes << ", " << st.relResidual << ", "
        << st.nIterations[0] + st.nIterations[1] + st.nIterations[2] << nl << endl;

    downloadState();

    // pointD / pointDD as the models leave them (patch-mode interpolation inside the loop), from the device
    pointVectorField& pf = incrementalModel() ? pointDD() : pointD();
    gpu_.check
    (
        s4fgpu_interpolate_to_points
        (
            gpu_.handle(), incrementalModel() ? S4F_FIELD_DD : S4F_FIELD_D, S4F_POINT_INTERP_PATCH,
            reinterpret_cast<double*>(pf.primitiveFieldRef().data())
        ),
        "evolve()"
    );
    if (incrementalModel())
    {
        pointD() = pointD().oldTime() + pointDD();
    }
    else
    {
        DD() = D() - D().oldTime();
        pointDD() = pointD() - pointD().oldTime();
    }
    U() = fvc::ddt(D());

    return st.converged;
}


tmp<vectorField> gpuUnsSolidBase::tractionBoundarySnGrad
(
    const vectorField& traction,
    const scalarField& pressure,
    const fvPatch& patch
) const
{
    // host versions of unsLinGeomSolid.C:193-230, unsNonLinGeomTotalLagSolid.C:420-488 and unsNonLinGeomUpdatedLagSolid.C:373-419 for
    // boundary conditions evaluated on the host; the device evaluates its traction patches itself (k_bc_update_uns)
    const label patchID = patch.index();
    const scalarField& pImpK = impK_.boundaryField()[patchID];
    const scalarField& pRImpK = rImpK_.boundaryField()[patchID];
    const tensorField& pGrad = gradDf_.boundaryField()[patchID];
    const symmTensorField& pSigma = sigmaf_.boundaryField()[patchID];
    const vectorField n(patch.nf());

    if (modelEnum_ == S4F_MODEL_UNS_LIN_GEOM)
    {
        return tmp<vectorField>
        (
            new vectorField(((traction - n*pressure) - (n & (pSigma - pImpK*pGrad)))*pRImpK)
        );
    }

    // finite strain: the deformed area vector per unit reference area, J F^-T & n, from the face gradient (for the
    // updated-Lagrangian model pGrad is grad(DD)f, so these are the relative tensors)
    const tensorField F(I + pGrad.T());
    const vectorField nCurrent(det(F)*(inv(F).T() & n));
    if (modelEnum_ == S4F_MODEL_UNS_NONLIN_TL)
    {
        return tmp<vectorField>
        (
            new vectorField(((traction - nCurrent*pressure) - (nCurrent & pSigma) + (n & (pImpK*pGrad)))*pRImpK)
        );
    }
    return tmp<vectorField>
    (
        new vectorField(((traction - n*pressure) - (nCurrent & pSigma) + (n & (pImpK*pGrad)))*pRImpK)
    );
}


void gpuUnsSolidBase::setTraction(const label interfaceI, const label patchID, const vectorField& faceZoneTraction)
{
    solidModel::setTraction(interfaceI, patchID, faceZoneTraction);
    const solidTractionFvPatchVectorField& t =
        refCast<const solidTractionFvPatchVectorField>(solved().boundaryField()[patchID]);
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


void gpuUnsSolidBase::updateTotalFields()
{
    if (!incrementalModel())
    {
        gpu_.check(s4fgpu_update_total_fields(gpu_.handle()), "updateTotalFields()");
        solidModel::updateTotalFields();
        return;
    }

    // unsNonLinGeomUpdatedLagSolid.C:423-439: rho_ = rho_.oldTime()/relJ_, moveMesh(oldPoints, DD(), pointDD()),
    // solidModel::updateTotalFields().  The point interpolation, the density and gradient updates and the new geometry of the
    // device copy run on the device (s4fgpu_move_points: also collective on decomposed meshes); the host moves its own
    // polyMesh with the same point displacement for output and the FSI coupler.
    pointVectorField& pDD = pointDD();
    gpu_.check
    (
        s4fgpu_interpolate_to_points
        (
            gpu_.handle(), S4F_FIELD_DD, S4F_POINT_INTERP_PATCH, reinterpret_cast<double*>(pDD.primitiveFieldRef().data())
        ),
        "updateTotalFields()"
    );
    gpu_.check(s4fgpu_update_total_fields(gpu_.handle()), "updateTotalFields()");
    gpu_.check(s4fgpu_move_points(gpu_.handle(), NULL), "updateTotalFields()");

    const vectorField oldPoints(mesh().points());
    moveMesh(oldPoints, DD(), pDD);

    solidModel::updateTotalFields();
}

} // End namespace solidModels
} // End namespace Foam
