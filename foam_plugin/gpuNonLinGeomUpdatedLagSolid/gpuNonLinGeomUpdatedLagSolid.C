/*---------------------------------------------------------------------------*\
  See gpuNonLinGeomUpdatedLagSolid.H.  Source only: needs OpenFOAM + solids4foam to compile.
  The mesh / boundary-condition mirroring is the one of gpuLinGeomTotalDispSolid.C (same C-ABI
  calls); only what differs for the updated-Lagrangian model is spelled out here.
\*---------------------------------------------------------------------------*/
#include "gpuNonLinGeomUpdatedLagSolid.H"
#include "addToRunTimeSelectionTable.H"
#include "fvc.H"
#include "fvm.H"
#include "emptyPolyPatch.H"
#include "symmetryPolyPatch.H"
#include "processorFvPatch.H"
#include "solidTractionFvPatchVectorField.H"
#include "fixedDisplacementFvPatchVectorField.H"
#include "solidSymmetryFvPatchVectorField.H"

namespace Foam
{
namespace solidModels
{

defineTypeNameAndDebug(gpuNonLinGeomUpdatedLagSolid, 0);
addToRunTimeSelectionTable(solidModel, gpuNonLinGeomUpdatedLagSolid, dictionary);   // nonLinGeomUpdatedLagSolid.C:40-43


void gpuNonLinGeomUpdatedLagSolid::check(const int rc, const char* where) const
{
    if (rc != 0)
    {
        FatalErrorIn(where) << "libs4fgpu: " << s4fgpu_last_error(gpu_) << abort(FatalError);
    }
}


// mirrorMesh(), mirrorBoundaryConditions(): identical to gpuLinGeomTotalDispSolid.C (s4fgpu_set_mesh / s4fgpu_set_bc)
// with DD().boundaryField() in place of D().boundaryField(): the fixedDisplacement patches hand over the TOTAL
// displacement, the device subtracts D.oldTime() (fixedDisplacementFvPatchVectorField.C:279-287).


void gpuNonLinGeomUpdatedLagSolid::mirrorGeometry()
{
    // ... the nine geometry arrays exactly as gpuLinGeomTotalDispSolid::mirrorGeometry() builds them, then:
    // check(s4fgpu_set_geometry(gpu_, C, V, Sf, magSf, Cf, w, nod, corr, CnbrB), "mirrorGeometry()");

    // points() and faces() for the vol->point interpolation (enhancedVolPointInterpolation): CSR of the fv faces
    const fvMesh& m = mesh();
    const faceList& fs = m.faces();
    labelList ptr(1, 0), verts;
    DynamicList<label> v;
    for (label faceI = 0; faceI < m.nFaces(); faceI++)
    {
        if (faceI >= m.nInternalFaces() && isA<emptyPolyPatch>(m.boundaryMesh()[m.boundaryMesh().whichPatch(faceI)])) continue;
        forAll(fs[faceI], fp) v.append(fs[faceI][fp]);
        ptr.append(v.size());
    }
    verts.transfer(v);
    check
    (
        s4fgpu_set_points
        (
            gpu_, m.nPoints(), reinterpret_cast<const double*>(m.points().cdata()), ptr.begin(), verts.begin()
        ),
        "mirrorGeometry()"
    );
}


void gpuNonLinGeomUpdatedLagSolid::mirrorLawAndControls()
{
    // as gpuLinGeomTotalDispSolid::mirrorLawAndControls() with
    //   law.kind = S4F_LAW_NEO_HOOKEAN_ELASTIC (or S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC + the hardening table),
    //   mu, K from neoHookeanElastic.C:51-85,  law.solvePressureEqn / pressureSmoothingScaleFactor from the law dict,
    //   c.solidModel = S4F_MODEL_NONLIN_UL,  c.d2dt2Scheme from d2dt2Schemes (steadyState | Euler | backward),
    //   c.fieldRelaxD = fieldRelaxationFactor("DD"),  solver controls from solverDict("DD").
}


void gpuNonLinGeomUpdatedLagSolid::downloadState()
{
    // D, DD, gradD, gradDD, sigma (+ boundary values) as in gpuLinGeomTotalDispSolid::downloadState(), and
    check(s4fgpu_download(gpu_, S4F_FIELD_F, reinterpret_cast<double*>(F_.primitiveFieldRef().data())), "downloadState()");
    check(s4fgpu_download(gpu_, S4F_FIELD_J, J_.primitiveFieldRef().data()), "downloadState()");
    check(s4fgpu_download(gpu_, S4F_FIELD_RHO, rho_.primitiveFieldRef().data()), "downloadState()");
}


gpuNonLinGeomUpdatedLagSolid::gpuNonLinGeomUpdatedLagSolid(Time& runTime, const word& region)
:
    solidModel(typeName, runTime, region),
    F_(IOobject("F", runTime.timeName(), mesh(), IOobject::READ_IF_PRESENT, IOobject::AUTO_WRITE), mesh(), dimensionedTensor("I", dimless, I)),
    J_(IOobject("J", runTime.timeName(), mesh(), IOobject::READ_IF_PRESENT, IOobject::NO_WRITE), det(F_)),
    rho_(IOobject("rho", runTime.timeName(), mesh(), IOobject::READ_IF_PRESENT, IOobject::AUTO_WRITE), mechanical().rho()),
    impK_(mechanical().impK()),
    rImpK_(1.0/impK_),
    gpu_(NULL),
    patchStart_()
{
    DDisRequired();
    fvm::d2dt2(rho_, DD());                     // old-time levels on the host, as the CPU model (:143-145)
    fvc::d2dt2(rho_, D().oldTime());

    if (s4fgpu_create(&gpu_, Pstream::parRun() ? Pstream::myProcNo() % 8 : 0) != 0)
    {
        FatalErrorIn("gpuNonLinGeomUpdatedLagSolid::gpuNonLinGeomUpdatedLagSolid(...)")
            << s4fgpu_last_error(NULL) << abort(FatalError);
    }
    mirrorMesh();
    mirrorGeometry();
    mirrorLawAndControls();
    mirrorBoundaryConditions();
    check(s4fgpu_upload(gpu_, S4F_FIELD_D, reinterpret_cast<const double*>(D().internalField().cdata())), "ctor");
    check(s4fgpu_upload(gpu_, S4F_FIELD_D_OLD, reinterpret_cast<const double*>(D().oldTime().internalField().cdata())), "ctor");
    check(s4fgpu_upload(gpu_, S4F_FIELD_F, reinterpret_cast<const double*>(F_.internalField().cdata())), "ctor");
    check(s4fgpu_initialise(gpu_), "ctor");
}


gpuNonLinGeomUpdatedLagSolid::~gpuNonLinGeomUpdatedLagSolid()
{
    s4fgpu_destroy(gpu_);
}


bool gpuNonLinGeomUpdatedLagSolid::evolve()
{
    Info<< "Evolving solid solver on the GPU" << nl
        << "Solving the updated Lagrangian form of the momentum equation for DD" << endl;

    check(s4fgpu_new_timestep(gpu_, runTime().deltaTValue()), "evolve()");
    mirrorBoundaryConditions();

    s4fgpu_stats st;
    check(s4fgpu_evolve(gpu_, &st), "evolve()");      // the do-while loop nonLinGeomUpdatedLagSolid.C:166-240

    downloadState();

    // mechanical().interpolate(DD(), gradDD(), pointDD()) (:249), on the device
    check
    (
        s4fgpu_interpolate_to_points
        (
            gpu_, S4F_FIELD_DD, S4F_POINT_INTERP_GRAD, reinterpret_cast<double*>(pointDD().primitiveFieldRef().data())
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
    check
    (
        s4fgpu_interpolate_to_points
        (
            gpu_, S4F_FIELD_DD, S4F_POINT_INTERP_PATCH, reinterpret_cast<double*>(pDD.primitiveFieldRef().data())
        ),
        "updateTotalFields()"
    );

    // rho_ = rho_.oldTime()/relJ_, gradD = fvc::grad(D.oldTime() + DD), law history: on the device
    check(s4fgpu_update_total_fields(gpu_), "updateTotalFields()");

    // symmetry-plane / empty corrections and mesh().movePoints(newPoints): the reference's own host code.  Note that
    // solidModel::moveMesh starts by interpolating DD to the points again on the host (solidModel.C:2020); the device
    // values above are the same field (tests/test_gpu_parity.py: vol->point parity), so either may be kept.
    const vectorField oldPoints(mesh().points());
    moveMesh(oldPoints, DD(), pDD);

    // the moved geometry goes back to the device; fields, boundary data and history are kept
    mirrorGeometry();

    solidModel::updateTotalFields();
}

} // End namespace solidModels
} // End namespace Foam
