/*---------------------------------------------------------------------------*\
  See gpuLinGeomTotalDispSolid.H.  Every s4fgpu_* call below is declared in include/s4fgpu.h
  with the reference interface it replaces.
\*---------------------------------------------------------------------------*/
#include "gpuLinGeomTotalDispSolid.H"
#include "addToRunTimeSelectionTable.H"
#include "fvm.H"
#include "fvc.H"
#include "processorFvPatch.H"
#include "symmetryPolyPatch.H"
#include "emptyPolyPatch.H"
#include "solidTractionFvPatchVectorField.H"
#include "fixedDisplacementFvPatchVectorField.H"
#include "solidSymmetryFvPatchVectorField.H"
#include "linearElastic.H"
#include "Pstream.H"

namespace Foam
{
namespace solidModels
{

defineTypeNameAndDebug(gpuLinGeomTotalDispSolid, 0);
addToRunTimeSelectionTable(solidModel, gpuLinGeomTotalDispSolid, dictionary);


void gpuLinGeomTotalDispSolid::check(const int rc, const char* where) const
{
    if (rc != 0)
    {
        FatalErrorIn(where)
            << "libs4fgpu: " << s4fgpu_last_error(gpu_) << abort(FatalError);
    }
}


void gpuLinGeomTotalDispSolid::mirrorMesh()
{
    const fvMesh& m = mesh();
    const lduAddressing& addr = m.lduAddr();

    const label nPatches = m.boundary().size();
    labelList pStart(nPatches), pSize(nPatches), pKind(nPatches), pNbr(nPatches, -1);
    labelList faceCells(m.nFaces() - m.nInternalFaces());

    forAll(m.boundary(), patchI)
    {
        const fvPatch& p = m.boundary()[patchI];
        pStart[patchI] = p.start() - m.nInternalFaces();
        pSize[patchI] = isA<emptyPolyPatch>(p.patch()) ? 0 : p.size();
        pKind[patchI] = S4F_PATCH_GENERIC;
        if (isA<emptyPolyPatch>(p.patch())) pKind[patchI] = S4F_PATCH_EMPTY;
        if (isA<symmetryPolyPatch>(p.patch())) pKind[patchI] = S4F_PATCH_SYMMETRY;
        if (isA<processorFvPatch>(p))
        {
            pKind[patchI] = S4F_PATCH_PROCESSOR;
            pNbr[patchI] = refCast<const processorFvPatch>(p).neighbProcNo();
        }
        const labelUList& fc = p.faceCells();
        forAll(fc, i) faceCells[pStart[patchI] + i] = fc[i];
    }
    patchStart_ = pStart;

    // empty directions are not solved (fvMatrix::solveSegregated skips them)
    const Vector<label>& sD = m.solutionD();
    int solD[3] = {sD[0] > 0, sD[1] > 0, sD[2] > 0};

    check
    (
        s4fgpu_set_mesh
        (
            gpu_, m.nCells(), m.nInternalFaces(),
            addr.lowerAddr().begin(), addr.upperAddr().begin(),
            nPatches, pStart.begin(), pSize.begin(), pKind.begin(), pNbr.begin(),
            faceCells.begin(), solD
        ),
        "gpuLinGeomTotalDispSolid::mirrorMesh()"
    );
}


void gpuLinGeomTotalDispSolid::mirrorGeometry()
{
    // vector = 3 contiguous scalars, so List<vector>::cdata() is the AoS array the C-ABI expects.
    const fvMesh& m = mesh();
    const label nF = m.nFaces(), nI = m.nInternalFaces(), nB = nF - nI;

    vectorField Sf(nF), Cf(nF), corr(nF, vector::zero), CnbrB(nB);
    scalarField magSf(nF), w(nF, 1.0), nod(nF);

    SubList<vector>(Sf, nI) = m.Sf().internalField();
    SubList<vector>(Cf, nI) = m.Cf().internalField();
    SubList<scalar>(magSf, nI) = m.magSf().internalField();
    SubList<scalar>(w, nI) = m.weights().internalField();
    SubList<scalar>(nod, nI) = m.nonOrthDeltaCoeffs().internalField();
    SubList<vector>(corr, nI) = m.nonOrthCorrectionVectors().internalField();

    forAll(m.boundary(), patchI)
    {
        const fvPatch& p = m.boundary()[patchI];
        if (isA<emptyPolyPatch>(p.patch())) continue;
        const label s = p.start();
        SubList<vector>(Sf, p.size(), s) = p.Sf();
        SubList<vector>(Cf, p.size(), s) = p.Cf();
        SubList<scalar>(magSf, p.size(), s) = p.magSf();
        SubList<scalar>(nod, p.size(), s) = m.nonOrthDeltaCoeffs().boundaryField()[patchI];
        if (p.coupled())
        {
            SubList<scalar>(w, p.size(), s) = m.weights().boundaryField()[patchI];
            SubList<vector>(corr, p.size(), s) = m.nonOrthCorrectionVectors().boundaryField()[patchI];
            SubList<vector>(CnbrB, p.size(), s - nI) =
                m.C().boundaryField()[patchI].patchNeighbourField();
        }
        else
        {
            SubList<vector>(CnbrB, p.size(), s - nI) = p.Cf();
        }
    }

    check
    (
        s4fgpu_set_geometry
        (
            gpu_,
            reinterpret_cast<const double*>(m.C().internalField().cdata()),
            m.V().field().cdata(),
            reinterpret_cast<const double*>(Sf.cdata()), magSf.cdata(),
            reinterpret_cast<const double*>(Cf.cdata()), w.cdata(), nod.cdata(),
            reinterpret_cast<const double*>(corr.cdata()),
            reinterpret_cast<const double*>(CnbrB.cdata())
        ),
        "gpuLinGeomTotalDispSolid::mirrorGeometry()"
    );
}


void gpuLinGeomTotalDispSolid::mirrorLawAndControls()
{
    // The law shell (gpuLinearElastic, registered in the linGeomMechLaw table) has parsed
    // mechanicalProperties exactly as linearElastic.C:62-133 does and exposes mu, K, lambda.
    s4fgpu_law law;
    memset(&law, 0, sizeof(law));
    const dictionary& lawDict =
        mechanical().mechanicalProperties().subDict("mechanical").subDict(mechanical()[0].name());
    law.kind = S4F_LAW_LINEAR_ELASTIC;
    law.rho = mechanical().rho()().internalField()[0];
    const scalar E = dimensionedScalar(lawDict.lookup("E")).value();
    const scalar nu = dimensionedScalar(lawDict.lookup("nu")).value();
    law.mu = E/(2.0*(1.0 + nu));
    law.lambda = mechanical().planeStress()
      ? nu*E/((1.0 + nu)*(1.0 - nu)) : nu*E/((1.0 + nu)*(1.0 - 2.0*nu));
    law.K = mechanical().planeStress() ? E/(3.0*(1.0 - nu)) : E/(3.0*(1.0 - 2.0*nu));
    law.updateBEbarConsistent = 1;
    law.DEpsilonPRelax = 1.0;
    check(s4fgpu_set_law(gpu_, &law), "gpuLinGeomTotalDispSolid::mirrorLawAndControls()");

    s4fgpu_controls c;
    memset(&c, 0, sizeof(c));
    c.solidModel = S4F_MODEL_LIN_GEOM_TOTAL_DISP;
    const word gradScheme(mesh().gradSchemes().lookupOrDefault<word>("default", "leastSquares"));
    c.gradScheme = (gradScheme == "Gauss") ? S4F_GRAD_GAUSS_LINEAR : S4F_GRAD_LEAST_SQUARES;
    const word d2dt2(mesh().d2dt2Schemes().lookupOrDefault<word>("default", "steadyState"));
    c.d2dt2Scheme = (d2dt2 == "Euler") ? S4F_D2DT2_EULER : S4F_D2DT2_STEADY_STATE;
    const dictionary& stab = solidModelDict().subOrEmptyDict("stabilisation");
    c.stabilisation = (stab.lookupOrDefault<word>("type", "RhieChow") == "none") ? S4F_STAB_NONE : S4F_STAB_RHIE_CHOW;
    c.stabScaleFactor = stab.lookupOrDefault<scalar>("scaleFactor", 0.1);
    c.relaxationMethod =
        (solidModelDict().lookupOrDefault<word>("relaxationMethod", "fixed") == "Aitken") ? S4F_RELAX_AITKEN : S4F_RELAX_FIXED;
    c.fieldRelaxD = mesh().relaxField("D") ? mesh().fieldRelaxationFactor("D") : 1.0;
    const dictionary& sol = mesh().solverDict("D");
    c.solver = S4F_SOLVER_PCG;
    const dictionary& gpuDict = solidModelDict().subOrEmptyDict("gpu");
    const word pre(gpuDict.lookupOrDefault<word>("preconditioner", "GAMG"));
    c.preconditioner = (pre == "diagonal") ? S4F_PRECOND_DIAGONAL : (pre == "none") ? S4F_PRECOND_NONE : S4F_PRECOND_GAMG;
    c.gamgSinglePrecision = gpuDict.lookupOrDefault<Switch>("gamgSinglePrecision", false);
    c.gamgOverCorrection = gpuDict.lookupOrDefault<scalar>("gamgOverCorrection", 2.2);
    c.gamgSmootherDegree = gpuDict.lookupOrDefault<label>("gamgSmootherDegree", 3);
    c.gamgCycle = gpuDict.lookupOrDefault<label>("gamgCycle", 2);          // K-cycle on level 1
    c.tolerance = sol.lookupOrDefault<scalar>("tolerance", 1e-6);
    c.relTol = sol.lookupOrDefault<scalar>("relTol", 0);
    c.maxIter = sol.lookupOrDefault<label>("maxIter", 1000);
    c.nCorrectors = nCorr();
    c.solutionTolerance = solutionTol();
    c.alternativeTolerance = alternativeTol();
    c.materialTolerance = materialTol();
    c.g[0] = g().value().x(); c.g[1] = g().value().y(); c.g[2] = g().value().z();
    c.deltaT = runTime().deltaTValue();
    c.deltaT0 = runTime().deltaT0Value();
    c.checkEvery = gpuDict.lookupOrDefault<label>("checkEvery", 4);
    check(s4fgpu_set_controls(gpu_, &c), "gpuLinGeomTotalDispSolid::mirrorLawAndControls()");
}


void gpuLinGeomTotalDispSolid::mirrorBoundaryConditions()
{
    forAll(D().boundaryField(), patchI)
    {
        const fvPatchVectorField& pf = D().boundaryField()[patchI];
        if (isA<solidTractionFvPatchVectorField>(pf))
        {
            const solidTractionFvPatchVectorField& t = refCast<const solidTractionFvPatchVectorField>(pf);
            check
            (
                s4fgpu_set_bc
                (
                    gpu_, patchI, S4F_BC_SOLID_TRACTION,
                    reinterpret_cast<const double*>(t.traction().cdata()), t.pressure().cdata()
                ),
                "mirrorBoundaryConditions()"
            );
        }
        else if (isA<fixedDisplacementFvPatchVectorField>(pf))
        {
            check
            (
                s4fgpu_set_bc
                (
                    gpu_, patchI, S4F_BC_FIXED_DISPLACEMENT,
                    reinterpret_cast<const double*>(pf.cdata()), NULL
                ),
                "mirrorBoundaryConditions()"
            );
        }
        else if (isA<solidSymmetryFvPatchVectorField>(pf))
        {
            check(s4fgpu_set_bc(gpu_, patchI, S4F_BC_SOLID_SYMMETRY, NULL, NULL), "mirrorBoundaryConditions()");
        }
        else if (pf.coupled())
        {
            check(s4fgpu_set_bc(gpu_, patchI, S4F_BC_PROCESSOR, NULL, NULL), "mirrorBoundaryConditions()");
        }
        else if (pf.size())
        {
            FatalErrorIn("gpuLinGeomTotalDispSolid::mirrorBoundaryConditions()")
                << "Patch " << pf.patch().name() << ": boundary condition " << pf.type()
                << " is not available on the GPU path (solidTraction, fixedDisplacement, "
                << "solidSymmetry, processor)" << abort(FatalError);
        }
    }
}


void gpuLinGeomTotalDispSolid::uploadState()
{
    check(s4fgpu_upload(gpu_, S4F_FIELD_D, reinterpret_cast<const double*>(D().internalField().cdata())), "uploadState()");
    check(s4fgpu_upload(gpu_, S4F_FIELD_D_OLD, reinterpret_cast<const double*>(D().oldTime().internalField().cdata())), "uploadState()");
    check
    (
        s4fgpu_upload(gpu_, S4F_FIELD_D_OLDOLD, reinterpret_cast<const double*>(D().oldTime().oldTime().internalField().cdata())),
        "uploadState()"
    );
}


void gpuLinGeomTotalDispSolid::downloadState()
{
    // host fields stay the source of truth for I/O, function objects and FSI coupling
    vectorField& Di = D().primitiveFieldRef();
    check(s4fgpu_download(gpu_, S4F_FIELD_D, reinterpret_cast<double*>(Di.data())), "downloadState()");
    check(s4fgpu_download(gpu_, S4F_FIELD_GRAD_D, reinterpret_cast<double*>(gradD().primitiveFieldRef().data())), "downloadState()");
    check(s4fgpu_download(gpu_, S4F_FIELD_SIGMA, reinterpret_cast<double*>(sigma().primitiveFieldRef().data())), "downloadState()");

    const label nB = mesh().nFaces() - mesh().nInternalFaces();
    vectorField Db(nB); tensorField gDb(nB); symmTensorField sb(nB);
    check(s4fgpu_download(gpu_, S4F_FIELD_D_B, reinterpret_cast<double*>(Db.data())), "downloadState()");
    check(s4fgpu_download(gpu_, S4F_FIELD_GRAD_D_B, reinterpret_cast<double*>(gDb.data())), "downloadState()");
    check(s4fgpu_download(gpu_, S4F_FIELD_SIGMA_B, reinterpret_cast<double*>(sb.data())), "downloadState()");
    forAll(D().boundaryField(), patchI)
    {
        const label n = D().boundaryField()[patchI].size(), s = patchStart_[patchI];
        if (n == 0 || D().boundaryField()[patchI].coupled()) continue;
        D().boundaryFieldRef()[patchI] == SubList<vector>(Db, n, s);
        gradD().boundaryFieldRef()[patchI] = SubList<tensor>(gDb, n, s);
        sigma().boundaryFieldRef()[patchI] = SubList<symmTensor>(sb, n, s);
    }
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
    gpu_(NULL),
    patchStart_()
{
    DisRequired();

    // old-time fields exist on the host exactly as for the CPU model (linGeomTotalDispSolid.C:79)
    fvm::d2dt2(D());

    const dictionary& gpuDict = solidModelDict().subOrEmptyDict("gpu");
    const label device = gpuDict.lookupOrDefault<label>("device", Pstream::parRun() ? Pstream::myProcNo() % 8 : 0);
    if (s4fgpu_create(&gpu_, device) != 0)
    {
        FatalErrorIn("gpuLinGeomTotalDispSolid::gpuLinGeomTotalDispSolid(...)")
            << s4fgpu_last_error(NULL) << abort(FatalError);
    }

    if (Pstream::parRun())
    {
        // NCCL bootstrap over the existing Pstream: the master makes the id, everyone gets it
        List<char> id(128);
        if (Pstream::master()) s4fgpu_get_unique_id(id.begin());
        Pstream::scatter(id);
        check(s4fgpu_comm_init(gpu_, Pstream::nProcs(), Pstream::myProcNo(), id.begin()), "comm_init");
    }

    mirrorMesh();
    mirrorGeometry();
    mirrorLawAndControls();
    mirrorBoundaryConditions();
    uploadState();

    // D.correctBoundaryConditions(); D.storePrevIter(); mechanical().grad(D, gradD)  (:82-84)
    check(s4fgpu_initialise(gpu_), "gpuLinGeomTotalDispSolid::gpuLinGeomTotalDispSolid(...)");
}


gpuLinGeomTotalDispSolid::~gpuLinGeomTotalDispSolid()
{
    s4fgpu_destroy(gpu_);
}


bool gpuLinGeomTotalDispSolid::evolve()
{
    Info<< "Evolving solid solver on the GPU" << endl;

    check(s4fgpu_new_timestep(gpu_, runTime().deltaTValue()), "evolve()");
    mirrorBoundaryConditions();     // time-varying tractions / displacements

    s4fgpu_stats st;
    check(s4fgpu_evolve(gpu_, &st), "evolve()");

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
    check
    (
        s4fgpu_set_bc
        (
            gpu_, patchID, S4F_BC_SOLID_TRACTION,
            reinterpret_cast<const double*>(t.traction().cdata()), t.pressure().cdata()
        ),
        "setTraction()"
    );
}


void gpuLinGeomTotalDispSolid::updateTotalFields()
{
    check(s4fgpu_update_total_fields(gpu_), "updateTotalFields()");
    solidModel::updateTotalFields();
}


void gpuLinGeomTotalDispSolid::writeFields(const Time& runTime)
{
    solidModel::writeFields(runTime);
}

} // End namespace solidModels
} // End namespace Foam
