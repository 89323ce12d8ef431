/*---------------------------------------------------------------------------*\
  See gpuSolidBridge.H.  Every s4fgpu_* call below is declared in include/s4fgpu.h with the reference
  interface it replaces.  Source only: needs OpenFOAM + solids4foam to compile.
\*---------------------------------------------------------------------------*/
#include "gpuSolidBridge.H"
#include "gpuLawParameters.H"
#include "processorFvPatch.H"
#include "symmetryPolyPatch.H"
#include "symmetryPlanePolyPatch.H"
#include "emptyPolyPatch.H"
#include "solidTractionFvPatchVectorField.H"
#include "fixedDisplacementFvPatchVectorField.H"
#include "solidSymmetryFvPatchVectorField.H"
#include "Pstream.H"

namespace Foam
{

gpuSolidBridge::gpuSolidBridge(const fvMesh& mesh, const dictionary& gpuDict)
:
    mesh_(mesh),
    gpu_(NULL),
    patchStart_(),
    timeIndex_(-1)
{
    // one rank <-> one GPU of the box
    const label device = gpuDict.lookupOrDefault<label>("device", Pstream::parRun() ? Pstream::myProcNo() % 8 : 0);
    if (s4fgpu_create(&gpu_, device) != 0)
    {
        FatalErrorIn("gpuSolidBridge::gpuSolidBridge(...)") << s4fgpu_last_error(NULL) << abort(FatalError);
    }
    if (Pstream::parRun())
    {
        // bootstrap over the existing Pstream: the master makes the id, everyone gets it (collective)
        List<char> id(128);
        if (Pstream::master()) s4fgpu_get_unique_id(id.begin());
        Pstream::scatter(id);
        check(s4fgpu_comm_init(gpu_, Pstream::nProcs(), Pstream::myProcNo(), id.begin()), "gpuSolidBridge: comm_init");
    }
}


gpuSolidBridge::~gpuSolidBridge()
{
    s4fgpu_destroy(gpu_);
}


void gpuSolidBridge::check(const int rc, const char* where) const
{
    if (rc != 0)
    {
        FatalErrorIn(where) << "libs4fgpu: " << s4fgpu_last_error(gpu_) << abort(FatalError);
    }
}


void gpuSolidBridge::mirrorMesh()
{
    const fvMesh& m = mesh_;
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
        if (isA<symmetryPolyPatch>(p.patch()) || isA<symmetryPlanePolyPatch>(p.patch())) pKind[patchI] = S4F_PATCH_SYMMETRY;
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
        "gpuSolidBridge::mirrorMesh()"
    );
}


void gpuSolidBridge::mirrorGeometry(const bool withPoints)
{
    // vector = 3 contiguous scalars, so List<vector>::cdata() is the AoS array the C-ABI expects.
    const fvMesh& m = mesh_;
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
            SubList<vector>(CnbrB, p.size(), s - nI) = m.C().boundaryField()[patchI].patchNeighbourField();
        }
        else
        {
            SubList<vector>(CnbrB, p.size(), s - nI) = p.Cf();
        }
    }

    // Decomposed case: the points first -- s4fgpu_set_geometry then identifies the processor-patch points across ranks and
    // reserves the point-neighbour ghost slots of the point stencils (include/s4fgpu.h, s4fgpu_set_points).  Serial: the
    // library builds the vol->point weights inside set_points, which needs the geometry.
    if (withPoints && Pstream::parRun()) mirrorPoints();

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
        "gpuSolidBridge::mirrorGeometry()"
    );

    if (withPoints && !Pstream::parRun()) mirrorPoints();
}


void gpuSolidBridge::mirrorPoints()
{
    const fvMesh& m = mesh_;
    const label nF = m.nFaces(), nI = m.nInternalFaces();

    // points() and faces() of the fv faces (empty-patch faces are not fv faces of the mirror): CSR
    const faceList& fs = m.faces();
    DynamicList<label> ptr, verts;
    ptr.append(0);
    for (label faceI = 0; faceI < nF; faceI++)
    {
        if (faceI >= nI && isA<emptyPolyPatch>(m.boundaryMesh()[m.boundaryMesh().whichPatch(faceI)])) continue;
        forAll(fs[faceI], fp) verts.append(fs[faceI][fp]);
        ptr.append(verts.size());
    }
    check
    (
        s4fgpu_set_points(gpu_, m.nPoints(), reinterpret_cast<const double*>(m.points().cdata()), ptr.begin(), verts.begin()),
        "gpuSolidBridge::mirrorGeometry()"
    );
}


void gpuSolidBridge::mirrorLaw(const mechanicalModel& mechanical)
{
    if (mechanical.size() != 1)
    {
        FatalErrorIn("gpuSolidBridge::mirrorLaw(...)")
            << "The GPU path covers the single-law branch of mechanicalModel (mechanicalModel.C:476-483); "
            << mechanical.size() << " laws were specified" << abort(FatalError);
    }
    const gpuLawParameters* lawPtr = dynamic_cast<const gpuLawParameters*>(&mechanical[0]);
    if (!lawPtr)
    {
        FatalErrorIn("gpuSolidBridge::mirrorLaw(...)")
            << "mechanical law " << mechanical[0].type() << " has no device counterpart; select one of gpuLinearElastic, "
            << "gpuNeoHookeanElastic, gpuNeoHookeanElasticMisesPlastic, gpuLinearElasticMisesPlastic in mechanicalProperties"
            << abort(FatalError);
    }
    check(s4fgpu_set_law(gpu_, &lawPtr->pod()), "gpuSolidBridge::mirrorLaw(...)");
}


void gpuSolidBridge::mirrorControls
(
    const int solidModelEnum,
    const word& solvedField,
    const dictionary& solidModelDict,
    const loopControls& lc,
    const vector& g
)
{
    const fvMesh& m = mesh_;
    s4fgpu_controls c;
    memset(&c, 0, sizeof(c));
    c.solidModel = solidModelEnum;

    // fvSchemes gradSchemes: "leastSquares" | "pointCellsLeastSquares" | "Gauss linear" (the first word decides)
    {
        ITstream& is = m.gradSchemes().found("grad(" + solvedField + ")")
          ? m.gradSchemes().lookup("grad(" + solvedField + ")") : m.gradSchemes().lookup("default");
        const word scheme(is);
        if (scheme == "Gauss") c.gradScheme = S4F_GRAD_GAUSS_LINEAR;
        else if (scheme == "pointCellsLeastSquares" || scheme == "edgeCellsLeastSquares") c.gradScheme = S4F_GRAD_POINT_CELLS_LEAST_SQUARES;
        else if (scheme == "leastSquares" || scheme == "extendedLeastSquares") c.gradScheme = S4F_GRAD_LEAST_SQUARES;
        else
        {
            FatalErrorIn("gpuSolidBridge::mirrorControls(...)")
                << "gradScheme " << scheme << " is not available on the GPU path (leastSquares, pointCellsLeastSquares, Gauss linear)"
                << abort(FatalError);
        }
    }
    // d2dt2Schemes: steadyState | Euler | backward (numerics/backwardD2dt2Scheme)
    {
        const word d2dt2(m.d2dt2Schemes().lookupOrDefault<word>("default", "steadyState"));
        if (d2dt2 == "steadyState") c.d2dt2Scheme = S4F_D2DT2_STEADY_STATE;
        else if (d2dt2 == "Euler") c.d2dt2Scheme = S4F_D2DT2_EULER;
        else if (d2dt2 == "backward") c.d2dt2Scheme = S4F_D2DT2_BACKWARD;
        else
        {
            FatalErrorIn("gpuSolidBridge::mirrorControls(...)")
                << "d2dt2Scheme " << d2dt2 << " is not available on the GPU path (steadyState, Euler, backward)" << abort(FatalError);
        }
    }
    // momentumStabilisation.C:43-81
    {
        const dictionary& stab = solidModelDict.subOrEmptyDict("stabilisation");
        const word type(stab.lookupOrDefault<word>("type", "RhieChow"));
        if (type == "none") c.stabilisation = S4F_STAB_NONE;
        else if (type == "RhieChow") c.stabilisation = S4F_STAB_RHIE_CHOW;
        else
        {
            FatalErrorIn("gpuSolidBridge::mirrorControls(...)")
                << "stabilisation type " << type << " is not available on the GPU path (RhieChow, none)" << abort(FatalError);
        }
        c.stabScaleFactor = stab.lookupOrDefault<scalar>("scaleFactor", 0.1);
    }
    c.relaxationMethod =
        (solidModelDict.lookupOrDefault<word>("relaxationMethod", "fixed") == "Aitken") ? S4F_RELAX_AITKEN : S4F_RELAX_FIXED;
    c.fieldRelaxD = m.relaxField(solvedField) ? m.fieldRelaxationFactor(solvedField) : 1.0;
    if (m.relaxEquation(solvedField) && mag(m.equationRelaxationFactor(solvedField) - 1.0) > SMALL)
    {
        FatalErrorIn("gpuSolidBridge::mirrorControls(...)")
            << "equation relaxation of " << solvedField << " is not available on the GPU path" << abort(FatalError);
    }

    // fvSolution solvers <field>: solver PCG | PBiCGStab (GAMG as a solver is run as PCG + GAMG preconditioner);
    // preconditioner DIC | FDIC | DILU (exact, level-scheduled) | diagonal | none | GAMG.  The "gpu" sub-dictionary of
    // the solid model's coefficients may override the preconditioner: the fast one on the device is GAMG.
    const dictionary& sol = m.solverDict(solvedField);
    const dictionary& gpuDict = solidModelDict.subOrEmptyDict("gpu");
    {
        const word solver(sol.lookup("solver"));
        word pre("none");
        if (solver == "GAMG") pre = "GAMG";
        else if (sol.found("preconditioner"))
        {
            if (sol.isDict("preconditioner")) pre = word(sol.subDict("preconditioner").lookup("preconditioner"));
            else pre = word(sol.lookup("preconditioner"));
        }
        pre = gpuDict.lookupOrDefault<word>("preconditioner", pre);
        if (solver == "PCG" || solver == "GAMG" || solver == "PBiCG") c.solver = S4F_SOLVER_PCG;      // the matrix is symmetric
        else if (solver == "PBiCGStab") c.solver = S4F_SOLVER_PBICGSTAB;
        else
        {
            FatalErrorIn("gpuSolidBridge::mirrorControls(...)") << "solver " << solver << " is not available on the GPU path" << abort(FatalError);
        }
        if (pre == "DIC" || pre == "FDIC" || pre == "DILU") c.preconditioner = S4F_PRECOND_DIC;
        else if (pre == "diagonal") c.preconditioner = S4F_PRECOND_DIAGONAL;
        else if (pre == "none") c.preconditioner = S4F_PRECOND_NONE;
        else if (pre == "GAMG") c.preconditioner = S4F_PRECOND_GAMG;
        else
        {
            FatalErrorIn("gpuSolidBridge::mirrorControls(...)") << "preconditioner " << pre << " is not available on the GPU path" << abort(FatalError);
        }
    }
    c.gamgSinglePrecision = gpuDict.lookupOrDefault<Switch>("gamgSinglePrecision", false);
    c.gamgOverCorrection = gpuDict.lookupOrDefault<scalar>("gamgOverCorrection", 2.2);
    c.gamgSmootherDegree = gpuDict.lookupOrDefault<label>("gamgSmootherDegree", 3);
    c.gamgCycle = gpuDict.lookupOrDefault<label>("gamgCycle", 2);          // K-cycle on level 1
    c.tolerance = sol.lookupOrDefault<scalar>("tolerance", 1e-6);
    c.relTol = sol.lookupOrDefault<scalar>("relTol", 0);
    c.maxIter = sol.lookupOrDefault<label>("maxIter", 1000);
    c.nCorrectors = lc.nCorr;
    c.solutionTolerance = lc.solutionTol;
    c.alternativeTolerance = lc.alternativeTol;
    c.materialTolerance = lc.materialTol;
    c.g[0] = g.x(); c.g[1] = g.y(); c.g[2] = g.z();
    c.deltaT = m.time().deltaTValue();
    c.deltaT0 = m.time().deltaT0Value();
    c.chebyshevDegree = 4;
    c.checkEvery = gpuDict.lookupOrDefault<label>("checkEvery", 4);
    check(s4fgpu_set_controls(gpu_, &c), "gpuSolidBridge::mirrorControls(...)");
}


void gpuSolidBridge::mirrorBoundaryConditions(volVectorField& Dsolved)
{
    forAll(Dsolved.boundaryField(), patchI)
    {
        const fvPatchVectorField& pf = Dsolved.boundaryField()[patchI];
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
                "gpuSolidBridge::mirrorBoundaryConditions(...)"
            );
        }
        else if (isA<fixedDisplacementFvPatchVectorField>(pf))
        {
            // updateCoeffs() evaluates the displacement time series on the host (fixedDisplacement...C:258-294); for a DD field
            // the patch then holds the TOTAL displacement minus D.oldTime(): the device wants the total and subtracts itself
            fixedDisplacementFvPatchVectorField& fd =
                refCast<fixedDisplacementFvPatchVectorField>(Dsolved.boundaryFieldRef()[patchI]);
            fd.updateCoeffs();
            vectorField total(fd);
            if (Dsolved.name() == "DD")
            {
                total += mesh_.lookupObject<volVectorField>("D").oldTime().boundaryField()[patchI];
            }
            check
            (
                s4fgpu_set_bc(gpu_, patchI, S4F_BC_FIXED_DISPLACEMENT, reinterpret_cast<const double*>(total.cdata()), NULL),
                "gpuSolidBridge::mirrorBoundaryConditions(...)"
            );
        }
        else if (isA<solidSymmetryFvPatchVectorField>(pf))
        {
            check(s4fgpu_set_bc(gpu_, patchI, S4F_BC_SOLID_SYMMETRY, NULL, NULL), "gpuSolidBridge::mirrorBoundaryConditions(...)");
        }
        else if (pf.coupled())
        {
            check(s4fgpu_set_bc(gpu_, patchI, S4F_BC_PROCESSOR, NULL, NULL), "gpuSolidBridge::mirrorBoundaryConditions(...)");
        }
        else if (pf.size())
        {
            FatalErrorIn("gpuSolidBridge::mirrorBoundaryConditions(...)")
                << "Patch " << pf.patch().name() << ": boundary condition " << pf.type()
                << " is not available on the GPU path (solidTraction, fixedDisplacement, solidSymmetry, processor)"
                << abort(FatalError);
        }
    }
}


void gpuSolidBridge::newTimeStepIfNeeded()
{
    // runTime++ drives the old-time roll in the reference; evolve() itself may run several times within one time step
    // (FSI strong-coupling iterations call it once per coupling iteration with a new interface traction)
    const label ti = mesh_.time().timeIndex();
    if (ti != timeIndex_)
    {
        check(s4fgpu_new_timestep(gpu_, mesh_.time().deltaTValue()), "gpuSolidBridge::newTimeStepIfNeeded()");
        timeIndex_ = ti;
    }
}


void gpuSolidBridge::upload(const int fieldId, const double* host, const char* where)
{
    check(s4fgpu_upload(gpu_, fieldId, host), where);
}


void gpuSolidBridge::download(const int fieldId, double* host, const char* where) const
{
    check(s4fgpu_download(gpu_, fieldId, host), where);
}


// host fields stay the source of truth for I/O, function objects and FSI coupling: cells, then the patch values from
// the flat B-array of the C-ABI
#define S4F_DOWNLOAD_FIELD(NAME, FieldType, Type)                                                          \
void gpuSolidBridge::NAME(FieldType& f, const int idInternal, const int idBoundary) const                  \
{                                                                                                          \
    download(idInternal, reinterpret_cast<double*>(f.primitiveFieldRef().data()), #NAME);                  \
    if (idBoundary < 0) return;                                                                            \
    const label nB = mesh_.nFaces() - mesh_.nInternalFaces();                                              \
    Field<Type> b(nB);                                                                                     \
    download(idBoundary, reinterpret_cast<double*>(b.data()), #NAME);                                      \
    forAll(f.boundaryField(), patchI)                                                                      \
    {                                                                                                      \
        const label n = f.boundaryField()[patchI].size(), s = patchStart_[patchI];                         \
        if (n == 0 || f.boundaryField()[patchI].coupled()) continue;                                       \
        f.boundaryFieldRef()[patchI] == SubList<Type>(b, n, s);                                            \
    }                                                                                                      \
}

S4F_DOWNLOAD_FIELD(downloadVector, volVectorField, vector)
S4F_DOWNLOAD_FIELD(downloadTensor, volTensorField, tensor)
S4F_DOWNLOAD_FIELD(downloadSymmTensor, volSymmTensorField, symmTensor)

#undef S4F_DOWNLOAD_FIELD

} // End namespace Foam
