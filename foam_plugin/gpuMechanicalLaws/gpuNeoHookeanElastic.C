/*---------------------------------------------------------------------------*\
  See gpuNeoHookeanElastic.H.  Source only: needs OpenFOAM + solids4foam to compile.
\*---------------------------------------------------------------------------*/
#include "gpuNeoHookeanElastic.H"
#include "addToRunTimeSelectionTable.H"
#include "lookupSolidModel.H"

namespace Foam
{
    defineTypeNameAndDebug(gpuNeoHookeanElastic, 0);
    addToRunTimeSelectionTable(mechanicalLaw, gpuNeoHookeanElastic, nonLinGeomMechLaw);      // as neoHookeanElastic.C:27-31
}


Foam::gpuNeoHookeanElastic::gpuNeoHookeanElastic
(
    const word& name,
    const fvMesh& mesh,
    const dictionary& dict,
    const nonLinearGeometry::nonLinearType& nonLinGeom
)
:
    mechanicalLaw(name, mesh, dict, nonLinGeom),
    mu_("mu", dimPressure, 0.0),
    K_("K", dimPressure, 0.0)
{
    // the same two ways of giving the elastic constants, the same formulas (neoHookeanElastic.C:51-85)
    if (dict.found("E") && dict.found("nu") && !dict.found("mu") && !dict.found("K"))
    {
        const dimensionedScalar E = dimensionedScalar(dict.lookup("E"));
        const dimensionedScalar nu = dimensionedScalar(dict.lookup("nu"));
        mu_ = E/(2.0*(1.0 + nu));
        if (planeStress()) K_ = (nu*E/((1.0 + nu)*(1.0 - nu))) + (2.0/3.0)*mu_;
        else K_ = (nu*E/((1.0 + nu)*(1.0 - 2.0*nu))) + (2.0/3.0)*mu_;
    }
    else if (dict.found("mu") && dict.found("K") && !dict.found("E") && !dict.found("nu"))
    {
        mu_ = dimensionedScalar(dict.lookup("mu"));
        K_ = dimensionedScalar(dict.lookup("K"));
    }
    else
    {
        FatalErrorIn("gpuNeoHookeanElastic::gpuNeoHookeanElastic(...)") << "Either E and nu or mu and K should be specified" << abort(FatalError);
    }

    memset(&pod_, 0, sizeof(pod_));
    pod_.kind = S4F_LAW_NEO_HOOKEAN_ELASTIC;
    pod_.rho = rho()().internalField()[0];
    pod_.mu = mu_.value(); pod_.K = K_.value(); pod_.lambda = K_.value() - (2.0/3.0)*mu_.value();
    pod_.updateBEbarConsistent = dict.lookupOrDefault<Switch>("updateBEbarConsistent", true);
    pod_.DEpsilonPRelax = mesh.relaxField("DEpsilonP") ? mesh.fieldRelaxationFactor("DEpsilonP") : 1.0;
    pod_.solvePressureEqn = dict.lookupOrDefault<Switch>("solvePressureEqn", false);               // mechanicalLaw.C:1525-1532
    pod_.pressureSmoothingScaleFactor = dict.lookupOrDefault<scalar>("pressureSmoothingScaleFactor", 100.0);
    if (pod_.solvePressureEqn)
    {
        // sigmaHydEqn.solve(); sigmaHyd.relax()  (mechanicalLaw.C:1455-1459): fvSolution solvers / relaxationFactors "sigmaHyd"
        const dictionary& sd = mesh.solverDict("sigmaHyd");
        pod_.sigmaHydTolerance = sd.lookupOrDefault<scalar>("tolerance", 1e-6);
        pod_.sigmaHydRelTol = sd.lookupOrDefault<scalar>("relTol", 0);
        pod_.sigmaHydMaxIter = sd.lookupOrDefault<label>("maxIter", 1000);
        pod_.sigmaHydRelax = mesh.relaxField("sigmaHyd") ? mesh.fieldRelaxationFactor("sigmaHyd") : 1.0;
    }
}


Foam::gpuNeoHookeanElastic::~gpuNeoHookeanElastic()
{}


Foam::tmp<Foam::volScalarField> Foam::gpuNeoHookeanElastic::impK() const
{
    // 4/3 mu + K  (neoHookeanElastic.C:101-119)
    return tmp<volScalarField>
    (
        new volScalarField
        (
            IOobject("impK", mesh().time().timeName(), mesh(), IOobject::NO_READ, IOobject::NO_WRITE),
            mesh(),
            (4.0/3.0)*mu_ + K_
        )
    );
}


Foam::tmp<Foam::volScalarField> Foam::gpuNeoHookeanElastic::bulkModulus() const
{
    return tmp<volScalarField>
    (
        new volScalarField
        (
            IOobject("bulkModulus", mesh().time().timeName(), mesh(), IOobject::NO_READ, IOobject::NO_WRITE),
            mesh(),
            K_
        )
    );
}


void Foam::gpuNeoHookeanElastic::correct(volSymmTensorField& sigma)
{
    // a gpu* solidModel evaluates the law inside its device loop (k_law_neo_hookean) and fills sigma itself
    if (word(lookupSolidModel(mesh()).type()).substr(0, 3) == "gpu") return;

    FatalErrorIn("gpuNeoHookeanElastic::correct(volSymmTensorField&)")
        << "gpuNeoHookeanElastic keeps its state on the device and runs under the gpu* solid models only; with a CPU solidModel "
        << "select neoHookeanElastic" << abort(FatalError);
}


void Foam::gpuNeoHookeanElastic::correct(surfaceSymmTensorField& sigma)
{
    notImplemented("gpuNeoHookeanElastic::correct(surfaceSymmTensorField&): the face-stress form is not on the GPU path for this law");
}
