/*---------------------------------------------------------------------------*\
  See gpuNeoHookeanElasticMisesPlastic.H.  Source only: needs OpenFOAM + solids4foam to compile.
\*---------------------------------------------------------------------------*/
#include "gpuNeoHookeanElasticMisesPlastic.H"
#include "addToRunTimeSelectionTable.H"
#include "lookupSolidModel.H"

namespace Foam
{
    defineTypeNameAndDebug(gpuNeoHookeanElasticMisesPlastic, 0);
    addToRunTimeSelectionTable(mechanicalLaw, gpuNeoHookeanElasticMisesPlastic, nonLinGeomMechLaw);      // as neoHookeanElasticMisesPlastic.C:33-36
}


Foam::gpuNeoHookeanElasticMisesPlastic::gpuNeoHookeanElasticMisesPlastic
(
    const word& name,
    const fvMesh& mesh,
    const dictionary& dict,
    const nonLinearGeometry::nonLinearType& nonLinGeom
)
:
    mechanicalLaw(name, mesh, dict, nonLinGeom),
    mu_("mu", dimPressure, 0.0),
    K_("K", dimPressure, 0.0),
    stressPlasticStrainSeries_(dict)
{
    // the same two ways of giving the elastic constants, the same formulas (neoHookeanElasticMisesPlastic.C:868-906)
    if (dict.found("E") && dict.found("nu"))
    {
        const dimensionedScalar E = dimensionedScalar(dict.lookup("E"));
        const dimensionedScalar nu = dimensionedScalar(dict.lookup("nu"));
        mu_ = E/(2.0*(1.0 + nu));
        if (planeStress()) K_ = (nu*E/((1.0 + nu)*(1.0 - nu))) + (2.0/3.0)*mu_;
        else K_ = (nu*E/((1.0 + nu)*(1.0 - 2.0*nu))) + (2.0/3.0)*mu_;
    }
    else if (dict.found("mu") && dict.found("K"))
    {
        mu_ = dimensionedScalar(dict.lookup("mu"));
        K_ = dimensionedScalar(dict.lookup("K"));
    }
    else
    {
        FatalErrorIn("gpuNeoHookeanElasticMisesPlastic::gpuNeoHookeanElasticMisesPlastic(...)") << "Either E and nu or mu and K should be specified" << abort(FatalError);
    }

    memset(&pod_, 0, sizeof(pod_));
    pod_.kind = S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC;
    pod_.rho = rho()().internalField()[0];
    pod_.mu = mu_.value(); pod_.K = K_.value(); pod_.lambda = K_.value() - (2.0/3.0)*mu_.value();
    // the (epsilonP sigmaY) table: s4f's own interpolationTable<scalar> (numerics/interpolationTable), piece-wise linear, clamped;
    // 2 points = linear hardening (Hp), 1 point = perfect plasticity (neoHookeanElasticMisesPlastic.C:909-930): the device distinguishes by nTable
    if (stressPlasticStrainSeries_.size() < 1 || stressPlasticStrainSeries_.size() > 64)
    {
        FatalErrorIn("gpuNeoHookeanElasticMisesPlastic::gpuNeoHookeanElasticMisesPlastic(...)") << "the hardening table must hold 1 to 64 points on the GPU path" << abort(FatalError);
    }
    pod_.nTable = stressPlasticStrainSeries_.size();
    forAll(stressPlasticStrainSeries_, i)
    {
        pod_.tableEps[i] = stressPlasticStrainSeries_[i].first();
        pod_.tableSigY[i] = stressPlasticStrainSeries_[i].second();
    }
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


Foam::gpuNeoHookeanElasticMisesPlastic::~gpuNeoHookeanElasticMisesPlastic()
{}


Foam::tmp<Foam::volScalarField> Foam::gpuNeoHookeanElasticMisesPlastic::impK() const
{
    // 4/3 mu + K at DLambda = 0, as evaluated once by the solid model constructors (neoHookeanElasticMisesPlastic.C:949-988)
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


Foam::tmp<Foam::volScalarField> Foam::gpuNeoHookeanElasticMisesPlastic::bulkModulus() const
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


void Foam::gpuNeoHookeanElasticMisesPlastic::correct(volSymmTensorField& sigma)
{
    // a gpu* solidModel evaluates the law inside its device loop (k_law_mises) and fills sigma itself
    if (word(lookupSolidModel(mesh()).type()).substr(0, 3) == "gpu") return;

    FatalErrorIn("gpuNeoHookeanElasticMisesPlastic::correct(volSymmTensorField&)")
        << "gpuNeoHookeanElasticMisesPlastic keeps its state on the device and runs under the gpu* solid models only; with a CPU solidModel "
        << "select neoHookeanElasticMisesPlastic" << abort(FatalError);
}


void Foam::gpuNeoHookeanElasticMisesPlastic::correct(surfaceSymmTensorField& sigma)
{
    notImplemented("gpuNeoHookeanElasticMisesPlastic::correct(surfaceSymmTensorField&): the face-stress form is not on the GPU path for this law");
}
