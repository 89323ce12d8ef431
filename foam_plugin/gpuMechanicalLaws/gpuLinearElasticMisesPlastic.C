/*---------------------------------------------------------------------------*\
  See gpuLinearElasticMisesPlastic.H.  Source only: needs OpenFOAM + solids4foam to compile.
\*---------------------------------------------------------------------------*/
#include "gpuLinearElasticMisesPlastic.H"
#include "addToRunTimeSelectionTable.H"
#include "lookupSolidModel.H"

namespace Foam
{
    defineTypeNameAndDebug(gpuLinearElasticMisesPlastic, 0);
    addToRunTimeSelectionTable(mechanicalLaw, gpuLinearElasticMisesPlastic, linGeomMechLaw);      // as linearElasticMisesPlastic.C:33-36
}


Foam::gpuLinearElasticMisesPlastic::gpuLinearElasticMisesPlastic
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
    // the same two ways of giving the elastic constants, the same formulas as linearElastic (linearElasticMisesPlastic.C:586-636)
    if (dict.found("E") && dict.found("nu"))
    {
        const scalar E = dimensionedScalar(dict.lookup("E")).value();
        const scalar nu = dimensionedScalar(dict.lookup("nu")).value();
        mu_.value() = E/(2.0*(1.0 + nu));
        K_.value() = planeStress() ? E/(3.0*(1.0 - nu)) : E/(3.0*(1.0 - 2.0*nu));
    }
    else if (dict.found("mu") && dict.found("K"))
    {
        mu_ = dimensionedScalar(dict.lookup("mu"));
        K_ = dimensionedScalar(dict.lookup("K"));
    }
    else
    {
        FatalErrorIn("gpuLinearElasticMisesPlastic::gpuLinearElasticMisesPlastic(...)") << "Either E and nu or mu and K elastic parameters should be specified" << abort(FatalError);
    }

    memset(&pod_, 0, sizeof(pod_));
    pod_.kind = S4F_LAW_LINEAR_ELASTIC_MISES_PLASTIC;
    pod_.rho = rho()().internalField()[0];
    pod_.mu = mu_.value(); pod_.K = K_.value(); pod_.lambda = K_.value() - (2.0/3.0)*mu_.value();
    // the (epsilonP sigmaY) table: s4f's own interpolationTable<scalar> (numerics/interpolationTable), piece-wise linear, clamped;
    // 2 points = linear hardening (Hp), 1 point = perfect plasticity (linearElasticMisesPlastic.C:640-660): the device distinguishes by nTable
    if (stressPlasticStrainSeries_.size() < 1 || stressPlasticStrainSeries_.size() > 64)
    {
        FatalErrorIn("gpuLinearElasticMisesPlastic::gpuLinearElasticMisesPlastic(...)") << "the hardening table must hold 1 to 64 points on the GPU path" << abort(FatalError);
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


Foam::gpuLinearElasticMisesPlastic::~gpuLinearElasticMisesPlastic()
{}


Foam::tmp<Foam::volScalarField> Foam::gpuLinearElasticMisesPlastic::impK() const
{
    // 4/3 mu + K  (linearElasticMisesPlastic.C:676-715 at DLambda = 0)
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


Foam::tmp<Foam::volScalarField> Foam::gpuLinearElasticMisesPlastic::bulkModulus() const
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


void Foam::gpuLinearElasticMisesPlastic::correct(volSymmTensorField& sigma)
{
    // a gpu* solidModel evaluates the law inside its device loop (k_law_lin_mises) and fills sigma itself
    if (word(lookupSolidModel(mesh()).type()).substr(0, 3) == "gpu") return;

    FatalErrorIn("gpuLinearElasticMisesPlastic::correct(volSymmTensorField&)")
        << "gpuLinearElasticMisesPlastic keeps its state on the device and runs under the gpu* solid models only; with a CPU solidModel "
        << "select linearElasticMisesPlastic" << abort(FatalError);
}


void Foam::gpuLinearElasticMisesPlastic::correct(surfaceSymmTensorField& sigma)
{
    notImplemented("gpuLinearElasticMisesPlastic::correct(surfaceSymmTensorField&): the face-stress form is not on the GPU path for this law");
}
