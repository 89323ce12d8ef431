/*---------------------------------------------------------------------------*\
  See gpuLinearElastic.H.  Source only: needs OpenFOAM + solids4foam to compile.
\*---------------------------------------------------------------------------*/
#include "gpuLinearElastic.H"
#include "addToRunTimeSelectionTable.H"
#include "lookupSolidModel.H"

namespace Foam
{
    defineTypeNameAndDebug(gpuLinearElastic, 0);
    addToRunTimeSelectionTable(mechanicalLaw, gpuLinearElastic, linGeomMechLaw);      // as linearElastic.C:30-34
}


Foam::gpuLinearElastic::gpuLinearElastic
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
    lambda_("lambda", dimPressure, 0.0),
    gpu_(NULL)
{
    // the same two ways of giving the elastic constants, the same formulas (linearElastic.C:62-133)
    if (dict.found("E") && dict.found("nu"))
    {
        const scalar E = dimensionedScalar(dict.lookup("E")).value();
        const scalar nu = dimensionedScalar(dict.lookup("nu")).value();
        if (nu < -1.0 || nu > 0.5)
        {
            FatalErrorIn("gpuLinearElastic::gpuLinearElastic(...)")
                << "Unphysical Poisson's ratio: nu should be >= -1.0 and <= 0.5" << abort(FatalError);
        }
        mu_.value() = E/(2.0*(1.0 + nu));
        if (nu < 0.5)
        {
            lambda_.value() = planeStress() ? nu*E/((1.0 + nu)*(1.0 - nu)) : nu*E/((1.0 + nu)*(1.0 - 2.0*nu));
            K_.value() = planeStress() ? E/(3.0*(1.0 - nu)) : E/(3.0*(1.0 - 2.0*nu));
        }
        else
        {
            lambda_.value() = GREAT;
            K_.value() = GREAT;
        }
    }
    else if (dict.found("mu") && dict.found("K"))
    {
        mu_ = dimensionedScalar(dict.lookup("mu"));
        K_ = dimensionedScalar(dict.lookup("K"));
        const scalar E = 9.0*K_.value()*mu_.value()/(3.0*K_.value() + mu_.value());
        const scalar nu = (3.0*K_.value() - 2.0*mu_.value())/(2.0*(3.0*K_.value() + mu_.value()));
        if (planeStress())             // linearElastic.C:107-113: lambda and K are reset for plane stress
        {
            lambda_.value() = nu*E/((1.0 + nu)*(1.0 - nu));
            K_.value() = E/(3.0*(1.0 - nu));
        }
        else
        {
            lambda_.value() = nu*E/((1.0 + nu)*(1.0 - 2.0*nu));
        }
    }
    else
    {
        FatalErrorIn("gpuLinearElastic::gpuLinearElastic(...)")
            << "Either E and nu or mu and K elastic parameters should be specified" << abort(FatalError);
    }

    memset(&pod_, 0, sizeof(pod_));
    pod_.kind = S4F_LAW_LINEAR_ELASTIC;
    pod_.rho = rho()().internalField()[0];
    pod_.mu = mu_.value(); pod_.K = K_.value(); pod_.lambda = lambda_.value();
    pod_.updateBEbarConsistent = 1; pod_.DEpsilonPRelax = 1.0;
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


Foam::gpuLinearElastic::~gpuLinearElastic()
{
    if (gpu_) s4fgpu_destroy(gpu_);
}


Foam::tmp<Foam::volScalarField> Foam::gpuLinearElastic::impK() const
{
    return tmp<volScalarField>
    (
        new volScalarField
        (
            IOobject("impK", mesh().time().timeName(), mesh(), IOobject::NO_READ, IOobject::NO_WRITE),
            mesh(),
            (lambda_.value() < 0.1*GREAT) ? 2.0*mu_ + lambda_ : 2.0*mu_          // linearElastic.C:204-245 (nu = 0.5: 2 mu)
        )
    );
}


void Foam::gpuLinearElastic::correct(volSymmTensorField& sigma)
{
    // a gpu* solidModel evaluates the law inside its device loop and fills sigma itself
    if (word(lookupSolidModel(mesh()).type()).substr(0, 3) == "gpu") return;

    // CPU solidModel: Hooke's law for cells and patch values on the device.  The first call mirrors the mesh
    // (s4fgpu_set_mesh / s4fgpu_set_geometry exactly as gpuLinGeomTotalDispSolid::mirrorMesh/mirrorGeometry) and the law.
    const volTensorField& gradD = mesh().lookupObject<volTensorField>("grad(D)");          // mechanicalLaw.C:988-997
    if (!gpu_)
    {
        if (s4fgpu_create(&gpu_, 0) != 0) FatalErrorIn("gpuLinearElastic::correct(...)") << s4fgpu_last_error(NULL) << abort(FatalError);
        // ... mirrorMesh(gpu_, mesh()); mirrorGeometry(gpu_, mesh());  s4fgpu_set_controls (linearGeometryTotalDisplacement) ...
        s4fgpu_set_law(gpu_, &pod_);
    }
    s4fgpu_upload(gpu_, S4F_FIELD_GRAD_D, reinterpret_cast<const double*>(gradD.internalField().cdata()));
    // patch values of grad(D): one flat [B] array in patch order
    // ... s4fgpu_upload(gpu_, S4F_FIELD_GRAD_D_B, flatBoundary(gradD)) ...
    s4fgpu_op_correct(gpu_);                                                                // k_law_linear_elastic
    s4fgpu_download(gpu_, S4F_FIELD_SIGMA, reinterpret_cast<double*>(sigma.primitiveFieldRef().data()));
    // ... s4fgpu_download(gpu_, S4F_FIELD_SIGMA_B, flat) and scatter to sigma.boundaryFieldRef() ...
}


void Foam::gpuLinearElastic::correct(surfaceSymmTensorField& sigma)
{
    // the face-stress form is evaluated by the gpuUnsLinearGeometry solid model (k_uns_face_stress); a CPU uns* model keeps
    // the reference's own law
    notImplemented("gpuLinearElastic::correct(surfaceSymmTensorField&): select gpuUnsLinearGeometry in solidProperties");
}
