/*
 * s4fgpu.h -- C-ABI of libs4fgpu.so: the B200-resident hot path of solids4foam's segregated
 * cell-centred finite-volume solid solver.
 *
 * This is the drop-in boundary.  The OpenFOAM-side plugin
 * (foam_plugin/: class gpuLinGeomTotalDispSolid : public solidModel, ... registered in the
 * reference's run-time-selection tables) mirrors the fvMesh / lduAddressing and the fields into
 * device-resident SoA arrays ONCE per mesh through these calls and then drives the
 * momentum-correction loop on the device.  Every entry point names the reference interface it
 * replaces (paths relative to /root/reference; SM = src/solids4FoamModels/solidModels,
 * ML = src/solids4FoamModels/materialModels/mechanicalModel/mechanicalLaws, [OF-ext] = behaviour of
 * the OpenFOAM library itself, which is not vendored in the reference).
 *
 * Conventions
 *  - plain C, opaque handle, caller-owned host buffers, callee-owned device buffers;
 *  - every function returns 0 on success, non-zero on error; s4fgpu_last_error() gives the text
 *    (the plugin turns that into FatalErrorIn(...) << abort(FatalError));
 *  - host arrays are in OpenFOAM's AoS layouts: vector = 3 doubles, symmTensor = 6 doubles
 *    (XX XY XZ YY YZ ZZ), tensor = 9 doubles row-major (XX XY XZ YX YY YZ ZX ZY ZZ),
 *    label = int32;  gradD_ij = d_i D_j as in OpenFOAM;
 *  - faces: internal faces [0,F) in lduAddressing order, then boundary faces [F,F+B) patch by
 *    patch (empty patches have size 0);  "B-arrays" are indexed by boundary face (face - F);
 *  - one handle per rank / GPU; a handle is not thread-safe; streams are internal.
 *  - there is NO CPU fallback anywhere behind this interface.
 */
#ifndef S4FGPU_H
#define S4FGPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct s4fgpu_ctx* s4fgpu_handle;

/* ---- enumerations ------------------------------------------------------------------------- */

/* fvPatch kinds (polyMesh/boundary "type") */
enum { S4F_PATCH_GENERIC = 0, S4F_PATCH_EMPTY = 1, S4F_PATCH_SYMMETRY = 2, S4F_PATCH_PROCESSOR = 3 };

/* boundary conditions of D (0/D "type"); SM/fvPatchFields/... */
enum {
    S4F_BC_FIXED_DISPLACEMENT = 0, /* fixedDisplacement/fixedDisplacementFvPatchVectorField.C:258-356 */
    S4F_BC_SOLID_TRACTION = 1,     /* solidTraction/solidTractionFvPatchVectorField.C:321-463 (also the
                                      carrier for analyticalPlateHoleTraction: host supplies the tractions) */
    S4F_BC_SOLID_SYMMETRY = 2,     /* solidSymmetry/solidSymmetryFvPatchVectorField.C:148-260 */
    S4F_BC_PROCESSOR = 3           /* [OF-ext] processorFvPatchField */
};

/* solidModel (constant/solidProperties "solidModel") */
enum {
    S4F_MODEL_LIN_GEOM_TOTAL_DISP = 0,   /* SM/linGeomTotalDispSolid/linGeomTotalDispSolid.C:111-232 */
    S4F_MODEL_NONLIN_TL_TOTAL_DISP = 1,  /* SM/nonLinGeomTotalLagTotalDispSolid/...C:173-281 */
    S4F_MODEL_NONLIN_TL = 2,             /* SM/nonLinGeomTotalLagSolid/nonLinGeomTotalLagSolid.C:125-260 (solves DD) */
    S4F_MODEL_NONLIN_UL = 3,             /* SM/nonLinGeomUpdatedLagSolid/...C:159-273 */
    S4F_MODEL_UNS_LIN_GEOM = 4,          /* SM/unsLinGeomSolid/unsLinGeomSolid.C:100-175 ("unsLinearGeometry"): face stresses from face
                                            gradients built on the vertex displacements, fvc::div(mesh().Sf() & sigmaf);
                                            linearElastic, needs s4fgpu_set_points; decomposed meshes: processor faces take the
                                            corrected snGrad of an internal face, vertex values see the other ranks' cells */
    S4F_MODEL_UNS_NONLIN_TL = 5,         /* SM/unsNonLinGeomTotalLagSolid/unsNonLinGeomTotalLagSolid.C:218-405
                                            ("unsNonLinearGeometryTotalLagrangian"): the same face gradients, Ff = I + gradDf.T()
                                            on the faces, neoHookeanElastic::correct(surfaceSymmTensorField&)
                                            (neoHookeanElastic.C:306-352), fvc::div((Jf Finvf.T() & Sf) & sigmaf); its own
                                            convergence criterion (:49-76, :333-378); neoHookeanElastic; the enforceLinear
                                            fall-back of the reference is not implemented */
    S4F_MODEL_UNS_NONLIN_UL = 6          /* SM/unsNonLinGeomUpdatedLagSolid/unsNonLinGeomUpdatedLagSolid.C:247-345
                                            ("unsNonLinearGeometryUpdatedLagrangian"): solves DD on the updated configuration;
                                            relFf = I + gradDDf.T(), Ff = relFf & Ff.oldTime() on the faces,
                                            fvc::div((relJf relFinvf.T() & Sf) & sigmaf), the updated-Lagrangian inertia terms,
                                            solidModel::converged; mesh motion as S4F_MODEL_NONLIN_UL; neoHookeanElastic */
};

/* mechanicalLaw (constant/mechanicalProperties "type") */
enum {
    S4F_LAW_LINEAR_ELASTIC = 0,                 /* ML/linearGeometryLaws/linearElastic/linearElastic.C */
    S4F_LAW_NEO_HOOKEAN_ELASTIC = 1,            /* ML/nonLinearGeometryLaws/neoHookeanElastic/neoHookeanElastic.C */
    S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC = 2,      /* ML/nonLinearGeometryLaws/neoHookeanElasticMisesPlastic/...C */
    S4F_LAW_LINEAR_ELASTIC_MISES_PLASTIC = 3    /* ML/linearGeometryLaws/linearElasticMisesPlastic/...C */
};

/* system/fvSchemes gradSchemes */
enum {
    S4F_GRAD_LEAST_SQUARES = 0, /* 1/|d|^2-weighted LS: numerics/extendedLeastSquaresGrad/extendedLeastSquaresVectors.C:121-158,229-272 */
    S4F_GRAD_GAUSS_LINEAR = 1,  /* [OF-ext] gaussGrad + linear */
    S4F_GRAD_POINT_CELLS_LEAST_SQUARES = 2 /* [OF-ext] LeastSquaresGrad<centredCPCCellToCellStencilObject> ("pointCellsLeastSquares",
                                   the scheme the tutorials use on OpenFOAM.com/.org, applications/scripts/solids4FoamScripts.sh:162-176):
                                   1/|d|^2-weighted least squares over the cells sharing a point with the cell and the boundary faces
                                   at its points; needs s4fgpu_set_points; on decomposed meshes the stencil includes the cells and
                                   boundary faces of other ranks at shared points (see s4fgpu_set_points) */
};
enum { S4F_D2DT2_STEADY_STATE = 0, S4F_D2DT2_EULER = 1, S4F_D2DT2_BACKWARD = 2 };
enum { S4F_STAB_NONE = 0, S4F_STAB_RHIE_CHOW = 1 };        /* SM/solidModel/momentumStabilisation/momentumStabilisation.C:210-217 */
enum { S4F_RELAX_FIXED = 0, S4F_RELAX_AITKEN = 1 };        /* SM/solidModel/solidModel.C:823-906 */
enum { S4F_SOLVER_PCG = 0, S4F_SOLVER_PBICGSTAB = 1 };     /* [OF-ext] PCG.C / PBiCGStab.C */
enum {
    S4F_PRECOND_NONE = 0,
    S4F_PRECOND_DIAGONAL = 1,   /* [OF-ext] diagonalPreconditioner (Jacobi) */
    S4F_PRECOND_DIC = 2,        /* [OF-ext] DICPreconditioner / FDIC: the sequential face sweeps evaluated level by level on the
                                   device (the CPU solver's iteration counts) */
    S4F_PRECOND_CHEBYSHEV = 3,  /* GPU polynomial preconditioner (no reference counterpart) */
    S4F_PRECOND_GAMG = 4        /* [OF-ext] GAMG (agglomeration multigrid) used as PCG preconditioner: pair-wise
                                   agglomeration like faceAreaPair, Galerkin coarse matrices, V- or K-cycle with a
                                   Chebyshev-Jacobi smoother, dense solve on the coarsest level; set up on the device
                                   (one rank) or per rank with distributed levels (decomposed), coefficients refreshed
                                   on the device when a later matrix has the same graph */
};

/* field ids for upload / download */
enum {
    S4F_FIELD_D = 0,        /* vector [N]   */
    S4F_FIELD_D_OLD = 1,    /* vector [N]   D.oldTime()          */
    S4F_FIELD_D_OLDOLD = 2, /* vector [N]   D.oldTime().oldTime()*/
    S4F_FIELD_GRAD_D = 3,   /* tensor [N]   */
    S4F_FIELD_SIGMA = 4,    /* symmTensor [N] */
    S4F_FIELD_D_B = 5,      /* vector [B]   boundary values of D */
    S4F_FIELD_GRAD_D_B = 6, /* tensor [B]   */
    S4F_FIELD_SIGMA_B = 7,  /* symmTensor [B] */
    S4F_FIELD_SOURCE = 8,   /* vector [N]   fvMatrix source after addBoundarySource (debug/parity) */
    S4F_FIELD_DIAG = 9,     /* vector [N]   per-component diagonal after addBoundaryDiag */
    S4F_FIELD_UPPER = 10,   /* scalar [F]   */
    S4F_FIELD_EPSILON_P_EQ = 11, /* scalar [N] */
    S4F_FIELD_SIGMA_Y = 12,      /* scalar [N] */
    S4F_FIELD_BEBAR = 13,        /* symmTensor [N] */
    S4F_FIELD_DLAMBDA = 14,      /* scalar [N] */
    S4F_FIELD_J = 15,            /* scalar [N] */
    S4F_FIELD_F = 16,            /* tensor [N] */
    S4F_FIELD_GRAD_D_OLD = 17,   /* tensor [N] */
    S4F_FIELD_DEPSILON_P = 18,   /* symmTensor [N] */
    S4F_FIELD_TRACTION_GRADIENT_B = 19, /* vector [B]: fixedGradient gradient() of traction patches */
    S4F_FIELD_EPSILON_P = 20,    /* symmTensor [N] */
    S4F_FIELD_DD = 21,           /* vector [N]  displacement increment (incremental solid models only) */
    S4F_FIELD_GRAD_DD = 22,      /* tensor [N] */
    S4F_FIELD_RHO = 23,          /* scalar [N]  density field of the updated-Lagrangian model (rho_ = rho_.oldTime()/relJ_) */
    S4F_FIELD_DD_B = 24,         /* vector [B]  boundary values of DD */
    S4F_FIELD_SIGMA_HYD = 25,    /* scalar [N]  hydrostatic stress of the law (mechanicalLaw::sigmaHyd()) */
    S4F_FIELD_GRAD_SIGMA_HYD = 26, /* vector [N] */
    S4F_FIELD_SIGMA_F = 27,      /* symmTensor [F+B]  face stress sigmaf of the uns* model (internal faces, then boundary faces) */
    S4F_FIELD_GRAD_D_F = 28      /* tensor [F+B]      face gradient gradDf */
};

/* ---- parameter blocks ---------------------------------------------------------------------- */

/* mechanicalProperties entry, already reduced by the law shell exactly as the CPU law constructors
 * do (linearElastic.C:62-133, neoHookeanElastic.C:51-85, neoHookeanElasticMisesPlastic.C:868-930). */
typedef struct {
    int kind;            /* S4F_LAW_* */
    double rho;
    double mu, K, lambda; /* lambda used by linearElastic impK (2mu+lambda) */
    double sigma0[6];    /* linearElastic initial stress */
    int nTable;          /* plasticity: points of (epsilonP, sigmaY) table, <= 64 */
    double tableEps[64];
    double tableSigY[64];
    int updateBEbarConsistent; /* neoHookeanElasticMisesPlastic.C:846-853, default 1 */
    double DEpsilonPRelax;     /* fvSolution relaxationFactors fields DEpsilonP (1 = none) */
    int solvePressureEqn;      /* mechanicalLaw.C:1525-1528: smooth sigmaHyd with the pressure Poisson equation :1374-1468 */
    double pressureSmoothingScaleFactor; /* :1529-1532, default 100 */
    /* fvSolution "solvers sigmaHyd" and "relaxationFactors fields sigmaHyd" of the pressure equation (sigmaHydEqn.solve();
     * sigmaHyd.relax(); mechanicalLaw.C:1455-1459).  tolerance <= 0: the D solver's tolerance / relTol / maxIter;
     * relax <= 0 or 1: no relaxation */
    double sigmaHydTolerance, sigmaHydRelTol;
    int sigmaHydMaxIter;
    double sigmaHydRelax;
} s4fgpu_law;

/* solidProperties <model>Coeffs + fvSchemes + fvSolution entries used on the path */
typedef struct {
    int solidModel;        /* S4F_MODEL_* */
    int gradScheme;        /* S4F_GRAD_* */
    int d2dt2Scheme;       /* S4F_D2DT2_* */
    int stabilisation;     /* S4F_STAB_* */
    double stabScaleFactor;/* default RhieChow 0.1: SM/solidModel/solidModel.C:1307-1315 */
    int relaxationMethod;  /* S4F_RELAX_* */
    double fieldRelaxD;    /* fvSolution relaxationFactors fields D (1 = none) */
    int solver;            /* S4F_SOLVER_* */
    int preconditioner;    /* S4F_PRECOND_* */
    double tolerance;      /* fvSolution solvers D tolerance */
    double relTol;
    int maxIter;           /* [OF-ext] lduMatrix::solver default 1000 */
    int nCorrectors;       /* SM/solidModel/solidModel.C:1167 */
    double solutionTolerance, alternativeTolerance, materialTolerance; /* :1147-1162 */
    double g[3];           /* constant/g */
    double deltaT, deltaT0;
    int chebyshevDegree;   /* S4F_PRECOND_CHEBYSHEV only */
    int checkEvery;        /* stream-launched solves (DIC, Chebyshev, PBiCGStab): host polls the device-side convergence flags every n iterations;
                              PCG with the diagonal / GAMG preconditioner runs as one CUDA graph with a device-side loop instead */
    int gamgSinglePrecision;   /* S4F_PRECOND_GAMG: 1 = V-cycle in fp32 (PCG itself stays fp64), 0 = fp64 */
    double gamgOverCorrection; /* S4F_PRECOND_GAMG: fixed scaling of the coarse-grid correction (<= 0: 2.2) */
    int gamgSmootherDegree;    /* S4F_PRECOND_GAMG: Chebyshev-Jacobi degree of the pre- and post-smoother (<= 0: 3) */
    int gamgCycle;             /* S4F_PRECOND_GAMG: 0 = V-cycle, 1 = W-cycle, 2 = K-cycle on level 1 (two flexible-CG steps) */
    double gamgSmootherRatio;  /* S4F_PRECOND_GAMG: lower end of the Chebyshev interval as a fraction of the upper (<= 0: 0.3) */
} s4fgpu_controls;

/* result of one outer (momentum-correction) iteration / of evolve(); the numbers of the
 * "Corr, res, relRes, matRes, iters" log line, SM/solidModel/solidModelTemplates.C:153-163 */
typedef struct {
    int nCorr;                 /* outer iterations done in this evolve() */
    int converged;
    double initialResidual[3]; /* SolverPerformance<vector>::initialResidual() of the last solve */
    double finalResidual[3];
    int nIterations[3];
    double solverPerfInitRes;  /* mag(initialResidual) */
    double relResidual;        /* residualvf */
    double materialResidual;
    long long totalInnerIterations; /* sum over components and outer iterations */
} s4fgpu_stats;

/* ---- life cycle ---------------------------------------------------------------------------- */

int s4fgpu_create(s4fgpu_handle* h, int device);
int s4fgpu_destroy(s4fgpu_handle h);
const char* s4fgpu_last_error(s4fgpu_handle h);   /* h may be NULL: last creation error */
int s4fgpu_version(void);

/* Parallel runs ([OF-ext] Pstream / processor patches).  Rank 0 obtains an id, the host
 * broadcasts it (Pstream in the plugin, torch.distributed in the python harness), every rank calls
 * comm_init (collective).  Without comm_init the handle is a serial run.  NCCL carries the set-up
 * traffic only; halos, reductions and the coarse-level gather of the iteration run over cudaIpc-mapped
 * peer memory (NVLink), so every GPU of the run must be peer-accessible from every other one. */
int s4fgpu_get_unique_id(char id[128]);
int s4fgpu_comm_init(s4fgpu_handle h, int nRanks, int rank, const char id[128]);

/* ---- mesh mirror (once per mesh) ------------------------------------------------------------ */

/* fvMesh::owner()/neighbour() (lduAddressing lower/upper), fvBoundaryMesh patches with faceCells.
 * Replaces nothing in solids4foam itself: this is the data solidModel's base constructor reaches
 * through mesh() (SM/solidModel/solidModel.C:954-1330).  solutionD: 1 solved / 0 empty direction. */
int s4fgpu_set_mesh(s4fgpu_handle h, int nCells, int nInternalFaces,
                    const int* owner, const int* neighbour,
                    int nPatches, const int* patchStart, const int* patchSize,
                    const int* patchKind, const int* patchNbrRank,
                    const int* faceCells, const int* solutionD);

/* fvMesh geometry: C() V() Sf() magSf() Cf() weights() nonOrthDeltaCoeffs()
 * nonOrthCorrectionVectors(); CnbrB = C().boundaryField().patchNeighbourField() on processor faces
 * (ignored elsewhere).  Re-callable after mesh motion (nonLinGeomUpdatedLagSolid.C:360-374). */
int s4fgpu_set_geometry(s4fgpu_handle h, const double* C, const double* V, const double* Sf,
                        const double* magSf, const double* Cf, const double* weights,
                        const double* nonOrthDeltaCoeffs, const double* nonOrthCorrVec,
                        const double* CnbrB);

/* ---- point mesh: vol -> point interpolation and mesh motion --------------------------------- */

/* polyMesh points() and faces() of the fv faces mirrored by set_mesh (internal faces, then the boundary
 * faces in patch order; faceVertsPtr [F+B+1] CSR offsets into faceVerts).  Builds pointCells and the
 * boundary pointFaces addressing that volPointInterpolation needs.  Re-callable with moved points (same
 * topology): the inverse-distance weights are recomputed from the geometry of the last set_geometry.
 * Decomposed meshes (s4fgpu_comm_init, processor patches): call this BEFORE s4fgpu_set_geometry.  set_geometry then
 * identifies the points of the processor patches across ranks by their coordinates (bit for bit -- decomposePar writes
 * them from one set of points), and reserves, behind the boundary slots of the field index space, one slot per cell and
 * per boundary face of ANOTHER rank that shares a point with this rank's cells ("point-neighbour ghosts"; also ranks that
 * touch this one only along an edge or at a corner).  One peer-memory exchange fills them before every operator with a
 * point stencil (pointCellsLeastSquares gradient, vol->point interpolation, the uns face gradients), which each rank then
 * evaluates completely by itself: where OpenFOAM synchronises partial sums over globalMeshData's shared points
 * ([OF-ext] volPointInterpolation / syncTools), every rank here holds the same point value by construction.
 * Reference: mesh().points()/faces() as used by enhancedVolPointInterpolation.C:60-250
 * (src/blockCoupledSolids4FoamTools/enhancedVolPointInterpolation) and solidModel::moveMesh
 * (SM/solidModel/solidModel.C:2008-2148). */
int s4fgpu_set_points(s4fgpu_handle h, int nPoints, const double* points, const int* faceVertsPtr,
                      const int* faceVerts);

/* mode S4F_POINT_INTERP_PATCH: mechanicalModel::interpolate(D, pointD) (mechanicalModel.C:786-826) ->
 * volToPoint().interpolate(vf, pf) (enhancedVolPointInterpolate.C:425-447): inverse-distance weighting from the
 * cell centres for internal points, from the boundary-face values for points on non-empty, non-coupled patches,
 * then the symmetry-plane point constraint (the interpolation solidModel::moveMesh uses).
 * mode S4F_POINT_INTERP_GRAD: mechanicalModel::interpolate(D, gradD, pointD) (mechanicalModel.C:829-877) ->
 * volToPoint().interpolate(vf, gradVf, pf) (enhancedVolPointInterpolate.C:351-418): every point from its
 * pointCells with the cell gradient extrapolation, pf = sum w (vf + delta & gradVf) / sum w, w = 1/|delta| (the
 * pointD / pointDD the solid models hand to the FSI coupling, e.g. nonLinGeomUpdatedLagSolid.C:249).
 * field = S4F_FIELD_D or S4F_FIELD_DD; pointField [3*nPoints] host. */
enum { S4F_POINT_INTERP_PATCH = 0, S4F_POINT_INTERP_GRAD = 1 };
int s4fgpu_interpolate_to_points(s4fgpu_handle h, int field, int mode, double* pointField);

/* solidModel::moveMesh (SM/solidModel/solidModel.C:2008-2148; called from nonLinGeomUpdatedLagSolid::updateTotalFields,
 * ...C:360-374) with everything fvMesh::movePoints invalidates, ON THE DEVICE: newPoints = points + pointDD (the points
 * of an axis-aligned symmetry plane keep their plane), then face centres / areas, cell centres / volumes, interpolation
 * weights, delta coefficients, correction vectors, least-squares vectors, patch correction vectors and the vol->point
 * weights are recomputed from the new points; fields, boundary data and law history stay, the GAMG hierarchy keeps its
 * aggregates and re-sums its coefficients.  pointDD: host [3*nPoints], or NULL = the point field the last
 * s4fgpu_interpolate_to_points left on the device (no host round trip).  Needs s4fgpu_set_points; meshes with empty
 * patches (2-D cases) are refused: mirror their moved geometry with set_geometry / set_points instead.
 * Decomposed meshes: collective; every rank moves its own points (the interpolated point displacement is identical on
 * all ranks that hold a point), the cell and boundary-face centres are exchanged over the processor patches and the
 * point-neighbour ghosts; the GAMG hierarchy keeps its aggregates there too. */
int s4fgpu_move_points(s4fgpu_handle h, const double* pointDD);

/* ---- models ---------------------------------------------------------------------------------- */

int s4fgpu_set_law(s4fgpu_handle h, const s4fgpu_law* law);
int s4fgpu_set_controls(s4fgpu_handle h, const s4fgpu_controls* c);

/* Boundary condition data of one patch of D.  value: [3*size] prescribed displacement
 * (fixedDisplacement totalDisp_) or traction (solidTraction traction_); pressure: [size] or NULL.
 * Re-callable any time (solidModel::setTraction, SM/solidModel/solidModel.C:1752-1817). */
int s4fgpu_set_bc(s4fgpu_handle h, int patch, int kind, const double* value, const double* pressure);

/* ---- state ------------------------------------------------------------------------------------ */

int s4fgpu_upload(s4fgpu_handle h, int field, const double* host);
int s4fgpu_download(s4fgpu_handle h, int field, double* host);

/* The solidModel constructor's consistent start (linGeomTotalDispSolid.C:82-84):
 * D.correctBoundaryConditions(); D.storePrevIter(); mechanical().grad(D, gradD); builds the
 * device matrix (fvm::laplacian(impKf, D) coefficients, constant while impKf is). */
int s4fgpu_initialise(s4fgpu_handle h);

/* runTime++ : roll D -> D.oldTime() -> oldOld, gradD.oldTime(), law old-time fields. */
int s4fgpu_new_timestep(s4fgpu_handle h, double deltaT);

/* ---- the hot path ------------------------------------------------------------------------------ */

/* One momentum-correction iteration: the body of the do-loop linGeomTotalDispSolid.C:135-192
 * (or the TL/UL equivalents), rows 1-17 of SURVEY.md 3.2, plus the residuals converged() needs. */
int s4fgpu_outer_iteration(s4fgpu_handle h, s4fgpu_stats* stats);

/* solidModel::evolve(): the whole loop with the convergence logic of
 * solidModelTemplates.C:27-188 evaluated from device-side reductions. */
int s4fgpu_evolve(s4fgpu_handle h, s4fgpu_stats* stats);

/* solidModel::updateTotalFields() (solidModel.C:1629-1632) -> law history commit
 * (neoHookeanElasticMisesPlastic.C:1526-1601). */
int s4fgpu_update_total_fields(s4fgpu_handle h);

/* ---- single operators (parity tests, roofline measurement) ------------------------------------- */

/* mechanicalModel::grad(D, gradD)  (mechanicalModel.C:571-582 -> fvc::grad + boundary correction) */
int s4fgpu_op_grad(s4fgpu_handle h);
/* mechanicalModel::correct(sigma)  (mechanicalModel.C:476-483 -> law correct, cells + boundary) */
int s4fgpu_op_correct(s4fgpu_handle h);
/* Assemble the momentum equation of the current state (matrix once, source each call) and leave
 * it in FIELD_SOURCE / FIELD_DIAG / FIELD_UPPER.  linGeomTotalDispSolid.C:141-149. */
int s4fgpu_op_assemble(s4fgpu_handle h);
/* [OF-ext] lduMatrix::Amul for component cmpt: y = A x, host vectors of nCells. */
int s4fgpu_op_amul(s4fgpu_handle h, int cmpt, const double* x, double* y);
/* [OF-ext] fvMatrix<vector>::solveSegregated on the assembled system with host psi in/out
 * ([3N] AoS) and host source ([3N] AoS, after addBoundarySource). */
int s4fgpu_op_solve(s4fgpu_handle h, double* psi, const double* source, s4fgpu_stats* stats);

/* Time `reps` back-to-back launches of one kernel with CUDA events on the library stream.
 * kernel ids: S4F_KERNEL_*.  Returns mean milliseconds per launch and the algorithmic bytes per
 * launch.  flushL2 != 0 writes a >L2 buffer between launches. */
enum {
    S4F_KERNEL_SPMV1 = 0,      /* scalar Amul */
    S4F_KERNEL_SPMV3 = 1,      /* 3-component fused Amul */
    S4F_KERNEL_PCG_ITER = 2,   /* one fused 3-component PCG iteration (all kernels) */
    S4F_KERNEL_GRAD = 3,
    S4F_KERNEL_LAW = 4,
    S4F_KERNEL_RHS = 5,
    S4F_KERNEL_PCG_P = 6,      /* pA = rD rA + beta pA */
    S4F_KERNEL_PCG_XR = 7,     /* psi += alpha pA; rA -= alpha wA; residual sums */
    S4F_KERNEL_GAMG_VCYCLE = 8,/* one application of the GAMG preconditioner (all levels) */
    S4F_KERNEL_GAMG_STEP0 = 9, /* the fine-level Chebyshev-Jacobi smoothing step of the V-cycle, alone */
    S4F_KERNEL_HALO3 = 10,     /* one processor-patch halo exchange of a 3-component field (decomposed runs; peer memory) */
    S4F_KERNEL_DOT_REDUCE = 11 /* a reducing kernel (z.r of the PCG) with the all-reduce fused into its last block */
};
int s4fgpu_time_kernel(s4fgpu_handle h, int kernel, int reps, int flushL2,
                       double* msPerLaunch, double* algoBytesPerLaunch);

/* The GAMG hierarchy of the current matrix (built on first use): number of levels, cells per level
 * (sizes[maxLevels]), algorithmic bytes of one V-cycle and the host set-up time. */
int s4fgpu_gamg_info(s4fgpu_handle h, int* nLevels, int* sizes, int maxLevels, double* bytesPerApply,
                     double* setupSeconds);
/* decomposed runs: how many levels (from the finest) are distributed, one part per rank with halo exchanges
 * ([OF-ext] GAMG processor interfaces); the levels below are gathered and replicated.  sizes[] of
 * s4fgpu_gamg_info holds this rank's part on the distributed levels. */
int s4fgpu_gamg_distributed_levels(s4fgpu_handle h);

/* CUDA-event timer on the library's own stream (torch.cuda.Event only sees torch's stream), and a
 * stream synchronise.  timer_stop returns the device time in milliseconds since timer_start. */
int s4fgpu_timer_start(s4fgpu_handle h);
int s4fgpu_timer_stop(s4fgpu_handle h, double* ms);
int s4fgpu_synchronize(s4fgpu_handle h);

/* number of kernel launches issued by this handle since creation (bench.py gpu_launches) */
long long s4fgpu_launch_count(s4fgpu_handle h);

#ifdef __cplusplus
}
#endif
#endif /* S4FGPU_H */
