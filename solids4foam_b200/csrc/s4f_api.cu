// s4f_api.cu -- the extern "C" boundary declared in include/s4fgpu.h, and the host side of the
// momentum-correction loop (linGeomTotalDispSolid::evolve, linGeomTotalDispSolid.C:111-232; the
// convergence logic of solidModel::converged, solidModelTemplates.C:27-188).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "s4f_comm.h"

int s4f_update_total_fields_impl(s4fgpu_ctx* c);

static std::string g_createError;

#define S4F_REQUIRE(c, cond, msg) \
    do { if (!(cond)) { (c)->err = (msg); return 1; } } while (0)

static int d2d(s4fgpu_ctx* c, double* dst, const double* src, size_t n) {
    if (!dst || !src || n == 0) return 0;
    S4F_CHECK_CUDA(c, cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

int s4f_setup_law(s4fgpu_ctx* c) {
    const s4fgpu_law& L = c->law;
    // impK: linearElastic.C:204-245 (2 mu + lambda); neoHookeanElastic.C:101-119 and the plastic laws at
    // DLambda = 0 (construction): 4/3 mu + K.   Evaluated once, as the solidModel constructors do.
    const double impK = (L.kind == S4F_LAW_LINEAR_ELASTIC) ? 2.0 * L.mu + L.lambda : (4.0 / 3.0) * L.mu + L.K;
    std::vector<double> h(c->ld, impK);
    S4F_CHECK_CUDA(c, c->impK.upload(h));
    c->impK0 = impK; c->mValid = false;
    c->Hp = 0;
    if (L.nTable == 2) c->Hp = (L.tableSigY[1] - L.tableSigY[0]) / (L.tableEps[1] - L.tableEps[0]);
    int rc = s4f_alloc_model_fields(c);
    if (rc) return rc;
    if (L.nTable >= 1 && c->sigmaY.p) {
        std::vector<double> sy(c->ld, L.tableSigY[0]);     // stressPlasticStrainSeries_(0.0), clamp
        double x0 = L.tableSigY[0];
        if (L.nTable > 1 && L.tableEps[0] < 0.0) {          // general lookup at 0
            for (int i = 0; i + 1 < L.nTable; i++)
                if (0.0 >= L.tableEps[i] && 0.0 < L.tableEps[i + 1])
                    x0 = L.tableSigY[i] + (L.tableSigY[i + 1] - L.tableSigY[i]) * (0.0 - L.tableEps[i]) / (L.tableEps[i + 1] - L.tableEps[i]);
            std::fill(sy.begin(), sy.end(), x0);
        }
        S4F_CHECK_CUDA(c, c->sigmaY.upload(sy)); S4F_CHECK_CUDA(c, c->sigmaYOld.upload(sy));
    }
    c->matrixValid = false;
    return 0;
}


int s4f_outer_iteration(s4fgpu_ctx* c, int iCorr) {
    int rc;
    if ((rc = d2d(c, c->Dprev.p, c->D.p, 3 * (size_t)c->ld))) return rc;        // D().storePrevIter()
    if ((rc = s4f_bc_update_coeffs(c))) return rc;                               // fvMatrix ctor -> updateCoeffs()
    if (!c->matrixValid && (rc = s4f_assemble_matrix(c))) return rc;             // fvm::laplacian(impKf, D): constant while impKf is
    if ((rc = s4f_assemble_source(c))) return rc;                                // explicit terms
    if ((rc = s4f_solve_segregated(c, c->D.p, c->source.p, true))) return rc;    // DEqn.solve(); statistics read with the residuals
    if ((rc = s4f_bc_evaluate(c))) return rc;                                    // D.correctBoundaryConditions()
    if ((rc = s4f_relax_and_residual(c, iCorr))) return rc;                      // relaxField + residual reductions
    if ((rc = s4f_update_totals(c, true, false))) return rc;                     // incremental: D = D.oldTime() + DD
    if ((rc = s4f_grad(c))) return rc;                                           // mechanical().grad(D, gradD)
    if ((rc = s4f_update_totals(c, false, true))) return rc;                     // incremental: gradD = gradD.oldTime() + gradDD
    if ((rc = s4f_law_correct(c))) return rc;                                    // mechanical().correct(sigma)
    return 0;
}

// converged(): solidModelTemplates.C:27-188, evaluated from the device-side reductions
int s4f_read_outer_scalars(s4fgpu_ctx* c, s4fgpu_stats* st, bool* converged, int iCorr) {
    S4F_CHECK_CUDA(c, cudaMemcpyAsync(c->hOutS, c->outS.p, sizeof(OuterScalars), cudaMemcpyDeviceToHost, c->stream));
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));      // the one host synchronisation of an outer iteration
    { int rf = s4f_finish_solve(c); if (rf) return rf; }
    const OuterScalars& o = *c->hOutS;
    if (c->unsTL()) {
        // unsNonLinGeomTotalLagSolid.C:49-76, :333-378: res = max|D - D.prevIter| / max(max|D - D.oldTime|, SMALL), tolerance
        // max(maxRes * relativeTol_, solutionTol) with relativeTol_ = solutionTolerance (:206-213); never on the first iteration
        const double res = o.maxDelta / std::max(o.maxIncr, 1e-15);
        if (iCorr == 0) c->unsMaxRes = 0;
        c->unsMaxRes = std::max(c->unsMaxRes, res);
        const double tol = std::max(c->unsMaxRes * c->ctl.solutionTolerance, c->ctl.solutionTolerance);
        const bool conv = iCorr > 0 && !(res > tol);
        const double* ir = c->last.initialResidual;
        c->last.solverPerfInitRes = std::sqrt(ir[0] * ir[0] + ir[1] * ir[1] + ir[2] * ir[2]); c->last.relResidual = res;
        c->last.materialResidual = 0; c->last.converged = conv ? 1 : 0; c->last.totalInnerIterations = c->totalInner;
        if (st) *st = c->last;
        *converged = conv;
        return 0;
    }
    const bool incremental = c->incremental();
    double denom = incremental ? o.maxMag : o.maxIncr;
    if (denom < 1e-15) denom = std::max(o.maxMag, 1e-15);
    const double residualvf = o.maxDelta / denom;
    double matRes = 0.0;
    if (c->law.kind == S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC || c->law.kind == S4F_LAW_LINEAR_ELASTIC_MISES_PLASTIC) matRes = o.matNum / o.matDen;
    const double* ir = c->last.initialResidual;
    const double spir = std::sqrt(ir[0] * ir[0] + ir[1] * ir[1] + ir[2] * ir[2]);
    bool conv = false;
    if (iCorr > 1 && matRes < c->ctl.materialTolerance) {
        if (spir < c->ctl.solutionTolerance && residualvf < c->ctl.solutionTolerance) conv = true;
        else if (residualvf < c->ctl.alternativeTolerance) conv = true;
        else if (spir < c->ctl.alternativeTolerance) conv = true;
    }
    c->last.solverPerfInitRes = spir; c->last.relResidual = residualvf; c->last.materialResidual = matRes;
    c->last.converged = conv ? 1 : 0; c->last.totalInnerIterations = c->totalInner;
    if (st) *st = c->last;
    *converged = conv;
    return 0;
}

extern "C" {

int s4fgpu_version(void) { return 100; }

const char* s4fgpu_last_error(s4fgpu_handle h) { return h ? h->err.c_str() : g_createError.c_str(); }

int s4fgpu_create(s4fgpu_handle* out, int device) {
    if (!out) return 1;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_createError = std::string("s4fgpu_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        return 2;
    }
    if (device < 0 || device >= n) { g_createError = "s4fgpu_create: bad device index"; return 1; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_createError = cudaGetErrorString(e); return 2; }
    s4fgpu_ctx* c = new s4fgpu_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_createError = cudaGetErrorString(e); delete c; return 2; }
    c->numSMs = prop.multiProcessorCount;
    // A BLOCKING stream: DevBuf::alloc / upload and s4fgpu_set_bc use cudaMemset / cudaMemcpy on the legacy default stream,
    // which is ordered against blocking streams only.  With cudaStreamNonBlocking a fresh buffer's zero fill could still be
    // running when the first kernel on c->stream wrote into it (latent at test sizes, real at 8 M cells).
    if ((e = cudaStreamCreate(&c->stream)) != cudaSuccess) { g_createError = cudaGetErrorString(e); delete c; return 2; }
    *out = c;
    return 0;
}

int s4fgpu_destroy(s4fgpu_handle c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    s4f_amg_destroy(c);
    s4f_uns_destroy(c);
    s4f_dic_destroy(c);
    s4f_solve_graphs_destroy(c);
    s4f_halo_plan_destroy(c->halo0); c->halo0 = nullptr;
    s4f_halo_plan_destroy(c->haloX); c->haloX = nullptr;
    s4f_comm_destroy(c);
    if (c->comm) ncclCommDestroy(c->comm);
    if (c->hPcgS) cudaFreeHost(c->hPcgS);
    if (c->hOutS) cudaFreeHost(c->hOutS);
    if (c->ev0) { cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); }
    cudaStream_t s = c->stream;
    delete c;
    cudaStreamDestroy(s);
    return 0;
}

int s4fgpu_get_unique_id(char id[128]) {
    ncclUniqueId u;
    if (ncclGetUniqueId(&u) != ncclSuccess) return 3;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    std::memcpy(id, &u, 128);
    return 0;
}

int s4fgpu_comm_init(s4fgpu_handle c, int nRanks, int rank, const char id[128]) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    ncclUniqueId u; std::memcpy(&u, id, 128);
    S4F_CHECK_NCCL(c, ncclCommInitRank(&c->comm, nRanks, u, rank));
    c->nRanks = nRanks; c->rank = rank; c->nGlobalCells = -1;
    return s4f_comm_setup(c);           // all-reduce mailboxes over peer memory (collective)
}

int s4fgpu_set_mesh(s4fgpu_handle c, int nCells, int nInternalFaces, const int* owner, const int* neighbour, int nPatches,
                    const int* patchStart, const int* patchSize, const int* patchKind, const int* patchNbrRank, const int* faceCells,
                    const int* solutionD) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, nCells > 0 && nInternalFaces >= 0 && nPatches >= 0, "set_mesh: bad sizes");
    c->N = nCells; c->F = nInternalFaces; c->nPatches = nPatches;
    c->own.assign(owner, owner + c->F); c->nei.assign(neighbour, neighbour + c->F);
    c->pStart.assign(patchStart, patchStart + nPatches); c->pSize.assign(patchSize, patchSize + nPatches);
    c->pKind.assign(patchKind, patchKind + nPatches); c->pNbr.assign(patchNbrRank, patchNbrRank + nPatches);
    int B = 0;
    for (int p = 0; p < nPatches; p++) B = std::max(B, patchStart[p] + patchSize[p]);
    c->B = B;
    c->faceCells.assign(faceCells, faceCells + B);
    for (int i = 0; i < 3; i++) c->solD[i] = solutionD[i] ? 1 : 0;
    for (int f = 0; f < c->F; f++)
        S4F_REQUIRE(c, owner[f] >= 0 && owner[f] < neighbour[f] && neighbour[f] < nCells, "set_mesh: lduAddressing must be upper-triangular (owner < neighbour)");
    for (int b = 0; b < B; b++) S4F_REQUIRE(c, faceCells[b] >= 0 && faceCells[b] < nCells, "set_mesh: faceCells out of range");
    c->bcKind.assign(nPatches, S4F_BC_SOLID_TRACTION);
    s4f_amg_destroy(c); c->amgRefresh = false;          // another graph: the aggregates go with it
    c->nPoints = 0; c->X = 0; c->extPtr.clear();          // another mesh: its points arrive with the next set_points
    c->meshSet = true; c->geomSet = false; c->matrixValid = false; c->nGlobalCells = -1;
    return 0;
}

int s4fgpu_set_geometry(s4fgpu_handle c, const double* C, const double* V, const double* Sf, const double* magSf, const double* Cf,
                        const double* weights, const double* nod, const double* corr, const double* CnbrB) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, c->meshSet, "set_geometry: call set_mesh first");
    const size_t N = c->N, FB = (size_t)c->F + c->B;
    c->hC.assign(C, C + 3 * N); c->hV.assign(V, V + N); c->hSf.assign(Sf, Sf + 3 * FB); c->hMagSf.assign(magSf, magSf + FB);
    c->hCf.assign(Cf, Cf + 3 * FB); c->hW.assign(weights, weights + FB); c->hNod.assign(nod, nod + FB); c->hCorr.assign(corr, corr + 3 * FB);
    c->hCnbrB.assign(CnbrB, CnbrB + 3 * (size_t)c->B);
    // a second call on the same mesh is a geometry refresh after mesh motion (nonLinGeomUpdatedLagSolid.C:360-374 ->
    // solidModel::moveMesh): fields, boundary data and the law history stay, everything derived from geometry is rebuilt
    const bool again = c->geomSet;
    int rc = 0;
    // decomposed mesh with points: the values of other ranks' cells and boundary faces at shared points get slots of their
    // own in the field index space, so they must be known before the rows fix the leading dimension
    if (c->nRanks > 1 && c->nPoints > 0 && (rc = s4f_build_point_ghosts(c))) return rc;
    rc = s4f_build_rows(c); if (rc) return rc;
    rc = s4f_alloc_fields(c); if (rc) return rc;
    if (c->nRanks > 1 && c->nPoints > 0 && (rc = s4f_build_point_weights(c, c->hPoints.data()))) return rc;
    // host geometry copies are only needed to build the rows (cell centres stay for the vol->point weights)
    std::vector<double>().swap(c->hSf); std::vector<double>().swap(c->hCf); std::vector<double>().swap(c->hCorr);
    std::vector<double>().swap(c->hW); std::vector<double>().swap(c->hNod);
    c->hostGeomStale = false;
    c->geomSet = true; c->matrixValid = false; c->amgValid = false; c->histValid = false; c->mValid = false; c->gValid = false; c->unsValid = false;
    if (c->lawSet && !again) { rc = s4f_setup_law(c); if (rc) return rc; }
    return 0;
}

int s4fgpu_set_law(s4fgpu_handle c, const s4fgpu_law* law) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, law && law->kind >= 0 && law->kind <= 3, "set_law: unknown law");
    S4F_REQUIRE(c, law->nTable >= 0 && law->nTable <= 64, "set_law: table too long");
    if (law->kind == S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC || law->kind == S4F_LAW_LINEAR_ELASTIC_MISES_PLASTIC)
        S4F_REQUIRE(c, law->nTable >= 1, "set_law: plasticity law needs the (epsilonP sigmaY) table");
    c->law = *law; c->lawSet = true; c->histValid = false; c->unsValid = false;
    if (c->geomSet) return s4f_setup_law(c);
    return 0;
}

int s4fgpu_set_controls(s4fgpu_handle c, const s4fgpu_controls* ctl) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, ctl, "set_controls: null");
    S4F_REQUIRE(c, ctl->solidModel >= S4F_MODEL_LIN_GEOM_TOTAL_DISP && ctl->solidModel <= S4F_MODEL_UNS_NONLIN_UL, "set_controls: unknown solidModel");
    S4F_REQUIRE(c, ctl->solidModel != S4F_MODEL_NONLIN_TL || ctl->d2dt2Scheme == S4F_D2DT2_STEADY_STATE,
                "set_controls: nonLinearGeometryTotalLagrangian is available with the steadyState d2dt2 scheme");
    S4F_REQUIRE(c, ctl->solver == S4F_SOLVER_PCG || ctl->solver == S4F_SOLVER_PBICGSTAB, "set_controls: solver PCG or PBiCGStab");
    S4F_REQUIRE(c, ctl->d2dt2Scheme >= S4F_D2DT2_STEADY_STATE && ctl->d2dt2Scheme <= S4F_D2DT2_BACKWARD, "set_controls: unknown d2dt2 scheme");
    S4F_REQUIRE(c, ctl->gradScheme >= S4F_GRAD_LEAST_SQUARES && ctl->gradScheme <= S4F_GRAD_POINT_CELLS_LEAST_SQUARES, "set_controls: unknown gradScheme");
    if (ctl->d2dt2Scheme == S4F_D2DT2_BACKWARD && ctl->deltaT0 > 0)     // backwardD2dt2Scheme.C:316-322
        S4F_REQUIRE(c, std::fabs(ctl->deltaT - ctl->deltaT0) <= 1e-15 + 1e-12 * ctl->deltaT, "set_controls: backwardD2dt2Scheme not implemented for variable time steps");
    c->unsValid = false;
    if (c->ctlSet && (ctl->gamgSinglePrecision != c->ctl.gamgSinglePrecision || ctl->gamgSmootherDegree != c->ctl.gamgSmootherDegree ||
                      ctl->gamgCycle != c->ctl.gamgCycle || ctl->gamgOverCorrection != c->ctl.gamgOverCorrection ||
                      ctl->gamgSmootherRatio != c->ctl.gamgSmootherRatio)) { s4f_amg_destroy(c); c->amgRefresh = false; }
    c->ctl = *ctl; c->ctlSet = true; c->matrixValid = false; c->amgValid = false; c->histValid = false; c->mValid = false;
    if (c->geomSet) return s4f_alloc_model_fields(c);
    return 0;
}

int s4fgpu_set_bc(s4fgpu_handle c, int patch, int kind, const double* value, const double* pressure) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, c->geomSet, "set_bc: call set_geometry first");
    S4F_REQUIRE(c, patch >= 0 && patch < c->nPatches, "set_bc: bad patch");
    S4F_REQUIRE(c, kind >= 0 && kind <= 3, "set_bc: unknown boundary condition");
    if (c->bcKind[patch] != kind) c->matrixValid = false;
    c->bcKind[patch] = kind;
    const int s = c->pStart[patch], n = c->pSize[patch], B = c->B;
    if (n == 0) return s4f_upload_bc(c);
    std::vector<double> v(n, 0.0);
    for (int q = 0; q < 3; q++) {
        for (int i = 0; i < n; i++) v[i] = value ? value[3 * i + q] : 0.0;
        S4F_CHECK_CUDA(c, cudaMemcpy(c->bcValue.p + (size_t)q * B + s, v.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    }
    for (int i = 0; i < n; i++) v[i] = pressure ? pressure[i] : 0.0;
    S4F_CHECK_CUDA(c, cudaMemcpy(c->bcPressure.p + s, v.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    return s4f_upload_bc(c);
}

static int field_lookup(s4fgpu_ctx* c, int field, double** p, int* ncomp, int* offset, int* count) {
    *offset = 0; *count = c->N;
    switch (field) {
        case S4F_FIELD_D: *p = c->incremental() ? c->Dtot.p : c->D.p; *ncomp = 3; break;
        case S4F_FIELD_DD: *p = c->incremental() ? c->D.p : nullptr; *ncomp = 3; break;
        case S4F_FIELD_GRAD_DD: *p = c->incremental() ? c->gradD.p : nullptr; *ncomp = 9; break;
        case S4F_FIELD_D_OLD: *p = c->Dold.p; *ncomp = 3; break;
        case S4F_FIELD_D_OLDOLD: *p = c->DoldOld.p; *ncomp = 3; break;
        case S4F_FIELD_GRAD_D: *p = c->incremental() ? c->gradDtot.p : c->gradD.p; *ncomp = 9; break;
        case S4F_FIELD_GRAD_D_OLD: *p = c->gradDold.p; *ncomp = 9; break;
        case S4F_FIELD_SIGMA: *p = c->sigma.p; *ncomp = 6; break;
        case S4F_FIELD_D_B: *p = c->incremental() ? c->Dtot.p : c->D.p; *ncomp = 3; *offset = c->bOff(); *count = c->B; break;
        case S4F_FIELD_GRAD_D_B: *p = c->incremental() ? c->gradDtot.p : c->gradD.p; *ncomp = 9; *offset = c->bOff(); *count = c->B; break;
        case S4F_FIELD_SIGMA_B: *p = c->sigma.p; *ncomp = 6; *offset = c->bOff(); *count = c->B; break;
        case S4F_FIELD_SOURCE: *p = c->source.p; *ncomp = 3; break;
        case S4F_FIELD_DIAG: *p = c->diagC.p; *ncomp = 3; break;
        case S4F_FIELD_EPSILON_P_EQ: *p = c->epsPEq.p; *ncomp = 1; break;
        case S4F_FIELD_SIGMA_Y: *p = c->sigmaY.p; *ncomp = 1; break;
        case S4F_FIELD_BEBAR: *p = c->bEbar.p; *ncomp = 6; break;
        case S4F_FIELD_DLAMBDA: *p = c->DLambda.p; *ncomp = 1; break;
        case S4F_FIELD_J: *p = c->lawJ.p; *ncomp = 1; break;
        case S4F_FIELD_F: *p = c->lawF.p; *ncomp = 9; break;
        case S4F_FIELD_DEPSILON_P: *p = c->DEpsP.p; *ncomp = 6; break;
        case S4F_FIELD_EPSILON_P: *p = c->epsP.p; *ncomp = 6; break;
        case S4F_FIELD_RHO: *p = c->rhoF.p; *ncomp = 1; break;
        case S4F_FIELD_SIGMA_HYD: *p = c->sigmaHyd.p; *ncomp = 1; break;
        case S4F_FIELD_GRAD_SIGMA_HYD: *p = c->gradP.p; *ncomp = 3; break;
        case S4F_FIELD_DD_B: *p = c->incremental() ? c->D.p : nullptr; *ncomp = 3; *offset = c->bOff(); *count = c->B; break;
        default: c->err = "unknown / unsupported field id"; return 1;
    }
    if (!*p) { c->err = "field not allocated for the selected model / law"; return 1; }
    return 0;
}

int s4fgpu_upload(s4fgpu_handle c, int field, const double* host) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, c->geomSet, "upload: call set_geometry first");
    double* p; int nc, off, cnt;
    int rc = field_lookup(c, field, &p, &nc, &off, &cnt); if (rc) return rc;
    if (field == S4F_FIELD_D_OLD || field == S4F_FIELD_D_OLDOLD) c->histValid = false;
    c->mValid = false;
    return s4f_aos_to_soa(c, host, p, cnt, nc, off);
}

int s4fgpu_download(s4fgpu_handle c, int field, double* host) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, c->geomSet, "download: call set_geometry first");
    if (field == S4F_FIELD_UPPER) {
        if (!c->matrixValid) { int rc = s4f_assemble_matrix(c); if (rc) return rc; }
        return s4f_download_upper(c, host);
    }
    if (field == S4F_FIELD_SIGMA_F || field == S4F_FIELD_GRAD_D_F) return s4f_uns_download(c, field, host);
    if (field == S4F_FIELD_TRACTION_GRADIENT_B) {
        std::vector<double> t(3 * (size_t)c->B);
        S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
        S4F_CHECK_CUDA(c, cudaMemcpy(t.data(), c->tracGrad.p, t.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int b = 0; b < c->B; b++) for (int q = 0; q < 3; q++) host[3 * b + q] = t[(size_t)q * c->B + b];
        return 0;
    }
    double* p; int nc, off, cnt;
    int rc = field_lookup(c, field, &p, &nc, &off, &cnt); if (rc) return rc;
    return s4f_soa_to_aos(c, p, host, cnt, nc, off);
}

int s4fgpu_initialise(s4fgpu_handle c) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, c->geomSet && c->lawSet && c->ctlSet, "initialise: mesh, geometry, law and controls must be set");
    S4F_REQUIRE(c, c->nRanks == 1 || c->comm, "initialise: parallel run without communicator");
    S4F_REQUIRE(c, !c->UL() || c->law.kind == S4F_LAW_NEO_HOOKEAN_ELASTIC || c->law.kind == S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC,
                "initialise: nonLinearGeometryUpdatedLagrangian needs a finite-strain law (neoHookeanElastic, neoHookeanElasticMisesPlastic)");
    int rc;
    if ((rc = s4f_alloc_model_fields(c))) return rc;
    if ((rc = s4f_upload_bc(c))) return rc;
    if ((rc = s4f_bc_update_coeffs(c))) return rc;
    if ((rc = s4f_bc_evaluate(c))) return rc;
    if ((rc = d2d(c, c->Dprev.p, c->D.p, 3 * (size_t)c->ld))) return rc;
    if ((rc = s4f_halo_exchange(c, c->D.p, 3))) return rc;
    c->mValid = false;
    c->unsGradientsOnly = true;
    rc = s4f_grad(c);
    c->unsGradientsOnly = false;
    if (rc) return rc;
    if ((rc = s4f_update_totals(c, false, true))) return rc;
    if ((rc = s4f_kinematics(c))) return rc;                     // F, Finv, J of the finite-strain models (ctor, restart branch)
    if ((rc = s4f_assemble_matrix(c))) return rc;
    c->iCorr = 0;
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int s4fgpu_new_timestep(s4fgpu_handle c, double deltaT) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    const size_t ld = c->ld;
    if (c->ctl.d2dt2Scheme == S4F_D2DT2_BACKWARD && c->timeIndex >= 1)
        S4F_REQUIRE(c, std::fabs(deltaT - c->ctl.deltaT) <= 1e-15 + 1e-12 * deltaT, "new_timestep: backwardD2dt2Scheme not implemented for variable time steps");
    c->ctl.deltaT0 = c->ctl.deltaT; c->ctl.deltaT = deltaT;
    int rc = 0;
    if (c->UL()) {     // GeometricField::storeOldTimes over the chains created in the constructor (nonLinGeomUpdatedLagSolid.C:143-145)
        if (c->timeIndex == 0) {
            for (DevBuf<double>* b : {&c->Dooo, &c->Doooo, &c->Dooooo}) rc |= d2d(c, b->p, c->DoldOld.p, 3 * ld);
            for (DevBuf<double>* b : {&c->DDo, &c->DDoo, &c->DDooo, &c->DDoooo}) rc |= d2d(c, b->p, c->D.p, 3 * ld);
            rc |= d2d(c, c->rhoO.p, c->rhoF.p, ld); rc |= d2d(c, c->rhoOO.p, c->rhoF.p, ld);
        }
        rc |= d2d(c, c->Dooooo.p, c->Doooo.p, 3 * ld); rc |= d2d(c, c->Doooo.p, c->Dooo.p, 3 * ld); rc |= d2d(c, c->Dooo.p, c->DoldOld.p, 3 * ld);
        rc |= d2d(c, c->DDoooo.p, c->DDooo.p, 3 * ld); rc |= d2d(c, c->DDooo.p, c->DDoo.p, 3 * ld); rc |= d2d(c, c->DDoo.p, c->DDo.p, 3 * ld);
        rc |= d2d(c, c->DDo.p, c->D.p, 3 * ld);
        rc |= d2d(c, c->rhoOO.p, c->rhoO.p, ld); rc |= d2d(c, c->rhoO.p, c->rhoF.p, ld);
        c->matrixValid = false;
    } else if (c->ctl.d2dt2Scheme == S4F_D2DT2_BACKWARD) {       // GeometricField::storeOldTimes over the four-level chain
        rc |= d2d(c, c->Doooo.p, c->timeIndex >= 2 ? c->Dooo.p : c->DoldOld.p, 3 * ld);
        rc |= d2d(c, c->Dooo.p, c->DoldOld.p, 3 * ld);
    }
    c->timeIndex++; c->histValid = false;
    rc |= d2d(c, c->DoldOld.p, c->Dold.p, 3 * ld);
    if (c->incremental()) {     // the total fields roll; DD keeps its value as the initial guess of the next step
        rc |= d2d(c, c->Dold.p, c->Dtot.p, 3 * ld); rc |= d2d(c, c->gradDold.p, c->gradDtot.p, 9 * ld);
    } else {
        rc |= d2d(c, c->Dold.p, c->D.p, 3 * ld); rc |= d2d(c, c->gradDold.p, c->gradD.p, 9 * ld);
    }
    rc |= d2d(c, c->sigmaOld.p, c->sigma.p, 6 * ld);
    if (c->lawF.p) { rc |= d2d(c, c->lawFold.p, c->lawF.p, 9 * ld); rc |= d2d(c, c->lawJold.p, c->lawJ.p, ld); }
    rc |= s4f_uns_new_timestep(c);
    if (c->bEbar.p) {
        rc |= d2d(c, c->bEbarOld.p, c->bEbar.p, 6 * ld); rc |= d2d(c, c->epsPOld.p, c->epsP.p, 6 * ld);
        rc |= d2d(c, c->epsPEqOld.p, c->epsPEq.p, ld); rc |= d2d(c, c->sigmaYOld.p, c->sigmaY.p, ld);
    }
    if (c->ctl.d2dt2Scheme != S4F_D2DT2_STEADY_STATE) c->matrixValid = false;
    c->iCorr = 0;
    return rc;
}

int s4fgpu_outer_iteration(s4fgpu_handle c, s4fgpu_stats* st) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    int rc = s4f_outer_iteration(c, c->iCorr); if (rc) return rc;
    bool conv;
    rc = s4f_read_outer_scalars(c, st, &conv, c->iCorr); if (rc) return rc;
    c->iCorr++;
    c->last.nCorr = c->iCorr;
    if (st) st->nCorr = c->iCorr;
    return 0;
}

int s4fgpu_evolve(s4fgpu_handle c, s4fgpu_stats* st) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    int iCorr = 0; bool conv = false;
    do {                                                       // linGeomTotalDispSolid.C:135-209
        int rc = s4f_outer_iteration(c, iCorr); if (rc) return rc;
        rc = s4f_read_outer_scalars(c, nullptr, &conv, iCorr); if (rc) return rc;
    } while (!conv && ++iCorr < c->ctl.nCorrectors);
    c->last.nCorr = conv ? iCorr + 1 : iCorr;
    c->iCorr = 0;
    if (st) *st = c->last;
    return 0;
}

int s4fgpu_set_points(s4fgpu_handle c, int nPoints, const double* points, const int* faceVertsPtr, const int* faceVerts) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, c->meshSet, "set_points: call set_mesh first");
    S4F_REQUIRE(c, nPoints > 0 && points && faceVertsPtr && faceVerts, "set_points: bad arguments");
    // single rank: after set_geometry (the weights are built here).  Decomposed: BEFORE set_geometry, which then lays out
    // the point-neighbour ghosts with the rows and builds the weights (s4f_build_point_ghosts)
    S4F_REQUIRE(c, c->nRanks > 1 ? (!c->geomSet || nPoints == c->nPoints) : c->geomSet,
                c->nRanks > 1 ? "set_points: on a decomposed mesh call set_points before set_geometry" : "set_points: call set_geometry first");
    const int nF = c->F + c->B;
    c->nPoints = nPoints;
    c->hFvPtr.assign(faceVertsPtr, faceVertsPtr + nF + 1);
    c->hFv.assign(faceVerts, faceVerts + faceVertsPtr[nF]);
    c->hPoints.assign(points, points + 3 * (size_t)nPoints);
    c->gValid = false; c->unsValid = false;
    if (c->nRanks > 1) return 0;
    return s4f_build_point_weights(c, points);
}

int s4fgpu_interpolate_to_points(s4fgpu_handle c, int field, int mode, double* pointField) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, c->nPoints > 0, "interpolate_to_points: call set_points first");
    S4F_REQUIRE(c, mode == S4F_POINT_INTERP_PATCH || mode == S4F_POINT_INTERP_GRAD, "interpolate_to_points: unknown mode");
    const double* X = nullptr; const double* G = nullptr;
    if (field == S4F_FIELD_D) { X = c->incremental() ? c->Dtot.p : c->D.p; G = c->incremental() ? c->gradDtot.p : c->gradD.p; }
    else if (field == S4F_FIELD_DD && c->incremental()) { X = c->D.p; G = c->gradD.p; }
    S4F_REQUIRE(c, X, "interpolate_to_points: field must be D or DD");
    return s4f_interpolate_to_points(c, X, mode == S4F_POINT_INTERP_GRAD ? G : nullptr, pointField);
}

int s4fgpu_move_points(s4fgpu_handle c, const double* pointDD) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, c->geomSet, "move_points: call set_geometry first");
    return s4f_move_points_device(c, pointDD);
}

int s4fgpu_update_total_fields(s4fgpu_handle c) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    return s4f_update_total_fields_impl(c);
}

int s4fgpu_op_grad(s4fgpu_handle c) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    int rc = s4f_halo_exchange(c, c->D.p, 3); if (rc) return rc;
    c->mValid = false;
    rc = s4f_grad(c); if (rc) return rc;
    rc = s4f_update_totals(c, false, true); if (rc) return rc;
    rc = s4f_kinematics(c); if (rc) return rc;
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int s4fgpu_op_correct(s4fgpu_handle c) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    int rc = s4f_law_correct(c); if (rc) return rc;
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int s4fgpu_op_assemble(s4fgpu_handle c) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = s4f_bc_update_coeffs(c))) return rc;
    if ((rc = s4f_assemble_matrix(c))) return rc;
    if ((rc = s4f_assemble_source(c))) return rc;
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int s4fgpu_op_amul(s4fgpu_handle c, int cmpt, const double* x, double* y) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, cmpt >= 0 && cmpt < 3, "op_amul: bad component");
    if (!c->matrixValid) { int rc = s4f_assemble_matrix(c); if (rc) return rc; }
    const size_t ld = c->ld;
    S4F_CHECK_CUDA(c, cudaMemsetAsync(c->pA.p, 0, 3 * ld * sizeof(double), c->stream));
    S4F_CHECK_CUDA(c, cudaMemcpyAsync(c->pA.p + cmpt * ld, x, c->N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    int rc = s4f_amul_device(c, c->pA.p, c->wA.p, 1 << cmpt); if (rc) return rc;
    S4F_CHECK_CUDA(c, cudaMemcpyAsync(y, c->wA.p + cmpt * ld, c->N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    S4F_CHECK_CUDA(c, cudaMemsetAsync(c->pA.p, 0, 3 * ld * sizeof(double), c->stream));
    return 0;
}

int s4fgpu_op_solve(s4fgpu_handle c, double* psi, const double* source, s4fgpu_stats* st) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    if (!c->matrixValid) { int rc = s4f_assemble_matrix(c); if (rc) return rc; }
    DevBuf<double> x, b;
    S4F_CHECK_CUDA(c, x.alloc(3 * (size_t)c->ld)); S4F_CHECK_CUDA(c, b.alloc(3 * (size_t)c->ld));
    int rc;
    if ((rc = s4f_aos_to_soa(c, psi, x.p, c->N, 3, 0))) return rc;
    if ((rc = s4f_aos_to_soa(c, source, b.p, c->N, 3, 0))) return rc;
    if ((rc = s4f_solve_segregated(c, x.p, b.p))) return rc;
    if ((rc = s4f_soa_to_aos(c, x.p, psi, c->N, 3, 0))) return rc;
    if (st) *st = c->last;
    return 0;
}

int s4fgpu_time_kernel(s4fgpu_handle c, int kernel, int reps, int flushL2, double* msPerLaunch, double* algoBytesPerLaunch) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, c->geomSet && c->matrixValid, "time_kernel: initialise first");
    S4F_REQUIRE(c, reps > 0, "time_kernel: reps");
    if (kernel == S4F_KERNEL_SPMV1 || kernel == S4F_KERNEL_SPMV3 || kernel == S4F_KERNEL_PCG_ITER || kernel == S4F_KERNEL_PCG_P ||
        kernel == S4F_KERNEL_PCG_XR || kernel == S4F_KERNEL_HALO3 || kernel == S4F_KERNEL_DOT_REDUCE)
        return s4f_time_pcg_kernels(c, kernel, reps, flushL2, msPerLaunch, algoBytesPerLaunch);
    return s4f_time_fv_kernels(c, kernel, reps, flushL2, msPerLaunch, algoBytesPerLaunch);
}

long long s4fgpu_launch_count(s4fgpu_handle c) { return c ? c->launches : 0; }

int s4fgpu_gamg_info(s4fgpu_handle c, int* nLevels, int* sizes, int maxLevels, double* bytesPerApply, double* setupSeconds) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, c->geomSet && c->matrixValid, "gamg_info: initialise first");
    if (!c->amgValid) { int rc = c->amgRefresh ? s4f_amg_refresh(c) : s4f_amg_setup(c); if (rc) return rc; c->amgValid = true; c->amgRefresh = false; }
    return s4f_amg_info(c, nLevels, sizes, maxLevels, bytesPerApply, setupSeconds);
}

int s4fgpu_gamg_distributed_levels(s4fgpu_handle c) { return c ? s4f_amg_distributed_levels(c) : 0; }

int s4fgpu_timer_start(s4fgpu_handle c) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    if (!c->ev0) { S4F_CHECK_CUDA(c, cudaEventCreate(&c->ev0)); S4F_CHECK_CUDA(c, cudaEventCreate(&c->ev1)); }
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    S4F_CHECK_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    return 0;
}
int s4fgpu_timer_stop(s4fgpu_handle c, double* ms) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_REQUIRE(c, c->ev0, "timer_stop without timer_start");
    S4F_CHECK_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    S4F_CHECK_CUDA(c, cudaEventSynchronize(c->ev1));
    float f; S4F_CHECK_CUDA(c, cudaEventElapsedTime(&f, c->ev0, c->ev1));
    *ms = f;
    return 0;
}
int s4fgpu_synchronize(s4fgpu_handle c) {
    S4F_CHECK_CUDA(c, cudaSetDevice(c->device));
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
