// s4f_ctx.h -- host-side context of libs4fgpu.so: device-resident mirror of one fvMesh + fields.
//
// Index space of every "vol field" on the device (SoA, one array per component, leading dimension ld):
//   [0, N)            cells of this rank
//   [N, N+G)          ghost cells = neighbour cells across processor-patch faces (filled by halo exchange)
//   [N+G, N+G+B)      boundary-face values (one slot per boundary face; processor faces unused)
// so that OpenFOAM's "field algebra acts on internal field and boundary field alike" becomes one
// kernel over one index range, and boundary faces can sit in the cell-centric rows as ordinary
// entries whose column is the boundary-value slot.
//
// Rows (cell-centric, atomic-free): SELL-32 ("sliced ELLPACK", slice = one warp of 32 consecutive
// cells).  Entry (slice s, k, lane) lives at slicePtr[s] + 32*k + lane, so a warp reads every
// per-entry array fully coalesced.  Entries of a row: internal-face neighbours, processor-face
// neighbours (ghost columns), boundary faces (boundary-slot columns).  Padding entries have
// col = row and all coefficients zero.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <string>
#include <vector>

#include "../../include/s4fgpu.h"

#define S4F_CHECK_CUDA(ctx, call)                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                         std::to_string(__LINE__) + ")";                                       \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)

#define S4F_CHECK_NCCL(ctx, call)                                                              \
    do {                                                                                       \
        ncclResult_t e_ = (call);                                                              \
        if (e_ != ncclSuccess) {                                                               \
            (ctx)->err = std::string(#call) + ": " + ncclGetErrorString(e_);                   \
            return 3;                                                                          \
        }                                                                                      \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr; n = 0;
    }
    void swap(DevBuf& o) { T* tp = p; p = o.p; o.p = tp; size_t tn = n; n = o.n; o.n = tn; }
    cudaError_t alloc(size_t count, bool zero = true) {
        if (count == n && p) { return zero ? cudaMemset(p, 0, n * sizeof(T)) : cudaSuccess; }
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return e; }
        n = count;
        return zero ? cudaMemset(p, 0, n * sizeof(T)) : cudaSuccess;
    }
    cudaError_t upload(const std::vector<T>& h) {
        cudaError_t e = alloc(h.size(), false);
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
};

#define S4F_MAX_RANKS 16
#define S4F_MAX_NBRS 16
#define S4F_RED_MAX 16

// all-reduce mailboxes of one context (device memory; null pointer = single rank)
struct PeerRed {
    int nRanks, rank;
    unsigned long long* box[S4F_MAX_RANKS];   // box[r]: rank r's mailbox of LL words  [2 parities][nRanks][2 * S4F_RED_MAX]
    unsigned int seq;                    // reductions completed so far
};

struct RedCtx { double* partials; unsigned int* ticket; PeerRed* peer; };

// scalar block of the fused 3-component PCG, one per solve, in device memory.
// All per-component quantities are [3].
struct PcgScalars {
    double part[16];       // finished reduction results of the last reducing kernel (raw sums)
    double rho[3];         // wArA
    double rhoOld[3];      // wArAold
    double wApA[3];
    double normFactor[3];
    double initRes[3];
    double finalRes[3];
    double avg[3];         // gAverage(psi)
    double alpha[3], beta[3];
    double omega[3];       // PBiCGStab
    int half[3];           // PBiCGStab: sA converged on the half step -> psi += alpha yA, then stop
    int active[3];         // component still iterating
    int nIter[3];
    int anyActive;
    int pad;
};

// reduction scalars of the outer loop (converged(): solidModelTemplates.C:27-188)
struct OuterScalars {
    double maxDelta;       // gMax |D - D.prevIter|
    double maxIncr;        // gMax |D - D.oldTime|
    double maxMag;         // gMax |D|
    double matNum, matDen; // material residual numerator / denominator
    double maxMagBE;       // gMax |bEbarTrial| (neoHookeanElasticMisesPlastic.C:1030)
    double minJ, maxJ;
};

struct s4fgpu_ctx {
    std::string err;
    int device = 0;
    cudaStream_t stream = nullptr;
    int numSMs = 148;
    long long launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // ---- parallel ----
    ncclComm_t comm = nullptr;            // set-up collectives only (IPC handles, agglomeration tables)
    int nRanks = 1, rank = 0;
    double nGlobalCells = -1;
    DevBuf<char> redArena;                // this rank's all-reduce mailbox (peers write into it over NVLink)
    DevBuf<PeerRed> redDev;               // device descriptor with every rank's mailbox pointers (s4f_comm.cu)
    std::vector<void*> ipcOpened;
    struct S4fHaloPlan* halo0 = nullptr;  // processor-patch halo of the fine mesh
    struct S4fHaloPlan* haloX = nullptr;  // point-neighbour ghosts (cells and boundary faces of other ranks at shared points)
    long long graphSerial = 0;            // bumped whenever anything a captured solve points at is rebuilt

    // ---- host copy of the mesh description ----
    int N = 0, F = 0, B = 0, G = 0, nPatches = 0;
    // decomposed meshes with points: X slots after the boundary slots hold the values of the cells and boundary faces of
    // OTHER ranks that share a point with this rank's cells (s4f_build_point_ghosts); filled by haloX before every
    // operator with a point stencil.  extPtr/extSlot/extCtr/extIsB: per local point those slots, their centres, cell or face.
    int X = 0;
    int xOff() const { return N + G + B; }
    std::vector<int> extPtr, extSlot; std::vector<double> extCtr; std::vector<char> extIsB;
    std::vector<double> extSymN; std::vector<int> extFixAxis;      // per point: a symmetry-plane normal / fixed axis known to another rank only
    int ld = 0;                       // leading dimension of vol-field component arrays (>= N+G+B)
    std::vector<int> own, nei, faceCells, pStart, pSize, pKind, pNbr, bcKind;
    int solD[3] = {1, 1, 1};
    std::vector<int> ghostOfFace;     // [B] ghost cell index (N+g) for processor faces, -1 otherwise
    bool meshSet = false, geomSet = false, lawSet = false, ctlSet = false;
    bool nonOrth = false;
    std::vector<double> hC, hV, hSf, hMagSf, hCf, hW, hNod, hCorr, hCnbrB;

    s4fgpu_law law{};
    s4fgpu_controls ctl{};
    double Hp = 0;                    // linear hardening modulus (2-point table)

    // ---- SELL-32 rows ----
    int nSlices = 0;
    long long nEntries = 0;           // padded
    long long nnzOff = 0;             // true off-diagonal entries (internal + processor faces)
    DevBuf<int> slicePtr;             // [nSlices+1] (entry offsets, multiples of 32), fits int for < 2^31 entries
    DevBuf<int> col;                  // [nEntries]
    DevBuf<double> eW;                // interpolation weight of the row cell's own value
    DevBuf<double> eSf;               // [3*nEntries] outward area vector (SoA: x | y | z)
    DevBuf<double> eLs;               // [3*nEntries] least-squares vector (SoA)
    DevBuf<double> eDn;               // magSf*nonOrthDeltaCoeffs (0 on boundary-face entries)
    DevBuf<double> eCorr;             // [3*nEntries] magSf*nonOrthCorrectionVector, outward sense (only if nonOrth)
    DevBuf<double> eA;                // laplacian coefficient impKf*magSf*delta  (= -upper)
    DevBuf<double> eGam;              // RhieChow gamma_f
    DevBuf<double> eU, eC0, eVc;      // factored RHS coefficients (k_source_g / k_source_m): (1-w) Sf [3*nE], gamma magSf delta - a, gamma (1-w) corr [3*nE]
    DevBuf<double> rowK;              // [6*ld] per-row sums U = sum w Sf (3), Vc = sum gamma w corr (3) of the factored RHS
    DevBuf<double> V, rV;             // cell volumes [ld]
    DevBuf<int> faceEntry;            // [F] entry index of internal face f in its owner's row (-> lduMatrix upper())
    DevBuf<int> procEntry;            // [G] entry index of processor-patch face g (ghost order) in its cell's row

    // ---- boundary faces (B-arrays, SoA) and boundary-cell lists ----
    DevBuf<int> bFaceCell, bKind;     // [B] ; bKind = S4F_BC_* of the face's patch
    DevBuf<double> bN, bK, bSf;       // [3B] unit normal, patch correction vector k, area vector
    DevBuf<double> bDelta, bNod, bMagSf; // [B]
    DevBuf<double> bcValue, bcPressure, tracGrad, bSn; // [3B],[B],[3B],[3B]
    DevBuf<int> bcCells, bcPtr, bcFaces; // boundary cells, CSR of their (non-processor) boundary faces
    int nBCells = 0;

    // ---- halo exchange ----
    struct Nbr { int rank; int patch; int count; int sendOff; int ghostOff; };
    std::vector<Nbr> nbrs;
    DevBuf<int> sendCells;            // [G] local cells adjacent to processor faces, in patch order

    // ---- fields (SoA, ld per component) ----
    DevBuf<double> D, Dprev, Dold, DoldOld;       // 3*ld
    DevBuf<double> Dooo, Doooo;                   // 3*ld: third / fourth old-time level (backward d2dt2 only)
    DevBuf<double> d2Hist;                        // 3*ld: old-time part of rho*fvm::d2dt2(D) per unit volume (transient schemes)
    // updated-Lagrangian model: the chains fvm::d2dt2(rho, DD) + fvc::d2dt2(rho, D.oldTime()) reach, and the density field
    DevBuf<double> Dooooo, DDo, DDoo, DDooo, DDoooo;   // 3*ld each
    DevBuf<double> rhoF, rhoO, rhoOO;                  // ld each
    // ---- point mesh (vol->point interpolation): CSR over points of source slots in the vol-field index space ----
    int nPoints = 0;
    std::vector<int> hFvPtr, hFv;                 // faces() of the fv faces
    std::vector<double> hBSfHost;                 // [3B] boundary area vectors (patch normals of the point constraints)
    bool rhoInit = false;
    std::vector<double> hCfB;                     // [3B] boundary face centres of the last set_geometry (hC is kept too)
    // device mesh motion (s4f_geom.cu): points, face -> vertex CSR, signed face of every row entry, face / cell geometry
    DevBuf<double> dPoints;                       // [3*nPoints] AoS
    DevBuf<int> dFvPtr, dFv, eFaceS, ptFixAxis;   // [F+B+1], [sum verts], [nEntries] +-(face+1) (0 = padding), [nPoints] axis a symmetry plane fixes (-1: none)
    DevBuf<double> fCtr, fSf, Cc;                 // [3*(F+B)] SoA, [3*(F+B)] SoA, [3*ld] cell centres (ghosts by halo exchange)
    bool amgRefresh = false;                      // the hierarchy's aggregates are still good: re-sum the coefficients only
    bool hostGeomStale = false;                   // the mesh moved on the device: hC / hCfB / hPoints are those of the last host call
    DevBuf<int> ptPtr, ptCol;                     // [nPoints+1], [nnzP]
    DevBuf<double> ptW, ptN;                      // [nnzP] normalised inverse-distance weights; [3*nPoints] constraint normal (0 = none)
    DevBuf<double> ptOut;                         // [3*nPoints]
    // pointCellsLeastSquares gradient: its own (wider) SELL-32 rows: cells sharing a point + boundary faces at the cell's points
    DevBuf<int> gSlicePtr, gCol; DevBuf<double> gLs; long long gNE = 0; bool gValid = false;
    std::vector<double> hPoints;
    DevBuf<int> pgPtr, pgCol;                     // gradient-extrapolated variant: every point from its pointCells
    DevBuf<double> pgW, pgDelta;                  // [nnz] normalised weights, [3*nnz] point - cell centre
    bool histValid = false;
    int timeIndex = 0;                            // new_timestep() calls = runTime.timeIndex()
    DevBuf<double> gradD, gradDold;               // 9*ld
    // D / gradD hold the SOLUTION field: D for the total-displacement models, DD for the incremental ones
    // (nonLinGeomTotalLagSolid.C:152-161), which keep D = D.oldTime() + DD and gradD = gradD.oldTime() + gradDD here
    DevBuf<double> Dtot, gradDtot;                // 3*ld, 9*ld (incremental models only)
    DevBuf<double> sigma, sigmaOld;               // 6*ld
    DevBuf<double> impK;                          // ld
    DevBuf<double> T9;                            // 9*ld: the combined tensor M = T - gamma grad(D) of the factored right-hand side, T = sigma or
                                                  // J Finv & sigma / relJ relFinv & sigma  [cells + ghosts + boundary slots: M_b = T_b]
    bool mValid = false;                          // T9 holds M of the current sigma / grad(D)
    double impK0 = 0;                             // the (uniform) implicit stiffness
    DevBuf<double> Finv, Jt;                      // 9*ld, ld (TL solver kinematics)
    // law history
    DevBuf<double> lawF, lawFold, lawJ, lawJold, bEbar, bEbarOld, sigmaY, sigmaYOld, DSigmaY, epsPEq, epsPEqOld,
        DEpsPEq, epsP, epsPOld, DEpsP, DEpsPprev, DLambda, plasticN, epsilon;
    // ---- pressure smoothing (mechanicalLaw::updateSigmaHyd, solvePressureEqn) ----
    DevBuf<double> sigmaHyd, pExp, pRatio;        // ld: sigmaHyd, explicit hydrostatic stress, impK/DEqnA
    DevBuf<double> gradP;                         // 3*ld: grad(sigmaHyd)
    DevBuf<double> eP;                            // [nEntries] rDAf magSf delta
    DevBuf<double> pDiag, pRDiag, pX, pB;         // 3*ld each (the scalar system rides the fused 3-component PCG, components 1,2 idle)
    s4fgpu_stats lastP{};
    // ---- fvMatrix ----
    DevBuf<double> diag0;             // ld: sum of laplacian coefficients + d2dt2
    DevBuf<double> diagC;             // 3*ld: per-component diagonal after addBoundaryDiag
    DevBuf<double> rDiagC;            // 3*ld: 1/diagC ([OF-ext] diagonalPreconditioner rD)
    DevBuf<double> source;            // 3*ld
    bool matrixValid = false;
    // ---- PCG work vectors ----
    DevBuf<double> pA, wA, rA;        // 3*ld each
    DevBuf<double> zA;                // 3*ld: M^-1 rA of the non-local preconditioners
    struct S4fSolveGraphs* solveGraphs = nullptr;   // captured solves (s4f_pcg.cu)
    bool solvePending = false;        // a solve's statistics are still on their way to hPcgS
    long long pendingPre = 0, pendingBody = 0;
    DevBuf<double> cheb0, cheb1;      // 3*ld polynomial-preconditioner work
    DevBuf<double> bi[6];             // 3*ld each: PBiCGStab rA0, yA, AyA, sA, zA, tA
    DevBuf<double> aitRes, aitResPrev, aitAlpha;  // Aitken relaxation state (solidModel.C:842-897)
    DevBuf<PcgScalars> pcgS;
    DevBuf<OuterScalars> outS;
    DevBuf<double> partials;          // per-block partial sums
    DevBuf<unsigned int> ticket;
    DevBuf<double> staging;           // AoS <-> SoA staging, 9*max(N,B,F)
    DevBuf<double> flushBuf;          // L2 flush buffer for time_kernel
    PcgScalars* hPcgS = nullptr;      // pinned mirrors
    OuterScalars* hOutS = nullptr;
    double lambdaMax = 2.0;           // Chebyshev: bound of the Jacobi-scaled spectrum
    struct S4fAmg* amg = nullptr;     // GAMG hierarchy (s4f_amg.cu), rebuilt with the matrix
    struct S4fDic* dic = nullptr;     // level-scheduled DIC (s4f_dic.cu), rebuilt with the matrix
    bool dicValid = false;
    struct S4fUns* uns = nullptr;     // face-based data of the unsLinearGeometry model (s4f_uns.cu)
    bool unsValid = false;
    DevBuf<int> ones3;                // {1,1,1}
    const int* amgAct = nullptr;      // device int[3] of components the V-cycle works on (null: all); the PCG passes its active flags
    bool amgValid = false;

    int iCorr = 0;
    long long totalInner = 0;
    s4fgpu_stats last{};

    bool incremental() const { return ctl.solidModel == S4F_MODEL_NONLIN_TL || ctl.solidModel == S4F_MODEL_NONLIN_UL || ctl.solidModel == S4F_MODEL_UNS_NONLIN_UL; }
    bool UL() const { return ctl.solidModel == S4F_MODEL_NONLIN_UL || ctl.solidModel == S4F_MODEL_UNS_NONLIN_UL; }
    bool unsModel() const { return ctl.solidModel == S4F_MODEL_UNS_LIN_GEOM || ctl.solidModel == S4F_MODEL_UNS_NONLIN_TL || ctl.solidModel == S4F_MODEL_UNS_NONLIN_UL; }
    bool unsTL() const { return ctl.solidModel == S4F_MODEL_UNS_NONLIN_TL; }
    bool unsUL() const { return ctl.solidModel == S4F_MODEL_UNS_NONLIN_UL; }
    bool unsFinite() const { return unsTL() || unsUL(); }
    double unsMaxRes = 0;                 // unsNonLinGeomTotalLagSolid::evolve: largest relative residual of the time step
    bool unsGradientsOnly = false;        // s4fgpu_initialise: the uns models' constructors update the gradients, not sigmaf
    // cell-level F / Finv / J of the solver (the uns updated-Lagrangian model needs relJ of the cells for its density update)
    bool finiteStrain() const { return ctl.solidModel == S4F_MODEL_NONLIN_TL_TOTAL_DISP || ctl.solidModel == S4F_MODEL_NONLIN_TL || UL(); }
    bool pointCellsGrad() const { return ctl.gradScheme == S4F_GRAD_POINT_CELLS_LEAST_SQUARES; }
    // rows the least-squares gradient kernels run over
    const int* gradSlicePtr() const { return pointCellsGrad() ? gSlicePtr.p : slicePtr.p; }
    const int* gradCol() const { return pointCellsGrad() ? gCol.p : col.p; }
    const double* gradLs() const { return pointCellsGrad() ? gLs.p : eLs.p; }
    long long gradNE() const { return pointCellsGrad() ? gNE : nEntries; }
    double gamma0() const { return ctl.stabilisation == S4F_STAB_RHIE_CHOW ? ctl.stabScaleFactor * impK0 : 0.0; }
    const double* gradForLaw() const { return incremental() ? gradDtot.p : gradD.p; }   // the registered "grad(D)"
    RedCtx red() const { return RedCtx{partials.p, ticket.p, nRanks > 1 ? redDev.p : nullptr}; }
    int NT() const { return N + G + B; }
    int bOff() const { return N + G; }
};

// ---- internal entry points (defined across the .cu files) ----
int s4f_build_rows(s4fgpu_ctx* c);                   // SELL rows + boundary lists + geometry upload
int s4f_alloc_fields(s4fgpu_ctx* c);
int s4f_setup_law(s4fgpu_ctx* c);
int s4f_assemble_matrix(s4fgpu_ctx* c);
int s4f_assemble_source(s4fgpu_ctx* c);
int s4f_d2dt2_history(s4fgpu_ctx* c);               // d2Hist from the old-time levels (lazy, once per time step)
int s4f_bc_update_coeffs(s4fgpu_ctx* c);
int s4f_bc_evaluate(s4fgpu_ctx* c);
int s4f_relax_and_residual(s4fgpu_ctx* c, int iCorr);
int s4f_grad(s4fgpu_ctx* c);
int s4f_kinematics(s4fgpu_ctx* c);
int s4f_pressure_smooth(s4fgpu_ctx* c);              // updateSigmaHyd with the pressure equation; fixes sigma in place
int s4f_dic_setup(s4fgpu_ctx* c);                    // s4f_dic.cu: exact DIC / FDIC by level scheduling
int s4f_dic_apply(s4fgpu_ctx* c, const double* r3, double* z3);
void s4f_dic_destroy(s4fgpu_ctx* c);
int s4f_uns_setup(s4fgpu_ctx* c);                    // s4f_uns.cu: the face-stress ("uns") discretisation
int s4f_uns_gradients(s4fgpu_ctx* c);
int s4f_uns_bc_update(s4fgpu_ctx* c);
int s4f_uns_source(s4fgpu_ctx* c);
int s4f_uns_new_timestep(s4fgpu_ctx* c);        // Ff.oldTime() of the uns updated-Lagrangian model
int s4f_uns_download(s4fgpu_ctx* c, int field, double* host);
void s4f_uns_destroy(s4fgpu_ctx* c);
int s4f_bc_sngrad_store(s4fgpu_ctx* c);              // snGrad() of every boundary face into bSn
int s4f_grad_calculated_interior(s4fgpu_ctx* c, const double* X, double* gradOut);   // cell values of fvc::grad(X) only
int s4f_make_m(s4fgpu_ctx* c);                       // lin-geom: M = sigma - gamma grad(D) when no law kernel produced it
int s4f_update_totals(s4fgpu_ctx* c, bool disp, bool grad);
int s4f_law_correct(s4fgpu_ctx* c);
double s4f_law_bytes(const s4fgpu_ctx* c);           // algorithmic bytes of one s4f_law_correct (roofline report)
int s4f_solve_segregated(s4fgpu_ctx* c, double* psi, const double* source, bool defer = false);   // device SoA pointers
int s4f_finish_solve(s4fgpu_ctx* c);
int s4f_halo_exchange(s4fgpu_ctx* c, double* field, int ncomp);
void s4f_solve_graphs_destroy(s4fgpu_ctx* c);       // captured PCG solves (s4f_pcg.cu)
int s4f_outer_iteration(s4fgpu_ctx* c, int iCorr);
int s4f_read_outer_scalars(s4fgpu_ctx* c, s4fgpu_stats* st, bool* converged, int iCorr);
int s4f_aos_to_soa(s4fgpu_ctx* c, const double* hostAoS, double* devSoA, int count, int ncomp, int offset);
int s4f_soa_to_aos(s4fgpu_ctx* c, const double* devSoA, double* hostAoS, int count, int ncomp, int offset);
int s4f_time_pcg_kernels(s4fgpu_ctx* c, int kernel, int reps, int flushL2, double* ms, double* bytes);
int s4f_time_fv_kernels(s4fgpu_ctx* c, int kernel, int reps, int flushL2, double* ms, double* bytes);
int s4f_amul_device(s4fgpu_ctx* c, const double* x3, double* w3, int mask);
int s4f_alloc_model_fields(s4fgpu_ctx* c);
int s4f_upload_bc(s4fgpu_ctx* c);
int s4f_amg_setup(s4fgpu_ctx* c);                                   // after s4f_assemble_matrix
int s4f_amg_refresh(s4fgpu_ctx* c);                                 // same aggregates, Galerkin sums of the new fine matrix (device)
int s4f_move_points_device(s4fgpu_ctx* c, const double* hostPointDD);   // s4f_geom.cu
int s4f_refresh_host_geometry(s4fgpu_ctx* c);
int s4f_amg_apply(s4fgpu_ctx* c, const double* r3, double* z3);     // z = M^-1 r, 3 components, stride ld
int s4f_amg_step0(s4fgpu_ctx* c, const double* r3, double* bytes);
void s4f_amg_destroy(s4fgpu_ctx* c);
int s4f_amg_distributed_levels(s4fgpu_ctx* c);
int s4f_amg_info(s4fgpu_ctx* c, int* nLevels, int* sizes, int maxLevels, double* bytesPerApply, double* setupSeconds);
int s4f_download_upper(s4fgpu_ctx* c, double* hostUpper);           // lduMatrix upper() [F]
int s4f_build_point_stencil(s4fgpu_ctx* c);                         // rows of the pointCellsLeastSquares gradient (s4f_setup.cu)
int s4f_build_point_ghosts(s4fgpu_ctx* c);                          // decomposed meshes: collective, from set_geometry
int s4f_point_ghost_exchange(s4fgpu_ctx* c, double* field, int ncomp);
int s4f_build_point_weights(s4fgpu_ctx* c, const double* points);   // vol->point CSR + weights (s4f_setup.cu)
int s4f_interpolate_to_points(s4fgpu_ctx* c, const double* field3, const double* grad9 /* null: patch mode */, double* hostOut);
int s4f_grad_calculated(s4fgpu_ctx* c, const double* X, double* gradOut);   // fvc::grad of a field with calculated patches
