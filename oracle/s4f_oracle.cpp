/*
 * s4f_oracle.cpp -- CPU restatement of the solids4foam segregated solid-solver hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under solids4foam_b200/ (the product) may include, link,
 * load or call this file; it is the checker used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * PARITY STATUS: "parity unpinned" at the OpenFOAM operator boundary.  The reference
 * (/root/reference) holds no golden fields or unit tests for this path, and the OpenFOAM library
 * that implements fvc::grad / fvc::div / fvm::laplacian / lduMatrix / PCG / DIC is neither vendored
 * nor installed.  The oracle is pinned instead against the closed-form known-answer
 * tests the reference ships (Kirsch plate-hole, patch test, single-cell law identities, the
 * neckingBar hardening table) in tests/test_oracle_*.py.
 *
 * Style: deliberately the reference's own data layout and loop shapes -- AoS fields, LDU
 * owner/neighbour face-loop scatter, boundary patches as separate face ranges, one segregated PCG
 * solve per component -- NOT the cell-centric SoA gather design of the CUDA path.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference;
 * SM = src/solids4FoamModels/solidModels, ML = src/solids4FoamModels/materialModels/
 * mechanicalModel/mechanicalLaws, NUM = src/solids4FoamModels/numerics).  "[OF-ext]" marks
 * behaviour of the OpenFOAM library restated from its published algorithms.
 */
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <omp.h>

#include "../include/s4fgpu.h"

namespace {

const double SMALL = 1e-15;   // [OF-ext] Foam::SMALL (double)
const double VSMALL = 1e-300;

typedef std::vector<double> dvec;
typedef std::vector<int> ivec;

// ---- small tensor algebra (OpenFOAM component orders) ------------------------------------------
inline void symm(const double* T, double* S) {             // symm(T) = (T+T^T)/2
    S[0] = T[0]; S[1] = 0.5 * (T[1] + T[3]); S[2] = 0.5 * (T[2] + T[6]);
    S[3] = T[4]; S[4] = 0.5 * (T[5] + T[7]); S[5] = T[8];
}
inline double trS(const double* S) { return S[0] + S[3] + S[5]; }
inline void devS(const double* S, double* D) {
    double t = trS(S) / 3.0;
    D[0] = S[0] - t; D[1] = S[1]; D[2] = S[2]; D[3] = S[3] - t; D[4] = S[4]; D[5] = S[5] - t;
}
inline double magSqrS(const double* S) {
    return S[0] * S[0] + 2 * S[1] * S[1] + 2 * S[2] * S[2] + S[3] * S[3] + 2 * S[4] * S[4] + S[5] * S[5];
}
inline double detS(const double* S) {
    return S[0] * S[3] * S[5] + S[1] * S[4] * S[2] + S[2] * S[1] * S[4]
         - S[0] * S[4] * S[4] - S[1] * S[1] * S[5] - S[2] * S[3] * S[2];
}
inline double detT(const double* T) {
    return T[0] * (T[4] * T[8] - T[5] * T[7]) - T[1] * (T[3] * T[8] - T[5] * T[6])
         + T[2] * (T[3] * T[7] - T[4] * T[6]);
}
inline void invT(const double* T, double* R) {
    double d = detT(T);
    R[0] = (T[4] * T[8] - T[5] * T[7]) / d; R[1] = (T[2] * T[7] - T[1] * T[8]) / d; R[2] = (T[1] * T[5] - T[2] * T[4]) / d;
    R[3] = (T[5] * T[6] - T[3] * T[8]) / d; R[4] = (T[0] * T[8] - T[2] * T[6]) / d; R[5] = (T[2] * T[3] - T[0] * T[5]) / d;
    R[6] = (T[3] * T[7] - T[4] * T[6]) / d; R[7] = (T[1] * T[6] - T[0] * T[7]) / d; R[8] = (T[0] * T[4] - T[1] * T[3]) / d;
}
inline void invS(const double* S, double* R) {
    double d = detS(S);
    R[0] = (S[3] * S[5] - S[4] * S[4]) / d; R[1] = (S[2] * S[4] - S[1] * S[5]) / d; R[2] = (S[1] * S[4] - S[2] * S[3]) / d;
    R[3] = (S[0] * S[5] - S[2] * S[2]) / d; R[4] = (S[1] * S[2] - S[0] * S[4]) / d; R[5] = (S[0] * S[3] - S[1] * S[1]) / d;
}
inline void mulTT(const double* A, const double* B, double* R) {   // A & B
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j];
        R[3 * i + j] = s;
    }
}
inline void S2T(const double* S, double* T) {
    T[0] = S[0]; T[1] = S[1]; T[2] = S[2]; T[3] = S[1]; T[4] = S[3]; T[5] = S[4]; T[6] = S[2]; T[7] = S[4]; T[8] = S[5];
}
inline void transposeT(const double* A, double* R) {
    R[0] = A[0]; R[1] = A[3]; R[2] = A[6]; R[3] = A[1]; R[4] = A[4]; R[5] = A[7]; R[6] = A[2]; R[7] = A[5]; R[8] = A[8];
}
inline void SvS(const double* S, const double* v, double* r) {      // S & v
    r[0] = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
    r[1] = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
    r[2] = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
}
inline void vT(const double* v, const double* T, double* r) {       // v & T : r_j = v_i T_ij
    for (int j = 0; j < 3; j++) r[j] = v[0] * T[j] + v[1] * T[3 + j] + v[2] * T[6 + j];
}
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double mag3(const double* a) { return std::sqrt(dot3(a, a)); }

// s4f interpolationTable<scalar>::operator(), clamp; NUM/interpolationTable/interpolationTable.C:493-632
double tableLookup(const s4fgpu_law& L, double x) {
    int n = L.nTable;
    if (n <= 1) return L.tableSigY[0];
    if (x < L.tableEps[0]) return L.tableSigY[0];
    if (x >= L.tableEps[n - 1]) return L.tableSigY[n - 1];
    int lo = 0, hi = 0;
    for (int i = 0; i < n; i++) {
        if (x >= L.tableEps[i]) { lo = hi = i; } else { hi = i; break; }
    }
    if (lo == hi) return L.tableSigY[hi];
    return L.tableSigY[lo] + (L.tableSigY[hi] - L.tableSigY[lo]) * (x - L.tableEps[lo]) / (L.tableEps[hi] - L.tableEps[lo]);
}

struct SolverPerf { double initRes, finalRes; int nIter; };

// ------------------------------------------------------------------------------------------------
// CPU agglomeration multigrid for the SAME preconditioner family as the GPU path's default: [OF-ext] GAMG-style pair-wise
// agglomeration by strongest face coefficient (three passes per level), Galerkin sums for piece-wise constant transfer,
// Chebyshev-Jacobi smoothing in three-term form, fixed over-correction, dense coarsest solve, optional K-cycle on level 1
// (two flexible-CG steps).  NOT part of the reference's algorithm (its tutorials run PCG + DIC): it exists so that
// bench.py can time the CPU on the algorithm the GPU runs ("a second CPU figure with the same preconditioner
// family"), and it is its own restatement -- nothing here is shared with solids4foam_b200/csrc/s4f_amg.cu.
// Enabled per handle with s4fo_set_cpu_gamg; otherwise S4F_PRECOND_GAMG keeps mapping onto DIC (the parity preconditioner).
// ------------------------------------------------------------------------------------------------
struct OLevel {
    int n = 0;
    std::vector<int> ptr, col; dvec val;      // rows: positive couplings a_ij (A_ij = -a_ij), both triangles
    dvec diag[3];
    std::vector<int> parent;                  // cell -> cell of the next level
    std::vector<int> cptr, child;             // next level's cells -> their children here
    dvec b, x, xp, xn, t;                     // work vectors of one scalar solve
};
struct OGamg {
    std::vector<OLevel> lv;
    dvec inv[3]; int nC = 0;
    int deg = 3, cycle = 2; double omega = 2.2, theta = 1.3, delta = 0.7;
    dvec kc1, kv1, kr, kc2;                   // K-cycle work on level 1
    bool valid = false;
};

void ogPairPass(const OLevel& L, std::vector<int>& agg, int& nc) {
    agg.assign(L.n, -1); nc = 0;
    for (int i = 0; i < L.n; i++) {
        if (agg[i] >= 0) continue;
        int best = -1; float bw = 0.f;
        for (int e = L.ptr[i]; e < L.ptr[i + 1]; e++) {
            const int j = L.col[e];
            if (agg[j] < 0 && j != i && (float)L.val[e] > bw * 1.0000001f) { bw = (float)L.val[e]; best = j; }
        }
        agg[i] = nc; if (best >= 0) agg[best] = nc;
        nc++;
    }
}

// coarse level of the aggregates agg (piece-wise constant transfer): couplings between aggregates add up, couplings
// inside an aggregate leave the diagonal (A_c[I][I] = sum_i A_ii - 2 sum_{faces inside} a)
void ogGalerkin(const OLevel& L, const std::vector<int>& agg, int nc, OLevel& C) {
    C.n = nc;
    for (int q = 0; q < 3; q++) C.diag[q].assign(nc, 0.0);
    std::vector<int> cnt(nc + 1, 0);
    for (int i = 0; i < L.n; i++) cnt[agg[i] + 1]++;
    for (int I = 0; I < nc; I++) cnt[I + 1] += cnt[I];
    std::vector<int> kids(L.n), cur(cnt.begin(), cnt.end() - 1);
    for (int i = 0; i < L.n; i++) kids[cur[agg[i]]++] = i;
    C.ptr.assign(nc + 1, 0); C.col.clear(); C.val.clear();
    std::vector<std::pair<int, double>> row;
    for (int I = 0; I < nc; I++) {
        row.clear();
        double inside = 0;
        for (int k = cnt[I]; k < cnt[I + 1]; k++) {
            const int i = kids[k];
            for (int q = 0; q < 3; q++) C.diag[q][I] += L.diag[q][i];
            for (int e = L.ptr[i]; e < L.ptr[i + 1]; e++) {
                const int J = agg[L.col[e]];
                if (J == I) inside += L.val[e]; else row.emplace_back(J, L.val[e]);
            }
        }
        for (int q = 0; q < 3; q++) C.diag[q][I] -= inside;
        std::sort(row.begin(), row.end(), [](const std::pair<int, double>& a, const std::pair<int, double>& b) { return a.first < b.first; });
        for (size_t k = 0; k < row.size();) {
            size_t m = k; double sum = 0;
            while (m < row.size() && row[m].first == row[k].first) sum += row[m++].second;
            C.col.push_back(row[k].first); C.val.push_back(sum);
            k = m;
        }
        C.ptr[I + 1] = (int)C.col.size();
    }
}


}  // namespace
struct s4f_oracle;
namespace { void updateSigmaHydSmoothed(s4f_oracle& o, double impK); }
extern "C" int s4fo_interpolate_to_points_impl(s4f_oracle* o, const std::vector<double>& X, double* out);

struct s4f_oracle {
    std::string err;
    // mesh ([OF-ext] lduAddressing + fvBoundaryMesh)
    int N = 0, F = 0, B = 0, nPatches = 0;
    ivec own, nei, faceCells, pStart, pSize, pKind, bcKind;
    int solD[3] = {1, 1, 1};
    // geometry
    dvec C, V, Sf, magSf, Cf, w, nod, corr, CnbrB;
    dvec lsP, lsN;                 // least-squares vectors (F+B)*3, F*3
    bool nonOrth = false;
    // models
    s4fgpu_law law{};
    s4fgpu_controls ctl{};
    double Hp = 0;
    // BC data (B-arrays)
    dvec bcValue, bcPressure, tracGrad;
    // vol fields: internal [0,N) then boundary [N,N+B)
    // D / gradD hold the SOLUTION field of the model: D for the total-displacement models, DD for the
    // incremental ones (nonLinGeomTotalLagSolid solves DD: SM/nonLinGeomTotalLagSolid/...C:152-161); the
    // incremental models keep the total displacement and its gradient in Dtot / gradDtot
    // (D = D.oldTime() + DD :190, gradD = gradD.oldTime() + gradDD :196).
    dvec D, Dprev, Dold, DoldOld, gradD, gradDold, sigma, sigmaOld, Dtot, gradDtot;
    dvec Dooo, Doooo;              // third / fourth old-time level (backward d2dt2 only)
    // updated-Lagrangian model (SM/nonLinGeomUpdatedLagSolid): the old-time chains that
    // fvm::d2dt2(rho, DD) + fvc::d2dt2(rho, D.oldTime()) reach, and the density field with its old times
    dvec Dooooo, DDo, DDoo, DDooo, DDoooo, rho, rhoO, rhoOO;
    dvec gradSigmaHyd, sigmaHydExp;     // pressure smoothing (mechanicalLaw.C:1366-1476)
    dvec pointD, gradDf, sigmaf;        // unsLinGeomSolid: vertex displacements [3 nPoints], face gradient [9 (F+B)], face stress [6 (F+B)]
    SolverPerf perfP{0, 0, 0};
    // polyMesh points/faces for vol->point interpolation
    int nPoints = 0;
    dvec points; ivec fvPtr, fv, pcPtr, pcCells, pbPtr, pbFaces;
    // pointCellsLeastSquares: per cell the stencil slots (cell index, or N + boundary face) and the least-squares vectors
    ivec psPtr, psSlot; dvec psLs; bool psValid = false;
    int timeIndex = 0;             // number of new_timestep() calls (runTime.timeIndex())
    dvec impK, impKf;              // impK (N+B), impKf (F+B)
    dvec Ft, Finv, Jt;             // solver-level F, Finv, J of the TL models
    // law history
    dvec lawF, lawFold, relF, lawJ, lawJold, bEbar, bEbarOld, bEbarTrial, sigmaY, DSigmaY, epsPEq, DEpsPEq,
         epsP, DEpsP, DEpsPprev, DLambda, plasticN, epsilon, sigmaHyd, epsPOld, epsPEqOld, sigmaYOld;
    // fvMatrix
    dvec upper, diag, diagC, source, intCoeffs, bouCoeffs;
    bool matrixValid = false;
    bool cpuGamg = false;          // S4F_PRECOND_GAMG runs the CPU multigrid below instead of mapping onto DIC (s4fo_set_cpu_gamg)
    OGamg gamg;
    int curComp = 0;               // component being solved (the CPU multigrid keeps one diagonal per component)
    // Aitken
    dvec aitkenRes, aitkenResPrev, aitkenAlpha;
    // last solve
    SolverPerf perf[3];
    long long totalInner = 0;
    int iCorrLast = 0;

    // Host threading for the timing baseline only (nThreads == 1: the literal serial LDU loops).
    // Cells are cut into nThreads contiguous index ranges ("ranks"); a face loop runs each range's
    // own faces in parallel and the few faces that straddle two ranges afterwards, serially; DIC is
    // applied per range and ignores the straddling faces, which is how OpenFOAM's DIC behaves across
    // MPI ranks (block Jacobi).
    int nThreads = 1;
    ivec cellStart, faceStart, crossFaces;

    int NB() const { return N + B; }
    bool incremental() const { return ctl.solidModel == S4F_MODEL_NONLIN_TL || ctl.solidModel == S4F_MODEL_NONLIN_UL || ctl.solidModel == S4F_MODEL_UNS_NONLIN_UL; }
    bool UL() const { return ctl.solidModel == S4F_MODEL_NONLIN_UL || ctl.solidModel == S4F_MODEL_UNS_NONLIN_UL; }
    bool uns() const { return ctl.solidModel == S4F_MODEL_UNS_LIN_GEOM || ctl.solidModel == S4F_MODEL_UNS_NONLIN_TL || ctl.solidModel == S4F_MODEL_UNS_NONLIN_UL; }
    bool unsTL() const { return ctl.solidModel == S4F_MODEL_UNS_NONLIN_TL; }
    bool unsUL() const { return ctl.solidModel == S4F_MODEL_UNS_NONLIN_UL; }
    bool unsFinite() const { return unsTL() || unsUL(); }
    dvec Ff, FfOld;                // unsNonLinGeomUpdatedLagSolid: total deformation gradient on the faces and its oldTime() [9 (F+B)]
    double unsMaxRes = 0;          // unsNonLinGeomTotalLagSolid::evolve: the largest relative residual of this time step
    const dvec& gradForLaw() const { return incremental() ? gradDtot : gradD; }   // the registered "grad(D)"
};

namespace {


void buildPartition(s4f_oracle& o) {
    const int T = o.nThreads, N = o.N, F = o.F;
    o.cellStart.assign(T + 1, 0); o.faceStart.assign(T + 1, 0); o.crossFaces.clear();
    for (int t = 0; t <= T; t++) o.cellStart[t] = (int)((long long)N * t / T);
    int f = 0;
    for (int t = 0; t < T; t++) {            // faces are sorted by owner
        o.faceStart[t] = f;
        while (f < F && o.own[f] < o.cellStart[t + 1]) { if (o.nei[f] >= o.cellStart[t + 1]) o.crossFaces.push_back(f); f++; }
    }
    o.faceStart[T] = F;
}

// internal-face loop with owner/neighbour scatter: body(f) may write to own[f] and nei[f]
template <class Body>
inline void forAllInternalFaces(const s4f_oracle& o, Body body) {
    if (o.nThreads <= 1) { for (int f = 0; f < o.F; f++) body(f); return; }
#pragma omp parallel num_threads(o.nThreads)
    {
        const int t = omp_get_thread_num();
        const int cEnd = o.cellStart[t + 1];
        for (int f = o.faceStart[t]; f < o.faceStart[t + 1]; f++) if (o.nei[f] < cEnd) body(f);
    }
    for (size_t i = 0; i < o.crossFaces.size(); i++) body(o.crossFaces[i]);
}
#define S4FO_PAR_FOR _Pragma("omp parallel for schedule(static) num_threads(o.nThreads) if (o.nThreads > 1)")

// ------------------------------------------------------------------------------------------------
// least-squares vectors: NUM/extendedLeastSquaresGrad/extendedLeastSquaresVectors.C:121-158 (dd),
// :217 (inv), :229-272 (lsP, lsN).  1/|d|^2 weights, true boundary deltas Cf - Cn.  Empty
// directions: [OF-ext] inv(symmTensorField) adds 1 on the diagonal of the singular direction,
// inverts, and removes it again.
// ------------------------------------------------------------------------------------------------
void makeLeastSquaresVectors(s4f_oracle& o) {
    const int N = o.N, F = o.F, B = o.B;
    dvec dd(6 * N, 0.0);
    auto addwdd = [&](int c, const double* d) {
        double r = 1.0 / dot3(d, d);
        double* t = &dd[6 * c];
        t[0] += r * d[0] * d[0]; t[1] += r * d[0] * d[1]; t[2] += r * d[0] * d[2];
        t[3] += r * d[1] * d[1]; t[4] += r * d[1] * d[2]; t[5] += r * d[2] * d[2];
    };
    for (int f = 0; f < F; f++) {
        int P = o.own[f], Nn = o.nei[f];
        double d[3] = {o.C[3 * Nn] - o.C[3 * P], o.C[3 * Nn + 1] - o.C[3 * P + 1], o.C[3 * Nn + 2] - o.C[3 * P + 2]};
        addwdd(P, d); addwdd(Nn, d);
    }
    for (int b = 0; b < B; b++) {
        int P = o.faceCells[b];
        const double* x = &o.CnbrB[3 * b];   // Cf on ordinary patches (neighbour centre on processor faces)
        double d[3] = {x[0] - o.C[3 * P], x[1] - o.C[3 * P + 1], x[2] - o.C[3 * P + 2]};
        addwdd(P, d);
    }
    dvec invDd(6 * N);
    for (int c = 0; c < N; c++) {
        double t[6]; for (int k = 0; k < 6; k++) t[k] = dd[6 * c + k];
        if (!o.solD[0]) t[0] += 1; if (!o.solD[1]) t[3] += 1; if (!o.solD[2]) t[5] += 1;
        double r[6]; invS(t, r);
        if (!o.solD[0]) r[0] -= 1; if (!o.solD[1]) r[3] -= 1; if (!o.solD[2]) r[5] -= 1;
        for (int k = 0; k < 6; k++) invDd[6 * c + k] = r[k];
    }
    o.lsP.assign(3 * (F + B), 0.0); o.lsN.assign(3 * F, 0.0);
    for (int f = 0; f < F; f++) {
        int P = o.own[f], Nn = o.nei[f];
        double d[3] = {o.C[3 * Nn] - o.C[3 * P], o.C[3 * Nn + 1] - o.C[3 * P + 1], o.C[3 * Nn + 2] - o.C[3 * P + 2]};
        double r = 1.0 / dot3(d, d), a[3], b2[3];
        SvS(&invDd[6 * P], d, a); SvS(&invDd[6 * Nn], d, b2);
        for (int k = 0; k < 3; k++) { o.lsP[3 * f + k] = r * a[k]; o.lsN[3 * f + k] = -r * b2[k]; }
    }
    for (int b = 0; b < B; b++) {
        int P = o.faceCells[b];
        const double* x = &o.CnbrB[3 * b];
        double d[3] = {x[0] - o.C[3 * P], x[1] - o.C[3 * P + 1], x[2] - o.C[3 * P + 2]};
        double r = 1.0 / dot3(d, d), a[3];
        SvS(&invDd[6 * P], d, a);
        for (int k = 0; k < 3; k++) o.lsP[3 * (F + b) + k] = r * a[k];
    }
}

// patchCorrectionVectors: k = (I - nn) & (Cf - Cn); NUM/patchCorrectionVectors/patchCorrectionVectors.C:24-36
inline void patchGeom(const s4f_oracle& o, int b, double* n, double* k, double& delta) {
    const int f = o.F + b, P = o.faceCells[b];
    for (int i = 0; i < 3; i++) n[i] = o.Sf[3 * f + i] / o.magSf[f];
    double d[3] = {o.Cf[3 * f] - o.C[3 * P], o.Cf[3 * f + 1] - o.C[3 * P + 1], o.Cf[3 * f + 2] - o.C[3 * P + 2]};
    double nd = dot3(n, d);
    for (int i = 0; i < 3; i++) k[i] = d[i] - n[i] * nd;
    delta = o.nod[f];   // patch().deltaCoeffs(): 1/max(n.d, 0.05|d|)
}

// ------------------------------------------------------------------------------------------------
// Boundary conditions of D
// ------------------------------------------------------------------------------------------------

// tractionBoundarySnGrad: SM/linGeomTotalDispSolid/linGeomTotalDispSolid.C:235-271 and the
// total-Lagrangian form SM/nonLinGeomTotalLagTotalDispSolid/...C:284-328 (deformed normal).
void tractionSnGrad(const s4f_oracle& o, int b, double* g) {
    const int N = o.N;
    double n[3], k[3], delta; patchGeom(o, b, n, k, delta);
    const double* t = &o.bcValue[3 * b];
    const double p = o.bcPressure[b];
    const double impK = o.impK[N + b], rImpK = 1.0 / impK;
    const double* gD = &o.gradD[9 * (N + b)];
    const double* sg = &o.sigma[6 * (N + b)];
    if (o.unsUL()) {
        // unsNonLinGeomUpdatedLagSolid::tractionBoundarySnGrad, unsNonLinGeomUpdatedLagSolid.C:373-419: nCurrent = relJf relFinvf.T() & n
        // on the stress term only -- the pressure acts along n: ((t - n p) - (nCurrent & sigmaf) + (n & (impK gradDDf))) rImpK
        const double* gf = &o.gradDf[9 * (size_t)(o.F + b)];
        double rF[9]; transposeT(gf, rF); rF[0] += 1; rF[4] += 1; rF[8] += 1;
        const double rJ = detT(rF);
        double Fi[9], FiT[9]; invT(rF, Fi); transposeT(Fi, FiT);
        double nc[3];
        for (int i = 0; i < 3; i++) nc[i] = rJ * (FiT[3 * i] * n[0] + FiT[3 * i + 1] * n[1] + FiT[3 * i + 2] * n[2]);
        double ns[3]; SvS(&o.sigmaf[6 * (size_t)(o.F + b)], nc, ns);
        double ng[3]; vT(n, gf, ng);
        for (int i = 0; i < 3; i++) g[i] = ((t[i] - n[i] * p) - ns[i] + impK * ng[i]) * rImpK;
    } else if (o.unsTL()) {
        // unsNonLinGeomTotalLagSolid::tractionBoundarySnGrad, unsNonLinGeomTotalLagSolid.C:420-488 (enforceLinear off):
        // nCurrent = Jf Finvf.T() & n (not normalised); ((t - nCurrent p) - (nCurrent & sigmaf) + (n & (impK gradDf))) rImpK
        const double* gf = &o.gradDf[9 * (size_t)(o.F + b)];
        double Ff[9]; transposeT(gf, Ff); Ff[0] += 1; Ff[4] += 1; Ff[8] += 1;
        const double J = detT(Ff);
        double Fi[9], FiT[9]; invT(Ff, Fi); transposeT(Fi, FiT);
        double nc[3];
        for (int i = 0; i < 3; i++) nc[i] = J * (FiT[3 * i] * n[0] + FiT[3 * i + 1] * n[1] + FiT[3 * i + 2] * n[2]);
        double ns[3]; SvS(&o.sigmaf[6 * (size_t)(o.F + b)], nc, ns);
        double ng[3]; vT(n, gf, ng);
        for (int i = 0; i < 3; i++) g[i] = ((t[i] - nc[i] * p) - ns[i] + impK * ng[i]) * rImpK;
    } else if (o.uns()) {
        // unsLinGeomSolid::tractionBoundarySnGrad, unsLinGeomSolid.C:193-230: the same expression on the FACE fields
        // sigmaf_ and gradDf_ of the patch
        const double* gf = &o.gradDf[9 * (size_t)(o.F + b)];
        double M[9]; S2T(&o.sigmaf[6 * (size_t)(o.F + b)], M);
        for (int i = 0; i < 9; i++) M[i] -= impK * gf[i];
        double nM[3]; vT(n, M, nM);
        for (int i = 0; i < 3; i++) g[i] = ((t[i] - n[i] * p) - nM[i]) * rImpK;
    } else if (o.ctl.solidModel == S4F_MODEL_LIN_GEOM_TOTAL_DISP) {
        // ((traction - n*pressure) - (n & (pSigma - impK*pGradD)))*rImpK
        double M[9]; S2T(sg, M);
        for (int i = 0; i < 9; i++) M[i] -= impK * gD[i];
        double nM[3]; vT(n, M, nM);
        for (int i = 0; i < 3; i++) g[i] = ((t[i] - n[i] * p) - nM[i]) * rImpK;
    } else {
        // nCurrent = Finv.T() & n / mag;  ((t - nCur*p) - (nCur & pSigma) + impK*(n & pGradD))*rImpK
        const double* Fi = &o.Finv[9 * (N + b)];
        double FiT[9]; transposeT(Fi, FiT);
        double nc[3] = {FiT[0] * n[0] + FiT[1] * n[1] + FiT[2] * n[2], FiT[3] * n[0] + FiT[4] * n[1] + FiT[5] * n[2],
                        FiT[6] * n[0] + FiT[7] * n[1] + FiT[8] * n[2]};
        double m = mag3(nc); for (int i = 0; i < 3; i++) nc[i] /= m;
        double ns[3]; SvS(sg, nc, ns);          // nCur & sigma (symmetric)
        double ng[3]; vT(n, gD, ng);
        for (int i = 0; i < 3; i++) g[i] = ((t[i] - nc[i] * p) - ns[i] + impK * ng[i]) * rImpK;
    }
}

// updateCoeffs() of every patch, as triggered by the fvMatrix constructor [OF-ext]:
//  solidTraction: gradient() = tractionBoundarySnGrad (relaxFac 1)  solidTractionFvPatchVectorField.C:384-392
//  fixedDisplacement: value = totalDisp                              fixedDisplacementFvPatchVectorField.C:258-294
void bcUpdateCoeffs(s4f_oracle& o) {
    for (int p = 0; p < o.nPatches; p++) for (int i = 0; i < o.pSize[p]; i++) {
        int b = o.pStart[p] + i;
        if (o.bcKind[p] == S4F_BC_SOLID_TRACTION) tractionSnGrad(o, b, &o.tracGrad[3 * b]);
        else if (o.bcKind[p] == S4F_BC_FIXED_DISPLACEMENT) {
            for (int c = 0; c < 3; c++) {
                double v = o.bcValue[3 * b + c];
                if (o.incremental())
                    v -= o.Dold[3 * (o.N + b) + c];        // DD field: disp -= Dold  (:279-287)
                o.D[3 * (o.N + b) + c] = v;
            }
        }
    }
}

// snGrad() of patch face b.  gradRef = the registered "grad(D)" field at the time of the call.
//  solidTraction (fixedGradient): gradient()
//  fixedDisplacement: (D_b - (D_P + k & gradD_P))*deltaCoeffs         fixedDisplacement...C:297-326
//  solidSymmetry: (transform(I-2nn, DP) - DP)*deltaCoeffs/2           solidSymmetry...C:148-196
void bcSnGrad(const s4f_oracle& o, int p, int b, const dvec& gradRef, double* sn) {
    const int N = o.N, P = o.faceCells[b];
    double n[3], k[3], delta; patchGeom(o, b, n, k, delta);
    if (o.bcKind[p] == S4F_BC_SOLID_TRACTION) { for (int i = 0; i < 3; i++) sn[i] = o.tracGrad[3 * b + i]; return; }
    double kg[3]; vT(k, &gradRef[9 * P], kg);
    double DP[3]; for (int i = 0; i < 3; i++) DP[i] = o.D[3 * P + i] + kg[i];
    if (o.bcKind[p] == S4F_BC_FIXED_DISPLACEMENT) {
        for (int i = 0; i < 3; i++) sn[i] = (o.D[3 * (N + b) + i] - DP[i]) * delta;
    } else {   // symmetry
        double nDP = dot3(n, DP);
        for (int i = 0; i < 3; i++) sn[i] = ((DP[i] - 2.0 * n[i] * nDP) - DP[i]) * (delta / 2.0);
    }
}

// evaluate() of every patch (D.correctBoundaryConditions()).  Uses the registered "grad(D)".
//  solidTraction: D_b = D_P + (k & gradD_P) + gradient()/deltaCoeffs          solidTraction...C:398-463
//  solidSymmetry: D_b = (DP + transform(I-2nn, DP))/2                         solidSymmetry...C:200-260
//  fixedDisplacement: value stays
void bcEvaluate(s4f_oracle& o) {
    const int N = o.N;
    for (int p = 0; p < o.nPatches; p++) for (int i = 0; i < o.pSize[p]; i++) {
        int b = o.pStart[p] + i, P = o.faceCells[b];
        double n[3], k[3], delta; patchGeom(o, b, n, k, delta);
        double kg[3]; vT(k, &o.gradD[9 * P], kg);
        if (o.bcKind[p] == S4F_BC_SOLID_TRACTION) {
            for (int c = 0; c < 3; c++) o.D[3 * (N + b) + c] = o.D[3 * P + c] + kg[c] + o.tracGrad[3 * b + c] / delta;
        } else if (o.bcKind[p] == S4F_BC_SOLID_SYMMETRY) {
            double DP[3]; for (int c = 0; c < 3; c++) DP[c] = o.D[3 * P + c] + kg[c];
            double nDP = dot3(n, DP);
            for (int c = 0; c < 3; c++) o.D[3 * (N + b) + c] = (DP[c] + (DP[c] - 2.0 * n[c] * nDP)) / 2.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fvc::grad(D): least squares (NUM/extendedLeastSquaresGrad/extendedLeastSquaresGrad.C:103-167) or
// [OF-ext] Gauss linear; then gaussGrad::correctBoundaryConditions: grad_b = grad_P + n (snGrad_b - n & grad_P).
// mechanicalModel::grad, mechanicalModel.C:571-582.
// ------------------------------------------------------------------------------------------------
// [OF-ext] LeastSquaresVectors<centredCPCCellToCellStencilObject>::calcLeastSquaresVectors ("pointCellsLeastSquares"):
// stencil of cell i = the cells sharing a point with i and the boundary faces (non-empty, non-coupled) at i's points
// (CPCCellToCellStencil); d_j = x_j - C_i (boundary faces: face centre); dd = dd0 + sum d d/|d|^2, dd0 = unit entries in the
// empty directions; ls_j = (inv(dd) - dd0) & d_j/|d|^2; grad_i = sum_j ls_j (phi_j - phi_i).   Restated from the OpenFOAM
// library (not vendored): consistent by construction (exact for linear fields), the weights are from the published source.
void makePointCellsStencil(s4f_oracle& o) {
    const int N = o.N, F = o.F, B = o.B;
    std::vector<std::vector<int>> cellPts(N);
    auto add = [](std::vector<int>& v, int x) { if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x); };
    for (int f = 0; f < F + B; f++) for (int j = o.fvPtr[f]; j < o.fvPtr[f + 1]; j++) {
        add(cellPts[f < F ? o.own[f] : o.faceCells[f - F]], o.fv[j]);
        if (f < F) add(cellPts[o.nei[f]], o.fv[j]);
    }
    o.psPtr.assign(1, 0); o.psSlot.clear(); o.psLs.clear();
    for (int i = 0; i < N; i++) {
        std::vector<int> st;
        for (int p : cellPts[i]) {
            for (int j = o.pcPtr[p]; j < o.pcPtr[p + 1]; j++) if (o.pcCells[j] != i) add(st, o.pcCells[j]);
            for (int j = o.pbPtr[p]; j < o.pbPtr[p + 1]; j++) add(st, N + o.pbFaces[j]);
        }
        std::sort(st.begin(), st.end());
        double dd[6] = {0, 0, 0, 0, 0, 0};
        if (!o.solD[0]) dd[0] += 1; if (!o.solD[1]) dd[3] += 1; if (!o.solD[2]) dd[5] += 1;
        std::vector<double> dl(3 * st.size());
        for (size_t k = 0; k < st.size(); k++) {
            const double* x = st[k] < N ? &o.C[3 * (size_t)st[k]] : &o.Cf[3 * (size_t)(F + st[k] - N)];
            double d[3] = {x[0] - o.C[3 * (size_t)i], x[1] - o.C[3 * (size_t)i + 1], x[2] - o.C[3 * (size_t)i + 2]};
            const double r = 1.0 / dot3(d, d);
            dd[0] += r * d[0] * d[0]; dd[1] += r * d[0] * d[1]; dd[2] += r * d[0] * d[2];
            dd[3] += r * d[1] * d[1]; dd[4] += r * d[1] * d[2]; dd[5] += r * d[2] * d[2];
            for (int q = 0; q < 3; q++) dl[3 * k + q] = r * d[q];
        }
        double iv[6]; invS(dd, iv);
        if (!o.solD[0]) iv[0] -= 1; if (!o.solD[1]) iv[3] -= 1; if (!o.solD[2]) iv[5] -= 1;
        for (size_t k = 0; k < st.size(); k++) {
            double a[3]; SvS(iv, &dl[3 * k], a);
            o.psSlot.push_back(st[k]);
            for (int q = 0; q < 3; q++) o.psLs.push_back(a[q]);
        }
        o.psPtr.push_back((int)o.psSlot.size());
    }
    o.psValid = true;
}

// cell values of fvc::grad(X) for a vol field X given as [internal | boundary] values
void gradInterior(const s4f_oracle& oc, const dvec& X, dvec& g) {
    s4f_oracle& o = const_cast<s4f_oracle&>(oc);
    const int N = o.N, F = o.F, B = o.B;
    g.assign(9 * (size_t)(N + B), 0.0);
    if (o.ctl.gradScheme == S4F_GRAD_POINT_CELLS_LEAST_SQUARES) {
        if (!o.psValid) makePointCellsStencil(o);
        for (int i = 0; i < N; i++)
            for (int k = o.psPtr[i]; k < o.psPtr[i + 1]; k++) {
                const int s = o.psSlot[k];
                for (int a = 0; a < 3; a++) for (int j = 0; j < 3; j++) g[9 * (size_t)i + 3 * a + j] += o.psLs[3 * (size_t)k + a] * (X[3 * (size_t)s + j] - X[3 * (size_t)i + j]);
            }
    } else if (o.ctl.gradScheme == S4F_GRAD_LEAST_SQUARES) {
        forAllInternalFaces(o, [&](int f) {
            int P = o.own[f], Nn = o.nei[f];
            double dv[3] = {X[3 * Nn] - X[3 * P], X[3 * Nn + 1] - X[3 * P + 1], X[3 * Nn + 2] - X[3 * P + 2]};
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
                g[9 * P + 3 * i + j] += o.lsP[3 * f + i] * dv[j];
                g[9 * Nn + 3 * i + j] -= o.lsN[3 * f + i] * dv[j];
            }
        });
        for (int b = 0; b < B; b++) {
            int P = o.faceCells[b];
            double dv[3] = {X[3 * (N + b)] - X[3 * P], X[3 * (N + b) + 1] - X[3 * P + 1], X[3 * (N + b) + 2] - X[3 * P + 2]};
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) g[9 * P + 3 * i + j] += o.lsP[3 * (F + b) + i] * dv[j];
        }
    } else {
        forAllInternalFaces(o, [&](int f) {
            int P = o.own[f], Nn = o.nei[f];
            double wf = o.w[f];
            for (int j = 0; j < 3; j++) {
                double vf = wf * X[3 * P + j] + (1 - wf) * X[3 * Nn + j];
                for (int i = 0; i < 3; i++) {
                    double t = o.Sf[3 * f + i] * vf;
                    g[9 * P + 3 * i + j] += t; g[9 * Nn + 3 * i + j] -= t;
                }
            }
        });
        for (int b = 0; b < B; b++) {
            int P = o.faceCells[b];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) g[9 * P + 3 * i + j] += o.Sf[3 * (F + b) + i] * X[3 * (N + b) + j];
        }
        for (int c = 0; c < N; c++) for (int q = 0; q < 9; q++) g[9 * c + q] /= o.V[c];
    }
}

void calcGrad(s4f_oracle& o) {
    const int N = o.N;
    dvec g; gradInterior(o, o.D, g);
    // boundary: extrapolate then correct the normal component with the BC's snGrad (which reads the
    // OLD registered grad(D) -- the assignment gradD = fvc::grad(D) happens after the evaluation)
    for (int p = 0; p < o.nPatches; p++) for (int i2 = 0; i2 < o.pSize[p]; i2++) {
        int b = o.pStart[p] + i2, P = o.faceCells[b];
        double n[3], k[3], delta; patchGeom(o, b, n, k, delta);
        double sn[3]; bcSnGrad(o, p, b, o.gradD, sn);
        double* gb = &g[9 * (N + b)];
        for (int q = 0; q < 9; q++) gb[q] = g[9 * P + q];
        double ng[3]; vT(n, gb, ng);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) gb[3 * i + j] += n[i] * (sn[j] - ng[j]);
    }
    o.gradD.swap(g);
}

// fvc::grad of a temporary with calculated patches, e.g. gradD() = fvc::grad(D().oldTime() + DD()) after the
// updated-Lagrangian loop (nonLinGeomUpdatedLagSolid.C:243): snGrad_b = deltaCoeffs (X_b - X_P) [OF-ext] fvPatchField::snGrad
void gradCalculated(const s4f_oracle& o, const dvec& X, dvec& g) {
    const int N = o.N, F = o.F;
    gradInterior(o, X, g);
    for (int b = 0; b < o.B; b++) {
        const int P = o.faceCells[b];
        double n[3], k[3], delta; patchGeom(o, b, n, k, delta);
        (void)F;
        double sn[3]; for (int j = 0; j < 3; j++) sn[j] = delta * (X[3 * (N + b) + j] - X[3 * P + j]);
        double* gb = &g[9 * (N + b)];
        for (int q = 0; q < 9; q++) gb[q] = g[9 * P + q];
        double ng[3]; vT(n, gb, ng);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) gb[3 * i + j] += n[i] * (sn[j] - ng[j]);
    }
}

// ------------------------------------------------------------------------------------------------
// unsLinGeomSolid: gradients built on the vertex displacements (NUM/fvc/fvcGradf.C)
// ------------------------------------------------------------------------------------------------
// fsGrad, fvcGradf.C:170-232 (internal faces) / :236-298 (patch faces) / :359-440 (patch fGrad): the in-plane gradient of a
// face from the edge-centre values, grad = (1/|Sf|) sum_edges Le fe, Le = (e - n (n & e)) ^ n, fe = (pf_start + pf_end)/2
void unsFaceTangentialGrad(const s4f_oracle& o, int f, double* T) {
    for (int q = 0; q < 9; q++) T[q] = 0;
    const double mag = o.magSf[f];
    double n[3] = {o.Sf[3 * (size_t)f] / mag, o.Sf[3 * (size_t)f + 1] / mag, o.Sf[3 * (size_t)f + 2] / mag};
    const int a = o.fvPtr[f], m = o.fvPtr[f + 1] - a;
    for (int i = 0; i < m; i++) {
        const int v0 = o.fv[a + i], v1 = o.fv[a + (i + 1) % m];
        const double* p0 = &o.points[3 * (size_t)v0]; const double* p1 = &o.points[3 * (size_t)v1];
        double e[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
        const double ne = dot3(n, e);
        for (int q = 0; q < 3; q++) e[q] -= n[q] * ne;
        const double Le[3] = {e[1] * n[2] - e[2] * n[1], e[2] * n[0] - e[0] * n[2], e[0] * n[1] - e[1] * n[0]};
        for (int j = 0; j < 3; j++) {
            const double fe = 0.5 * (o.pointD[3 * (size_t)v0 + j] + o.pointD[3 * (size_t)v1 + j]);
            for (int i2 = 0; i2 < 3; i2++) T[3 * i2 + j] += Le[i2] * fe;
        }
    }
    for (int q = 0; q < 9; q++) T[q] /= mag;
}

// one face's contribution to fvc::grad(vf, pf), fvcGradf.C:497-566 (:603-668 on patches): triangles about the vertex average,
// G = sum_t St ttcf (ttcf = (pf_a + pf_b + cf)/3), Vf = sum_t St & Ct; a triangular face is taken whole
void unsFaceGauss(const s4f_oracle& o, int f, double* G, double& Vf) {
    for (int q = 0; q < 9; q++) G[q] = 0;
    Vf = 0;
    const int a = o.fvPtr[f], m = o.fvPtr[f + 1] - a;
    auto P = [&](int i) { return &o.points[3 * (size_t)o.fv[a + i]]; };
    auto U = [&](int i) { return &o.pointD[3 * (size_t)o.fv[a + i]]; };
    if (m == 3) {
        const double e1[3] = {P(1)[0] - P(0)[0], P(1)[1] - P(0)[1], P(1)[2] - P(0)[2]}, e2[3] = {P(2)[0] - P(0)[0], P(2)[1] - P(0)[1], P(2)[2] - P(0)[2]};
        const double S[3] = {0.5 * (e1[1] * e2[2] - e1[2] * e2[1]), 0.5 * (e1[2] * e2[0] - e1[0] * e2[2]), 0.5 * (e1[0] * e2[1] - e1[1] * e2[0])};
        double c[3], u[3];
        for (int q = 0; q < 3; q++) { c[q] = (P(0)[q] + P(1)[q] + P(2)[q]) / 3.0; u[q] = (U(0)[q] + U(1)[q] + U(2)[q]) / 3.0; }
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) G[3 * i + j] = S[i] * u[j];
        Vf = dot3(S, c);
        return;
    }
    double cp[3] = {0, 0, 0}, cf[3] = {0, 0, 0};
    for (int i = 0; i < m; i++) for (int q = 0; q < 3; q++) { cp[q] += P(i)[q]; cf[q] += U(i)[q]; }
    for (int q = 0; q < 3; q++) { cp[q] /= m; cf[q] /= m; }
    for (int i = 0; i < m; i++) {
        const double* pa = P(i); const double* pb = P((i + 1) % m);
        const double* ua = U(i); const double* ub = U((i + 1) % m);
        const double ra[3] = {pa[0] - cp[0], pa[1] - cp[1], pa[2] - cp[2]}, rb[3] = {pb[0] - cp[0], pb[1] - cp[1], pb[2] - cp[2]};
        const double St[3] = {0.5 * (ra[1] * rb[2] - ra[2] * rb[1]), 0.5 * (ra[2] * rb[0] - ra[0] * rb[2]), 0.5 * (ra[0] * rb[1] - ra[1] * rb[0])};
        double Ct[3], tt[3];
        for (int q = 0; q < 3; q++) { Ct[q] = (cp[q] + pa[q] + pb[q]) / 3.0; tt[q] = (ua[q] + ub[q] + cf[q]) / 3.0; }
        for (int i2 = 0; i2 < 3; i2++) for (int j = 0; j < 3; j++) G[3 * i2 + j] += St[i2] * tt[j];
        Vf += dot3(St, Ct);
    }
}

void gradInterior(const s4f_oracle& oc, const dvec& X, dvec& g);

// mechanical().interpolate(D, pointD, false); mechanical().grad(D, pointD, gradD, gradDf)   (unsLinGeomSolid.C:146-150,
// mechanicalModel.C:722-731): gradD = fvc::grad(D, pointD) (fvcGradf.C:442-800), gradDf = fvc::fGrad(D, pointD) =
// fsGrad + n*fvc::snGrad(D) (fvcGradf.C:104-107; snGrad(D) corrected: the non-orthogonal part takes fvc::grad(D) of the gradScheme)
void unsUpdateGradients(s4f_oracle& o, bool interpolate = true) {
    const int N = o.N, F = o.F, B = o.B;
    if (interpolate) {
        o.pointD.assign(3 * (size_t)o.nPoints, 0.0);
        s4fo_interpolate_to_points_impl(&o, o.D, o.pointD.data());
    }
    // ---- gradD = fvc::grad(D, pointD)
    dvec g(9 * (size_t)(N + B), 0.0), V3(N, 0.0);
    for (int f = 0; f < F + B; f++) {
        double G[9], Vf; unsFaceGauss(o, f, G, Vf);
        const int P = f < F ? o.own[f] : o.faceCells[f - F];
        for (int q = 0; q < 9; q++) g[9 * (size_t)P + q] += G[q];
        V3[P] += Vf;
        if (f < F) { for (int q = 0; q < 9; q++) g[9 * (size_t)o.nei[f] + q] -= G[q]; V3[o.nei[f]] -= Vf; }
    }
    // the faces of empty patches are not mirrored (2-D cases): they are planar with normals along the empty direction, so
    // their share of sum St & Ct is the cell volume (prismatic cells) and their share of the gradient sum cancels front to back
    const int nEmpty = (o.solD[0] ? 0 : 1) + (o.solD[1] ? 0 : 1) + (o.solD[2] ? 0 : 1);
    for (int c = 0; c < N; c++) V3[c] += nEmpty * o.V[c];
    for (int c = 0; c < N; c++) for (int q = 0; q < 9; q++) g[9 * (size_t)c + q] /= (V3[c] / 3.0);
    // patches: the in-plane gradient of the patch face (fGrad(polyPatch, ppf) :693-718), then the normal gradient of the
    // boundary condition (:767-781; its snGrad() reads the registered grad(D), still the previous one)
    for (int p = 0; p < o.nPatches; p++) for (int i = 0; i < o.pSize[p]; i++) {
        const int b = o.pStart[p] + i;
        double n[3], k[3], delta; patchGeom(o, b, n, k, delta);
        double* gb = &g[9 * (size_t)(N + b)];
        unsFaceTangentialGrad(o, F + b, gb);
        double sn[3]; bcSnGrad(o, p, b, o.gradD, sn);
        double ng[3]; vT(n, gb, ng);
        for (int a = 0; a < 3; a++) for (int j = 0; j < 3; j++) gb[3 * a + j] += n[a] * (sn[j] - ng[j]);
    }
    o.gradD.swap(g);
    // ---- gradDf = fsGrad(D, pointD) + n*snGrad(D)
    dvec gLS;
    if (o.nonOrth) gradInterior(o, o.D, gLS);
    o.gradDf.assign(9 * (size_t)(F + B), 0.0);
    for (int f = 0; f < F; f++) {
        double* T = &o.gradDf[9 * (size_t)f];
        unsFaceTangentialGrad(o, f, T);
        const int P = o.own[f], Nn = o.nei[f];
        const double mag = o.magSf[f];
        double sn[3];
        for (int j = 0; j < 3; j++) sn[j] = o.nod[f] * (o.D[3 * (size_t)Nn + j] - o.D[3 * (size_t)P + j]);
        if (o.nonOrth) {
            double gf[9]; for (int q = 0; q < 9; q++) gf[q] = o.w[f] * gLS[9 * (size_t)P + q] + (1 - o.w[f]) * gLS[9 * (size_t)Nn + q];
            double cg[3]; vT(&o.corr[3 * (size_t)f], gf, cg);
            for (int j = 0; j < 3; j++) sn[j] += cg[j];
        }
        for (int a = 0; a < 3; a++) for (int j = 0; j < 3; j++) T[3 * a + j] += (o.Sf[3 * (size_t)f + a] / mag) * sn[j];
    }
    for (int p = 0; p < o.nPatches; p++) for (int i = 0; i < o.pSize[p]; i++) {
        const int b = o.pStart[p] + i;
        double n[3], k[3], delta; patchGeom(o, b, n, k, delta);
        double* T = &o.gradDf[9 * (size_t)(F + b)];
        unsFaceTangentialGrad(o, F + b, T);
        double sn[3]; bcSnGrad(o, p, b, o.gradD, sn);        // the patch snGrad() with the gradD just assigned
        for (int a = 0; a < 3; a++) for (int j = 0; j < 3; j++) T[3 * a + j] += n[a] * sn[j];
    }
}

// linearElastic::correct(surfaceSymmTensorField&), linearElastic.C:342-370: sigmaf = 2 mu epsilonf + lambda tr(epsilonf) I + sigma0f
void unsLawFaces(s4f_oracle& o) {
    const int nf = o.F + o.B;
    o.sigmaf.assign(6 * (size_t)nf, 0.0);
    if (o.unsFinite()) {
        // unsNonLinGeomTotalLagSolid.C:312-330: Ff = I + gradDf.T(); then neoHookeanElastic::correct(surfaceSymmTensorField&)
        // -> correctF (neoHookeanElastic.C:306-352): J = det F; bEbar = J^(-2/3) symm(F & F.T()); s = mu dev(bEbar);
        // sigma = (1/J) (0.5 K (J^2 - 1) I + s).
        // unsNonLinGeomUpdatedLagSolid.C:283-301: relFf = I + gradDDf.T(); Ff = relFf & Ff.oldTime() (the field the law's
        // updateF(surface) forms as well, mechanicalLaw.C updated-Lagrangian branch)
        if ((int)o.FfOld.size() != 9 * nf) { o.FfOld.assign(9 * (size_t)nf, 0.0); for (int f = 0; f < nf; f++) for (int d = 0; d < 3; d++) o.FfOld[9 * (size_t)f + 4 * d] = 1; }
        o.Ff.resize(9 * (size_t)nf);
        for (int f = 0; f < nf; f++) {
            double rF[9]; transposeT(&o.gradDf[9 * (size_t)f], rF); rF[0] += 1; rF[4] += 1; rF[8] += 1;
            double Ff[9];
            if (o.unsUL()) mulTT(rF, &o.FfOld[9 * (size_t)f], Ff); else for (int q = 0; q < 9; q++) Ff[q] = rF[q];
            for (int q = 0; q < 9; q++) o.Ff[9 * (size_t)f + q] = Ff[q];
            const double J = detT(Ff);
            double FT[9], FFT[9], b[6];
            transposeT(Ff, FT); mulTT(Ff, FT, FFT); symm(FFT, b);
            const double sc = std::pow(J, -2.0 / 3.0);
            for (int q = 0; q < 6; q++) b[q] *= sc;
            double dv[6]; devS(b, dv);
            const double sh = 0.5 * o.law.K * (std::pow(J, 2.0) - 1.0);
            double* sg = &o.sigmaf[6 * (size_t)f];
            for (int q = 0; q < 6; q++) sg[q] = o.law.mu * dv[q];
            sg[0] += sh; sg[3] += sh; sg[5] += sh;
            for (int q = 0; q < 6; q++) sg[q] *= 1.0 / J;
        }
        return;
    }
    for (int f = 0; f < nf; f++) {
        double e[6]; symm(&o.gradDf[9 * (size_t)f], e);
        const double tr = trS(e);
        double* s = &o.sigmaf[6 * (size_t)f];
        for (int q = 0; q < 6; q++) s[q] = 2.0 * o.law.mu * e[q] + o.law.sigma0[q];
        s[0] += o.law.lambda * tr; s[3] += o.law.lambda * tr; s[5] += o.law.lambda * tr;
    }
}

// ------------------------------------------------------------------------------------------------
// Mechanical laws (cells and boundary faces alike: OpenFOAM field algebra acts on both)
// ------------------------------------------------------------------------------------------------

// linearElastic::correct, ML/linearGeometryLaws/linearElastic/linearElastic.C:318-339;
// updateEpsilon ML/mechanicalLaw/mechanicalLaw.C:983-1001; updateSigmaHyd explicit branch :1469-1475
void lawLinearElastic(s4f_oracle& o) {
    const int n = o.NB();
    const double mu = o.law.mu, K = o.law.K;
    S4FO_PAR_FOR
    for (int c = 0; c < n; c++) {
        double e[6]; symm(&o.gradForLaw()[9 * c], e);
        for (int q = 0; q < 6; q++) o.epsilon[6 * c + q] = e[q];
        double sh = K * trS(e);
        o.sigmaHyd[c] = sh;
        double de[6]; devS(e, de);
        double* s = &o.sigma[6 * c];
        for (int q = 0; q < 6; q++) s[q] = 2.0 * mu * de[q] + o.law.sigma0[q];
        s[0] += sh; s[3] += sh; s[5] += sh;
    }
}

// mechanicalLaw::updateF, TL total displacement branch: F = I + gradD.T(); relF = F & inv(F.old)
// ML/mechanicalLaw/mechanicalLaw.C:1130-1140.  The incremental TL branch (:1097-1108) forms F = F.old + gradDD.T(),
// which is the same tensor up to round-off because gradD = gradD.old + gradDD; the total form is used for both.
void lawUpdateF(s4f_oracle& o) {
    const int n = o.NB();
    const dvec& gD = o.gradForLaw();
    if (o.UL()) {          // updated Lagrangian :1055-1072: relF = I + gradDD.T(); F = relF & F.oldTime()
        for (int c = 0; c < n; c++) {
            double* rF = &o.relF[9 * c];
            transposeT(&o.gradD[9 * c], rF);
            rF[0] += 1; rF[4] += 1; rF[8] += 1;
            mulTT(rF, &o.lawFold[9 * c], &o.lawF[9 * c]);
        }
        return;
    }
    for (int c = 0; c < n; c++) {
        double Ft[9]; transposeT(&gD[9 * c], Ft);
        Ft[0] += 1; Ft[4] += 1; Ft[8] += 1;
        for (int q = 0; q < 9; q++) o.lawF[9 * c + q] = Ft[q];
        double Fi[9]; invT(&o.lawFold[9 * c], Fi);
        mulTT(Ft, Fi, &o.relF[9 * c]);
    }
}

// neoHookeanElastic::correct, ML/nonLinearGeometryLaws/neoHookeanElastic/neoHookeanElastic.C:275-303
void lawNeoHookean(s4f_oracle& o) {
    lawUpdateF(o);
    const int n = o.NB();
    const double mu = o.law.mu, K = o.law.K;
    for (int c = 0; c < n; c++) {
        const double* Fm = &o.lawF[9 * c];
        double J = detT(Fm);
        double FT[9], FFT[9], b[6];
        transposeT(Fm, FT); mulTT(Fm, FT, FFT); symm(FFT, b);
        double sc = std::pow(J, -2.0 / 3.0);
        for (int q = 0; q < 6; q++) b[q] *= sc;
        double s[6]; devS(b, s);
        for (int q = 0; q < 6; q++) s[q] *= mu;
        double sh = 0.5 * K * (std::pow(J, 2.0) - 1.0);
        o.sigmaHyd[c] = sh; o.lawJ[c] = J;
        double* sg = &o.sigma[6 * c];
        for (int q = 0; q < 6; q++) sg[q] = s[q];
        sg[0] += sh; sg[3] += sh; sg[5] += sh;
        for (int q = 0; q < 6; q++) sg[q] *= (1.0 / J);
    }
}

// neoHookeanElasticMisesPlastic helpers: curYieldStress :139-150, yieldFunction :153-183,
// newtonLoop :186-247 (LoopTol 1e-8, MaxNewtonIter 200, finiteDiff 0.25e-6: :41-47)
const double sqrtTwoOverThree = std::sqrt(2.0 / 3.0);
inline double curYieldStress(const s4fgpu_law& L, double epsPEq, double J) { return J * tableLookup(L, std::max(epsPEq, SMALL)); }
inline double yieldFunction(const s4fgpu_law& L, double epsOld, double magSTrial, double DLambda, double muBar, double J) {
    return magSTrial - 2 * muBar * DLambda - sqrtTwoOverThree * curYieldStress(L, epsOld + sqrtTwoOverThree * DLambda, J);
}
void newtonLoop(const s4fgpu_law& L, double& DLambda, double& curSigmaY, double epsOld, double magSTrial, double muBar,
                double J, double maxMagDEpsilon) {
    const double LoopTol = 1e-8, finiteDiff = 0.25e-6; const int MaxNewtonIter = 200;
    int i = 0;
    double fTrial = yieldFunction(L, epsOld, magSTrial, DLambda, muBar, J);
    double residual = 1.0;
    do {
        double fStep = yieldFunction(L, epsOld, magSTrial, DLambda + finiteDiff, muBar, J);
        double deriv = (fStep - fTrial) / finiteDiff;
        residual = fTrial / deriv;
        DLambda -= residual;
        residual /= maxMagDEpsilon;
        fTrial = yieldFunction(L, epsOld, magSTrial, DLambda, muBar, J);
    } while ((std::fabs(residual) > LoopTol) && ++i < MaxNewtonIter);
    curSigmaY = curYieldStress(L, epsOld + sqrtTwoOverThree * DLambda, J) / J;
}
// Ibar: Rubin-Attia cubic for det(bEbar)=1, :250-395
inline double IbarOf(const double* devB) {
    double detd = detS(devB), dotp = magSqrS(devB), fac1 = 2.0 * dotp / 3.0, alpha1;
    if (std::fabs(fac1) < SMALL) alpha1 = 3.0;
    else {
        double fac2 = (4.0 * (1.0 - detd)) / std::pow(fac1, 1.5);
        if (fac2 >= 1.0) alpha1 = 3.0 * std::sqrt(fac1) * std::cosh(std::acosh(fac2) / 3.0);
        else alpha1 = 3.0 * std::sqrt(fac1) * std::cos(std::acos(fac2) / 3.0);
    }
    return alpha1 / 3.0;
}

// neoHookeanElasticMisesPlastic::correct, ...MisesPlastic.C:991-1223
void lawNeoHookeanMises(s4f_oracle& o) {
    lawUpdateF(o);
    const int N = o.N, n = o.NB();
    const double mu = o.law.mu, K = o.law.K;
    dvec sTrial(6 * n), IbarT(n), muBar(n), fTrial(n);
    o.DEpsPprev = o.DEpsP;                                   // DEpsilonP_.storePrevIter()
    double maxMagBE = 0;
    for (int c = 0; c < n; c++) {
        double J = detT(&o.lawF[9 * c]);
        o.lawJ[c] = J;
        double relJ = J / o.lawJold[c];
        double sc = std::pow(relJ, -1.0 / 3.0);
        double rFb[9]; for (int q = 0; q < 9; q++) rFb[q] = sc * o.relF[9 * c + q];
        double bo[9], t1[9], rT[9], t2[9];
        S2T(&o.bEbarOld[6 * c], bo); mulTT(rFb, bo, t1); transposeT(rFb, rT); mulTT(t1, rT, t2);
        double* bt = &o.bEbarTrial[6 * c];
        symm(t2, bt);                                          // transform(relFbar, bEbar.oldTime())
        double dv[6]; devS(bt, dv);
        for (int q = 0; q < 6; q++) sTrial[6 * c + q] = mu * dv[q];
        IbarT[c] = trS(bt) / 3.0; muBar[c] = IbarT[c] * mu;
        if (c < N) maxMagBE = std::max(maxMagBE, std::sqrt(magSqrS(bt)));   // gMax over the internal field :1030
        fTrial[c] = std::sqrt(magSqrS(&sTrial[6 * c])) - sqrtTwoOverThree * J * o.sigmaY[c];
    }
    maxMagBE = std::max(maxMagBE, SMALL);
    const bool nonLinearPlasticity = o.law.nTable > 2;
    const double magHp = std::fabs(o.Hp);
    for (int c = 0; c < n; c++) {                              // cells :1066-1118, boundary faces :1120-1193
        double magS = std::sqrt(magSqrS(&sTrial[6 * c]));
        if (magS > SMALL) for (int q = 0; q < 6; q++) o.plasticN[6 * c + q] = sTrial[6 * c + q] / magS;
        if (fTrial[c] < SMALL) { o.DSigmaY[c] = 0; o.DLambda[c] = 0; }
        else if (nonLinearPlasticity) {
            double curSigmaY = 0;
            newtonLoop(o.law, o.DLambda[c], curSigmaY, o.epsPEq[c], magS, muBar[c], o.lawJ[c], maxMagBE);
            o.DSigmaY[c] = curSigmaY - o.sigmaY[c];
        } else {
            o.DLambda[c] = fTrial[c] / (2 * muBar[c]);
            if (magHp > SMALL) { o.DLambda[c] /= 1.0 + o.Hp / (3 * muBar[c]); o.DSigmaY[c] = sqrtTwoOverThree * o.DLambda[c] * o.Hp; }
        }
    }
    const double relax = o.law.DEpsilonPRelax;
    for (int c = 0; c < n; c++) {
        o.DEpsPEq[c] = sqrtTwoOverThree * o.DLambda[c];
        double s[6], devB[6];
        for (int q = 0; q < 6; q++) {
            double v = IbarT[c] * o.DLambda[c] * o.plasticN[6 * c + q];
            // DEpsilonP_.relax(): prevIter + alpha*(new - prevIter)  [OF-ext] GeometricField::relax
            if (relax != 1.0) v = o.DEpsPprev[6 * c + q] + relax * (v - o.DEpsPprev[6 * c + q]);
            o.DEpsP[6 * c + q] = v;
            s[q] = sTrial[6 * c + q] - 2 * mu * v;
            devB[q] = s[q] / mu;
        }
        double Ib = o.law.updateBEbarConsistent ? IbarOf(devB) : IbarT[c];
        double* be = &o.bEbar[6 * c];
        for (int q = 0; q < 6; q++) be[q] = devB[q];
        be[0] += Ib; be[3] += Ib; be[5] += Ib;
        double J = o.lawJ[c];
        double sh = 0.5 * K * (std::pow(J, 2.0) - 1.0);
        o.sigmaHyd[c] = sh;
        double* sg = &o.sigma[6 * c];
        for (int q = 0; q < 6; q++) sg[q] = s[q];
        sg[0] += sh; sg[3] += sh; sg[5] += sh;
        for (int q = 0; q < 6; q++) sg[q] *= (1.0 / J);
    }
}

// linearElasticMisesPlastic::correct (small strain J2), ML/linearGeometryLaws/linearElasticMisesPlastic/
// linearElasticMisesPlastic.C:953-1078 with updatePlasticity :58-131, yieldFunction :144-171 and
// newtonLoop :174-232 (no J scaling, muBar = mu).  Total fields are rebuilt from their old-time
// values on every call (:1062-1066), so updateTotalFields has nothing to add for this law.
void lawLinearElasticMises(s4f_oracle& o) {
    const int N = o.N, n = o.NB();
    const double mu = o.law.mu, K = o.law.K;
    double maxMagBE = 0;
    for (int c = 0; c < n; c++) {
        symm(&o.gradForLaw()[9 * c], &o.epsilon[6 * c]);                  // updateEpsilon()
        if (c < N) maxMagBE = std::max(maxMagBE, std::sqrt(magSqrS(&o.epsilon[6 * c])));
    }
    maxMagBE = std::max(maxMagBE, SMALL);
    const bool nonLinearPlasticity = o.law.nTable > 2;
    o.DEpsPprev = o.DEpsP;                                                 // DEpsilonP_.storePrevIter()
    for (int c = 0; c < n; c++) {
        const double* eps = &o.epsilon[6 * c];
        double e[6], dpo[6], sT[6];
        devS(eps, e); devS(&o.epsPOld[6 * c], dpo);
        for (int q = 0; q < 6; q++) sT[q] = 2.0 * mu * (e[q] - dpo[q]);
        double fT = std::sqrt(magSqrS(sT)) - sqrtTwoOverThree * o.sigmaYOld[c];
        double* pn = &o.plasticN[6 * c];
        if (fT < SMALL) {
            pn[0] = 1; pn[1] = 0; pn[2] = 0; pn[3] = 1; pn[4] = 0; pn[5] = 1;
            o.DLambda[c] = 0; o.DSigmaY[c] = 0; o.sigmaY[c] = o.sigmaYOld[c];
        } else {
            double magS = std::sqrt(magSqrS(sT));
            if (magS > SMALL) for (int q = 0; q < 6; q++) pn[q] = sT[q] / magS;
            else { pn[0] = 1; pn[1] = 0; pn[2] = 0; pn[3] = 1; pn[4] = 0; pn[5] = 1; }
            if (nonLinearPlasticity) {
                // same Newton iteration as the finite-strain law with J = 1 (Kirchhoff == Cauchy)
                newtonLoop(o.law, o.DLambda[c], o.sigmaY[c], o.epsPEqOld[c], magS, mu, 1.0, maxMagBE);
                o.DSigmaY[c] = o.sigmaY[c] - o.sigmaYOld[c];
            } else {
                o.DLambda[c] = fT / (2 * mu);
                if (std::fabs(o.Hp) > SMALL) {
                    o.DLambda[c] /= 1.0 + o.Hp / (3 * mu);
                    o.DSigmaY[c] = sqrtTwoOverThree * o.DLambda[c] * o.Hp;
                    o.sigmaY[c] = o.sigmaYOld[c] + o.DSigmaY[c];
                }
            }
        }
        o.DEpsPEq[c] = sqrtTwoOverThree * o.DLambda[c];
        double s[6];
        for (int q = 0; q < 6; q++) {
            o.DEpsP[6 * c + q] = o.DLambda[c] * pn[q];
            o.epsP[6 * c + q] = o.epsPOld[6 * c + q] + o.DEpsP[6 * c + q];
            s[q] = sT[q] - 2 * mu * o.DEpsP[6 * c + q];
        }
        o.epsPEq[c] = o.epsPEqOld[c] + o.DEpsPEq[c];
        double sh = K * trS(eps);
        o.sigmaHyd[c] = sh;
        double* sg = &o.sigma[6 * c];
        for (int q = 0; q < 6; q++) sg[q] = s[q];
        sg[0] += sh; sg[3] += sh; sg[5] += sh;
    }
}

void lawCorrect(s4f_oracle& o) {
    dvec prevHyd; if (o.law.solvePressureEqn) prevHyd = o.sigmaHyd;       // sigmaHyd persists: initial guess of the pressure solve
    switch (o.law.kind) {
        case S4F_LAW_LINEAR_ELASTIC: lawLinearElastic(o); break;
        case S4F_LAW_NEO_HOOKEAN_ELASTIC: lawNeoHookean(o); break;
        case S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC: lawNeoHookeanMises(o); break;
        case S4F_LAW_LINEAR_ELASTIC_MISES_PLASTIC: lawLinearElasticMises(o); break;
    }
    if (o.law.solvePressureEqn) {
        // the laws above evaluated sigma with the explicit hydrostatic stress (held in sigmaHyd); updateSigmaHyd
        // (mechanicalLaw.C:1366-1468) replaces it by the solution of the pressure equation:
        //   linearElastic.C:337-340   sigma = 2 mu dev(eps) + sigmaHyd I + sigma0
        //   neoHookeanElastic.C:295-302, neoHookeanElasticMisesPlastic.C:1215-1222   sigma = (sigmaHyd I + s)/J
        const bool lin = (o.law.kind == S4F_LAW_LINEAR_ELASTIC || o.law.kind == S4F_LAW_LINEAR_ELASTIC_MISES_PLASTIC);
        const int n = o.NB();
        o.sigmaHydExp = o.sigmaHyd;
        o.sigmaHyd = prevHyd;
        updateSigmaHydSmoothed(o, o.impK[0]);
        for (int c = 0; c < n; c++) {
            const double d = (o.sigmaHyd[c] - o.sigmaHydExp[c]) / (lin ? 1.0 : o.lawJ[c]);
            o.sigma[6 * c] += d; o.sigma[6 * c + 3] += d; o.sigma[6 * c + 5] += d;
        }
    }
}

// residual(): neoHookeanElasticMisesPlastic.C:1468-1523 (internal field only); elastic laws: 0
double lawResidual(const s4f_oracle& o) {
    if (o.law.kind != S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC && o.law.kind != S4F_LAW_LINEAR_ELASTIC_MISES_PLASTIC) return 0.0;
    double num = 0, den = 0;
    for (int c = 0; c < o.N; c++) {
        double d[6]; for (int q = 0; q < 6; q++) d[q] = o.DEpsP[6 * c + q] - o.DEpsPprev[6 * c + q];
        num = std::max(num, std::sqrt(magSqrS(d)));
        den = std::max(den, SMALL + std::sqrt(magSqrS(&o.DEpsPprev[6 * c])));
    }
    return num / den;
}

// solver-level kinematics of the TL models: F = I + gradD.T(); Finv = inv(F); J = det(F)
// SM/nonLinGeomTotalLagTotalDispSolid/nonLinGeomTotalLagTotalDispSolid.C:225-232
void updateKinematics(s4f_oracle& o) {
    const int n = o.NB();
    // updated Lagrangian: the relative kinematics relF = I + gradDD.T(), relFinv, relJ take the place of F, Finv, J in
    // fvc::div(relJ*relFinv & sigma) and in the traction boundary (nonLinGeomUpdatedLagSolid.C:188, :199-215, :283-288)
    const dvec& gD = o.UL() ? o.gradD : o.gradForLaw();
    for (int c = 0; c < n; c++) {
        double* Fm = &o.Ft[9 * c];
        transposeT(&gD[9 * c], Fm); Fm[0] += 1; Fm[4] += 1; Fm[8] += 1;
        invT(Fm, &o.Finv[9 * c]);
        o.Jt[c] = detT(Fm);
    }
}


// ------------------------------------------------------------------------------------------------
// rho*fvm::d2dt2(D) as "diag coefficient * V" and "V * (h1 D.o + h2 D.oo + h3 D.ooo + h4 D.oooo)".
//  Euler   [OF-ext] EulerD2dt2Scheme::fvmD2dt2 (variable deltaT form)
//  backward  numerics/backwardD2dt2Scheme/backwardD2dt2Scheme.C:309-395: fvm = coefft*rho*rDeltaT*
//    backwardDdt.fvmDdt(D); source += rDeltaT*rho*V*(coefft0*backwardDdt.fvcDdt(D.o) - coefft00*
//    backwardDdt.fvcDdt(D.oo)), coefficients :350-356 with deltaT0_(vf) = GREAT while D.o and D.oo carry the
//    same time index (:48-68), i.e. during the first time step; constant deltaT only (:316-322).
//    [OF-ext] backwardDdtScheme: fvmDdt diag = b rDeltaT V, source = rDeltaT V (b0 D.o - b00 D.oo);
//    fvcDdt(X) = rDeltaT (b X - b0 X.o + b00 X.oo), b = 1.5, b0 = 2, b00 = 0.5 at constant deltaT (D is
//    constructed with two old-time levels, solidModel.C:1247).  The levels D.ooo and D.oooo that the nested
//    fvcDdt calls reach are created by GeometricField::oldTime() as copies of D.oo at the first evaluation.
// ------------------------------------------------------------------------------------------------
struct D2dt2Coeffs { double diag, h[4]; };
D2dt2Coeffs d2dt2Coeffs(const s4fgpu_controls& ctl, double rho, int timeIndex) {
    D2dt2Coeffs k{0, {0, 0, 0, 0}};
    const double dt = ctl.deltaT, dt0 = ctl.deltaT0 > 0 ? ctl.deltaT0 : dt;
    if (ctl.d2dt2Scheme == S4F_D2DT2_EULER) {
        const double coefft = (dt + dt0) / (2 * dt), coefft00 = (dt + dt0) / (2 * dt0), rDeltaT2 = 4.0 / ((dt + dt0) * (dt + dt0));
        k.diag = coefft * rDeltaT2 * rho;
        k.h[0] = rDeltaT2 * rho * (coefft + coefft00); k.h[1] = -rDeltaT2 * rho * coefft00;
    } else if (ctl.d2dt2Scheme == S4F_D2DT2_BACKWARD) {
        const bool first = timeIndex <= 1;            // deltaT0_(vf) == GREAT
        const double c = first ? 1.0 : 1.0 + dt / (dt + dt0), c00 = first ? 0.0 : dt * dt / (dt0 * (dt + dt0)), c0 = c + c00;
        const double b = 1.0 + dt / (dt + dt0), b00 = dt * dt / (dt0 * (dt + dt0)), b0 = b + b00;
        const double r = rho / (dt * dt);
        k.diag = r * c * b;
        k.h[0] = r * (c * b0 + c0 * b);
        k.h[1] = r * (-c * b00 - c0 * b0 - c00 * b);
        k.h[2] = r * (c0 * b00 + c00 * b0);
        k.h[3] = -r * c00 * b00;
    }
    return k;
}

// Updated-Lagrangian inertia, SM/nonLinGeomUpdatedLagSolid/nonLinGeomUpdatedLagSolid.C:173-176:
//     fvm::d2dt2(rho_, DD()) + fvc::d2dt2(rho_, D().oldTime())
// with the density FIELD rho_ (rho_ = rho_.oldTime()/relJ_ in updateTotalFields :362) on the mesh of the current
// (updated) configuration, mesh.moving() == false.  Returned: the diagonal coefficient (to be multiplied by
// rho_P V_P) and, per cell, the explicit part per unit volume (to be multiplied by V_P) INCLUDING rho_P*g.
//  backward  NUM/backwardD2dt2Scheme/backwardD2dt2Scheme.C:149-222 (fvc, rho field) and :391-470 (fvm, rho field):
//      fvm = c rho rDeltaT fvmDdt(DD);  source += rDeltaT V (c0 rho.o ddt(DD.o) - c00 rho.oo ddt(DD.oo))
//      fvc = rDeltaT (c rho ddt(Y) - c0 rho.o ddt(Y.o) + c00 rho.oo ddt(Y.oo)),  Y = D.oldTime()
//      ddt(X) = rDeltaT (b X - b0 X.o + b00 X.oo)  [OF-ext] backwardDdtScheme, b = 1.5, b0 = 2, b00 = 0.5
//    deltaT0_(vf) == GREAT (c = c0 = 1, c00 = 0) while vf.oldTime() and vf.oldTime().oldTime() carry the same time
//    index (:48-68): during time step 1 for vf = DD, during time steps 1 and 2 for vf = D.oldTime().  All old-time
//    levels are created in the constructor (:143-145) as copies of the initial field.
//  Euler     [OF-ext] EulerD2dt2Scheme, rho-field overloads on a static mesh:
//      fvm: diag = coefft rDeltaT2 rho V, source = rDeltaT2 V rho ((coefft + coefft00) DD.o - coefft00 DD.oo)
//      fvc: rDeltaT2 rho (coefft Y - (coefft + coefft00) Y.o + coefft00 Y.oo)
double ulD2dt2(const s4f_oracle& o, dvec& hist) {
    const int N = o.N;
    hist.assign(3 * (size_t)N, 0.0);
    // rho_*g() (nonLinGeomUpdatedLagSolid.C:190); the uns model writes rho()*g(), the reference density (unsNonLinGeomUpdatedLagSolid.C:263)
    for (int c = 0; c < N; c++) for (int q = 0; q < 3; q++) hist[3 * c + q] = (o.unsUL() ? o.law.rho : o.rho[c]) * o.ctl.g[q];
    if (o.ctl.d2dt2Scheme == S4F_D2DT2_STEADY_STATE) return 0.0;
    const double dt = o.ctl.deltaT, dt0 = o.ctl.deltaT0 > 0 ? o.ctl.deltaT0 : dt;
    if (o.ctl.d2dt2Scheme == S4F_D2DT2_EULER) {
        const double cf = (dt + dt0) / (2 * dt), cf00 = (dt + dt0) / (2 * dt0), r2 = 4.0 / ((dt + dt0) * (dt + dt0));
        for (int c = 0; c < N; c++) for (int q = 0; q < 3; q++) {
            const int i = 3 * c + q;
            hist[i] += r2 * o.rho[c] * ((cf + cf00) * o.DDo[i] - cf00 * o.DDoo[i])
                     - r2 * o.rho[c] * (cf * o.Dold[i] - (cf + cf00) * o.DoldOld[i] + cf00 * o.Dooo[i]);
        }
        return cf * r2;
    }
    const bool firstM = o.timeIndex <= 1, firstC = o.timeIndex <= 2;
    const double kb = 1.0 + dt / (dt + dt0), kb00 = dt * dt / (dt0 * (dt + dt0)), kb0 = kb + kb00;
    const double cm = firstM ? 1.0 : kb, cm00 = firstM ? 0.0 : kb00, cm0 = cm + cm00;
    const double cc = firstC ? 1.0 : kb, cc00 = firstC ? 0.0 : kb00, cc0 = cc + cc00;
    const double r = 1.0 / (dt * dt);
    auto ddt = [&](const dvec& X, const dvec& Xo, const dvec& Xoo, int i) { return kb * X[i] - kb0 * Xo[i] + kb00 * Xoo[i]; };
    for (int c = 0; c < N; c++) for (int q = 0; q < 3; q++) {
        const int i = 3 * c + q;
        hist[i] += r * (cm * o.rho[c] * (kb0 * o.DDo[i] - kb00 * o.DDoo[i])
                        + cm0 * o.rhoO[c] * ddt(o.DDo, o.DDoo, o.DDooo, i) - cm00 * o.rhoOO[c] * ddt(o.DDoo, o.DDooo, o.DDoooo, i)
                        - cc * o.rho[c] * ddt(o.Dold, o.DoldOld, o.Dooo, i) + cc0 * o.rhoO[c] * ddt(o.DoldOld, o.Dooo, o.Doooo, i)
                        - cc00 * o.rhoOO[c] * ddt(o.Dooo, o.Doooo, o.Dooooo, i));
    }
    return r * cm * kb;
}

// ------------------------------------------------------------------------------------------------
// Momentum equation  (SM/linGeomTotalDispSolid/linGeomTotalDispSolid.C:141-149)
//
//   rho*fvm::d2dt2(D) == fvm::laplacian(impKf,D) - fvc::laplacian(impKf,D) + fvc::div(sigma)
//                        + rho*g + stabilisation
//
// Assembled directly in the sign of the final system A D = b ( "A == B" is A - B, [OF-ext] ):
//   upper_f = -impKf_f*nonOrthDeltaCoeffs_f*magSf_f ;  diag_P = sum_f (-upper_f)  (+ d2dt2)
//   internalCoeffs_b = -impKf_b*magSf_b*gradientInternalCoeffs_b (added to diag per component)
//   boundaryCoeffs_b = +impKf_b*magSf_b*gradientBoundaryCoeffs_b (added to source)
//   source = d2dt2 source - V*fvcLaplacian + V*div(sigma) + V*rho*g + V*stabilisation
// The non-orthogonal correction of fvm::laplacian (source -= V div(gammaMagSf*correction)) and the one
// inside fvc::laplacian's corrected snGrad are the same field with opposite sign: they are left out.
// ------------------------------------------------------------------------------------------------
void assembleMatrix(s4f_oracle& o) {
    const int N = o.N, F = o.F, B = o.B;
    o.upper.assign(F, 0.0); o.diag.assign(N, 0.0);
    o.intCoeffs.assign(3 * B, 0.0);
    for (int f = 0; f < F; f++) {
        double a = o.impKf[f] * o.nod[f] * o.magSf[f];
        o.upper[f] = -a; o.diag[o.own[f]] += a; o.diag[o.nei[f]] += a;
    }
    if (o.UL()) {
        dvec hist; const double kd = ulD2dt2(o, hist);
        for (int c = 0; c < N; c++) o.diag[c] += kd * o.rho[c] * o.V[c];
    } else if (o.ctl.d2dt2Scheme != S4F_D2DT2_STEADY_STATE) {
        const D2dt2Coeffs k = d2dt2Coeffs(o.ctl, o.law.rho, o.timeIndex);
        for (int c = 0; c < N; c++) o.diag[c] += k.diag * o.V[c];
    }
    for (int p = 0; p < o.nPatches; p++) for (int i = 0; i < o.pSize[p]; i++) {
        int b = o.pStart[p] + i, f = F + b;
        double n[3], k[3], delta; patchGeom(o, b, n, k, delta);
        double gm = o.impKf[f] * o.magSf[f];
        double dlap = o.nod[f];   // tsnGradScheme_().deltaCoeffs(vf) boundary value
        if (o.bcKind[p] == S4F_BC_FIXED_DISPLACEMENT) {
            for (int c = 0; c < 3; c++) o.intCoeffs[3 * b + c] = gm * dlap;     // -gm*(-deltaCoeffs)
        } else if (o.bcKind[p] == S4F_BC_SOLID_SYMMETRY) {
            // [OF-ext] basicSymmetry: gradientInternalCoeffs = -deltaCoeffs*snGradTransformDiag, diag = |n_c|
            for (int c = 0; c < 3; c++) o.intCoeffs[3 * b + c] = gm * dlap * std::fabs(n[c]);
        }   // fixedGradient: 0
    }
    // per-component diagonal after addBoundaryDiag
    o.diagC.assign(3 * N, 0.0);
    for (int c = 0; c < N; c++) for (int q = 0; q < 3; q++) o.diagC[3 * c + q] = o.diag[c];
    for (int b = 0; b < B; b++) for (int q = 0; q < 3; q++) o.diagC[3 * o.faceCells[b] + q] += o.intCoeffs[3 * b + q];
    o.matrixValid = true;
    o.gamg.valid = false;
}

void assembleSource(s4f_oracle& o) {
    const int N = o.N, F = o.F, B = o.B;
    const bool TL = (o.ctl.solidModel != S4F_MODEL_LIN_GEOM_TOTAL_DISP && !o.uns());
    o.source.assign(3 * N, 0.0);
    dvec& s = o.source;
    // d2dt2 old-time terms
    if (o.UL()) {
        dvec hist; ulD2dt2(o, hist);
        for (int c = 0; c < N; c++) for (int q = 0; q < 3; q++) s[3 * c + q] += o.V[c] * hist[3 * c + q];
    } else if (o.ctl.d2dt2Scheme != S4F_D2DT2_STEADY_STATE) {
        const D2dt2Coeffs k = d2dt2Coeffs(o.ctl, o.law.rho, o.timeIndex);
        const bool deep = o.ctl.d2dt2Scheme == S4F_D2DT2_BACKWARD && o.timeIndex > 1;   // before: D.ooo = D.oooo = D.oo (copies)
        const dvec& D3 = deep ? o.Dooo : o.DoldOld;
        const dvec& D4 = deep ? o.Doooo : o.DoldOld;
        for (int c = 0; c < N; c++) for (int q = 0; q < 3; q++)
            s[3 * c + q] += o.V[c] * (k.h[0] * o.Dold[3 * c + q] + k.h[1] * o.DoldOld[3 * c + q] + k.h[2] * D3[3 * c + q] + k.h[3] * D4[3 * c + q]);
    }
    // - V*fvc::laplacian(impKf, D): compact part, face flux impKf*magSf*delta*(D_N - D_P)
    forAllInternalFaces(o, [&](int f) {
        int P = o.own[f], Nn = o.nei[f];
        double a = -o.upper[f];
        for (int q = 0; q < 3; q++) {
            double flux = a * (o.D[3 * Nn + q] - o.D[3 * P + q]);
            s[3 * P + q] -= flux; s[3 * Nn + q] += flux;
        }
    });
    // + V*fvc::div(sigma)  [OF-ext] gaussDivScheme, linear:  Sf & (w sigma_P + (1-w) sigma_N)
    // TL: the cell tensor is J*Finv & sigma (nonLinGeomTotalLagTotalDispSolid.C:206)
    dvec T;   // full tensor per cell/boundary face
    T.resize(9 * (N + B));
    S4FO_PAR_FOR
    for (int c = 0; c < N + B; c++) {
        double sg[9]; S2T(&o.sigma[6 * c], sg);
        if (TL) { double t[9]; mulTT(&o.Finv[9 * c], sg, t); for (int q = 0; q < 9; q++) T[9 * c + q] = o.Jt[c] * t[q]; }
        else for (int q = 0; q < 9; q++) T[9 * c + q] = sg[q];
    }
    if (o.uns()) {
        // unsLinGeomSolid.C:129: fvc::div(mesh().Sf() & sigmaf_): the face stress itself, no interpolation
        for (int f = 0; f < F + B; f++) {
            double fl[3];
            if (o.unsFinite()) {  // unsNonLinGeomTotalLagSolid.C:273: fvc::div((Jf Finvf.T() & Sf) & sigmaf); updated Lagrangian
                                  // (unsNonLinGeomUpdatedLagSolid.C:262) the same with relJf, relFinvf: gradDf is grad(DD)f there
                double Ff[9]; transposeT(&o.gradDf[9 * (size_t)f], Ff); Ff[0] += 1; Ff[4] += 1; Ff[8] += 1;
                const double J = detT(Ff);
                double Fi[9], FiT[9]; invT(Ff, Fi); transposeT(Fi, FiT);
                const double* S = &o.Sf[3 * (size_t)f];
                double a[3];
                for (int i = 0; i < 3; i++) a[i] = J * (FiT[3 * i] * S[0] + FiT[3 * i + 1] * S[1] + FiT[3 * i + 2] * S[2]);
                SvS(&o.sigmaf[6 * (size_t)f], a, fl);
            } else
            SvS(&o.sigmaf[6 * (size_t)f], &o.Sf[3 * (size_t)f], fl);      // Sf & sigmaf (symmetric)
            const int P = f < F ? o.own[f] : o.faceCells[f - F];
            for (int q = 0; q < 3; q++) s[3 * P + q] += fl[q];
            if (f < F) for (int q = 0; q < 3; q++) s[3 * o.nei[f] + q] -= fl[q];
        }
    } else {
    forAllInternalFaces(o, [&](int f) {
        int P = o.own[f], Nn = o.nei[f];
        double wf = o.w[f], Tf[9];
        for (int q = 0; q < 9; q++) Tf[q] = wf * T[9 * P + q] + (1 - wf) * T[9 * Nn + q];
        double fl[3]; vT(&o.Sf[3 * f], Tf, fl);
        for (int q = 0; q < 3; q++) { s[3 * P + q] += fl[q]; s[3 * Nn + q] -= fl[q]; }
    });
    for (int b = 0; b < B; b++) {
        double fl[3]; vT(&o.Sf[3 * (F + b)], &T[9 * (N + b)], fl);
        for (int q = 0; q < 3; q++) s[3 * o.faceCells[b] + q] += fl[q];
    }
    }
    // + V*rho*g  (updated Lagrangian: the rho_ field, already in hist)
    if (!o.UL()) for (int c = 0; c < N; c++) for (int q = 0; q < 3; q++) s[3 * c + q] += o.V[c] * o.law.rho * o.ctl.g[q];
    // + V*stabilisation: RhieChow, SM/solidModel/momentumStabilisation/momentumStabilisation.C:112-114
    // (gamma = scaleFactor*impK), :119 (linear interpolate), :198-206 (zero on non-coupled boundaries),
    // :210-217  fvc::laplacian(gammaf, D) - fvc::div(gammaf*(Sf & interpolate(gradD)))
    if (o.ctl.stabilisation == S4F_STAB_RHIE_CHOW && !o.uns()) {      // the uns momentum equation has no stabilisation term (:124-131)
        const double sf = o.ctl.stabScaleFactor;
        forAllInternalFaces(o, [&](int f) {
            int P = o.own[f], Nn = o.nei[f];
            double wf = o.w[f];
            double gP = sf * o.impK[P], gN = sf * o.impK[Nn];
            double gf = wf * gP + (1 - wf) * gN;
            if (std::fabs(o.impK[P] - o.impK[Nn]) > SMALL) gf = 0.01 * 0.5 * (o.impK[P] + o.impK[Nn]);   // :137-150
            double gradf[9];
            for (int q = 0; q < 9; q++) gradf[q] = wf * o.gradD[9 * P + q] + (1 - wf) * o.gradD[9 * Nn + q];
            double Sg[3], cg[3]; vT(&o.Sf[3 * f], gradf, Sg); vT(&o.corr[3 * f], gradf, cg);
            for (int q = 0; q < 3; q++) {
                double sn = o.nod[f] * (o.D[3 * Nn + q] - o.D[3 * P + q]) + cg[q];   // corrected snGrad
                double flux = gf * (o.magSf[f] * sn - Sg[q]);
                s[3 * P + q] += flux; s[3 * Nn + q] -= flux;
            }
        });
    }
    // boundary: - impKf_b*magSf_b*snGrad_b (from -V*fvc::laplacian) and + boundaryCoeffs (addBoundarySource)
    o.bouCoeffs.assign(3 * B, 0.0);
    for (int p = 0; p < o.nPatches; p++) for (int i = 0; i < o.pSize[p]; i++) {
        int b = o.pStart[p] + i, f = F + b, P = o.faceCells[b];
        double n[3], k[3], delta; patchGeom(o, b, n, k, delta);
        double gm = o.impKf[f] * o.magSf[f];
        double sn[3]; bcSnGrad(o, p, b, o.gradD, sn);
        double gbc[3];   // gradientBoundaryCoeffs
        if (o.bcKind[p] == S4F_BC_SOLID_TRACTION) { for (int q = 0; q < 3; q++) gbc[q] = o.tracGrad[3 * b + q]; }
        else if (o.bcKind[p] == S4F_BC_FIXED_DISPLACEMENT) {
            // deltaCoeffs*(*this - (k & gradD_P))   fixedDisplacement...C:328-356
            double kg[3]; vT(k, &o.gradD[9 * P], kg);
            for (int q = 0; q < 3; q++) gbc[q] = delta * (o.D[3 * (N + b) + q] - kg[q]);
        } else {
            // [OF-ext] transformFvPatchField: snGrad() - cmptMultiply(gradientInternalCoeffs, patchInternalField)
            for (int q = 0; q < 3; q++) gbc[q] = sn[q] + o.nod[f] * std::fabs(n[q]) * o.D[3 * P + q];
        }
        for (int q = 0; q < 3; q++) {
            o.bouCoeffs[3 * b + q] = gm * gbc[q];
            s[3 * P + q] += -gm * sn[q] + o.bouCoeffs[3 * b + q];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// [OF-ext] lduMatrix::Amul, PCG, DIC/FDIC, diagonal preconditioner, solverPerformance, normFactor
// ------------------------------------------------------------------------------------------------
void Amul(const s4f_oracle& o, const double* diag, const double* x, double* y) {
    const int N = o.N, F = o.F;
    S4FO_PAR_FOR
    for (int c = 0; c < N; c++) y[c] = diag[c] * x[c];
    (void)F;
    forAllInternalFaces(o, [&](int f) {
        int l = o.own[f], u = o.nei[f];
        y[u] += o.upper[f] * x[l];   // lower == upper (symmetric)
        y[l] += o.upper[f] * x[u];
    });
}


// ---- CPU GAMG: set-up and cycle (see OLevel / OGamg above) -----------------------------------------------------------------
bool ogBuild(const s4f_oracle& o, OGamg& G) {
    const int N = o.N, F = o.F;
    G = OGamg();
    G.deg = o.ctl.gamgSmootherDegree > 0 ? o.ctl.gamgSmootherDegree : 3;
    G.cycle = o.ctl.gamgCycle;
    G.omega = o.ctl.gamgOverCorrection > 0 ? o.ctl.gamgOverCorrection : 2.2;
    const double ratio = o.ctl.gamgSmootherRatio > 0 ? o.ctl.gamgSmootherRatio : 0.3, lmax = 2.0, lmin = ratio * lmax;   // Gershgorin bound of D^-1 A
    G.theta = 0.5 * (lmax + lmin); G.delta = 0.5 * (lmax - lmin);
    G.lv.emplace_back();
    {
        OLevel& L = G.lv[0];
        L.n = N;
        L.ptr.assign(N + 1, 0);
        for (int f = 0; f < F; f++) { L.ptr[o.own[f] + 1]++; L.ptr[o.nei[f] + 1]++; }
        for (int i = 0; i < N; i++) L.ptr[i + 1] += L.ptr[i];
        L.col.resize(2 * (size_t)F); L.val.resize(2 * (size_t)F);
        std::vector<int> cur(L.ptr.begin(), L.ptr.end() - 1);
        for (int f = 0; f < F; f++) { const int e = cur[o.nei[f]]++; L.col[e] = o.own[f]; L.val[e] = -o.upper[f]; }   // lower neighbours first
        for (int f = 0; f < F; f++) { const int e = cur[o.own[f]]++; L.col[e] = o.nei[f]; L.val[e] = -o.upper[f]; }
        for (int q = 0; q < 3; q++) { L.diag[q].resize(N); for (int c = 0; c < N; c++) L.diag[q][c] = o.diagC[3 * (size_t)c + q]; }
    }
    while (G.lv.back().n > 512 && G.lv.size() < 12) {
        const int nFine = G.lv.back().n;
        std::vector<int> total(nFine); for (int i = 0; i < nFine; i++) total[i] = i;
        OLevel curL; const OLevel* m = &G.lv.back();
        int nc = nFine;
        for (int pass = 0; pass < 3; pass++) {
            std::vector<int> agg; ogPairPass(*m, agg, nc);
            OLevel next; ogGalerkin(*m, agg, nc, next);
            for (int i = 0; i < nFine; i++) total[i] = agg[total[i]];
            curL = std::move(next); m = &curL;
            if (nc <= 128) break;
        }
        if (nc >= nFine) break;
        OLevel& Fn = G.lv.back();
        Fn.parent = total;
        Fn.cptr.assign(nc + 1, 0);
        for (int i = 0; i < nFine; i++) Fn.cptr[total[i] + 1]++;
        for (int I = 0; I < nc; I++) Fn.cptr[I + 1] += Fn.cptr[I];
        Fn.child.resize(nFine);
        { std::vector<int> cur(Fn.cptr.begin(), Fn.cptr.end() - 1); for (int i = 0; i < nFine; i++) Fn.child[cur[total[i]]++] = i; }
        G.lv.push_back(std::move(curL));
    }
    for (OLevel& L : G.lv) { L.b.assign(L.n, 0.0); L.x.assign(L.n, 0.0); L.xp.assign(L.n, 0.0); L.xn.assign(L.n, 0.0); L.t.assign(L.n, 0.0); }
    const OLevel& C = G.lv.back();
    const int n = C.n; G.nC = n;
    if (n > 4096) return false;
    for (int q = 0; q < 3; q++) {         // dense inverse by Cholesky
        dvec A((size_t)n * n, 0.0);
        for (int i = 0; i < n; i++) { A[(size_t)i * n + i] = C.diag[q][i]; for (int e = C.ptr[i]; e < C.ptr[i + 1]; e++) A[(size_t)i * n + C.col[e]] -= C.val[e]; }
        for (int j = 0; j < n; j++) {
            double d = A[(size_t)j * n + j];
            for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
            if (!(d > 0)) return false;
            d = std::sqrt(d); A[(size_t)j * n + j] = d;
            for (int i = j + 1; i < n; i++) {
                double v = A[(size_t)i * n + j];
                for (int k = 0; k < j; k++) v -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
                A[(size_t)i * n + j] = v / d;
            }
        }
        G.inv[q].assign((size_t)n * n, 0.0);
        dvec y(n);
        for (int c = 0; c < n; c++) {
            for (int i = 0; i < n; i++) { double v = (i == c) ? 1.0 : 0.0; for (int k = 0; k < i; k++) v -= A[(size_t)i * n + k] * y[k]; y[i] = v / A[(size_t)i * n + i]; }
            for (int i = n - 1; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < n; k++) v -= A[(size_t)k * n + i] * G.inv[q][(size_t)k * n + c]; G.inv[q][(size_t)i * n + c] = v / A[(size_t)i * n + i]; }
        }
    }
    if (G.lv.size() > 1) { const int n1 = G.lv[1].n; G.kc1.assign(n1, 0.0); G.kv1.assign(n1, 0.0); G.kr.assign(n1, 0.0); G.kc2.assign(n1, 0.0); }
    G.valid = true;
    return true;
}

// y = A_l x for component q
void ogAmul(const s4f_oracle& o, const OLevel& L, int q, const double* x, double* y) {
    S4FO_PAR_FOR
    for (int i = 0; i < L.n; i++) {
        double acc = 0;
        for (int e = L.ptr[i]; e < L.ptr[i + 1]; e++) acc += L.val[e] * x[L.col[e]];
        y[i] = L.diag[q][i] * x[i] - acc;
    }
}

// Chebyshev-Jacobi of degree G.deg in three-term form on level l: from the zero initial guess (pre) or from x (post)
void ogSmooth(const s4f_oracle& o, OGamg& G, OLevel& L, int q, const double* b, bool fromZero) {
    const double sigma = G.theta / G.delta;
    double rho = 1.0 / sigma;
    const int n = L.n;
    int k0 = 0; bool havePrev = false, prevZero = false;
    if (fromZero) {
        S4FO_PAR_FOR
        for (int i = 0; i < n; i++) L.x[i] = b[i] / (G.theta * L.diag[q][i]);
        k0 = 1; prevZero = true;
    }
    for (int k = k0; k < G.deg; k++) {
        double c1, c2;
        if (k == 0) { c1 = 0.0; c2 = 1.0 / G.theta; }
        else { const double rhon = 1.0 / (2.0 * sigma - rho); c1 = rhon * rho; c2 = 2.0 * rhon / G.delta; rho = rhon; }
        S4FO_PAR_FOR
        for (int i = 0; i < n; i++) {
            double acc = 0;
            for (int e = L.ptr[i]; e < L.ptr[i + 1]; e++) acc += L.val[e] * L.x[L.col[e]];
            const double d = L.diag[q][i], xv = L.x[i], r = b[i] - (d * xv - acc);
            double xn = xv + c2 * r / d;
            if (k > 0) { if (havePrev) xn += c1 * (xv - L.xp[i]); else if (prevZero) xn += c1 * xv; }
            L.xn[i] = xn;
        }
        L.xp.swap(L.x); L.x.swap(L.xn);       // previous <- current, current <- new
        havePrev = true; prevZero = false;
    }
}

void ogCycle(const s4f_oracle& o, OGamg& G, size_t l, int q, const double* b);

// two flexible-CG steps on A_1 x = b_1, each preconditioned by the V-cycle from level 1 (Notay's K-cycle); result in lv[1].x
void ogKcycle1(const s4f_oracle& o, OGamg& G, int q) {
    OLevel& C = G.lv[1];
    const int n = C.n;
    dvec b(C.b);                                        // the cycle below overwrites the deeper right-hand sides only, keep b
    ogCycle(o, G, 1, q, b.data());
    G.kc1 = C.x;
    ogAmul(o, C, q, G.kc1.data(), G.kv1.data());
    double rho1 = 0, alpha1 = 0;
    for (int i = 0; i < n; i++) { rho1 += G.kc1[i] * G.kv1[i]; alpha1 += G.kc1[i] * b[i]; }
    const double s1 = std::fabs(rho1) > 1e-300 ? alpha1 / rho1 : 0.0;
    for (int i = 0; i < n; i++) G.kr[i] = b[i] - s1 * G.kv1[i];
    ogCycle(o, G, 1, q, G.kr.data());
    G.kc2 = C.x;
    dvec v2(n); ogAmul(o, C, q, G.kc2.data(), v2.data());
    double beta = 0, alpha2 = 0, gamma = 0;
    for (int i = 0; i < n; i++) { beta += G.kc2[i] * v2[i]; alpha2 += G.kc2[i] * G.kr[i]; gamma += G.kc2[i] * G.kv1[i]; }
    double k1 = 0, k2 = 0;
    if (std::fabs(rho1) > 1e-300) {
        k1 = alpha1 / rho1;
        const double rho2 = beta - gamma * gamma / rho1;
        if (std::fabs(rho2) > 1e-300 * std::fabs(beta) && std::fabs(rho2) > 1e-300) { k2 = alpha2 / rho2; k1 -= gamma * k2 / rho1; }
    }
    for (int i = 0; i < n; i++) C.x[i] = k1 * G.kc1[i] + k2 * G.kc2[i];
}

// x_l ~ A_l^-1 b (V-cycle; K-cycle step on level 1 when called for level 0 with G.cycle == 2); result in lv[l].x
void ogCycle(const s4f_oracle& o, OGamg& G, size_t l, int q, const double* b) {
    OLevel& L = G.lv[l];
    if (l + 1 == G.lv.size()) {
        const int n = L.n;
        for (int i = 0; i < n; i++) { double v = 0; const double* row = &G.inv[q][(size_t)i * n]; for (int k = 0; k < n; k++) v += row[k] * b[k]; L.x[i] = v; }
        return;
    }
    OLevel& C = G.lv[l + 1];
    ogSmooth(o, G, L, q, b, true);
    {   // residual and restriction
        S4FO_PAR_FOR
        for (int i = 0; i < L.n; i++) {
            double acc = 0;
            for (int e = L.ptr[i]; e < L.ptr[i + 1]; e++) acc += L.val[e] * L.x[L.col[e]];
            L.t[i] = b[i] - (L.diag[q][i] * L.x[i] - acc);
        }
        S4FO_PAR_FOR
        for (int I = 0; I < C.n; I++) { double v = 0; for (int k = L.cptr[I]; k < L.cptr[I + 1]; k++) v += L.t[L.child[k]]; C.b[I] = v; }
    }
    if (G.cycle == 2 && l == 0 && G.lv.size() > 2) ogKcycle1(o, G, q);
    else ogCycle(o, G, l + 1, q, C.b.data());
    S4FO_PAR_FOR
    for (int i = 0; i < L.n; i++) L.x[i] += G.omega * C.x[L.parent[i]];
    ogSmooth(o, G, L, q, b, false);
}

struct Precond {
    int kind; dvec rD;
    void init(const s4f_oracle& o, const double* diag, int k) {
        kind = k; const int N = o.N, F = o.F;
        if (kind == S4F_PRECOND_DIC) {
            rD.assign(diag, diag + N);
            if (o.nThreads <= 1) {
                for (int f = 0; f < F; f++) rD[o.nei[f]] -= o.upper[f] * o.upper[f] / rD[o.own[f]];
            } else {
#pragma omp parallel num_threads(o.nThreads)
                {
                    const int t = omp_get_thread_num(), cEnd = o.cellStart[t + 1];
                    for (int f = o.faceStart[t]; f < o.faceStart[t + 1]; f++)
                        if (o.nei[f] < cEnd) rD[o.nei[f]] -= o.upper[f] * o.upper[f] / rD[o.own[f]];
                }
            }
            for (int c = 0; c < N; c++) rD[c] = 1.0 / rD[c];
        } else if (kind == S4F_PRECOND_DIAGONAL) {
            rD.resize(N); for (int c = 0; c < N; c++) rD[c] = 1.0 / diag[c];
        }
    }
    void apply(const s4f_oracle& o, double* w, const double* r) const {
        const int N = o.N, F = o.F;
        if (kind == S4F_PRECOND_NONE) { std::memcpy(w, r, sizeof(double) * N); return; }
        S4FO_PAR_FOR
        for (int c = 0; c < N; c++) w[c] = rD[c] * r[c];
        if (kind == S4F_PRECOND_DIC) {
            if (o.nThreads <= 1) {
                for (int f = 0; f < F; f++) w[o.nei[f]] -= rD[o.nei[f]] * o.upper[f] * w[o.own[f]];
                for (int f = F - 1; f >= 0; f--) w[o.own[f]] -= rD[o.own[f]] * o.upper[f] * w[o.nei[f]];
            } else {
#pragma omp parallel num_threads(o.nThreads)
                {
                    const int t = omp_get_thread_num(), cEnd = o.cellStart[t + 1];
                    for (int f = o.faceStart[t]; f < o.faceStart[t + 1]; f++)
                        if (o.nei[f] < cEnd) w[o.nei[f]] -= rD[o.nei[f]] * o.upper[f] * w[o.own[f]];
                    for (int f = o.faceStart[t + 1] - 1; f >= o.faceStart[t]; f--)
                        if (o.nei[f] < cEnd) w[o.own[f]] -= rD[o.own[f]] * o.upper[f] * w[o.nei[f]];
                }
            }
        }
    }
};

SolverPerf solvePCG(s4f_oracle& o, const double* diag, double* psi, const double* source) {
    const int N = o.N, F = o.F;
    SolverPerf perf{0, 0, 0};
    dvec pA(N, 0.0), wA(N), rA(N), sumA(N);
    double wArA = 1e300, wArAold = wArA;   // solverPerf.great_
    Amul(o, diag, psi, wA.data());
    for (int c = 0; c < N; c++) rA[c] = source[c] - wA[c];
    // normFactor: sumA * gAverage(psi)
    for (int c = 0; c < N; c++) sumA[c] = diag[c];
    for (int f = 0; f < F; f++) { sumA[o.own[f]] += o.upper[f]; sumA[o.nei[f]] += o.upper[f]; }
    double avg = 0; for (int c = 0; c < N; c++) avg += psi[c]; avg /= N;
    double nf = 0;
    for (int c = 0; c < N; c++) { double t = sumA[c] * avg; nf += std::fabs(wA[c] - t) + std::fabs(source[c] - t); }
    nf += 1e-20;
    double sm = 0; for (int c = 0; c < N; c++) sm += std::fabs(rA[c]);
    perf.initRes = sm / nf; perf.finalRes = perf.initRes;
    auto converged = [&](double fr) { return fr < o.ctl.tolerance || (o.ctl.relTol > 1e-20 && fr < o.ctl.relTol * perf.initRes); };
    if (!converged(perf.finalRes)) {
        // GPU-only preconditioners map onto the reference's own: Chebyshev -> diagonal, GAMG -> DIC
        int pk = o.ctl.preconditioner;
        if (pk == S4F_PRECOND_CHEBYSHEV) pk = S4F_PRECOND_DIAGONAL;
        const bool mg = (pk == S4F_PRECOND_GAMG && o.cpuGamg);
        if (pk == S4F_PRECOND_GAMG && !mg) pk = S4F_PRECOND_DIC;
        Precond pre;
        if (mg) { if (!o.gamg.valid && !ogBuild(o, o.gamg)) { o.err = "CPU GAMG: set-up failed"; return perf; } }
        else pre.init(o, diag, pk);
        const bool flexible = mg && o.gamg.cycle == 2;       // the K-cycle is a variable preconditioner: Polak-Ribiere beta
        dvec rOld; if (flexible) rOld.assign(N, 0.0);
        do {
            wArAold = wArA;
            if (mg) { ogCycle(o, o.gamg, 0, o.curComp, rA.data()); std::memcpy(wA.data(), o.gamg.lv[0].x.data(), sizeof(double) * N); }
            else pre.apply(o, wA.data(), rA.data());
            wArA = 0;
#pragma omp parallel for schedule(static) reduction(+ : wArA) num_threads(o.nThreads) if (o.nThreads > 1)
            for (int c = 0; c < N; c++) wArA += wA[c] * rA[c];
            if (perf.nIter == 0) {
                S4FO_PAR_FOR
                for (int c = 0; c < N; c++) pA[c] = wA[c];
            } else {
                double beta = wArA / wArAold;
                if (flexible) {
                    double num = 0;
#pragma omp parallel for schedule(static) reduction(+ : num) num_threads(o.nThreads) if (o.nThreads > 1)
                    for (int c = 0; c < N; c++) num += wA[c] * (rA[c] - rOld[c]);
                    beta = num / wArAold;
                }
                S4FO_PAR_FOR
                for (int c = 0; c < N; c++) pA[c] = wA[c] + beta * pA[c];
            }
            Amul(o, diag, pA.data(), wA.data());
            double wApA = 0;
#pragma omp parallel for schedule(static) reduction(+ : wApA) num_threads(o.nThreads) if (o.nThreads > 1)
            for (int c = 0; c < N; c++) wApA += wA[c] * pA[c];
            if (std::fabs(wApA) / nf < VSMALL) break;   // checkSingularity
            double alpha = wArA / wApA;
            if (flexible) std::memcpy(rOld.data(), rA.data(), sizeof(double) * N);
            sm = 0;
#pragma omp parallel for schedule(static) reduction(+ : sm) num_threads(o.nThreads) if (o.nThreads > 1)
            for (int c = 0; c < N; c++) { psi[c] += alpha * pA[c]; rA[c] -= alpha * wA[c]; sm += std::fabs(rA[c]); }
            perf.finalRes = sm / nf;
        } while (++perf.nIter < o.ctl.maxIter && !converged(perf.finalRes));
    }
    return perf;
}


// [OF-ext] PBiCGStab::scalarSolve (van der Vorst's preconditioned BiCGStab as OpenFOAM states it: the
// residual after the first half step, sA, is tested for convergence and ends the solve with psi += alpha yA;
// omega = (tA.sA)/(tA.tA); checkSingularity is applied to |rA0.rA| and |omega| un-normalised).  Selected with
// fvSolution "solver PBiCGStab" (12x PBiCG / 1x PBiCGStab in the tutorials).
SolverPerf solvePBiCGStab(s4f_oracle& o, const double* diag, double* psi, const double* source) {
    const int N = o.N, F = o.F;
    SolverPerf perf{0, 0, 0};
    dvec pA(N, 0.0), yA(N), rA(N), sumA(N);
    Amul(o, diag, psi, yA.data());
    for (int c = 0; c < N; c++) rA[c] = source[c] - yA[c];
    for (int c = 0; c < N; c++) sumA[c] = diag[c];
    for (int f = 0; f < F; f++) { sumA[o.own[f]] += o.upper[f]; sumA[o.nei[f]] += o.upper[f]; }
    double avg = 0; for (int c = 0; c < N; c++) avg += psi[c]; avg /= N;
    double nf = 0;
    for (int c = 0; c < N; c++) { double t = sumA[c] * avg; nf += std::fabs(yA[c] - t) + std::fabs(source[c] - t); }
    nf += 1e-20;
    double sm = 0; for (int c = 0; c < N; c++) sm += std::fabs(rA[c]);
    perf.initRes = sm / nf; perf.finalRes = perf.initRes;
    auto converged = [&](double fr) { return fr < o.ctl.tolerance || (o.ctl.relTol > 1e-20 && fr < o.ctl.relTol * perf.initRes); };
    if (converged(perf.finalRes)) return perf;
    int pk = o.ctl.preconditioner;
    if (pk == S4F_PRECOND_CHEBYSHEV) pk = S4F_PRECOND_DIAGONAL;
    if (pk == S4F_PRECOND_GAMG) pk = S4F_PRECOND_DIC;
    Precond pre; pre.init(o, diag, pk);
    dvec AyA(N), sA(N), zA(N), tA(N);
    const dvec rA0(rA);
    double rA0rA = 0, alpha = 0, omega = 0;
    auto dot = [&](const dvec& a, const dvec& b) {
        double s = 0;
#pragma omp parallel for schedule(static) reduction(+ : s) num_threads(o.nThreads) if (o.nThreads > 1)
        for (int c = 0; c < N; c++) s += a[c] * b[c];
        return s;
    };
    auto sumMag = [&](const dvec& a) {
        double s = 0;
#pragma omp parallel for schedule(static) reduction(+ : s) num_threads(o.nThreads) if (o.nThreads > 1)
        for (int c = 0; c < N; c++) s += std::fabs(a[c]);
        return s;
    };
    do {
        const double rA0rAold = rA0rA;
        rA0rA = dot(rA0, rA);
        if (!(std::fabs(rA0rA) > VSMALL)) break;
        if (perf.nIter == 0) pA = rA;
        else {
            if (!(std::fabs(omega) > VSMALL)) break;
            const double beta = (rA0rA / rA0rAold) * (alpha / omega);
            S4FO_PAR_FOR
            for (int c = 0; c < N; c++) pA[c] = rA[c] + beta * (pA[c] - omega * AyA[c]);
        }
        pre.apply(o, yA.data(), pA.data());
        Amul(o, diag, yA.data(), AyA.data());
        const double rA0AyA = dot(rA0, AyA);
        alpha = rA0rA / rA0AyA;
        S4FO_PAR_FOR
        for (int c = 0; c < N; c++) sA[c] = rA[c] - alpha * AyA[c];
        perf.finalRes = sumMag(sA) / nf;
        if (converged(perf.finalRes)) {
            S4FO_PAR_FOR
            for (int c = 0; c < N; c++) psi[c] += alpha * yA[c];
            perf.nIter++;
            return perf;
        }
        pre.apply(o, zA.data(), sA.data());
        Amul(o, diag, zA.data(), tA.data());
        const double tAtA = dot(tA, tA);
        omega = dot(tA, sA) / tAtA;
        S4FO_PAR_FOR
        for (int c = 0; c < N; c++) { psi[c] += alpha * yA[c] + omega * zA[c]; rA[c] = sA[c] - omega * tA[c]; }
        perf.finalRes = sumMag(rA) / nf;
    } while (++perf.nIter < o.ctl.maxIter && !converged(perf.finalRes));
    return perf;
}

// [OF-ext] fvMatrix<vector>::solveSegregated: per solved component, addBoundaryDiag, solve.
// mechanicalLaw::updateSigmaHyd, solvePressureEqn branch (ML/mechanicalLaw/mechanicalLaw.C:1374-1468):
//   AD = DEqnA = DEqn.A() ([OF-ext] fvMatrix::A(): (diag + component average of internalCoeffs)/V, zero-gradient patches)
//   rDAf = pressureSmoothingScaleFactor * linear interpolate(impK/AD)
//   fvm::Sp(1, p) - fvm::laplacian(rDAf, p) == pExplicit - fvc::div(rDAf*(interpolate(grad p) & Sf))
//   p: zeroGradient patches (:452-458) -> no boundary coefficients; the corrected-laplacian's non-orthogonal part is
//   explicit, gamma magSf (corr & interpolate(fvc::grad(p))) [OF-ext gaussLaplacianScheme], and fvc::grad(p) of the
//   current p is the stored grad(sigmaHyd) (:1467).  Then p.relax() (no factor given: none), grad p = fvc::grad(p).
void updateSigmaHydSmoothed(s4f_oracle& o, double /*impK: uniform, o.impK holds the field*/) {
    const int N = o.N, F = o.F, B = o.B;
    if ((int)o.gradSigmaHyd.size() != 3 * (N + B)) o.gradSigmaHyd.assign(3 * (size_t)(N + B), 0.0);
    dvec r(N + B);
    for (int c = 0; c < N; c++) {
        const double AD = (o.diagC[3 * c] + o.diagC[3 * c + 1] + o.diagC[3 * c + 2]) / 3.0 / o.V[c];
        r[c] = o.impK[c] / AD;
    }
    for (int b = 0; b < B; b++) r[N + b] = o.impK[N + b] / (o.impK[o.faceCells[b]] / r[o.faceCells[b]]);
    dvec up(F), dg(N), src(N);
    for (int c = 0; c < N; c++) { dg[c] = o.V[c]; src[c] = o.V[c] * o.sigmaHydExp[c]; }
    const double sc = o.law.pressureSmoothingScaleFactor;
    for (int f = 0; f < F; f++) {
        const int P = o.own[f], Nn = o.nei[f];
        const double wf = o.w[f];
        const double gam = sc * (wf * r[P] + (1 - wf) * r[Nn]);
        const double a = gam * o.magSf[f] * o.nod[f];
        up[f] = -a; dg[P] += a; dg[Nn] += a;
        double gf[3]; for (int q = 0; q < 3; q++) gf[q] = wf * o.gradSigmaHyd[3 * P + q] + (1 - wf) * o.gradSigmaHyd[3 * Nn + q];
        const double flux = gam * (o.magSf[f] * dot3(&o.corr[3 * f], gf) - dot3(&o.Sf[3 * f], gf));
        src[P] += flux; src[Nn] -= flux;
    }
    for (int b = 0; b < B; b++) {       // - div term on the boundary: rDAf_b (grad p)_b & Sf_b  (zero once grad p is boundary-corrected)
        const double gam = sc * r[N + b];
        src[o.faceCells[b]] -= gam * dot3(&o.Sf[3 * (size_t)(F + b)], &o.gradSigmaHyd[3 * (size_t)(N + b)]);
    }
    dvec x(o.sigmaHyd.begin(), o.sigmaHyd.begin() + N);
    o.upper.swap(up);
    {   // sigmaHydEqn.solve() with the fvSolution "sigmaHyd" entry when the case has one (mechanicalLaw.C:1455)
        const s4fgpu_controls keep = o.ctl;
        if (o.law.sigmaHydTolerance > 0) { o.ctl.tolerance = o.law.sigmaHydTolerance; o.ctl.relTol = o.law.sigmaHydRelTol; if (o.law.sigmaHydMaxIter > 0) o.ctl.maxIter = o.law.sigmaHydMaxIter; }
        o.perfP = solvePCG(o, dg.data(), x.data(), src.data());
        o.ctl = keep;
    }
    o.upper.swap(up);
    const double al = o.law.sigmaHydRelax > 0 ? o.law.sigmaHydRelax : 1.0;                 // sigmaHyd.relax() :1459
    for (int c = 0; c < N; c++) o.sigmaHyd[c] = (al == 1.0) ? x[c] : o.sigmaHyd[c] + al * (x[c] - o.sigmaHyd[c]);
    for (int b = 0; b < B; b++) o.sigmaHyd[N + b] = o.sigmaHyd[o.faceCells[b]];         // zeroGradient
    // grad(sigmaHyd) = fvc::grad(sigmaHyd): scalar gradient through the vector machinery (component 0)
    dvec X(3 * (size_t)(N + B), 0.0), g;
    for (int c = 0; c < N + B; c++) X[3 * c] = o.sigmaHyd[c];
    gradInterior(o, X, g);
    for (int c = 0; c < N; c++) for (int i = 0; i < 3; i++) o.gradSigmaHyd[3 * c + i] = g[9 * c + 3 * i];
    for (int b = 0; b < B; b++) {       // gaussGrad::correctBoundaryConditions with snGrad = 0
        const int P = o.faceCells[b];
        double n[3], k[3], delta; patchGeom(o, b, n, k, delta);
        const double ng = dot3(n, &o.gradSigmaHyd[3 * P]);
        for (int i = 0; i < 3; i++) o.gradSigmaHyd[3 * (N + b) + i] = o.gradSigmaHyd[3 * P + i] - n[i] * ng;
    }
}

void solveSegregated(s4f_oracle& o, double* psi /*AoS [3N]*/, const double* source /*AoS*/) {
    const int N = o.N;
    dvec x(N), b(N), dg(N);
    for (int q = 0; q < 3; q++) {
        o.perf[q] = SolverPerf{0, 0, 0};
        if (!o.solD[q]) continue;
        for (int c = 0; c < N; c++) { x[c] = psi[3 * c + q]; b[c] = source[3 * c + q]; dg[c] = o.diagC[3 * c + q]; }
        o.curComp = q;
        o.perf[q] = (o.ctl.solver == S4F_SOLVER_PBICGSTAB) ? solvePBiCGStab(o, dg.data(), x.data(), b.data())
                                                       : solvePCG(o, dg.data(), x.data(), b.data());
        for (int c = 0; c < N; c++) psi[3 * c + q] = x[c];
        o.totalInner += o.perf[q].nIter;
    }
}

// solidModel::relaxField, SM/solidModel/solidModel.C:823-906 (fixed: D.relax() incl. boundary values)
void relaxField(s4f_oracle& o, int iCorr) {
    const int n3 = 3 * o.NB();
    if (o.ctl.relaxationMethod == S4F_RELAX_FIXED) {
        double a = o.ctl.fieldRelaxD;
        if (a != 1.0) for (int i = 0; i < n3; i++) o.D[i] = o.Dprev[i] + a * (o.D[i] - o.Dprev[i]);
    } else {
        if ((int)o.aitkenRes.size() != n3) { o.aitkenRes.assign(n3, 0.0); o.aitkenResPrev.assign(n3, 0.0); o.aitkenAlpha.assign(o.NB(), 1.0); }
        o.aitkenResPrev = o.aitkenRes;
        for (int i = 0; i < n3; i++) o.aitkenRes[i] = o.Dprev[i] - o.D[i];
        if (iCorr == 0) { std::fill(o.aitkenAlpha.begin(), o.aitkenAlpha.end(), o.ctl.fieldRelaxD); }
        else for (int c = 0; c < o.NB(); c++) {
            double dl[3], num = 0, den = 0;
            for (int q = 0; q < 3; q++) { dl[q] = o.aitkenResPrev[3 * c + q] - o.aitkenRes[3 * c + q]; num += o.aitkenResPrev[3 * c + q] * dl[q]; den += dl[q] * dl[q]; }
            double a = o.aitkenAlpha[c] * num / (den + SMALL);
            o.aitkenAlpha[c] = std::max(0.0, std::min(2.0, a));
        }
        for (int c = 0; c < o.NB(); c++) for (int q = 0; q < 3; q++) o.D[3 * c + q] -= o.aitkenAlpha[c] * o.aitkenRes[3 * c + q];
    }
}

// solidModel::converged, SM/solidModel/solidModelTemplates.C:27-188 (total approach: incremental()==false
// for the total-displacement models)
bool convergedCheck(s4f_oracle& o, int iCorr, s4fgpu_stats* st) {
    const int N = o.N;
    if (o.unsTL()) {
        // unsNonLinGeomTotalLagSolid.C:49-76 (residual) and :333-378: res = max|D - D.prevIter| / max(max|D - D.oldTime|, SMALL);
        // tolerance = max(maxRes * relativeTol_, solutionTol) with relativeTol_ read from "solutionTolerance" as well (:206-213);
        // the first iteration never converges ("force at least one iteration")
        double num = 0, den = 0;
        for (int c = 0; c < N; c++) {
            double a[3], r[3];
            for (int q = 0; q < 3; q++) { a[q] = o.D[3 * c + q] - o.Dold[3 * c + q]; r[q] = o.D[3 * c + q] - o.Dprev[3 * c + q]; }
            num = std::max(num, mag3(r)); den = std::max(den, mag3(a));
        }
        const double res = num / std::max(den, SMALL);
        if (iCorr == 0) o.unsMaxRes = 0;
        o.unsMaxRes = std::max(o.unsMaxRes, res);
        const double tol = std::max(o.unsMaxRes * o.ctl.solutionTolerance, o.ctl.solutionTolerance);
        const bool conv = iCorr > 0 && !(res > tol);
        if (st) {
            for (int q = 0; q < 3; q++) { st->initialResidual[q] = o.perf[q].initRes; st->finalResidual[q] = o.perf[q].finalRes; st->nIterations[q] = o.perf[q].nIter; }
            const double ir[3] = {o.perf[0].initRes, o.perf[1].initRes, o.perf[2].initRes};
            st->solverPerfInitRes = mag3(ir); st->relResidual = res; st->materialResidual = 0; st->converged = conv;
            st->totalInnerIterations = o.totalInner;
        }
        return conv;
    }
    double denom = 0, dmax = 0, res = 0;
    const bool incremental = o.incremental();
    for (int c = 0; c < N; c++) {
        double a[3], m[3], r[3];
        for (int q = 0; q < 3; q++) { a[q] = o.D[3 * c + q] - o.Dold[3 * c + q]; m[q] = o.D[3 * c + q]; r[q] = o.D[3 * c + q] - o.Dprev[3 * c + q]; }
        denom = std::max(denom, incremental ? mag3(m) : mag3(a)); dmax = std::max(dmax, mag3(m)); res = std::max(res, mag3(r));
    }
    if (denom < SMALL) denom = std::max(dmax, SMALL);
    double residualvf = res / denom;
    double matRes = lawResidual(o);
    double ir[3] = {o.perf[0].initRes, o.perf[1].initRes, o.perf[2].initRes};
    double spir = mag3(ir);
    bool conv = false;
    if (iCorr > 1 && matRes < o.ctl.materialTolerance) {
        if (spir < o.ctl.solutionTolerance && residualvf < o.ctl.solutionTolerance) conv = true;
        else if (residualvf < o.ctl.alternativeTolerance) conv = true;
        else if (spir < o.ctl.alternativeTolerance) conv = true;
    }
    if (st) {
        for (int q = 0; q < 3; q++) { st->initialResidual[q] = o.perf[q].initRes; st->finalResidual[q] = o.perf[q].finalRes; st->nIterations[q] = o.perf[q].nIter; }
        st->solverPerfInitRes = spir; st->relResidual = residualvf; st->materialResidual = matRes; st->converged = conv;
        st->totalInnerIterations = o.totalInner;
    }
    return conv;
}

// incremental models: D = D.oldTime() + DD; gradD = gradD.oldTime() + gradDD  (nonLinGeomTotalLagSolid.C:190-196)
void updateTotals(s4f_oracle& o, bool disp, bool grad) {
    if (!o.incremental()) return;
    if (disp) for (size_t i = 0; i < o.Dtot.size(); i++) o.Dtot[i] = o.Dold[i] + o.D[i];
    if (grad) for (size_t i = 0; i < o.gradDtot.size(); i++) o.gradDtot[i] = o.gradDold[i] + o.gradD[i];
}

// one pass of the do-loop body, linGeomTotalDispSolid.C:135-192 / nonLinGeomTotalLagTotalDispSolid.C:195-236 /
// nonLinGeomTotalLagSolid.C:147-223
void outerIteration(s4f_oracle& o, int iCorr) {
    o.Dprev = o.D;                               // D().storePrevIter()
    bcUpdateCoeffs(o);                           // fvMatrix ctor -> updateCoeffs
    if (!o.matrixValid) assembleMatrix(o);       // impKf is constant (update commented out :186-192)
    assembleSource(o);
    solveSegregated(o, o.D.data(), o.source.data());
    bcEvaluate(o);                               // D.correctBoundaryConditions()
    relaxField(o, iCorr);
    updateTotals(o, true, false);
    if (o.uns()) {                               // unsLinGeomSolid.C:146-157
        unsUpdateGradients(o);                   // interpolate(D, pointD); grad(D, pointD, gradD, gradDf)
        unsLawFaces(o);                          // mechanical().correct(sigmaf)
        if (o.unsUL()) updateKinematics(o);      // after the loop (:318-330): relF, relFinv, relJ of the cells (density update)
        lawCorrect(o);                           // mechanical().correct(sigma)
        return;
    }
    calcGrad(o);                                 // mechanical().grad(D, gradD)
    updateTotals(o, false, true);
    if (o.ctl.solidModel != S4F_MODEL_LIN_GEOM_TOTAL_DISP && !o.uns()) updateKinematics(o);
    lawCorrect(o);                               // mechanical().correct(sigma)
}

void allocFields(s4f_oracle& o) {
    const int n = o.NB();
    auto z = [&](dvec& v, int nc) { if ((int)v.size() != nc * n) v.assign((size_t)nc * n, 0.0); };
    z(o.D, 3); z(o.Dprev, 3); z(o.Dold, 3); z(o.DoldOld, 3); z(o.gradD, 9); z(o.gradDold, 9); z(o.sigma, 6); z(o.sigmaOld, 6);
    z(o.Dtot, 3); z(o.gradDtot, 9);
    z(o.epsilon, 6); z(o.sigmaHyd, 1);
    if ((int)o.sigmaf.size() != 6 * (o.F + o.B)) { o.sigmaf.assign(6 * (size_t)(o.F + o.B), 0.0); o.gradDf.assign(9 * (size_t)(o.F + o.B), 0.0); }
    z(o.Dooo, 3); z(o.Doooo, 3); z(o.Dooooo, 3); z(o.DDo, 3); z(o.DDoo, 3); z(o.DDooo, 3); z(o.DDoooo, 3);
    if ((int)o.rho.size() != n) { o.rho.assign(n, o.law.rho); o.rhoO = o.rho; o.rhoOO = o.rho; }
    if ((int)o.Ft.size() != 9 * n) {
        o.Ft.assign(9 * n, 0.0); o.Finv.assign(9 * n, 0.0); o.Jt.assign(n, 1.0);
        o.lawF.assign(9 * n, 0.0); o.lawFold.assign(9 * n, 0.0); o.relF.assign(9 * n, 0.0);
        for (int c = 0; c < n; c++) for (int d = 0; d < 3; d++) { o.Ft[9 * c + 4 * d] = 1; o.Finv[9 * c + 4 * d] = 1; o.lawF[9 * c + 4 * d] = 1; o.lawFold[9 * c + 4 * d] = 1; o.relF[9 * c + 4 * d] = 1; }
        o.lawJ.assign(n, 1.0); o.lawJold.assign(n, 1.0);
        o.bEbar.assign(6 * n, 0.0); o.bEbarOld.assign(6 * n, 0.0); o.bEbarTrial.assign(6 * n, 0.0);
        for (int c = 0; c < n; c++) { for (int d : {0, 3, 5}) { o.bEbar[6 * c + d] = 1; o.bEbarOld[6 * c + d] = 1; o.bEbarTrial[6 * c + d] = 1; } }
        o.sigmaY.assign(n, 0.0); o.DSigmaY.assign(n, 0.0); o.epsPEq.assign(n, 0.0); o.DEpsPEq.assign(n, 0.0);
        o.epsPOld.assign(6 * n, 0.0); o.epsPEqOld.assign(n, 0.0); o.sigmaYOld.assign(n, 0.0);
        o.epsP.assign(6 * n, 0.0); o.DEpsP.assign(6 * n, 0.0); o.DEpsPprev.assign(6 * n, 0.0); o.DLambda.assign(n, 0.0); o.plasticN.assign(6 * n, 0.0);
    }
}

// impK and impKf = fvc::interpolate(impK) (mechanicalModel.C:409-415); the face values follow the mesh weights
void setupImpK(s4f_oracle& o) {
    const int n = o.NB(), F = o.F, B = o.B;
    // impK: linearElastic.C:204-245 (2mu+lambda), neoHookeanElastic.C:101-119 and Mises at DLambda=0: 4/3 mu + K
    double impK = (o.law.kind == S4F_LAW_LINEAR_ELASTIC || o.law.kind == S4F_LAW_LINEAR_ELASTIC_MISES_PLASTIC)
                      ? 2.0 * o.law.mu + o.law.lambda : (4.0 / 3.0) * o.law.mu + o.law.K;
    o.impK.assign(n, impK);
    o.impKf.assign(F + B, 0.0);
    for (int f = 0; f < F; f++) o.impKf[f] = o.w[f] * o.impK[o.own[f]] + (1 - o.w[f]) * o.impK[o.nei[f]];
    for (int b = 0; b < B; b++) o.impKf[F + b] = o.impK[o.N + b];
}

void setupLaw(s4f_oracle& o) {
    setupImpK(o);
    std::fill(o.rho.begin(), o.rho.end(), o.law.rho); o.rhoO = o.rho; o.rhoOO = o.rho;     // rho_(mechanical().rho())
    o.Hp = 0;
    if (o.law.nTable == 2) o.Hp = (o.law.tableSigY[1] - o.law.tableSigY[0]) / (o.law.tableEps[1] - o.law.tableEps[0]);
    if (o.law.nTable >= 1) { std::fill(o.sigmaY.begin(), o.sigmaY.end(), tableLookup(o.law, 0.0)); o.sigmaYOld = o.sigmaY; }
    o.matrixValid = false;
}

}  // namespace

// ================================================================================================
// C API (mirrors include/s4fgpu.h with the s4fo_ prefix)
// ================================================================================================
extern "C" {

s4f_oracle* s4fo_create() { return new s4f_oracle(); }
void s4fo_destroy(s4f_oracle* o) { delete o; }
const char* s4fo_last_error(s4f_oracle* o) { return o->err.c_str(); }

int s4fo_set_mesh(s4f_oracle* o, int nCells, int nInternalFaces, const int* owner, const int* neighbour, int nPatches,
                  const int* patchStart, const int* patchSize, const int* patchKind, const int* patchNbrRank,
                  const int* faceCells, const int* solutionD) {
    (void)patchNbrRank;
    o->N = nCells; o->F = nInternalFaces; o->nPatches = nPatches;
    o->own.assign(owner, owner + o->F); o->nei.assign(neighbour, neighbour + o->F);
    o->pStart.assign(patchStart, patchStart + nPatches); o->pSize.assign(patchSize, patchSize + nPatches);
    o->pKind.assign(patchKind, patchKind + nPatches);
    o->B = 0; for (int p = 0; p < nPatches; p++) o->B = std::max(o->B, patchStart[p] + patchSize[p]);
    o->faceCells.assign(faceCells, faceCells + o->B);
    for (int i = 0; i < 3; i++) o->solD[i] = solutionD[i];
    o->bcKind.assign(nPatches, S4F_BC_SOLID_TRACTION);
    o->bcValue.assign(3 * o->B, 0.0); o->bcPressure.assign(o->B, 0.0); o->tracGrad.assign(3 * o->B, 0.0);
    for (int f = 0; f < o->F; f++) if (!(owner[f] < neighbour[f])) { o->err = "owner >= neighbour"; return 1; }
    return 0;
}

int s4fo_set_geometry(s4f_oracle* o, const double* C, const double* V, const double* Sf, const double* magSf,
                      const double* Cf, const double* weights, const double* nod, const double* corr, const double* CnbrB) {
    const int N = o->N, FB = o->F + o->B;
    o->C.assign(C, C + 3 * N); o->V.assign(V, V + N); o->Sf.assign(Sf, Sf + 3 * FB); o->magSf.assign(magSf, magSf + FB);
    o->Cf.assign(Cf, Cf + 3 * FB); o->w.assign(weights, weights + FB); o->nod.assign(nod, nod + FB);
    o->corr.assign(corr, corr + 3 * FB); o->CnbrB.assign(CnbrB, CnbrB + 3 * o->B);
    const bool again = !o->impK.empty() && (int)o->D.size() == 3 * o->NB();   // mesh motion: fields, BC data and history stay
    o->psValid = false;
    o->nonOrth = false;
    for (size_t i = 0; i < o->corr.size(); i++) if (std::fabs(o->corr[i]) > 1e-12) { o->nonOrth = true; break; }
    makeLeastSquaresVectors(*o);
    allocFields(*o);
    if (again) setupImpK(*o);
    o->matrixValid = false;
    return 0;
}

// polyMesh points and faces; pointCells [OF-ext primitiveMesh::pointCells] and the boundary pointFaces
// (enhancedVolPointInterpolation.C:60-160: boundary faces of non-empty, non-coupled patches)
int s4fo_set_points(s4f_oracle* o, int nPoints, const double* points, const int* faceVertsPtr, const int* faceVerts) {
    const int F = o->F, B = o->B;
    o->nPoints = nPoints; o->points.assign(points, points + 3 * (size_t)nPoints); o->psValid = false;
    o->fvPtr.assign(faceVertsPtr, faceVertsPtr + F + B + 1); o->fv.assign(faceVerts, faceVerts + faceVertsPtr[F + B]);
    std::vector<std::vector<int>> pc(nPoints), pb(nPoints);
    auto add = [](std::vector<int>& v, int x) { if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x); };
    for (int f = 0; f < F + B; f++) for (int j = o->fvPtr[f]; j < o->fvPtr[f + 1]; j++) {
        const int p = o->fv[j];
        if (p < 0 || p >= nPoints) { o->err = "set_points: vertex out of range"; return 1; }
        add(pc[p], f < F ? o->own[f] : o->faceCells[f - F]);
        if (f < F) add(pc[p], o->nei[f]);
    }
    for (int ip = 0; ip < o->nPatches; ip++) {
        if (o->pKind[ip] == S4F_PATCH_PROCESSOR) continue;
        for (int i = 0; i < o->pSize[ip]; i++) {
            const int b = o->pStart[ip] + i;
            for (int j = o->fvPtr[F + b]; j < o->fvPtr[F + b + 1]; j++) pb[o->fv[j]].push_back(b);
        }
    }
    o->pcPtr.assign(1, 0); o->pcCells.clear(); o->pbPtr.assign(1, 0); o->pbFaces.clear();
    for (int p = 0; p < nPoints; p++) {
        std::sort(pc[p].begin(), pc[p].end());
        o->pcCells.insert(o->pcCells.end(), pc[p].begin(), pc[p].end()); o->pcPtr.push_back((int)o->pcCells.size());
        o->pbFaces.insert(o->pbFaces.end(), pb[p].begin(), pb[p].end()); o->pbPtr.push_back((int)o->pbFaces.size());
    }
    return 0;
}

// volToPoint().interpolate(vf, pf), enhancedVolPointInterpolate.C:425-447: interpolateInternalField :125-158 with the
// inverse-distance weights of enhancedVolPointInterpolation.C:165-198 for points off the patches, interpolateBoundaryField
// :262-330 with the weights of :201-245 from the boundary-face values for patch points, then pointConstraints::constrain
// ([OF-ext]: symmetryPlane points lose their normal component, transform(I - nn, pf)).
int s4fo_interpolate_to_points(s4f_oracle* o, int field, int mode, double* out);
int s4fo_interpolate_to_points_impl(s4f_oracle* o, const dvec& Xf, double* out) {
    // PATCH mode on an explicit field (used by the uns model every outer iteration)
    dvec save; const bool inc = o->incremental();
    (void)inc;
    dvec& slot = o->incremental() ? o->Dtot : o->D;
    const bool same = (&Xf == &slot);
    if (!same) { save = slot; slot = Xf; }
    const int rc = s4fo_interpolate_to_points(o, S4F_FIELD_D, S4F_POINT_INTERP_PATCH, out);
    if (!same) slot = save;
    return rc;
}
int s4fo_interpolate_to_points(s4f_oracle* o, int field, int mode, double* out) {
    if (o->nPoints == 0) { o->err = "interpolate_to_points: call set_points first"; return 1; }
    const dvec* X = nullptr; const dvec* G = nullptr;
    if (field == S4F_FIELD_D) { X = o->incremental() ? &o->Dtot : &o->D; G = o->incremental() ? &o->gradDtot : &o->gradD; }
    else if (field == S4F_FIELD_DD && o->incremental()) { X = &o->D; G = &o->gradD; }
    if (!X) { o->err = "interpolate_to_points: field must be D or DD"; return 1; }
    const int N = o->N, F = o->F;
    if (mode == S4F_POINT_INTERP_GRAD) {
        // interpolate(vf, gradVf, pf), enhancedVolPointInterpolate.C:351-418: all points, cells only, no constraints
        for (int p = 0; p < o->nPoints; p++) {
            const double* x = &o->points[3 * (size_t)p];
            double acc[3] = {0, 0, 0}, sw = 0;
            for (int j = o->pcPtr[p]; j < o->pcPtr[p + 1]; j++) {
                const int c = o->pcCells[j];
                const double d[3] = {x[0] - o->C[3 * (size_t)c], x[1] - o->C[3 * (size_t)c + 1], x[2] - o->C[3 * (size_t)c + 2]};
                const double w = 1.0 / mag3(d);
                double dg[3]; vT(d, &(*G)[9 * (size_t)c], dg);             // delta & gradVf
                for (int q = 0; q < 3; q++) acc[q] += w * ((*X)[3 * (size_t)c + q] + dg[q]);
                sw += w;
            }
            for (int q = 0; q < 3; q++) out[3 * (size_t)p + q] = acc[q] / sw;
        }
        return 0;
    }
    for (int p = 0; p < o->nPoints; p++) {
        const double* x = &o->points[3 * (size_t)p];
        double acc[3] = {0, 0, 0}, sw = 0;
        if (o->pbPtr[p + 1] > o->pbPtr[p]) {
            for (int j = o->pbPtr[p]; j < o->pbPtr[p + 1]; j++) {
                const int b = o->pbFaces[j];
                const double* cf = &o->Cf[3 * (size_t)(F + b)];
                const double d[3] = {x[0] - cf[0], x[1] - cf[1], x[2] - cf[2]};
                const double w = 1.0 / mag3(d);
                for (int q = 0; q < 3; q++) acc[q] += w * (*X)[3 * (size_t)(N + b) + q];
                sw += w;
            }
        } else {
            for (int j = o->pcPtr[p]; j < o->pcPtr[p + 1]; j++) {
                const int c = o->pcCells[j];
                const double d[3] = {x[0] - o->C[3 * (size_t)c], x[1] - o->C[3 * (size_t)c + 1], x[2] - o->C[3 * (size_t)c + 2]};
                const double w = 1.0 / mag3(d);
                for (int q = 0; q < 3; q++) acc[q] += w * (*X)[3 * (size_t)c + q];
                sw += w;
            }
        }
        for (int q = 0; q < 3; q++) out[3 * (size_t)p + q] = acc[q] / sw;
    }
    for (int ip = 0; ip < o->nPatches; ip++) {
        if (o->pKind[ip] != S4F_PATCH_SYMMETRY || o->pSize[ip] == 0) continue;
        double n[3] = {0, 0, 0};                      // symmetryPlanePolyPatch::n(): the (planar) patch normal
        for (int i = 0; i < o->pSize[ip]; i++) for (int q = 0; q < 3; q++) n[q] += o->Sf[3 * (size_t)(F + o->pStart[ip] + i) + q];
        const double m = mag3(n); for (int q = 0; q < 3; q++) n[q] /= m;
        for (int i = 0; i < o->pSize[ip]; i++) {
            const int b = o->pStart[ip] + i;
            for (int j = o->fvPtr[F + b]; j < o->fvPtr[F + b + 1]; j++) {
                double* v = &out[3 * (size_t)o->fv[j]];
                const double nv = dot3(n, v);
                for (int q = 0; q < 3; q++) v[q] -= n[q] * nv;
            }
        }
    }
    return 0;
}

int s4fo_set_law(s4f_oracle* o, const s4fgpu_law* law) { o->law = *law; allocFields(*o); setupLaw(*o); return 0; }
int s4fo_set_controls(s4f_oracle* o, const s4fgpu_controls* c) { o->ctl = *c; o->matrixValid = false; return 0; }

int s4fo_set_bc(s4f_oracle* o, int patch, int kind, const double* value, const double* pressure) {
    if (patch < 0 || patch >= o->nPatches) { o->err = "bad patch"; return 1; }
    o->bcKind[patch] = kind; o->matrixValid = false;
    const int s = o->pStart[patch], n = o->pSize[patch];
    for (int i = 0; i < n; i++) {
        for (int q = 0; q < 3; q++) o->bcValue[3 * (s + i) + q] = value ? value[3 * i + q] : 0.0;
        o->bcPressure[s + i] = pressure ? pressure[i] : 0.0;
    }
    return 0;
}

static dvec* fieldPtr(s4f_oracle* o, int field, int& ncomp, int& off, int& count) {
    const int N = o->N, B = o->B, F = o->F;
    off = 0; count = N;
    switch (field) {
        case S4F_FIELD_D: ncomp = 3; return o->incremental() ? &o->Dtot : &o->D;
        case S4F_FIELD_DD: ncomp = 3; return o->incremental() ? &o->D : nullptr;
        case S4F_FIELD_GRAD_DD: ncomp = 9; return o->incremental() ? &o->gradD : nullptr;
        case S4F_FIELD_D_OLD: ncomp = 3; return &o->Dold;
        case S4F_FIELD_D_OLDOLD: ncomp = 3; return &o->DoldOld;
        case S4F_FIELD_GRAD_D: ncomp = 9; return o->incremental() ? &o->gradDtot : &o->gradD;
        case S4F_FIELD_GRAD_D_OLD: ncomp = 9; return &o->gradDold;
        case S4F_FIELD_SIGMA: ncomp = 6; return &o->sigma;
        case S4F_FIELD_D_B: ncomp = 3; off = N; count = B; return o->incremental() ? &o->Dtot : &o->D;
        case S4F_FIELD_GRAD_D_B: ncomp = 9; off = N; count = B; return o->incremental() ? &o->gradDtot : &o->gradD;
        case S4F_FIELD_SIGMA_B: ncomp = 6; off = N; count = B; return &o->sigma;
        case S4F_FIELD_SOURCE: ncomp = 3; return &o->source;
        case S4F_FIELD_DIAG: ncomp = 3; return &o->diagC;
        case S4F_FIELD_UPPER: ncomp = 1; count = F; return &o->upper;
        case S4F_FIELD_EPSILON_P_EQ: ncomp = 1; return &o->epsPEq;
        case S4F_FIELD_SIGMA_Y: ncomp = 1; return &o->sigmaY;
        case S4F_FIELD_BEBAR: ncomp = 6; return &o->bEbar;
        case S4F_FIELD_DLAMBDA: ncomp = 1; return &o->DLambda;
        case S4F_FIELD_J: ncomp = 1; return &o->lawJ;
        case S4F_FIELD_F: ncomp = 9; return &o->lawF;
        case S4F_FIELD_DEPSILON_P: ncomp = 6; return &o->DEpsP;
        case S4F_FIELD_EPSILON_P: ncomp = 6; return &o->epsP;
        case S4F_FIELD_TRACTION_GRADIENT_B: ncomp = 3; count = B; return &o->tracGrad;
        case S4F_FIELD_RHO: ncomp = 1; return &o->rho;
        case S4F_FIELD_SIGMA_HYD: ncomp = 1; return &o->sigmaHyd;
        case S4F_FIELD_SIGMA_F: ncomp = 6; count = F + B; return &o->sigmaf;
        case S4F_FIELD_GRAD_D_F: ncomp = 9; count = F + B; return &o->gradDf;
        case S4F_FIELD_GRAD_SIGMA_HYD: ncomp = 3; return &o->gradSigmaHyd;
        case S4F_FIELD_DD_B: ncomp = 3; off = N; count = B; return o->incremental() ? &o->D : nullptr;
    }
    return nullptr;
}
int s4fo_upload(s4f_oracle* o, int field, const double* host) {
    int nc, off, cnt; dvec* v = fieldPtr(o, field, nc, off, cnt);
    if (!v) { o->err = "bad field"; return 1; }
    if ((int)v->size() < nc * (off + cnt)) v->resize((size_t)nc * (off + cnt), 0.0);
    std::memcpy(v->data() + (size_t)nc * off, host, sizeof(double) * nc * cnt); return 0;
}
int s4fo_download(s4f_oracle* o, int field, double* host) {
    int nc, off, cnt; dvec* v = fieldPtr(o, field, nc, off, cnt);
    if (!v || (int)v->size() < nc * (off + cnt)) { o->err = "bad field / not available"; return 1; }
    std::memcpy(host, v->data() + (size_t)nc * off, sizeof(double) * nc * cnt); return 0;
}

// linGeomTotalDispSolid.C:82-84: D.correctBoundaryConditions(); D.storePrevIter(); mechanical().grad(D, gradD)
int s4fo_initialise(s4f_oracle* o) {
    allocFields(*o);
    bcUpdateCoeffs(*o);
    bcEvaluate(*o);
    o->Dprev = o->D;
    if (o->uns()) {       // unsLinGeomSolid.C:86-89
        if (o->nPoints == 0) { o->err = "unsLinearGeometry needs the mesh points (set_points)"; return 1; }
        unsUpdateGradients(*o);
        assembleMatrix(*o);
        return 0;
    }
    calcGrad(*o);
    updateTotals(*o, false, true);
    if (o->ctl.solidModel != S4F_MODEL_LIN_GEOM_TOTAL_DISP) updateKinematics(*o);
    assembleMatrix(*o);
    return 0;
}

int s4fo_new_timestep(s4f_oracle* o, double deltaT) {
    o->ctl.deltaT0 = o->ctl.deltaT; o->ctl.deltaT = deltaT;
    if (o->ctl.d2dt2Scheme == S4F_D2DT2_BACKWARD && !o->UL()) {     // GeometricField::storeOldTimes over the four-level chain
        if (o->timeIndex >= 2) { o->Doooo = o->Dooo; o->Dooo = o->DoldOld; }
        else { o->Doooo = o->DoldOld; o->Dooo = o->DoldOld; }
    }
    if (o->UL()) {     // GeometricField::storeOldTimes over the chains created in the constructor (:143-145)
        if (o->timeIndex == 0) {
            o->Dooo = o->DoldOld; o->Doooo = o->DoldOld; o->Dooooo = o->DoldOld;
            o->DDo = o->D; o->DDoo = o->D; o->DDooo = o->D; o->DDoooo = o->D;
            o->rhoO = o->rho; o->rhoOO = o->rho;
        }
        o->Dooooo = o->Doooo; o->Doooo = o->Dooo; o->Dooo = o->DoldOld;
        o->DDoooo = o->DDooo; o->DDooo = o->DDoo; o->DDoo = o->DDo; o->DDo = o->D;
        o->rhoOO = o->rhoO; o->rhoO = o->rho;
        o->matrixValid = false;       // rho_ and the mesh changed in updateTotalFields
    }
    o->timeIndex++;
    o->DoldOld = o->Dold;
    if (o->incremental()) { o->Dold = o->Dtot; o->gradDold = o->gradDtot; }      // the total fields roll; DD keeps its value as the initial guess
    else { o->Dold = o->D; o->gradDold = o->gradD; }
    o->sigmaOld = o->sigma;
    o->lawFold = o->lawF; o->lawJold = o->lawJ; o->bEbarOld = o->bEbar;
    if (o->unsUL() && !o->Ff.empty()) o->FfOld = o->Ff;
    o->epsPOld = o->epsP; o->epsPEqOld = o->epsPEq; o->sigmaYOld = o->sigmaY;
    if (o->ctl.d2dt2Scheme != S4F_D2DT2_STEADY_STATE) o->matrixValid = false;
    return 0;
}

int s4fo_outer_iteration(s4f_oracle* o, s4fgpu_stats* st) {
    outerIteration(*o, o->iCorrLast);
    convergedCheck(*o, o->iCorrLast, st);
    o->iCorrLast++;
    if (st) st->nCorr = o->iCorrLast;
    return 0;
}

// solidModel::evolve loop: do { ... } while (!converged(iCorr,...) && ++iCorr < nCorr)
int s4fo_evolve(s4f_oracle* o, s4fgpu_stats* st) {
    int iCorr = 0; bool conv;
    s4fgpu_stats loc{}; if (!st) st = &loc;
    do { outerIteration(*o, iCorr); conv = convergedCheck(*o, iCorr, st); } while (!conv && ++iCorr < o->ctl.nCorrectors);
    st->nCorr = conv ? iCorr + 1 : iCorr; o->iCorrLast = 0;
    return 0;
}

// updateTotalFields: neoHookeanElasticMisesPlastic.C:1526-1536 (history commit)
int s4fo_update_total_fields(s4f_oracle* o) {
    if (o->UL()) {
        // after the loop, nonLinGeomUpdatedLagSolid.C:243: gradD() = fvc::grad(D().oldTime() + DD()) (calculated patches)
        gradCalculated(*o, o->Dtot, o->gradDtot);
        // updateTotalFields :360-374: rho_ = rho_.oldTime()/relJ_; the mesh motion itself is done by the host
        // (s4fo_interpolate_to_points -> new points -> s4fo_set_geometry), then solidModel::updateTotalFields()
        const int n = o->NB();
        for (int c = 0; c < n; c++) o->rho[c] = o->rhoO[c] / o->Jt[c];
        o->matrixValid = false;
    }
    if (o->law.kind == S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC) {
        const int n = o->NB();
        for (int c = 0; c < n; c++) { o->sigmaY[c] += o->DSigmaY[c]; o->epsPEq[c] += o->DEpsPEq[c]; }
        for (int i = 0; i < 6 * n; i++) o->epsP[i] += o->DEpsP[i];
    }
    return 0;
}

int s4fo_op_grad(s4f_oracle* o) { if (o->uns()) { unsUpdateGradients(*o); unsLawFaces(*o); return 0; } calcGrad(*o); updateTotals(*o, false, true); if (o->ctl.solidModel != S4F_MODEL_LIN_GEOM_TOTAL_DISP) updateKinematics(*o); return 0; }
int s4fo_op_correct(s4f_oracle* o) { lawCorrect(*o); return 0; }
int s4fo_op_assemble(s4f_oracle* o) { bcUpdateCoeffs(*o); assembleMatrix(*o); assembleSource(*o); return 0; }
int s4fo_op_amul(s4f_oracle* o, int cmpt, const double* x, double* y) {
    if (!o->matrixValid) assembleMatrix(*o);
    dvec dg(o->N); for (int c = 0; c < o->N; c++) dg[c] = o->diagC[3 * c + cmpt];
    Amul(*o, dg.data(), x, y); return 0;
}
int s4fo_op_solve(s4f_oracle* o, double* psi, const double* source, s4fgpu_stats* st) {
    if (!o->matrixValid) assembleMatrix(*o);
    solveSegregated(*o, psi, source);
    if (st) for (int q = 0; q < 3; q++) { st->initialResidual[q] = o->perf[q].initRes; st->finalResidual[q] = o->perf[q].finalRes; st->nIterations[q] = o->perf[q].nIter; }
    return 0;
}
// unsLinGeomSolid gradients from GIVEN vertex displacements (tests: the gradient formulas of fvcGradf.C on their own)
int s4fo_uns_grad_from_points(s4f_oracle* o, const double* pointD) {
    o->pointD.assign(pointD, pointD + 3 * (size_t)o->nPoints);
    unsUpdateGradients(*o, false); unsLawFaces(*o);
    return 0;
}
// least-squares vectors for inspection (lsP [3(F+B)], lsN [3F])
int s4fo_get_ls_vectors(s4f_oracle* o, double* lsP, double* lsN) {
    std::memcpy(lsP, o->lsP.data(), sizeof(double) * o->lsP.size());
    std::memcpy(lsN, o->lsN.data(), sizeof(double) * o->lsN.size()); return 0;
}
double s4fo_table_lookup(const s4fgpu_law* law, double x) { return tableLookup(*law, x); }
// host threads for the timing baseline (0 = all cores); 1 restores the literal serial loops
// S4F_PRECOND_GAMG as a CPU multigrid of the GPU path's preconditioner family (bench.py's like-for-like CPU figure) instead of
// the mapping onto DIC; returns the previous setting
int s4fo_set_cpu_gamg(s4f_oracle* o, int on) { const int was = o->cpuGamg ? 1 : 0; o->cpuGamg = on != 0; o->gamg.valid = false; return was; }
// levels of the CPU multigrid after its last set-up (sizes[maxLevels]); returns the number of levels
int s4fo_cpu_gamg_levels(s4f_oracle* o, int* sizes, int maxLevels) {
    if (!o->gamg.valid) return 0;
    for (size_t l = 0; l < o->gamg.lv.size() && (int)l < maxLevels; l++) sizes[l] = o->gamg.lv[l].n;
    return (int)o->gamg.lv.size();
}

int s4fo_set_threads(s4f_oracle* o, int n) {
    if (n <= 0) n = omp_get_max_threads();
    o->nThreads = n;
    if (n > 1) buildPartition(*o);
    return n;
}

}  // extern "C"
