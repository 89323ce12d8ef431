// s4f_setup.cu -- once-per-mesh mirror: lduAddressing -> SELL-32 cell-centric rows, per-entry
// geometry, least-squares vectors, boundary lists, halo lists; field allocation; AoS<->SoA transfer.
//
// Reference data this mirrors: fvMesh owner()/neighbour()/boundary() and the geometric fields
// ([OF-ext] surfaceInterpolation weights / nonOrthDeltaCoeffs / nonOrthCorrectionVectors); the
// least-squares vectors follow NUM/extendedLeastSquaresGrad/extendedLeastSquaresVectors.C:121-158
// (dd tensor, 1/|d|^2 weights, true boundary deltas) and :229-272 (lsP / lsN), restated per
// (cell, neighbour) pair:  ls = invDd_P & d / |d|^2  with d pointing from the row cell outwards.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "s4f_comm.h"
#include "s4f_dev.cuh"

namespace {

inline void invSymm(const double* S, double* R) {
    double d = S[0] * S[3] * S[5] + 2.0 * S[1] * S[4] * S[2] - S[0] * S[4] * S[4] - S[1] * S[1] * S[5] - S[2] * S[3] * S[2];
    R[0] = (S[3] * S[5] - S[4] * S[4]) / d; R[1] = (S[2] * S[4] - S[1] * S[5]) / d; R[2] = (S[1] * S[4] - S[2] * S[3]) / d;
    R[3] = (S[0] * S[5] - S[2] * S[2]) / d; R[4] = (S[1] * S[2] - S[0] * S[4]) / d; R[5] = (S[0] * S[3] - S[1] * S[1]) / d;
}

__global__ void k_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, int count, int ncomp, int ld, int offset) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = (long long)count * ncomp;
    if (i >= total) return;
    int cell = (int)(i / ncomp), q = (int)(i % ncomp);
    soa[(size_t)q * ld + offset + cell] = aos[i];
}
__global__ void k_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, int count, int ncomp, int ld, int offset) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long total = (long long)count * ncomp;
    if (i >= total) return;
    int cell = (int)(i / ncomp), q = (int)(i % ncomp);
    aos[i] = soa[(size_t)q * ld + offset + cell];
}
__global__ void k_fill(double* p, double v, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

int s4f_aos_to_soa(s4fgpu_ctx* c, const double* hostAoS, double* devSoA, int count, int ncomp, int offset) {
    if (count == 0) return 0;
    size_t n = (size_t)count * ncomp;
    if (c->staging.n < n) S4F_CHECK_CUDA(c, c->staging.alloc(n, false));
    S4F_CHECK_CUDA(c, cudaMemcpyAsync(c->staging.p, hostAoS, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_aos_to_soa<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->staging.p, devSoA, count, ncomp, c->ld, offset);
    c->launches++;
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int s4f_soa_to_aos(s4fgpu_ctx* c, const double* devSoA, double* hostAoS, int count, int ncomp, int offset) {
    if (count == 0) return 0;
    size_t n = (size_t)count * ncomp;
    if (c->staging.n < n) S4F_CHECK_CUDA(c, c->staging.alloc(n, false));
    k_soa_to_aos<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(devSoA, c->staging.p, count, ncomp, c->ld, offset);
    c->launches++;
    S4F_CHECK_CUDA(c, cudaMemcpyAsync(hostAoS, c->staging.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// symmetry-plane data of the points (shared by the point-ghost builder and the vol->point weights): hN = the plane normal
// at the points of a symmetryPlane patch (symmetryPlanePolyPatch::n(): the normalised sum of the patch's face areas), fixAxis
// = the coordinate such a plane fixes when it is aligned with an axis (solidModel::moveMesh, solidModel.C:2040-2080).
static void point_symmetry(const s4fgpu_ctx* c, const double* bSf, std::vector<double>& hN, std::vector<int>& fixAxis) {
    const int F = c->F, nP = c->nPoints;
    hN.assign(3 * (size_t)std::max(nP, 1), 0.0);
    fixAxis.assign(std::max(nP, 1), -1);
    for (int ip = 0; ip < c->nPatches; ip++) {
        if (c->pKind[ip] != S4F_PATCH_SYMMETRY || c->pSize[ip] == 0) continue;
        double n[3] = {0, 0, 0};
        for (int i = 0; i < c->pSize[ip]; i++) for (int q = 0; q < 3; q++) n[q] += bSf[3 * (size_t)(c->pStart[ip] + i) + q];
        const double m = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        std::vector<double> pn(3 * (size_t)nP, 0.0); std::vector<char> on(nP, 0);
        for (int i = 0; i < c->pSize[ip]; i++) {
            const int b = c->pStart[ip] + i;
            const double* sf = &bSf[3 * (size_t)b];
            const double ms = std::sqrt(sf[0] * sf[0] + sf[1] * sf[1] + sf[2] * sf[2]);
            for (int j = c->hFvPtr[F + b]; j < c->hFvPtr[F + b + 1]; j++) {
                const int p = c->hFv[j]; on[p] = 1;
                for (int q = 0; q < 3; q++) { hN[3 * (size_t)p + q] = n[q] / m; pn[3 * (size_t)p + q] += sf[q] / ms; }
            }
        }
        // the average of the point normals (each the normalised sum of the unit normals of its patch faces)
        double avg[3] = {0, 0, 0}; int cntP = 0;
        for (int p = 0; p < nP; p++) if (on[p]) {
            const double* v = &pn[3 * (size_t)p];
            const double mv = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            for (int q = 0; q < 3; q++) avg[q] += v[q] / mv;
            cntP++;
        }
        for (int ax = 0; ax < 3; ax++)
            if (std::fabs(avg[ax] / cntP) > 0.95) { for (int p = 0; p < nP; p++) if (on[p]) fixAxis[p] = ax; break; }
    }
}

// Decomposed meshes: the operators with a point stencil (pointCellsLeastSquares gradient, vol->point interpolation, the
// uns face gradients) need, at the points of the processor patches, the cells and the boundary faces of the OTHER ranks
// around the same point -- also of ranks that touch this one only along an edge or at a corner.  OpenFOAM does this with
// globalMeshData / syncTools point synchronisation of partial sums; here every rank holds those remote values in X
// extra slots behind its boundary slots (one peer-memory exchange, haloX, fills them) and evaluates the complete stencil
// itself: no partial sums, the same summation on every rank that shares the point.
//   1. every rank publishes, per processor-patch point: coordinates, the local cells around it with their centres, the
//      local (non-processor) boundary faces around it with their centres, its symmetry-plane data;
//   2. points are identified across ranks by their coordinates, bit for bit (processor patches are written from one set
//      of points; a mismatch is an error, not a tolerance);
//   3. requests go back (which of your cells / boundary slots I need), giving the send lists of the exchange plan.
// Collective; called from set_geometry, before the rows are laid out, because X enters the leading dimension.
int s4f_build_point_ghosts(s4fgpu_ctx* c) {
    const int N = c->N, F = c->F, B = c->B, nP = c->nPoints;
    s4f_halo_plan_destroy(c->haloX); c->haloX = nullptr;
    c->X = 0; c->extPtr.assign(nP + 1, 0); c->extSlot.clear(); c->extCtr.clear(); c->extIsB.clear();
    c->extSymN.assign(3 * (size_t)std::max(nP, 1), 0.0); c->extFixAxis.assign(std::max(nP, 1), -1);
    if (c->nRanks <= 1 || nP == 0) return 0;
    int G = 0;
    for (int ip = 0; ip < c->nPatches; ip++) if (c->pKind[ip] == S4F_PATCH_PROCESSOR) G += c->pSize[ip];
    const int bOff = N + G, xOff = N + G + B;
    // local point -> cells / boundary faces, for the processor-patch points only
    std::vector<char> isProc(nP, 0);
    for (int ip = 0; ip < c->nPatches; ip++) {
        if (c->pKind[ip] != S4F_PATCH_PROCESSOR) continue;
        for (int i = 0; i < c->pSize[ip]; i++) {
            const int b = c->pStart[ip] + i;
            for (int j = c->hFvPtr[F + b]; j < c->hFvPtr[F + b + 1]; j++) isProc[c->hFv[j]] = 1;
        }
    }
    std::vector<int> procPts, procIdx(nP, -1);
    for (int p = 0; p < nP; p++) if (isProc[p]) { procIdx[p] = (int)procPts.size(); procPts.push_back(p); }
    std::vector<std::vector<int>> pc(procPts.size()), pb(procPts.size());
    auto add = [](std::vector<int>& v, int x) { if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x); };
    for (int f = 0; f < F + B; f++) for (int j = c->hFvPtr[f]; j < c->hFvPtr[f + 1]; j++) {
        const int k = procIdx[c->hFv[j]];
        if (k < 0) continue;
        add(pc[k], f < F ? c->own[f] : c->faceCells[f - F]);
        if (f < F) add(pc[k], c->nei[f]);
    }
    for (int ip = 0; ip < c->nPatches; ip++) {
        if (c->pKind[ip] == S4F_PATCH_PROCESSOR) continue;
        for (int i = 0; i < c->pSize[ip]; i++) {
            const int b = c->pStart[ip] + i;
            for (int j = c->hFvPtr[F + b]; j < c->hFvPtr[F + b + 1]; j++) { const int k = procIdx[c->hFv[j]]; if (k >= 0) pb[k].push_back(b); }
        }
    }
    std::vector<double> symN; std::vector<int> fixAxis;
    point_symmetry(c, c->hSf.data() + 3 * (size_t)F, symN, fixAxis);
    // 1. publish
    std::vector<double> rec;
    for (size_t k = 0; k < procPts.size(); k++) {
        const int p = procPts[k];
        for (int q = 0; q < 3; q++) rec.push_back(c->hPoints[3 * (size_t)p + q] + 0.0);        // + 0.0: -0 -> +0
        rec.push_back((double)pc[k].size()); rec.push_back((double)pb[k].size());
        for (int q = 0; q < 3; q++) rec.push_back(symN[3 * (size_t)p + q]);
        rec.push_back((double)fixAxis[p]);
        for (int cell : pc[k]) { rec.push_back((double)cell); for (int q = 0; q < 3; q++) rec.push_back(c->hC[3 * (size_t)cell + q]); }
        for (int b : pb[k]) { rec.push_back((double)b); for (int q = 0; q < 3; q++) rec.push_back(c->hCf[3 * (size_t)(F + b) + q]); }
    }
    std::vector<std::vector<char>> all;
    int rc = s4f_allgatherv_host(c, rec.data(), rec.size() * sizeof(double), all); if (rc) return rc;
    // 2. match by coordinates
    struct Key { unsigned long long a, b, d; bool operator<(const Key& o) const { return a != o.a ? a < o.a : (b != o.b ? b < o.b : d < o.d); } };
    auto keyOf = [](const double* x) { Key k; std::memcpy(&k.a, x, 8); std::memcpy(&k.b, x + 1, 8); std::memcpy(&k.d, x + 2, 8); return k; };
    std::map<Key, int> mine;
    for (size_t k = 0; k < procPts.size(); k++) {
        const double x[3] = {c->hPoints[3 * (size_t)procPts[k]] + 0.0, c->hPoints[3 * (size_t)procPts[k] + 1] + 0.0, c->hPoints[3 * (size_t)procPts[k] + 2] + 0.0};
        mine[keyOf(x)] = (int)k;
    }
    struct Ext { int rank, id; bool isB; double ctr[3]; };
    std::vector<std::vector<Ext>> ext(procPts.size());
    for (int r = 0; r < c->nRanks; r++) {
        if (r == c->rank) continue;
        const double* d = (const double*)all[r].data();
        const size_t nd = all[r].size() / sizeof(double);
        for (size_t i = 0; i < nd;) {
            const int nC = (int)d[i + 3], nB = (int)d[i + 4];
            auto it = mine.find(keyOf(d + i));
            if (it != mine.end()) {
                const int k = it->second, p = procPts[k];
                if (d[i + 5] != 0 || d[i + 6] != 0 || d[i + 7] != 0) for (int q = 0; q < 3; q++) c->extSymN[3 * (size_t)p + q] = d[i + 5 + q];
                if ((int)d[i + 8] >= 0) c->extFixAxis[p] = (int)d[i + 8];
                for (int j = 0; j < nC + nB; j++) {
                    const double* e = d + i + 9 + 4 * (size_t)j;
                    Ext x; x.rank = r; x.id = (int)e[0]; x.isB = j >= nC; x.ctr[0] = e[1]; x.ctr[1] = e[2]; x.ctr[2] = e[3];
                    ext[k].push_back(x);
                }
            }
            i += 9 + 4 * (size_t)(nC + nB);
        }
    }
    for (size_t k = 0; k < procPts.size(); k++)
        if (ext[k].empty()) { c->err = "point ghosts: a processor-patch point has no counterpart on any other rank (the points of processor patches must coincide bit for bit)"; return 1; }
    // 3. slots: per source rank its cells (ascending), then its boundary faces (ascending)
    std::vector<std::vector<int>> needC(c->nRanks), needB(c->nRanks);
    for (auto& v : ext) for (auto& x : v) (x.isB ? needB : needC)[x.rank].push_back(x.id);
    std::vector<int> baseC(c->nRanks, 0), baseB(c->nRanks, 0);
    int X = 0;
    std::vector<int> req;                                   // [dest, nC, nB, ids...] per source rank
    for (int r = 0; r < c->nRanks; r++) {
        for (auto* v : {&needC[r], &needB[r]}) { std::sort(v->begin(), v->end()); v->erase(std::unique(v->begin(), v->end()), v->end()); }
        baseC[r] = X; X += (int)needC[r].size();
        baseB[r] = X; X += (int)needB[r].size();
        if (needC[r].empty() && needB[r].empty()) continue;
        req.push_back(r); req.push_back((int)needC[r].size()); req.push_back((int)needB[r].size());
        req.insert(req.end(), needC[r].begin(), needC[r].end()); req.insert(req.end(), needB[r].begin(), needB[r].end());
    }
    std::vector<std::vector<char>> allReq;
    rc = s4f_allgatherv_host(c, req.data(), req.size() * sizeof(int), allReq); if (rc) return rc;
    std::vector<std::vector<int>> sendTo(c->nRanks);         // index-space positions of what rank r wants from me
    for (int r = 0; r < c->nRanks; r++) {
        if (r == c->rank) continue;
        const int* q = (const int*)allReq[r].data();
        const size_t nq = allReq[r].size() / sizeof(int);
        for (size_t i = 0; i < nq;) {
            const int dest = q[i], nC = q[i + 1], nB = q[i + 2];
            if (dest == c->rank) {
                for (int j = 0; j < nC; j++) {
                    if (q[i + 3 + j] < 0 || q[i + 3 + j] >= N) { c->err = "point ghosts: a requested cell is out of range"; return 1; }
                    sendTo[r].push_back(q[i + 3 + j]);
                }
                for (int j = 0; j < nB; j++) {
                    if (q[i + 3 + nC + j] < 0 || q[i + 3 + nC + j] >= B) { c->err = "point ghosts: a requested boundary face is out of range"; return 1; }
                    sendTo[r].push_back(bOff + q[i + 3 + nC + j]);
                }
            }
            i += 3 + (size_t)nC + nB;
        }
    }
    std::vector<int> nbrRank, sendCount, recvCount, sendCells;
    for (int r = 0; r < c->nRanks; r++) {
        const int nr = (int)(needC[r].size() + needB[r].size()), ns = (int)sendTo[r].size();
        if (nr == 0 && ns == 0) continue;
        nbrRank.push_back(r); sendCount.push_back(ns); recvCount.push_back(nr);
        sendCells.insert(sendCells.end(), sendTo[r].begin(), sendTo[r].end());
    }
    rc = s4f_halo_plan_create_asym(c, nbrRank, sendCount, recvCount, sendCells, 9, &c->haloX); if (rc) return rc;
    // per local point: the slots of its remote cells / faces
    c->X = X;
    for (int p = 0; p < nP; p++) {
        c->extPtr[p] = (int)c->extSlot.size();
        const int k = procIdx[p];
        if (k < 0) continue;
        for (const Ext& x : ext[k]) {
            const std::vector<int>& v = x.isB ? needB[x.rank] : needC[x.rank];
            const int pos = (int)(std::lower_bound(v.begin(), v.end(), x.id) - v.begin());
            const int slot = xOff + (x.isB ? baseB : baseC)[x.rank] + pos;
            if (std::find(c->extSlot.begin() + c->extPtr[p], c->extSlot.end(), slot) != c->extSlot.end()) continue;
            c->extSlot.push_back(slot); c->extIsB.push_back(x.isB ? 1 : 0);
            for (int q = 0; q < 3; q++) c->extCtr.push_back(x.ctr[q]);
        }
    }
    c->extPtr[nP] = (int)c->extSlot.size();
    c->graphSerial++;
    return 0;
}

int s4f_build_rows(s4fgpu_ctx* c) {
    const int N = c->N, F = c->F, B = c->B;
    // ---- ghosts: one per processor-patch face, in patch order ----
    c->ghostOfFace.assign(B, -1);
    c->nbrs.clear();
    int G = 0;
    std::vector<int> sendCells;
    for (int p = 0; p < c->nPatches; p++) {
        if (c->pKind[p] != S4F_PATCH_PROCESSOR) continue;
        s4fgpu_ctx::Nbr nb;
        nb.rank = c->pNbr[p]; nb.patch = p; nb.count = c->pSize[p]; nb.sendOff = G; nb.ghostOff = N + G;
        for (int i = 0; i < c->pSize[p]; i++) {
            int b = c->pStart[p] + i;
            c->ghostOfFace[b] = N + G + i;
            sendCells.push_back(c->faceCells[b]);
        }
        G += c->pSize[p];
        c->nbrs.push_back(nb);
    }
    c->G = G;
    const int bOff = N + G;
    c->ld = ((N + G + B + c->X + 31) / 32) * 32;           // X: point-neighbour ghosts of a decomposed mesh (s4f_build_point_ghosts)
    if (c->ld == 0) c->ld = 32;

    // ---- rows: lower neighbours, upper neighbours, boundary/processor faces ----
    std::vector<int> cnt(N, 0);
    for (int f = 0; f < F; f++) { cnt[c->own[f]]++; cnt[c->nei[f]]++; }
    for (int b = 0; b < B; b++) cnt[c->faceCells[b]]++;
    std::vector<long long> rowPtr(N + 1, 0);
    for (int i = 0; i < N; i++) rowPtr[i + 1] = rowPtr[i] + cnt[i];
    const long long nnz = rowPtr[N];
    std::vector<int> rCol(nnz), rFace(nnz);
    std::vector<signed char> rSign(nnz);
    std::vector<long long> cur(rowPtr.begin(), rowPtr.end() - 1);
    for (int f = 0; f < F; f++) { long long e = cur[c->nei[f]]++; rCol[e] = c->own[f]; rFace[e] = f; rSign[e] = -1; }
    for (int f = 0; f < F; f++) { long long e = cur[c->own[f]]++; rCol[e] = c->nei[f]; rFace[e] = f; rSign[e] = +1; }
    for (int b = 0; b < B; b++) {
        long long e = cur[c->faceCells[b]]++;
        rCol[e] = (c->ghostOfFace[b] >= 0) ? c->ghostOfFace[b] : bOff + b;
        rFace[e] = F + b; rSign[e] = +1;
    }
    c->nnzOff = 2LL * F + G;

    // ---- SELL-32 ----
    const int nSlices = (N + 31) / 32;
    c->nSlices = nSlices;
    std::vector<int> slicePtr(nSlices + 1, 0);
    for (int s = 0; s < nSlices; s++) {
        int w = 0;
        for (int r = s * 32; r < std::min(N, s * 32 + 32); r++) w = std::max(w, cnt[r]);
        long long next = (long long)slicePtr[s] + 32LL * w;
        if (next > 2147483647LL) { c->err = "mesh too large for int32 entry offsets"; return 1; }
        slicePtr[s + 1] = (int)next;
    }
    const long long nE = slicePtr[nSlices];
    c->nEntries = nE;

    // least-squares dd tensor per cell
    const double* C = c->hC.data();
    auto otherPoint = [&](long long e, int P, double* d) {
        int f = rFace[e];
        const double* X;
        if (f < F) X = &C[3 * (size_t)rCol[e]];
        else X = &c->hCnbrB[3 * (size_t)(f - F)];     // Cf on ordinary patches, neighbour centre on processor faces
        d[0] = X[0] - C[3 * (size_t)P]; d[1] = X[1] - C[3 * (size_t)P + 1]; d[2] = X[2] - C[3 * (size_t)P + 2];
    };
    std::vector<double> invDd(6 * (size_t)N);
    for (int P = 0; P < N; P++) {
        double t[6] = {0, 0, 0, 0, 0, 0};
        for (long long e = rowPtr[P]; e < rowPtr[P + 1]; e++) {
            double d[3]; otherPoint(e, P, d);
            double r = 1.0 / (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            t[0] += r * d[0] * d[0]; t[1] += r * d[0] * d[1]; t[2] += r * d[0] * d[2];
            t[3] += r * d[1] * d[1]; t[4] += r * d[1] * d[2]; t[5] += r * d[2] * d[2];
        }
        // [OF-ext] inv(symmTensorField): regularise the empty directions of 2-D cases
        if (!c->solD[0]) t[0] += 1; if (!c->solD[1]) t[3] += 1; if (!c->solD[2]) t[5] += 1;
        double r[6]; invSymm(t, r);
        if (!c->solD[0]) r[0] -= 1; if (!c->solD[1]) r[3] -= 1; if (!c->solD[2]) r[5] -= 1;
        for (int k = 0; k < 6; k++) invDd[6 * (size_t)P + k] = r[k];
    }

    std::vector<int> hCol(nE), hFaceEntry(std::max(F, 1), 0), hProcEntry(std::max(G, 1), 0);
    std::vector<double> hW(nE, 1.0), hSf(3 * nE, 0.0), hLs(3 * nE, 0.0), hDn(nE, 0.0), hCorr;
    bool nonOrth = false;
    for (size_t i = 0; i < c->hCorr.size(); i++) if (std::fabs(c->hCorr[i]) > 1e-12) { nonOrth = true; break; }
    c->nonOrth = nonOrth;
    if (nonOrth) hCorr.assign(3 * nE, 0.0);
    for (int s = 0; s < nSlices; s++) {
        const int width = (slicePtr[s + 1] - slicePtr[s]) / 32;
        for (int lane = 0; lane < 32; lane++) {
            const int P = s * 32 + lane;
            for (int k = 0; k < width; k++) {
                const long long E = (long long)slicePtr[s] + 32LL * k + lane;
                if (P >= N) { hCol[E] = 0; continue; }
                if (k >= cnt[P]) { hCol[E] = P; continue; }        // padding
                const long long e = rowPtr[P] + k;
                const int f = rFace[e];
                const double sg = rSign[e];
                hCol[E] = rCol[e];
                if (f < F && sg > 0) hFaceEntry[f] = (int)E;
                if (f >= F && c->ghostOfFace[f - F] >= 0) hProcEntry[c->ghostOfFace[f - F] - N] = (int)E;
                const bool bnd = (f >= F) && (c->ghostOfFace[f - F] < 0);
                if (bnd) hW[E] = 0.0;
                else hW[E] = (sg > 0) ? c->hW[f] : 1.0 - c->hW[f];
                for (int q = 0; q < 3; q++) hSf[(size_t)q * nE + E] = sg * c->hSf[3 * (size_t)f + q];
                if (!bnd) {
                    hDn[E] = c->hMagSf[f] * c->hNod[f];
                    if (nonOrth) {
                        // correction vectors at round-off level (|corr| < 1e-13 of a unit vector) are orthogonal faces: exactly
                        // zero lets the right-hand side skip their grad(D) gather (k_source_g)
                        const double* cv = &c->hCorr[3 * (size_t)f];
                        const bool tiny = std::fabs(cv[0]) < 1e-13 && std::fabs(cv[1]) < 1e-13 && std::fabs(cv[2]) < 1e-13;
                        for (int q = 0; q < 3; q++) hCorr[(size_t)q * nE + E] = tiny ? 0.0 : sg * c->hMagSf[f] * cv[q];
                    }
                }
                double d[3]; otherPoint(e, P, d);
                const double r = 1.0 / (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                const double* iv = &invDd[6 * (size_t)P];
                hLs[0 * nE + E] = r * (iv[0] * d[0] + iv[1] * d[1] + iv[2] * d[2]);
                hLs[1 * nE + E] = r * (iv[1] * d[0] + iv[3] * d[1] + iv[4] * d[2]);
                hLs[2 * nE + E] = r * (iv[2] * d[0] + iv[4] * d[1] + iv[5] * d[2]);
            }
        }
    }
    S4F_CHECK_CUDA(c, c->slicePtr.upload(slicePtr));
    S4F_CHECK_CUDA(c, c->col.upload(hCol));
    S4F_CHECK_CUDA(c, c->faceEntry.upload(hFaceEntry));
    S4F_CHECK_CUDA(c, c->procEntry.upload(hProcEntry));
    S4F_CHECK_CUDA(c, c->eW.upload(hW));
    S4F_CHECK_CUDA(c, c->eSf.upload(hSf));
    S4F_CHECK_CUDA(c, c->eLs.upload(hLs));
    S4F_CHECK_CUDA(c, c->eDn.upload(hDn));
    if (nonOrth) S4F_CHECK_CUDA(c, c->eCorr.upload(hCorr)); else c->eCorr.release();
    S4F_CHECK_CUDA(c, c->eA.alloc(nE)); S4F_CHECK_CUDA(c, c->eGam.alloc(nE));
    S4F_CHECK_CUDA(c, c->eU.alloc(3 * (size_t)nE)); S4F_CHECK_CUDA(c, c->eC0.alloc(nE));
    if (nonOrth) S4F_CHECK_CUDA(c, c->eVc.alloc(3 * (size_t)nE)); else c->eVc.release();
    S4F_CHECK_CUDA(c, c->rowK.alloc(6 * (size_t)c->ld));

    // volumes
    std::vector<double> hV(c->ld, 1.0), hrV(c->ld, 1.0);
    for (int i = 0; i < N; i++) { hV[i] = c->hV[i]; hrV[i] = 1.0 / c->hV[i]; }
    S4F_CHECK_CUDA(c, c->V.upload(hV)); S4F_CHECK_CUDA(c, c->rV.upload(hrV));

    // ---- boundary faces ----
    std::vector<int> hFaceCell(std::max(B, 1), 0);
    std::vector<double> hN(3 * (size_t)std::max(B, 1), 0.0), hK(3 * (size_t)std::max(B, 1), 0.0), hBSf(3 * (size_t)std::max(B, 1), 0.0),
        hDelta(std::max(B, 1), 0.0), hMag(std::max(B, 1), 0.0);
    for (int b = 0; b < B; b++) {
        const int f = F + b, P = c->faceCells[b];
        hFaceCell[b] = P;
        double n[3], d[3], nd = 0;
        for (int q = 0; q < 3; q++) { n[q] = c->hSf[3 * (size_t)f + q] / c->hMagSf[f]; d[q] = c->hCf[3 * (size_t)f + q] - C[3 * (size_t)P + q]; nd += n[q] * d[q]; }
        for (int q = 0; q < 3; q++) {
            hN[(size_t)q * B + b] = n[q];
            hK[(size_t)q * B + b] = d[q] - n[q] * nd;           // patchCorrectionVectors.C:24-36
            hBSf[(size_t)q * B + b] = c->hSf[3 * (size_t)f + q];
        }
        hDelta[b] = c->hNod[f]; hMag[b] = c->hMagSf[f];
    }
    S4F_CHECK_CUDA(c, c->bFaceCell.upload(hFaceCell));
    S4F_CHECK_CUDA(c, c->bN.upload(hN)); S4F_CHECK_CUDA(c, c->bK.upload(hK)); S4F_CHECK_CUDA(c, c->bSf.upload(hBSf));
    S4F_CHECK_CUDA(c, c->bDelta.upload(hDelta)); S4F_CHECK_CUDA(c, c->bMagSf.upload(hMag));
    if (!c->geomSet || c->bcValue.n != 3 * (size_t)std::max(B, 1)) {      // a geometry refresh (mesh motion) keeps the boundary data
        S4F_CHECK_CUDA(c, c->bcValue.alloc(3 * (size_t)std::max(B, 1))); S4F_CHECK_CUDA(c, c->bcPressure.alloc(std::max(B, 1)));
        S4F_CHECK_CUDA(c, c->tracGrad.alloc(3 * (size_t)std::max(B, 1))); S4F_CHECK_CUDA(c, c->bSn.alloc(3 * (size_t)std::max(B, 1)));
        S4F_CHECK_CUDA(c, c->bKind.alloc(std::max(B, 1)));
    }
    c->hCfB.assign(c->hCf.begin() + 3 * (size_t)F, c->hCf.end());
    c->hBSfHost.assign(c->hSf.begin() + 3 * (size_t)F, c->hSf.end());

    // boundary cells and their (non-processor) faces, ascending face order
    {
        std::vector<int> nb(N, 0);
        for (int b = 0; b < B; b++) if (c->ghostOfFace[b] < 0) nb[c->faceCells[b]]++;
        std::vector<int> cells, ptr(1, 0), faces;
        std::vector<int> slot(N, -1);
        for (int i = 0; i < N; i++) if (nb[i] > 0) { slot[i] = (int)cells.size(); cells.push_back(i); ptr.push_back(ptr.back() + nb[i]); }
        faces.resize(ptr.back());
        std::vector<int> cur2(ptr.begin(), ptr.end() - 1);
        for (int b = 0; b < B; b++) if (c->ghostOfFace[b] < 0) faces[cur2[slot[c->faceCells[b]]]++] = b;
        c->nBCells = (int)cells.size();
        if (cells.empty()) { cells.push_back(0); faces.push_back(0); }
        S4F_CHECK_CUDA(c, c->bcCells.upload(cells)); S4F_CHECK_CUDA(c, c->bcPtr.upload(ptr)); S4F_CHECK_CUDA(c, c->bcFaces.upload(faces));
    }
    // halo
    if (G > 0) S4F_CHECK_CUDA(c, c->sendCells.upload(sendCells));
    if (c->nRanks > 1) {       // collective: every rank (also one without processor patches) builds its plan here
        s4f_halo_plan_destroy(c->halo0); c->halo0 = nullptr;
        std::vector<int> nbrRank, nbrCount;
        for (const auto& nb : c->nbrs) { nbrRank.push_back(nb.rank); nbrCount.push_back(nb.count); }
        int rc = s4f_halo_plan_create(c, nbrRank, nbrCount, sendCells, 9, &c->halo0); if (rc) return rc;
        c->graphSerial++;
    }
    return 0;
}

int s4f_alloc_fields(s4fgpu_ctx* c) {
    const size_t ld = c->ld;
    if (c->D.n == 3 * ld && c->gradD.n == 9 * ld) return 0;
    auto A = [&](DevBuf<double>& b, int nc) { return b.alloc(nc * ld); };
    S4F_CHECK_CUDA(c, A(c->D, 3)); S4F_CHECK_CUDA(c, A(c->Dprev, 3)); S4F_CHECK_CUDA(c, A(c->Dold, 3)); S4F_CHECK_CUDA(c, A(c->DoldOld, 3));
    S4F_CHECK_CUDA(c, A(c->gradD, 9)); S4F_CHECK_CUDA(c, A(c->gradDold, 9));
    S4F_CHECK_CUDA(c, A(c->sigma, 6)); S4F_CHECK_CUDA(c, A(c->sigmaOld, 6));
    S4F_CHECK_CUDA(c, A(c->impK, 1));
    S4F_CHECK_CUDA(c, A(c->diag0, 1)); S4F_CHECK_CUDA(c, A(c->diagC, 3)); S4F_CHECK_CUDA(c, A(c->rDiagC, 3)); S4F_CHECK_CUDA(c, A(c->source, 3));
    S4F_CHECK_CUDA(c, A(c->pA, 3)); S4F_CHECK_CUDA(c, A(c->wA, 3)); S4F_CHECK_CUDA(c, A(c->rA, 3));
    S4F_CHECK_CUDA(c, c->pcgS.alloc(1)); S4F_CHECK_CUDA(c, c->outS.alloc(1));
    S4F_CHECK_CUDA(c, c->partials.alloc(32 * 4096)); S4F_CHECK_CUDA(c, c->ticket.alloc(8));
    S4F_CHECK_CUDA(c, c->ones3.upload(std::vector<int>{1, 1, 1}));
    if (!c->hPcgS) S4F_CHECK_CUDA(c, cudaMallocHost((void**)&c->hPcgS, sizeof(PcgScalars)));
    if (!c->hOutS) S4F_CHECK_CUDA(c, cudaMallocHost((void**)&c->hOutS, sizeof(OuterScalars)));
    return 0;
}

// fields that only finite-strain models / plastic laws need; called from set_law / set_controls
int s4f_alloc_model_fields(s4fgpu_ctx* c) {
    const size_t ld = c->ld;
    const bool TL = c->ctlSet && c->finiteStrain();
    const int kind = c->law.kind;
    auto A = [&](DevBuf<double>& b, int nc) { return (b.n == nc * ld) ? cudaSuccess : b.alloc(nc * ld); };
    auto fillI = [&](DevBuf<double>& b, int nc, const int* diagIdx, int nd) {
        for (int i = 0; i < nd; i++) {
            k_fill<<<(unsigned)((ld + 255) / 256), 256, 0, c->stream>>>(b.p + (size_t)diagIdx[i] * ld, 1.0, (long long)ld);
            c->launches++;
        }
    };
    const int dT[3] = {0, 4, 8}, dS[3] = {0, 3, 5}, d1[1] = {0};
    if (c->ctlSet && c->incremental()) { S4F_CHECK_CUDA(c, A(c->Dtot, 3)); S4F_CHECK_CUDA(c, A(c->gradDtot, 9)); }
    if (c->ctlSet && c->ctl.d2dt2Scheme != S4F_D2DT2_STEADY_STATE) { S4F_CHECK_CUDA(c, A(c->d2Hist, 3)); c->histValid = false; }
    if (c->ctlSet && c->ctl.d2dt2Scheme == S4F_D2DT2_BACKWARD) { S4F_CHECK_CUDA(c, A(c->Dooo, 3)); S4F_CHECK_CUDA(c, A(c->Doooo, 3)); }
    if (c->ctlSet && c->UL()) {
        S4F_CHECK_CUDA(c, A(c->d2Hist, 3)); c->histValid = false;
        S4F_CHECK_CUDA(c, A(c->Dooo, 3)); S4F_CHECK_CUDA(c, A(c->Doooo, 3)); S4F_CHECK_CUDA(c, A(c->Dooooo, 3));
        S4F_CHECK_CUDA(c, A(c->DDo, 3)); S4F_CHECK_CUDA(c, A(c->DDoo, 3)); S4F_CHECK_CUDA(c, A(c->DDooo, 3)); S4F_CHECK_CUDA(c, A(c->DDoooo, 3));
        if (c->rhoF.n != ld) {
            S4F_CHECK_CUDA(c, A(c->rhoF, 1)); S4F_CHECK_CUDA(c, A(c->rhoO, 1)); S4F_CHECK_CUDA(c, A(c->rhoOO, 1));
            c->rhoInit = false;
        }
        if (c->lawSet && !c->rhoInit) {      // rho_(mechanical().rho()), nonLinGeomUpdatedLagSolid.C:124-135
            for (DevBuf<double>* b : {&c->rhoF, &c->rhoO, &c->rhoOO}) {
                k_fill<<<(unsigned)((ld + 255) / 256), 256, 0, c->stream>>>(b->p, c->law.rho, (long long)ld);
                c->launches++;
            }
            c->rhoInit = true;
        }
    }
    if (c->ctlSet && !TL) S4F_CHECK_CUDA(c, A(c->T9, 9));      // lin-geom: the combined tensor M of the factored right-hand side
    if (TL) {
        if (c->Finv.n != 9 * ld) { S4F_CHECK_CUDA(c, A(c->Finv, 9)); fillI(c->Finv, 9, dT, 3); }
        if (c->Jt.n != ld) { S4F_CHECK_CUDA(c, A(c->Jt, 1)); fillI(c->Jt, 1, d1, 1); }
        S4F_CHECK_CUDA(c, A(c->T9, 9));
    }
    if (c->lawSet && kind != S4F_LAW_LINEAR_ELASTIC) {
        if (kind != S4F_LAW_LINEAR_ELASTIC_MISES_PLASTIC) {
            if (c->lawF.n != 9 * ld) { S4F_CHECK_CUDA(c, A(c->lawF, 9)); fillI(c->lawF, 9, dT, 3); S4F_CHECK_CUDA(c, A(c->lawFold, 9)); fillI(c->lawFold, 9, dT, 3); }
            if (c->lawJ.n != ld) { S4F_CHECK_CUDA(c, A(c->lawJ, 1)); fillI(c->lawJ, 1, d1, 1); S4F_CHECK_CUDA(c, A(c->lawJold, 1)); fillI(c->lawJold, 1, d1, 1); }
        }
        if (kind == S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC || kind == S4F_LAW_LINEAR_ELASTIC_MISES_PLASTIC) {
            if (c->bEbar.n != 6 * ld) {
                S4F_CHECK_CUDA(c, A(c->bEbar, 6)); fillI(c->bEbar, 6, dS, 3); S4F_CHECK_CUDA(c, A(c->bEbarOld, 6)); fillI(c->bEbarOld, 6, dS, 3);
            }
            S4F_CHECK_CUDA(c, A(c->sigmaY, 1)); S4F_CHECK_CUDA(c, A(c->sigmaYOld, 1)); S4F_CHECK_CUDA(c, A(c->DSigmaY, 1));
            S4F_CHECK_CUDA(c, A(c->epsPEq, 1)); S4F_CHECK_CUDA(c, A(c->epsPEqOld, 1)); S4F_CHECK_CUDA(c, A(c->DEpsPEq, 1));
            S4F_CHECK_CUDA(c, A(c->epsP, 6)); S4F_CHECK_CUDA(c, A(c->epsPOld, 6)); S4F_CHECK_CUDA(c, A(c->DEpsP, 6)); S4F_CHECK_CUDA(c, A(c->DEpsPprev, 6));
            S4F_CHECK_CUDA(c, A(c->DLambda, 1)); S4F_CHECK_CUDA(c, A(c->plasticN, 6)); S4F_CHECK_CUDA(c, A(c->epsilon, 6));
        }
    }
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// vol -> point interpolation (enhancedVolPointInterpolation, src/blockCoupledSolids4FoamTools): per point a list of
// source slots in the vol-field index space -- the cells around an internal point, the boundary-value slots of the
// patch faces around a patch point -- with normalised inverse-distance weights
// (enhancedVolPointInterpolation.C:165-198 makeInternalWeights, :201-245 makeBoundaryWeights); symmetry-plane points
// carry the plane normal for the point constraint.  One thread per point, atomic-free gather.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void k_vol_to_point(const int* __restrict__ ptPtr, const int* __restrict__ ptCol, const double* __restrict__ ptW,
                               const double* __restrict__ ptN, const double* __restrict__ X, double* __restrict__ out, int nPoints, int ld) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nPoints) return;
    double a[3] = {0, 0, 0};
    for (int j = ptPtr[p]; j < ptPtr[p + 1]; j++) {
        const int s = ptCol[j];
        const double w = ptW[j];
        a[0] += w * X[s]; a[1] += w * X[(size_t)ld + s]; a[2] += w * X[2 * (size_t)ld + s];
    }
    const double n[3] = {ptN[3 * (size_t)p], ptN[3 * (size_t)p + 1], ptN[3 * (size_t)p + 2]};
    const double na = n[0] * a[0] + n[1] * a[1] + n[2] * a[2];
    out[3 * (size_t)p] = a[0] - n[0] * na; out[3 * (size_t)p + 1] = a[1] - n[1] * na; out[3 * (size_t)p + 2] = a[2] - n[2] * na;
}
// interpolate(vf, gradVf, pf), enhancedVolPointInterpolate.C:351-418: pf = sum w (vf + delta & gradVf), w normalised
__global__ void k_vol_to_point_grad(const int* __restrict__ ptr, const int* __restrict__ colc, const double* __restrict__ w,
                                    const double* __restrict__ delta, const double* __restrict__ X, const double* __restrict__ G,
                                    double* __restrict__ out, int nPoints, int ld, long long nnz) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nPoints) return;
    double a[3] = {0, 0, 0};
    for (int j = ptr[p]; j < ptr[p + 1]; j++) {
        const int s = colc[j];
        const double wj = w[j], d0 = delta[j], d1 = delta[nnz + j], d2 = delta[2 * nnz + j];
#pragma unroll
        for (int q = 0; q < 3; q++)
            a[q] += wj * (X[(size_t)q * ld + s] + d0 * G[(size_t)q * ld + s] + d1 * G[(size_t)(3 + q) * ld + s] + d2 * G[(size_t)(6 + q) * ld + s]);
    }
    out[3 * (size_t)p] = a[0]; out[3 * (size_t)p + 1] = a[1]; out[3 * (size_t)p + 2] = a[2];
}
}  // namespace

// [OF-ext] LeastSquaresVectors<centredCPCCellToCellStencilObject> ("pointCellsLeastSquares"): per cell the cells sharing a
// point with it and the boundary faces at its points, 1/|d|^2 weights, ls = (inv(dd) - dd0) & d/|d|^2 -- see the oracle's
// makePointCellsStencil for the restatement.  Laid out as SELL-32 rows of their own (about 26 entries per hex cell) that the
// gradient kernels gather over exactly as they do over the face rows.
int s4f_build_point_stencil(s4fgpu_ctx* c) {
    const int N = c->N, F = c->F, B = c->B, nP = c->nPoints, bOff = c->bOff();
    if (nP == 0) { c->err = "pointCellsLeastSquares needs the mesh points (s4fgpu_set_points)"; return 1; }
    if (c->hostGeomStale) { int rg = s4f_refresh_host_geometry(c); if (rg) return rg; }
    if (c->nRanks > 1 && c->extPtr.size() != (size_t)nP + 1) {
        c->err = "pointCellsLeastSquares on a decomposed mesh: call s4fgpu_set_points before s4fgpu_set_geometry (the point-neighbour ghosts are laid out with the rows)";
        return 1;
    }
    const int xOff = c->xOff();
    const bool ext = c->nRanks > 1;
    // stencil entries are positions in the field index space: cell < N, boundary slot bOff + b, remote cell/face >= xOff
    std::vector<std::vector<int>> pc(nP), pb(nP), cellPts(N);
    auto add = [](std::vector<int>& v, int x) { if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x); };
    for (int f = 0; f < F + B; f++) for (int j = c->hFvPtr[f]; j < c->hFvPtr[f + 1]; j++) {
        const int p = c->hFv[j];
        const int o = f < F ? c->own[f] : c->faceCells[f - F];
        add(pc[p], o); add(cellPts[o], p);
        if (f < F) { add(pc[p], c->nei[f]); add(cellPts[c->nei[f]], p); }
    }
    for (int ip = 0; ip < c->nPatches; ip++) {
        if (c->pKind[ip] == S4F_PATCH_PROCESSOR) continue;
        for (int i = 0; i < c->pSize[ip]; i++) {
            const int b = c->pStart[ip] + i;
            for (int j = c->hFvPtr[F + b]; j < c->hFvPtr[F + b + 1]; j++) pb[c->hFv[j]].push_back(bOff + b);
        }
    }
    std::vector<double> xCtr;                       // centre of every extra slot (a slot may be listed at several points)
    if (ext) {
        xCtr.assign(3 * (size_t)std::max(c->X, 1), 0.0);
        for (size_t e = 0; e < c->extSlot.size(); e++) for (int q = 0; q < 3; q++) xCtr[3 * (size_t)(c->extSlot[e] - xOff) + q] = c->extCtr[3 * e + q];
    }
    std::vector<int> rowPtr(N + 1, 0), slot; std::vector<double> ls[3];
    for (int i = 0; i < N; i++) {
        std::vector<int> st;
        for (int p : cellPts[i]) {
            for (int cc : pc[p]) if (cc != i) add(st, cc);
            for (int b : pb[p]) add(st, b);
            if (ext) for (int e = c->extPtr[p]; e < c->extPtr[p + 1]; e++) add(st, c->extSlot[e]);
        }
        std::sort(st.begin(), st.end());
        double dd[6] = {0, 0, 0, 0, 0, 0};
        if (!c->solD[0]) dd[0] += 1; if (!c->solD[1]) dd[3] += 1; if (!c->solD[2]) dd[5] += 1;
        std::vector<double> dl(3 * st.size());
        const double* Ci = &c->hC[3 * (size_t)i];
        for (size_t k = 0; k < st.size(); k++) {
            const double* x = st[k] < N ? &c->hC[3 * (size_t)st[k]] : (st[k] >= xOff ? &xCtr[3 * (size_t)(st[k] - xOff)] : &c->hCfB[3 * (size_t)(st[k] - bOff)]);
            const double d[3] = {x[0] - Ci[0], x[1] - Ci[1], x[2] - Ci[2]};
            const double r = 1.0 / (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            dd[0] += r * d[0] * d[0]; dd[1] += r * d[0] * d[1]; dd[2] += r * d[0] * d[2];
            dd[3] += r * d[1] * d[1]; dd[4] += r * d[1] * d[2]; dd[5] += r * d[2] * d[2];
            for (int q = 0; q < 3; q++) dl[3 * k + q] = r * d[q];
        }
        double iv[6]; invSymm(dd, iv);
        if (!c->solD[0]) iv[0] -= 1; if (!c->solD[1]) iv[3] -= 1; if (!c->solD[2]) iv[5] -= 1;
        for (size_t k = 0; k < st.size(); k++) {
            const double* d = &dl[3 * k];
            slot.push_back(st[k]);
            ls[0].push_back(iv[0] * d[0] + iv[1] * d[1] + iv[2] * d[2]);
            ls[1].push_back(iv[1] * d[0] + iv[3] * d[1] + iv[4] * d[2]);
            ls[2].push_back(iv[2] * d[0] + iv[4] * d[1] + iv[5] * d[2]);
        }
        rowPtr[i + 1] = (int)slot.size();
    }
    const int nSlices = c->nSlices;
    std::vector<int> sp(nSlices + 1, 0);
    for (int s = 0; s < nSlices; s++) {
        int w = 0;
        for (int r = s * 32; r < std::min(N, s * 32 + 32); r++) w = std::max(w, rowPtr[r + 1] - rowPtr[r]);
        const long long next = (long long)sp[s] + 32LL * w;
        if (next > 2147483647LL) { c->err = "mesh too large for int32 entry offsets (pointCellsLeastSquares rows)"; return 1; }
        sp[s + 1] = (int)next;
    }
    const long long nE = sp[nSlices];
    std::vector<int> hc((size_t)std::max<long long>(nE, 1), 0); std::vector<double> hl(3 * (size_t)std::max<long long>(nE, 1), 0.0);
    for (int s = 0; s < nSlices; s++) {
        const int w = (sp[s + 1] - sp[s]) / 32;
        for (int lane = 0; lane < 32; lane++) {
            const int P = s * 32 + lane;
            for (int k = 0; k < w; k++) {
                const size_t E = (size_t)sp[s] + 32 * (size_t)k + lane;
                if (P >= N) { hc[E] = 0; continue; }
                if (k >= rowPtr[P + 1] - rowPtr[P]) { hc[E] = P; continue; }           // padding: zero vector, own column
                const size_t e = (size_t)rowPtr[P] + k;
                hc[E] = slot[e];
                for (int q = 0; q < 3; q++) hl[(size_t)q * nE + E] = ls[q][e];
            }
        }
    }
    S4F_CHECK_CUDA(c, c->gSlicePtr.upload(sp)); S4F_CHECK_CUDA(c, c->gCol.upload(hc)); S4F_CHECK_CUDA(c, c->gLs.upload(hl));
    c->gNE = nE; c->gValid = true;
    return 0;
}

int s4f_build_point_weights(s4fgpu_ctx* c, const double* points) {
    const int N = c->N, F = c->F, B = c->B, nP = c->nPoints, bOff = c->bOff(), xOff = c->xOff();
    if (c->hostGeomStale) { int rg = s4f_refresh_host_geometry(c); if (rg) return rg; }
    const bool ext = c->nRanks > 1;
    if (ext && c->extPtr.size() != (size_t)nP + 1) {
        c->err = "set_points on a decomposed mesh: call s4fgpu_set_points before s4fgpu_set_geometry (the point-neighbour ghosts are laid out with the rows)";
        return 1;
    }
    std::vector<std::vector<int>> pc(nP), pb(nP);
    auto add = [](std::vector<int>& v, int x) { if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x); };
    for (int f = 0; f < F + B; f++) for (int j = c->hFvPtr[f]; j < c->hFvPtr[f + 1]; j++) {
        const int p = c->hFv[j];
        if (p < 0 || p >= nP) { c->err = "set_points: vertex out of range"; return 1; }
        add(pc[p], f < F ? c->own[f] : c->faceCells[f - F]);
        if (f < F) add(pc[p], c->nei[f]);
    }
    for (int ip = 0; ip < c->nPatches; ip++) {
        if (c->pKind[ip] == S4F_PATCH_PROCESSOR) continue;
        for (int i = 0; i < c->pSize[ip]; i++) {
            const int b = c->pStart[ip] + i;
            for (int j = c->hFvPtr[F + b]; j < c->hFvPtr[F + b + 1]; j++) pb[c->hFv[j]].push_back(bOff + b);
        }
    }
    std::vector<double> hN; std::vector<int> fixAxis;
    point_symmetry(c, c->hBSfHost.data(), hN, fixAxis);
    std::vector<double> xCtr;                       // centre of every extra slot
    if (ext) {
        xCtr.assign(3 * (size_t)std::max(c->X, 1), 0.0);
        for (size_t e = 0; e < c->extSlot.size(); e++) for (int q = 0; q < 3; q++) xCtr[3 * (size_t)(c->extSlot[e] - xOff) + q] = c->extCtr[3 * e + q];
        for (int p = 0; p < nP; p++) {              // cells and boundary faces of the other ranks around the same point
            for (int e = c->extPtr[p]; e < c->extPtr[p + 1]; e++) (c->extIsB[e] ? pb : pc)[p].push_back(c->extSlot[e]);
            const double* n = &c->extSymN[3 * (size_t)p];
            if (hN[3 * (size_t)p] == 0 && hN[3 * (size_t)p + 1] == 0 && hN[3 * (size_t)p + 2] == 0) for (int q = 0; q < 3; q++) hN[3 * (size_t)p + q] = n[q];
            if (fixAxis[p] < 0) fixAxis[p] = c->extFixAxis[p];
        }
    }
    auto centre = [&](int slot) -> const double* {
        return slot < N ? &c->hC[3 * (size_t)slot] : (slot >= xOff ? &xCtr[3 * (size_t)(slot - xOff)] : &c->hCfB[3 * (size_t)(slot - bOff)]);
    };
    {   // gradient-extrapolated variant: all points from their pointCells
        std::vector<int> gptr(1, 0), gcol; std::vector<double> gw, gd[3];
        for (int p = 0; p < nP; p++) {
            const double* x = &points[3 * (size_t)p];
            std::sort(pc[p].begin(), pc[p].end());
            const size_t s0 = gw.size();
            double sw = 0;
            for (int cell : pc[p]) {
                const double* cc = centre(cell);
                const double d[3] = {x[0] - cc[0], x[1] - cc[1], x[2] - cc[2]};
                const double m = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                gcol.push_back(cell); gw.push_back(1.0 / m); sw += 1.0 / m;
                for (int q = 0; q < 3; q++) gd[q].push_back(d[q]);
            }
            for (size_t j = s0; j < gw.size(); j++) gw[j] /= sw;
            gptr.push_back((int)gw.size());
        }
        std::vector<double> gdel; gdel.reserve(3 * gw.size());
        for (int q = 0; q < 3; q++) gdel.insert(gdel.end(), gd[q].begin(), gd[q].end());
        if (gcol.empty()) { gcol.push_back(0); gw.push_back(0.0); gdel.assign(3, 0.0); }
        S4F_CHECK_CUDA(c, c->pgPtr.upload(gptr)); S4F_CHECK_CUDA(c, c->pgCol.upload(gcol)); S4F_CHECK_CUDA(c, c->pgW.upload(gw));
        S4F_CHECK_CUDA(c, c->pgDelta.upload(gdel));
    }
    // volPointInterpolation as the solid models use it: a point of a (non-empty, non-processor) patch takes the patch face
    // values around it, any other point the cell values around it; normalised inverse-distance weights
    std::vector<int> ptr(1, 0), col; std::vector<double> w;
    for (int p = 0; p < nP; p++) {
        const double* x = &points[3 * (size_t)p];
        const size_t s0 = w.size();
        double sw = 0;
        std::vector<int>& src = pb[p].empty() ? pc[p] : pb[p];
        std::sort(src.begin(), src.end());
        for (int sl : src) {
            const double* cc = centre(sl);
            const double d = std::sqrt((x[0] - cc[0]) * (x[0] - cc[0]) + (x[1] - cc[1]) * (x[1] - cc[1]) + (x[2] - cc[2]) * (x[2] - cc[2]));
            col.push_back(sl); w.push_back(1.0 / d); sw += 1.0 / d;
        }
        for (size_t j = s0; j < w.size(); j++) w[j] /= sw;
        ptr.push_back((int)w.size());
    }
    if (col.empty()) { col.push_back(0); w.push_back(0.0); }
    S4F_CHECK_CUDA(c, c->ptPtr.upload(ptr)); S4F_CHECK_CUDA(c, c->ptCol.upload(col)); S4F_CHECK_CUDA(c, c->ptW.upload(w));
    S4F_CHECK_CUDA(c, c->ptN.upload(hN));
    if (c->ptOut.n != 3 * (size_t)std::max(nP, 1)) S4F_CHECK_CUDA(c, c->ptOut.alloc(3 * (size_t)std::max(nP, 1)));
    // what the device-side mesh motion needs (s4f_geom.cu): the points, the face -> vertex CSR, the symmetry-plane constraint
    S4F_CHECK_CUDA(c, c->ptFixAxis.upload(fixAxis));
    S4F_CHECK_CUDA(c, c->dPoints.upload(std::vector<double>(points, points + 3 * (size_t)nP)));
    S4F_CHECK_CUDA(c, c->dFvPtr.upload(c->hFvPtr)); S4F_CHECK_CUDA(c, c->dFv.upload(c->hFv.empty() ? std::vector<int>(1, 0) : c->hFv));
    return 0;
}

// decomposed meshes: the values of the other ranks' cells and boundary faces at shared points (slots >= xOff)
int s4f_point_ghost_exchange(s4fgpu_ctx* c, double* field, int ncomp) {
    if (c->nRanks <= 1 || !c->haloX) return 0;
    return s4f_halo_run(c, c->haloX, field, c->ld, ncomp, c->xOff());
}

int s4f_interpolate_to_points(s4fgpu_ctx* c, const double* X, const double* G, double* hostOut) {
    const int nP = c->nPoints;
    int rx = s4f_point_ghost_exchange(c, const_cast<double*>(X), 3); if (rx) return rx;
    if (G && (rx = s4f_point_ghost_exchange(c, const_cast<double*>(G), 9))) return rx;
    if (G) k_vol_to_point_grad<<<(nP + 127) / 128, 128, 0, c->stream>>>(c->pgPtr.p, c->pgCol.p, c->pgW.p, c->pgDelta.p, X, G, c->ptOut.p, nP, c->ld, (long long)c->pgW.n);
    else k_vol_to_point<<<(nP + 127) / 128, 128, 0, c->stream>>>(c->ptPtr.p, c->ptCol.p, c->ptW.p, c->ptN.p, X, c->ptOut.p, nP, c->ld);
    c->launches++;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    S4F_CHECK_CUDA(c, cudaMemcpyAsync(hostOut, c->ptOut.p, 3 * (size_t)nP * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
