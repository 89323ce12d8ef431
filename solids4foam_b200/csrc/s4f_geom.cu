// s4f_geom.cu -- mesh motion on the device: solidModel::moveMesh (SM/solidModel/solidModel.C:2008-2148) as called at the
// end of every updated-Lagrangian time step (nonLinGeomUpdatedLagSolid::updateTotalFields, ...C:360-374), followed by
// what fvMesh::movePoints invalidates and OpenFOAM recomputes on demand:
//   primitiveMesh::makeFaceCentresAndAreas     -> k_geo_faces     (polygons: triangles about the vertex average)
//   primitiveMesh::makeCellCentresAndVols      -> k_geo_cells     (pyramids about the face-centre average)
//   surfaceInterpolation weights / nonOrthDeltaCoeffs / nonOrthCorrectionVectors, the least-squares vectors
//   (extendedLeastSquaresVectors.C:121-158, :229-272) and the patch correction vectors (patchCorrectionVectors.C:24-36)
//                                              -> k_geo_entries   (one row per lane over the SELL-32 rows)
//   the inverse-distance weights of enhancedVolPointInterpolation (…C:165-245)      -> k_geo_point_weights
// Round 1 recomputed all of this on the host and re-uploaded it (and re-agglomerated the GAMG hierarchy on one CPU
// thread); here the points never leave the device: the new geometry is a function of the point displacement that
// k_vol_to_point has just produced.  The GAMG hierarchy keeps its aggregates and re-sums its Galerkin coefficients
// (s4f_amg_refresh).  [OF-ext] formulas are restated in solids4foam_b200/mesh.py (the host mirror) and the oracle.
#include <algorithm>
#include <cmath>
#include <vector>

#include "s4f_comm.h"
#include "s4f_dev.cuh"

namespace {

__global__ void k_geo_move_points(double* __restrict__ pts, const double* __restrict__ dd, const int* __restrict__ fixAxis, int nP) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nP) return;
    const int ax = fixAxis[p];       // symmetry-plane points keep their plane (solidModel.C:2040-2080)
#pragma unroll
    for (int q = 0; q < 3; q++) if (q != ax) pts[3 * (size_t)p + q] += dd[3 * (size_t)p + q];
}

// face centre and area vector of polygon f: [OF-ext] primitiveMesh::makeFaceCentresAndAreas
__global__ void k_geo_faces(const int* __restrict__ fvPtr, const int* __restrict__ fv, const double* __restrict__ pts, double* __restrict__ fCtr,
                            double* __restrict__ fSf, int nF) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    const int j0 = fvPtr[f], n = fvPtr[f + 1] - j0;
    double fc[3] = {0, 0, 0};
    for (int j = 0; j < n; j++) { const double* p = pts + 3 * (size_t)fv[j0 + j]; fc[0] += p[0]; fc[1] += p[1]; fc[2] += p[2]; }
    fc[0] /= n; fc[1] /= n; fc[2] /= n;
    double sN[3] = {0, 0, 0}, sA = 0, sAc[3] = {0, 0, 0};
    for (int j = 0; j < n; j++) {
        const double* a0 = pts + 3 * (size_t)fv[j0 + j];
        const double* a1 = pts + 3 * (size_t)fv[j0 + (j + 1 == n ? 0 : j + 1)];
        const double e[3] = {a1[0] - a0[0], a1[1] - a0[1], a1[2] - a0[2]}, g[3] = {fc[0] - a0[0], fc[1] - a0[1], fc[2] - a0[2]};
        const double nn[3] = {e[1] * g[2] - e[2] * g[1], e[2] * g[0] - e[0] * g[2], e[0] * g[1] - e[1] * g[0]};
        const double a = sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
        sA += a;
#pragma unroll
        for (int q = 0; q < 3; q++) { sN[q] += nn[q]; sAc[q] += a * (a0[q] + a1[q] + fc[q]); }
    }
#pragma unroll
    for (int q = 0; q < 3; q++) {
        fCtr[(size_t)q * nF + f] = sA > 1e-300 ? sAc[q] / (3.0 * sA) : fc[q];
        fSf[(size_t)q * nF + f] = 0.5 * sN[q];
    }
}

// cell centre and volume: [OF-ext] primitiveMesh::makeCellCentresAndVols over the faces of the row (signed face list)
__global__ void __launch_bounds__(S4F_BLOCK) k_geo_cells(const int* __restrict__ slicePtr, const int* __restrict__ eFaceS, const double* __restrict__ fCtr,
                                                         const double* __restrict__ fSf, double* __restrict__ C, double* __restrict__ V,
                                                         double* __restrict__ rV, int N, int ld, int nF, int nSlices) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nSlices; s += nWarps) {
        const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5, row = s * 32 + lane;
        double ce[3] = {0, 0, 0}; int nf = 0;
        for (int k = 0; k < width; k++) {
            const int fs = eFaceS[base + 32 * k + lane];
            if (fs == 0) continue;
            const int f = abs(fs) - 1;
            ce[0] += fCtr[f]; ce[1] += fCtr[(size_t)nF + f]; ce[2] += fCtr[2 * (size_t)nF + f];
            nf++;
        }
        if (row >= N || nf == 0) continue;
        ce[0] /= nf; ce[1] /= nf; ce[2] /= nf;
        double vol = 0, ctr[3] = {0, 0, 0};
        for (int k = 0; k < width; k++) {
            const int fs = eFaceS[base + 32 * k + lane];
            if (fs == 0) continue;
            const int f = abs(fs) - 1;
            const double sg = fs > 0 ? 1.0 : -1.0;
            const double fc[3] = {fCtr[f], fCtr[(size_t)nF + f], fCtr[2 * (size_t)nF + f]};
            const double pyr = sg * (fSf[f] * (fc[0] - ce[0]) + fSf[(size_t)nF + f] * (fc[1] - ce[1]) + fSf[2 * (size_t)nF + f] * (fc[2] - ce[2]));
            vol += pyr;
#pragma unroll
            for (int q = 0; q < 3; q++) ctr[q] += pyr * (0.75 * fc[q] + 0.25 * ce[q]);
        }
#pragma unroll
        for (int q = 0; q < 3; q++) C[(size_t)q * ld + row] = ctr[q] / vol;
        V[row] = vol / 3.0; rV[row] = 3.0 / vol;
    }
}

__device__ __forceinline__ void inv_symm(const double* S, double* R) {
    const double d = S[0] * S[3] * S[5] + 2.0 * S[1] * S[4] * S[2] - S[0] * S[4] * S[4] - S[1] * S[1] * S[5] - S[2] * S[3] * S[2];
    R[0] = (S[3] * S[5] - S[4] * S[4]) / d; R[1] = (S[2] * S[4] - S[1] * S[5]) / d; R[2] = (S[1] * S[4] - S[2] * S[3]) / d;
    R[3] = (S[0] * S[5] - S[2] * S[2]) / d; R[4] = (S[1] * S[2] - S[0] * S[4]) / d; R[5] = (S[0] * S[3] - S[1] * S[1]) / d;
}

// per-entry interpolation geometry and least-squares vectors; boundary-face arrays where the entry is a boundary face
struct GeoOut {
    double *eW, *eSf, *eDn, *eCorr, *eLs;
    double *bN, *bK, *bSf, *bDelta, *bMagSf;
};
__global__ void __launch_bounds__(S4F_BLOCK) k_geo_entries(const int* __restrict__ slicePtr, const int* __restrict__ col, const int* __restrict__ eFaceS,
                                                           const double* __restrict__ fCtr, const double* __restrict__ fSf, const double* __restrict__ C,
                                                           GeoOut o, int N, int bOff, int B, int ld, int nF, long long nE, int nSlices, int sd0,
                                                           int sd1, int sd2) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nSlices; s += nWarps) {
        const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5, row = s * 32 + lane;
        const int r = row < N ? row : 0;
        const double CP[3] = {C[r], C[(size_t)ld + r], C[2 * (size_t)ld + r]};
        double t[6] = {0, 0, 0, 0, 0, 0};
        for (int k = 0; k < width; k++) {
            const long long e = (long long)base + 32 * k + lane;
            const int fs = eFaceS[e];
            if (fs == 0 || row >= N) {      // padding: zero coefficients, own column
                o.eW[e] = 1.0; o.eDn[e] = 0.0;
#pragma unroll
                for (int q = 0; q < 3; q++) { o.eSf[(size_t)q * nE + e] = 0.0; o.eLs[(size_t)q * nE + e] = 0.0; o.eCorr[(size_t)q * nE + e] = 0.0; }
                continue;
            }
            const int f = abs(fs) - 1, cc = col[e];
            const double sg = fs > 0 ? 1.0 : -1.0;
            const bool bnd = cc >= bOff;
            const double Sf[3] = {sg * fSf[f], sg * fSf[(size_t)nF + f], sg * fSf[2 * (size_t)nF + f]};
            const double fc[3] = {fCtr[f], fCtr[(size_t)nF + f], fCtr[2 * (size_t)nF + f]};
            double X[3];
            if (bnd) { X[0] = fc[0]; X[1] = fc[1]; X[2] = fc[2]; }
            else { X[0] = C[cc]; X[1] = C[(size_t)ld + cc]; X[2] = C[2 * (size_t)ld + cc]; }
            const double d[3] = {X[0] - CP[0], X[1] - CP[1], X[2] - CP[2]};
            const double magSf = sqrt(Sf[0] * Sf[0] + Sf[1] * Sf[1] + Sf[2] * Sf[2]);
            const double n[3] = {Sf[0] / magSf, Sf[1] / magSf, Sf[2] / magSf};
            const double magd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            const double nd = n[0] * d[0] + n[1] * d[1] + n[2] * d[2];
            const double nod = 1.0 / fmax(nd, 0.05 * magd);
#pragma unroll
            for (int q = 0; q < 3; q++) o.eSf[(size_t)q * nE + e] = Sf[q];
            if (!bnd && f >= nF - B) {       // processor-patch face: its normal and area also live in the boundary-face arrays (uns kernels)
                const int b = f - (nF - B);
#pragma unroll
                for (int q = 0; q < 3; q++) { o.bN[(size_t)q * B + b] = n[q]; o.bSf[(size_t)q * B + b] = Sf[q]; }
                o.bMagSf[b] = magSf; o.bDelta[b] = nod;
            }
            if (bnd) {
                const int b = cc - bOff;
                o.eW[e] = 0.0; o.eDn[e] = 0.0;
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    o.eCorr[(size_t)q * nE + e] = 0.0;
                    o.bN[(size_t)q * B + b] = n[q]; o.bK[(size_t)q * B + b] = d[q] - n[q] * nd; o.bSf[(size_t)q * B + b] = Sf[q];
                }
                o.bDelta[b] = nod; o.bMagSf[b] = magSf;
            } else {
                // weight of the row cell's own value: |Sf.(X - Cf)| / (|Sf.(Cf - C_P)| + |Sf.(X - Cf)|)   [OF-ext] surfaceInterpolation::makeWeights
                const double own = fabs(Sf[0] * (fc[0] - CP[0]) + Sf[1] * (fc[1] - CP[1]) + Sf[2] * (fc[2] - CP[2]));
                const double oth = fabs(Sf[0] * (X[0] - fc[0]) + Sf[1] * (X[1] - fc[1]) + Sf[2] * (X[2] - fc[2]));
                o.eW[e] = oth / (own + oth);
                o.eDn[e] = magSf * nod;
                double cv[3]; bool tiny = true;
#pragma unroll
                for (int q = 0; q < 3; q++) { cv[q] = n[q] - d[q] * nod; tiny = tiny && fabs(cv[q]) < 1e-13; }
#pragma unroll
                for (int q = 0; q < 3; q++) o.eCorr[(size_t)q * nE + e] = tiny ? 0.0 : magSf * cv[q];
            }
            const double rr = 1.0 / (magd * magd);
            t[0] += rr * d[0] * d[0]; t[1] += rr * d[0] * d[1]; t[2] += rr * d[0] * d[2];
            t[3] += rr * d[1] * d[1]; t[4] += rr * d[1] * d[2]; t[5] += rr * d[2] * d[2];
        }
        if (row >= N) continue;
        if (!sd0) t[0] += 1; if (!sd1) t[3] += 1; if (!sd2) t[5] += 1;     // [OF-ext] inv(symmTensorField): empty directions of 2-D cases
        double iv[6]; inv_symm(t, iv);
        if (!sd0) iv[0] -= 1; if (!sd1) iv[3] -= 1; if (!sd2) iv[5] -= 1;
        for (int k = 0; k < width; k++) {
            const long long e = (long long)base + 32 * k + lane;
            const int fs = eFaceS[e];
            if (fs == 0) continue;
            const int f = abs(fs) - 1, cc = col[e];
            double X[3];
            if (cc >= bOff) { X[0] = fCtr[f]; X[1] = fCtr[(size_t)nF + f]; X[2] = fCtr[2 * (size_t)nF + f]; }
            else { X[0] = C[cc]; X[1] = C[(size_t)ld + cc]; X[2] = C[2 * (size_t)ld + cc]; }
            const double d[3] = {X[0] - CP[0], X[1] - CP[1], X[2] - CP[2]};
            const double rr = 1.0 / (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            o.eLs[e] = rr * (iv[0] * d[0] + iv[1] * d[1] + iv[2] * d[2]);
            o.eLs[nE + e] = rr * (iv[1] * d[0] + iv[3] * d[1] + iv[4] * d[2]);
            o.eLs[2 * nE + e] = rr * (iv[2] * d[0] + iv[4] * d[1] + iv[5] * d[2]);
        }
    }
}

// vol->point weights (enhancedVolPointInterpolation.C:165-245): normalised inverse distances from the point to its donors
// (cell centres for internal points, boundary-face centres for patch points); and, for the gradient-extrapolated variant,
// delta = point - cell centre with weights 1/|delta|
__global__ void k_geo_point_weights(const int* __restrict__ ptPtr, const int* __restrict__ ptCol, double* __restrict__ ptW, const int* __restrict__ pgPtr,
                                    const int* __restrict__ pgCol, double* __restrict__ pgW, double* __restrict__ pgDelta, long long nnzG,
                                    const double* __restrict__ pts, const double* __restrict__ C, const double* __restrict__ fCtr, int nP, int bOff,
                                    int ld, int F, int nF) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nP) return;
    const double x[3] = {pts[3 * (size_t)p], pts[3 * (size_t)p + 1], pts[3 * (size_t)p + 2]};
    double sw = 0;
    for (int j = ptPtr[p]; j < ptPtr[p + 1]; j++) {
        const int s = ptCol[j];
        // the centre field holds cell centres, boundary-face centres in the boundary slots (k_geo_boundary_centres) and, on a
        // decomposed mesh, the centres of the other ranks' cells / boundary faces in the point-neighbour ghost slots
        const double X[3] = {C[s], C[(size_t)ld + s], C[2 * (size_t)ld + s]};
        const double w = 1.0 / sqrt((x[0] - X[0]) * (x[0] - X[0]) + (x[1] - X[1]) * (x[1] - X[1]) + (x[2] - X[2]) * (x[2] - X[2]));
        ptW[j] = w; sw += w;
    }
    for (int j = ptPtr[p]; j < ptPtr[p + 1]; j++) ptW[j] /= sw;
    sw = 0;
    for (int j = pgPtr[p]; j < pgPtr[p + 1]; j++) {
        const int s = pgCol[j];
        const double d[3] = {x[0] - C[s], x[1] - C[(size_t)ld + s], x[2] - C[2 * (size_t)ld + s]};
        const double w = 1.0 / sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        pgW[j] = w; sw += w;
        pgDelta[j] = d[0]; pgDelta[nnzG + j] = d[1]; pgDelta[2 * nnzG + j] = d[2];
    }
    for (int j = pgPtr[p]; j < pgPtr[p + 1]; j++) pgW[j] /= sw;
}

__global__ void k_geo_boundary_centres(const double* __restrict__ fCtr, double* __restrict__ C, int F, int B, int nF, int bOff, int ld) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
#pragma unroll
    for (int q = 0; q < 3; q++) C[(size_t)q * ld + bOff + b] = fCtr[(size_t)q * nF + F + b];
}

__global__ void k_soa3_to_aos(const double* __restrict__ soa, double* __restrict__ aos, int n, size_t stride, int offset) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int q = 0; q < 3; q++) aos[3 * (size_t)i + q] = soa[(size_t)q * stride + offset + i];
}

}  // namespace

// signed face of every SELL entry, in the row order of s4f_build_rows (lower neighbours, upper neighbours, boundary faces)
static int build_entry_faces(s4fgpu_ctx* c) {
    const int N = c->N, F = c->F, B = c->B;
    std::vector<int> cnt(N, 0);
    for (int f = 0; f < F; f++) { cnt[c->own[f]]++; cnt[c->nei[f]]++; }
    for (int b = 0; b < B; b++) cnt[c->faceCells[b]]++;
    std::vector<long long> rowPtr(N + 1, 0);
    for (int i = 0; i < N; i++) rowPtr[i + 1] = rowPtr[i] + cnt[i];
    std::vector<int> rFace(rowPtr[N]);
    std::vector<long long> cur(rowPtr.begin(), rowPtr.end() - 1);
    for (int f = 0; f < F; f++) rFace[cur[c->nei[f]]++] = -(f + 1);
    for (int f = 0; f < F; f++) rFace[cur[c->own[f]]++] = f + 1;
    for (int b = 0; b < B; b++) rFace[cur[c->faceCells[b]]++] = F + b + 1;
    std::vector<int> sp(c->nSlices + 1);
    S4F_CHECK_CUDA(c, cudaMemcpy(sp.data(), c->slicePtr.p, sp.size() * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> h((size_t)std::max<long long>(c->nEntries, 1), 0);
    for (int s = 0; s < c->nSlices; s++) {
        const int w = (sp[s + 1] - sp[s]) / 32;
        for (int lane = 0; lane < 32; lane++) {
            const int P = s * 32 + lane;
            if (P >= N) continue;
            for (int k = 0; k < w && k < cnt[P]; k++) h[(size_t)sp[s] + 32 * (size_t)k + lane] = rFace[rowPtr[P] + k];
        }
    }
    S4F_CHECK_CUDA(c, c->eFaceS.upload(h));
    return 0;
}

// host copies that the lazily built point stencils (pointCellsLeastSquares rows, the uns model) read
int s4f_refresh_host_geometry(s4fgpu_ctx* c) {
    const int N = c->N, F = c->F, B = c->B, nF = F + B, nP = c->nPoints;
    DevBuf<double> tmp;
    S4F_CHECK_CUDA(c, tmp.alloc(3 * (size_t)std::max(std::max(N, B), nP), false));
    k_soa3_to_aos<<<(N + 255) / 256, 256, 0, c->stream>>>(c->Cc.p, tmp.p, N, (size_t)c->ld, 0);
    c->hC.resize(3 * (size_t)N);
    S4F_CHECK_CUDA(c, cudaMemcpyAsync(c->hC.data(), tmp.p, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    if (B > 0) {
        k_soa3_to_aos<<<(B + 255) / 256, 256, 0, c->stream>>>(c->fCtr.p, tmp.p, B, (size_t)nF, F);
        c->hCfB.resize(3 * (size_t)B);
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(c->hCfB.data(), tmp.p, 3 * (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
        k_soa3_to_aos<<<(B + 255) / 256, 256, 0, c->stream>>>(c->fSf.p, tmp.p, B, (size_t)nF, F);
        c->hBSfHost.resize(3 * (size_t)B);
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(c->hBSfHost.data(), tmp.p, 3 * (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    c->hPoints.resize(3 * (size_t)nP);
    S4F_CHECK_CUDA(c, cudaMemcpy(c->hPoints.data(), c->dPoints.p, 3 * (size_t)nP * sizeof(double), cudaMemcpyDeviceToHost));
    if (c->X > 0) {      // decomposed: the moved centres of the other ranks' cells / boundary faces at shared points
        std::vector<double> x(3 * (size_t)c->X);
        DevBuf<double> tx; S4F_CHECK_CUDA(c, tx.alloc(3 * (size_t)c->X, false));
        k_soa3_to_aos<<<(c->X + 255) / 256, 256, 0, c->stream>>>(c->Cc.p, tx.p, c->X, (size_t)c->ld, c->xOff());
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(x.data(), tx.p, x.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
        for (size_t e = 0; e < c->extSlot.size(); e++) for (int q = 0; q < 3; q++) c->extCtr[3 * e + q] = x[3 * (size_t)(c->extSlot[e] - c->xOff()) + q];
    }
    c->launches += 3;
    c->hostGeomStale = false;
    return 0;
}

// newPoints = oldPoints + pointDD (symmetry-plane points keep their plane), then everything derived from the geometry.
// pointDD: host [3*nPoints], or null = the point field the last s4fgpu_interpolate_to_points left on the device.
int s4f_move_points_device(s4fgpu_ctx* c, const double* hostPointDD) {
    const int N = c->N, F = c->F, B = c->B, nF = F + B, nP = c->nPoints, ld = c->ld;
    if (nP == 0) { c->err = "move_points: call set_points first"; return 1; }
    if (c->nRanks > 1 && c->extPtr.size() != (size_t)nP + 1) { c->err = "move_points on a decomposed mesh: call set_points before set_geometry"; return 1; }
    for (int p = 0; p < c->nPatches; p++)
        if (c->pKind[p] == S4F_PATCH_EMPTY) { c->err = "move_points: meshes with empty patches (2-D cases) move on the host (set_geometry / set_points)"; return 1; }
    if (c->eFaceS.n != (size_t)std::max<long long>(c->nEntries, 1)) { int rc = build_entry_faces(c); if (rc) return rc; }
    if (c->fCtr.n != 3 * (size_t)nF) { S4F_CHECK_CUDA(c, c->fCtr.alloc(3 * (size_t)nF)); S4F_CHECK_CUDA(c, c->fSf.alloc(3 * (size_t)nF)); }
    if (c->Cc.n != 3 * (size_t)ld) S4F_CHECK_CUDA(c, c->Cc.alloc(3 * (size_t)ld));
    if (!c->nonOrth) {      // an orthogonal mesh stops being one when it moves
        S4F_CHECK_CUDA(c, c->eCorr.alloc(3 * (size_t)c->nEntries)); S4F_CHECK_CUDA(c, c->eVc.alloc(3 * (size_t)c->nEntries));
        c->nonOrth = true;
    }
    const double* dd = c->ptOut.p;
    DevBuf<double> up;
    if (hostPointDD) {
        S4F_CHECK_CUDA(c, up.alloc(3 * (size_t)nP, false));
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(up.p, hostPointDD, 3 * (size_t)nP * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        dd = up.p;
    }
    k_geo_move_points<<<(nP + 255) / 256, 256, 0, c->stream>>>(c->dPoints.p, dd, c->ptFixAxis.p, nP);
    k_geo_faces<<<(nF + 127) / 128, 128, 0, c->stream>>>(c->dFvPtr.p, c->dFv.p, c->dPoints.p, c->fCtr.p, c->fSf.p, nF);
    const int gridM = s4f_grid(c->numSMs, (long long)c->nSlices * 32, 4);
    k_geo_cells<<<gridM, S4F_BLOCK, 0, c->stream>>>(c->slicePtr.p, c->eFaceS.p, c->fCtr.p, c->fSf.p, c->Cc.p, c->V.p, c->rV.p, N, ld, nF, c->nSlices);
    c->launches += 3;
    int rc = s4f_halo_exchange(c, c->Cc.p, 3); if (rc) return rc;       // neighbour cell centres across processor patches
    if (B > 0) { k_geo_boundary_centres<<<(B + 127) / 128, 128, 0, c->stream>>>(c->fCtr.p, c->Cc.p, F, B, nF, c->bOff(), ld); c->launches++; }
    if ((rc = s4f_point_ghost_exchange(c, c->Cc.p, 3))) return rc;      // decomposed: centres of the other ranks' cells / faces at shared points
    GeoOut o{c->eW.p, c->eSf.p, c->eDn.p, c->eCorr.p, c->eLs.p, c->bN.p, c->bK.p, c->bSf.p, c->bDelta.p, c->bMagSf.p};
    k_geo_entries<<<gridM, S4F_BLOCK, 0, c->stream>>>(c->slicePtr.p, c->col.p, c->eFaceS.p, c->fCtr.p, c->fSf.p, c->Cc.p, o, N, c->bOff(), B, ld, nF,
                                                     c->nEntries, c->nSlices, c->solD[0], c->solD[1], c->solD[2]);
    k_geo_point_weights<<<(nP + 127) / 128, 128, 0, c->stream>>>(c->ptPtr.p, c->ptCol.p, c->ptW.p, c->pgPtr.p, c->pgCol.p, c->pgW.p, c->pgDelta.p,
                                                                (long long)c->pgW.n, c->dPoints.p, c->Cc.p, c->fCtr.p, nP, c->bOff(), ld, F, nF);
    c->launches += 2;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    // as a second s4fgpu_set_geometry: fields, boundary data and law history stay; what depends on the geometry is rebuilt
    c->matrixValid = false; c->histValid = false; c->mValid = false; c->gValid = false; c->unsValid = false; c->dicValid = false;
    c->amgValid = false;        // the re-assembly that follows turns this into a coefficient refresh (s4f_amg_refresh)
    c->hostGeomStale = true;
    if (c->pointCellsGrad() || c->unsModel()) { rc = s4f_refresh_host_geometry(c); if (rc) return rc; }
    return 0;
}
