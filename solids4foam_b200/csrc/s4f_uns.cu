// s4f_uns.cu -- the "uns" discretisation of unsLinGeomSolid (SM/unsLinGeomSolid/unsLinGeomSolid.C:100-175): the stress
// lives on the FACES, sigmaf = law(gradDf), and the momentum equation takes fvc::div(mesh().Sf() & sigmaf) (:129).
//   pointD  = volToPoint().interpolate(D)                           (:146, k_vol_to_point)
//   gradD   = fvc::grad(D, pointD)      NUM/fvc/fvcGradf.C:442-800   Gauss sum over the faces' triangle fans of the vertex values
//   gradDf  = fvc::fGrad(D, pointD)     fvcGradf.C:44-112: fsGrad (:123-232, in-plane gradient from the edge-centre values)
//                                       + n*fvc::snGrad(D) (corrected snGrad; patch faces: the boundary condition's snGrad())
//   sigmaf  = 2 mu symm(gradDf) + lambda tr I + sigma0f             linearElastic.C:342-370
// unsNonLinGeomTotalLagSolid (SM/unsNonLinGeomTotalLagSolid/unsNonLinGeomTotalLagSolid.C:218-405) uses the same gradients at
// finite strain: Ff = I + gradDf.T() (:312), sigmaf = neoHookeanElastic::correct(surfaceSymmTensorField&)
// (neoHookeanElastic.C:306-352), and the divergence of (Jf Finvf.T() & Sf) & sigmaf (:273) -- formed per face as one
// traction vector in k_uns_face_stress, summed by k_source_uns.
// One thread per face for the face quantities (vertex gathers through the face->vertex CSR), the usual atomic-free row
// gathers for the cell gradient and the divergence.  CPU restatement: oracle/s4f_oracle.cpp (unsUpdateGradients, unsLawFaces).
#include <algorithm>
#include <cmath>

#include "s4f_ctx.h"
#include "s4f_dev.cuh"

struct S4fUns {
    int nF = 0, ldF = 0;                 // faces (internal + boundary), leading dimension of the face fields
    DevBuf<int> fPtr, fVerts;            // face -> vertices
    DevBuf<int> fOwn, fNei;              // [F]
    DevBuf<int> eFace;                   // [nEntries] +-(face+1): face of a row entry, sign + when the row cell owns it; 0 = padding
    DevBuf<double> pts;                  // [3*nPoints] AoS
    DevBuf<double> rV3;                  // [ld] 3 / sum St & Ct
    DevBuf<double> fT, fG;               // [9*ldF] in-plane gradient, Gauss sum of a face
    DevBuf<double> gradDf, sigmaf;       // [9*ldF], [6*ldF]
    DevBuf<double> gLS;                  // [9*ld] fvc::grad(D) of the gradScheme (non-orthogonal part of snGrad(D))
    DevBuf<int> procFace;                // [G] boundary face of each processor-patch ghost (decomposed meshes)
    DevBuf<double> faceT;                // [3*ldF] finite-strain models: (Jf Finvf.T() & Sf) & sigmaf, owner orientation
    DevBuf<double> Ff, FfOld;            // [9*ldF] updated-Lagrangian model: total deformation gradient of the faces, its oldTime()
};

namespace {

__global__ void k_uns_fill(double* p, double v, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_uns_face_pre(const int* __restrict__ fPtr, const int* __restrict__ fVerts, const double* __restrict__ pts,
                               const double* __restrict__ pD, const int* __restrict__ faceEntry, const double* __restrict__ eSf,
                               const double* __restrict__ bSf, double* __restrict__ fT, double* __restrict__ fG, int F, int B, int ldF,
                               long long nE) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F + B) return;
    double S[3];
    if (f < F) { const long long e = faceEntry[f]; S[0] = eSf[e]; S[1] = eSf[nE + e]; S[2] = eSf[2 * nE + e]; }
    else { const int b = f - F; S[0] = bSf[b]; S[1] = bSf[(size_t)B + b]; S[2] = bSf[2 * (size_t)B + b]; }
    const double mag = sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2]);
    const double n[3] = {S[0] / mag, S[1] / mag, S[2] / mag};
    const int a = fPtr[f], m = fPtr[f + 1] - a;
    double T[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double cp[3] = {0, 0, 0}, cf[3] = {0, 0, 0};
    for (int i = 0; i < m; i++) {
        const int v = fVerts[a + i];
#pragma unroll
        for (int q = 0; q < 3; q++) { cp[q] += pts[3 * (size_t)v + q]; cf[q] += pD[3 * (size_t)v + q]; }
    }
#pragma unroll
    for (int q = 0; q < 3; q++) { cp[q] /= m; cf[q] /= m; }
    for (int i = 0; i < m; i++) {
        const int v0 = fVerts[a + i], v1 = fVerts[a + (i + 1 == m ? 0 : i + 1)];
        double p0[3], p1[3], u0[3], u1[3];
#pragma unroll
        for (int q = 0; q < 3; q++) { p0[q] = pts[3 * (size_t)v0 + q]; p1[q] = pts[3 * (size_t)v1 + q]; u0[q] = pD[3 * (size_t)v0 + q]; u1[q] = pD[3 * (size_t)v1 + q]; }
        // fsGrad: Le = (e - n (n & e)) ^ n, fe = (u0 + u1)/2
        double e[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
        const double ne = n[0] * e[0] + n[1] * e[1] + n[2] * e[2];
#pragma unroll
        for (int q = 0; q < 3; q++) e[q] -= n[q] * ne;
        const double Le[3] = {e[1] * n[2] - e[2] * n[1], e[2] * n[0] - e[0] * n[2], e[0] * n[1] - e[1] * n[0]};
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double fe = 0.5 * (u0[j] + u1[j]);
#pragma unroll
            for (int i2 = 0; i2 < 3; i2++) T[3 * i2 + j] += Le[i2] * fe;
        }
        // Gauss sum over the triangle fan (a triangular face is taken whole below)
        if (m != 3) {
            const double ra[3] = {p0[0] - cp[0], p0[1] - cp[1], p0[2] - cp[2]}, rb[3] = {p1[0] - cp[0], p1[1] - cp[1], p1[2] - cp[2]};
            const double St[3] = {0.5 * (ra[1] * rb[2] - ra[2] * rb[1]), 0.5 * (ra[2] * rb[0] - ra[0] * rb[2]), 0.5 * (ra[0] * rb[1] - ra[1] * rb[0])};
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double tt = (u0[j] + u1[j] + cf[j]) / 3.0;
#pragma unroll
                for (int i2 = 0; i2 < 3; i2++) G[3 * i2 + j] += St[i2] * tt;
            }
        }
    }
    if (m == 3) {          // fvcGradf.C:501-513: SF = face normal * average of the vertex values
        const int v0 = fVerts[a], v1 = fVerts[a + 1], v2 = fVerts[a + 2];
        double e1[3], e2[3];
#pragma unroll
        for (int q = 0; q < 3; q++) { e1[q] = pts[3 * (size_t)v1 + q] - pts[3 * (size_t)v0 + q]; e2[q] = pts[3 * (size_t)v2 + q] - pts[3 * (size_t)v0 + q]; }
        const double Sn[3] = {0.5 * (e1[1] * e2[2] - e1[2] * e2[1]), 0.5 * (e1[2] * e2[0] - e1[0] * e2[2]), 0.5 * (e1[0] * e2[1] - e1[1] * e2[0])};
#pragma unroll
        for (int i2 = 0; i2 < 3; i2++)
#pragma unroll
            for (int j = 0; j < 3; j++) G[3 * i2 + j] = Sn[i2] * cf[j];
    }
#pragma unroll
    for (int q = 0; q < 9; q++) { fT[(size_t)q * ldF + f] = T[q] / mag; fG[(size_t)q * ldF + f] = G[q]; }
}

// gradD_P = (3 / sum St & Ct) sum_faces +-G_f : row gather over the faces of the cell
__global__ void __launch_bounds__(S4F_BLOCK) k_uns_cell_grad(const int* __restrict__ slicePtr, const int* __restrict__ eFace,
                                                            const double* __restrict__ fG, const double* __restrict__ rV3,
                                                            double* __restrict__ gradD, int N, int ld, int ldF, int nSlices) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nSlices; s += nWarps) {
        const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5;
        const int row = s * 32 + lane;
        double g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < width; k++) {
            const int ef = eFace[(long long)base + 32 * k + lane];
            if (ef == 0) continue;
            const int f = (ef > 0 ? ef : -ef) - 1;
            const double sg = ef > 0 ? 1.0 : -1.0;
#pragma unroll
            for (int q = 0; q < 9; q++) g[q] += sg * fG[(size_t)q * ldF + f];
        }
        if (row < N) {
            const double sc = rV3[row];
#pragma unroll
            for (int q = 0; q < 9; q++) gradD[(size_t)q * ld + row] = g[q] * sc;
        }
    }
}

// patch values of gradD: the in-plane gradient of the patch face, then the boundary condition's normal gradient (fvcGradf.C:693-781)
__global__ void k_uns_grad_boundary(const int* __restrict__ bKind, const double* __restrict__ bN, const double* __restrict__ bSn,
                                    const double* __restrict__ fT, double* __restrict__ gradD, int F, int B, int bOff, int ld, int ldF) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || bKind[b] == S4F_BC_PROCESSOR) return;
    const double n[3] = {bN[b], bN[(size_t)B + b], bN[2 * (size_t)B + b]};
    double g[9];
#pragma unroll
    for (int q = 0; q < 9; q++) g[q] = fT[(size_t)q * ldF + F + b];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const double ng = n[0] * g[j] + n[1] * g[3 + j] + n[2] * g[6 + j];
        const double corr = bSn[(size_t)j * B + b] - ng;
#pragma unroll
        for (int i = 0; i < 3; i++) g[3 * i + j] += n[i] * corr;
    }
#pragma unroll
    for (int q = 0; q < 9; q++) gradD[(size_t)q * ld + bOff + b] = g[q];
}

// gradDf = fsGrad + n*snGrad(D); sigmaf = 2 mu symm(gradDf) + lambda tr I + sigma0f
struct S6u { double v[6]; };
__global__ void k_uns_face_stress(const int* __restrict__ fOwn, const int* __restrict__ fNei, const int* __restrict__ faceEntry,
                                  const double* __restrict__ eSf, const double* __restrict__ eDn, const double* __restrict__ eW,
                                  const double* __restrict__ eCorr /* null when orthogonal */, const double* __restrict__ bN,
                                  const double* __restrict__ bSn, const double* __restrict__ D, const double* __restrict__ gLS,
                                  const double* __restrict__ fT, double* __restrict__ gradDf, double* __restrict__ sigmaf, int F, int B,
                                  int ld, int ldF, long long nE, double mu, double lambda, S6u s0,
                                  const double* __restrict__ bSf, double* __restrict__ faceT /* null: linear geometry */, double K,
                                  const double* __restrict__ FfOld /* updated Lagrangian only */, double* __restrict__ FfOut) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F + B) return;
    double n[3], sn[3], Sv[3];
    if (f < F) {
        const long long e = faceEntry[f];
        const double S[3] = {eSf[e], eSf[nE + e], eSf[2 * nE + e]};
        Sv[0] = S[0]; Sv[1] = S[1]; Sv[2] = S[2];
        const double mag = sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2]);
#pragma unroll
        for (int q = 0; q < 3; q++) n[q] = S[q] / mag;
        const int P = fOwn[f], Nn = fNei[f];
        const double nod = eDn[e] / mag;
#pragma unroll
        for (int j = 0; j < 3; j++) sn[j] = nod * (D[(size_t)j * ld + Nn] - D[(size_t)j * ld + P]);
        if (eCorr) {       // corrected snGrad: nonOrthCorrectionVectors & linear interpolate(fvc::grad(D))
            const double w = eW[e], w1 = 1.0 - w;
            const double c[3] = {eCorr[e] / mag, eCorr[nE + e] / mag, eCorr[2 * nE + e] / mag};
#pragma unroll
            for (int j = 0; j < 3; j++) {
                double t = 0;
#pragma unroll
                for (int i = 0; i < 3; i++) t += c[i] * (w * gLS[(size_t)(3 * i + j) * ld + P] + w1 * gLS[(size_t)(3 * i + j) * ld + Nn]);
                sn[j] += t;
            }
        }
    } else {
        const int b = f - F;
#pragma unroll
        for (int q = 0; q < 3; q++) { n[q] = bN[(size_t)q * B + b]; sn[q] = bSn[(size_t)q * B + b]; Sv[q] = bSf[(size_t)q * B + b]; }
    }
    double g[9];
#pragma unroll
    for (int q = 0; q < 9; q++) g[q] = fT[(size_t)q * ldF + f];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) g[3 * i + j] += n[i] * sn[j];
#pragma unroll
    for (int q = 0; q < 9; q++) gradDf[(size_t)q * ldF + f] = g[q];
    if (!sigmaf) return;            // constructor: the gradients only
    double s[6];
    if (faceT) {
        // Ff = I + gradDf.T(); neoHookeanElastic::correctF: bEbar = J^(-2/3) symm(F & F.T()), sigma = (0.5 K (J^2 - 1) I + mu dev(bEbar)) / J
        // updated Lagrangian (unsNonLinGeomUpdatedLagSolid.C:283-301): g is grad(DD)f, rF = relFf = I + g.T(), Ff = relFf & Ff.oldTime();
        // the stress takes Ff, the flux (relJf relFinvf.T() & Sf) & sigmaf the relative tensors
        double rF[9], Fm[9], FT[9], FFT[9], b6[6], Fi[9];
        t_transpose(g, rF); rF[0] += 1; rF[4] += 1; rF[8] += 1;
        if (FfOld) {
            double Fo[9];
#pragma unroll
            for (int q = 0; q < 9; q++) Fo[q] = FfOld[(size_t)q * ldF + f];
            t_mul(rF, Fo, Fm);
#pragma unroll
            for (int q = 0; q < 9; q++) FfOut[(size_t)q * ldF + f] = Fm[q];
        } else {
#pragma unroll
            for (int q = 0; q < 9; q++) Fm[q] = rF[q];
        }
        const double J = t_det(Fm);
        t_transpose(Fm, FT); t_mul(Fm, FT, FFT); t_symm(FFT, b6);
        const double sc = pow(J, -2.0 / 3.0);
#pragma unroll
        for (int q = 0; q < 6; q++) b6[q] *= sc;
        s_dev(b6, s);
        const double sh = 0.5 * K * (pow(J, 2.0) - 1.0), rJ = 1.0 / J;
#pragma unroll
        for (int q = 0; q < 6; q++) s[q] *= mu;
        s[0] += sh; s[3] += sh; s[5] += sh;
#pragma unroll
        for (int q = 0; q < 6; q++) s[q] *= rJ;
        // (Jf Finvf.T() & Sf) & sigmaf, relative tensors for the updated-Lagrangian model
        t_inv(rF, Fi);
        const double Jr = FfOld ? t_det(rF) : J;
        double a[3];
#pragma unroll
        for (int i = 0; i < 3; i++) a[i] = Jr * (Fi[i] * Sv[0] + Fi[3 + i] * Sv[1] + Fi[6 + i] * Sv[2]);      // Finv.T()_ij = Finv_ji
        faceT[f] = a[0] * s[0] + a[1] * s[1] + a[2] * s[2];
        faceT[(size_t)ldF + f] = a[0] * s[1] + a[1] * s[3] + a[2] * s[4];
        faceT[2 * (size_t)ldF + f] = a[0] * s[2] + a[1] * s[4] + a[2] * s[5];
    } else {
        double e6[6];
        t_symm(g, e6);
        const double tr = s_tr(e6);
#pragma unroll
        for (int q = 0; q < 6; q++) s[q] = 2.0 * mu * e6[q] + s0.v[q];
        s[0] += lambda * tr; s[3] += lambda * tr; s[5] += lambda * tr;
    }
#pragma unroll
    for (int q = 0; q < 6; q++) sigmaf[(size_t)q * ldF + f] = s[q];
}

// processor-patch faces: the corrected snGrad(D) of an internal face, with the cell across the cut in its ghost slot
// (what the coupled patch field's snGrad() + the processor patch's interpolate(fvc::grad(D)) give in OpenFOAM); both ranks
// evaluate the same expression with n and (D_N - D_P) reversed, so the face stress is the same on either side
__global__ void k_uns_proc_sngrad(const int* __restrict__ procFace, const int* __restrict__ procEntry, const int* __restrict__ bFaceCell,
                                  const double* __restrict__ eSf, const double* __restrict__ eDn, const double* __restrict__ eW,
                                  const double* __restrict__ eCorr /* null when orthogonal */, const double* __restrict__ D,
                                  const double* __restrict__ gLS, double* __restrict__ bSn, int G, int N, int B, int ld, long long nE) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int b = procFace[g], P = bFaceCell[b], Nn = N + g;
    const long long e = procEntry[g];
    const double S[3] = {eSf[e], eSf[nE + e], eSf[2 * nE + e]};
    const double mag = sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2]);
    const double nod = eDn[e] / mag;
    double sn[3];
#pragma unroll
    for (int j = 0; j < 3; j++) sn[j] = nod * (D[(size_t)j * ld + Nn] - D[(size_t)j * ld + P]);
    if (eCorr) {
        const double w = eW[e], w1 = 1.0 - w;
        const double c[3] = {eCorr[e] / mag, eCorr[nE + e] / mag, eCorr[2 * nE + e] / mag};
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double t = 0;
#pragma unroll
            for (int i = 0; i < 3; i++) t += c[i] * (w * gLS[(size_t)(3 * i + j) * ld + P] + w1 * gLS[(size_t)(3 * i + j) * ld + Nn]);
            sn[j] += t;
        }
    }
#pragma unroll
    for (int j = 0; j < 3; j++) bSn[(size_t)j * B + b] = sn[j];
}

// The face traction (relJf relFinvf.T() & Sf) & sigmaf again, from the stored grad(DD)f and sigmaf and the CURRENT face area
// vectors: the reference evaluates the product when it assembles the equation, i.e. after updateTotalFields moved the mesh,
// with the relative tensors still those of the last iteration of the previous step.
__global__ void k_uns_flux_refresh(const int* __restrict__ faceEntry, const double* __restrict__ eSf, const double* __restrict__ bSf,
                                   const double* __restrict__ gradDf, const double* __restrict__ sigmaf, double* __restrict__ faceT, int F,
                                   int B, int ldF, long long nE) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F + B) return;
    double S[3], g[9], s[6], rF[9], Fi[9];
    if (f < F) { const long long e = faceEntry[f]; S[0] = eSf[e]; S[1] = eSf[nE + e]; S[2] = eSf[2 * nE + e]; }
    else { const int b = f - F; S[0] = bSf[b]; S[1] = bSf[(size_t)B + b]; S[2] = bSf[2 * (size_t)B + b]; }
#pragma unroll
    for (int q = 0; q < 9; q++) g[q] = gradDf[(size_t)q * ldF + f];
#pragma unroll
    for (int q = 0; q < 6; q++) s[q] = sigmaf[(size_t)q * ldF + f];
    t_transpose(g, rF); rF[0] += 1; rF[4] += 1; rF[8] += 1;
    const double Jr = t_det(rF);
    t_inv(rF, Fi);
    double a[3];
#pragma unroll
    for (int i = 0; i < 3; i++) a[i] = Jr * (Fi[i] * S[0] + Fi[3 + i] * S[1] + Fi[6 + i] * S[2]);
    faceT[f] = a[0] * s[0] + a[1] * s[1] + a[2] * s[2];
    faceT[(size_t)ldF + f] = a[0] * s[1] + a[1] * s[3] + a[2] * s[4];
    faceT[2 * (size_t)ldF + f] = a[0] * s[2] + a[1] * s[4] + a[2] * s[5];
}

// traction patches: unsLinGeomSolid::tractionBoundarySnGrad (unsLinGeomSolid.C:193-230) on the face fields
__global__ void k_bc_update_uns(const int* __restrict__ bKind, const double* __restrict__ bN, const double* __restrict__ bcValue,
                                const double* __restrict__ bcPressure, const double* __restrict__ impK, const double* __restrict__ sigmaf,
                                const double* __restrict__ gradDf, double* __restrict__ tracGrad, double* __restrict__ D, int F, int B,
                                int bOff, int ld, int ldF, int TL /* 0 linear, 1 total, 2 updated Lagrangian */,
                                const double* __restrict__ Dold /* incremental model: the field solved for is DD */) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int kind = bKind[b];
    if (kind == S4F_BC_FIXED_DISPLACEMENT) {
#pragma unroll
        for (int c = 0; c < 3; c++)
            D[(size_t)c * ld + bOff + b] = bcValue[(size_t)c * B + b] - (Dold ? Dold[(size_t)c * ld + bOff + b] : 0.0);   // fixedDisplacement...C:279-287
    } else if (kind == S4F_BC_SOLID_TRACTION) {
        double n[3], t[3], g[9], s[6], M[9];
#pragma unroll
        for (int c = 0; c < 3; c++) { n[c] = bN[(size_t)c * B + b]; t[c] = bcValue[(size_t)c * B + b]; }
#pragma unroll
        for (int q = 0; q < 9; q++) g[q] = gradDf[(size_t)q * ldF + F + b];
#pragma unroll
        for (int q = 0; q < 6; q++) s[q] = sigmaf[(size_t)q * ldF + F + b];
        const double p = bcPressure[b], k = impK[bOff + b];
        s_to_t(s, M);
        if (TL) {
            // unsNonLinGeomTotalLagSolid::tractionBoundarySnGrad (:420-488): nCurrent = Jf Finvf.T() & n (not normalised);
            // ((t - nCurrent p) - (nCurrent & sigmaf) + (n & (impK gradDf))) / impK.  The updated-Lagrangian model
            // (unsNonLinGeomUpdatedLagSolid.C:373-419) takes the relative tensors (g is grad(DD)f) and lets the pressure act along n
            double Fm[9], Fi[9], nc[3];
            t_transpose(g, Fm); Fm[0] += 1; Fm[4] += 1; Fm[8] += 1;
            const double J = t_det(Fm);
            t_inv(Fm, Fi);
#pragma unroll
            for (int i = 0; i < 3; i++) nc[i] = J * (Fi[i] * n[0] + Fi[3 + i] * n[1] + Fi[6 + i] * n[2]);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double ns = nc[0] * M[c] + nc[1] * M[3 + c] + nc[2] * M[6 + c];
                const double ng = n[0] * g[c] + n[1] * g[3 + c] + n[2] * g[6 + c];
                tracGrad[(size_t)c * B + b] = ((t[c] - (TL == 2 ? n[c] : nc[c]) * p) - ns + k * ng) / k;
            }
            return;
        }
#pragma unroll
        for (int q = 0; q < 9; q++) M[q] -= k * g[q];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double nM = n[0] * M[c] + n[1] * M[3 + c] + n[2] * M[6 + c];
            tracGrad[(size_t)c * B + b] = ((t[c] - n[c] * p) - nM) / k;
        }
    }
}

// right-hand side: - V fvc::laplacian(impKf, D) [compact] + sum_faces Sf & sigmaf + V (rho g + d2dt2 history)
__global__ void __launch_bounds__(S4F_BLOCK) k_source_uns(const int* __restrict__ slicePtr, const int* __restrict__ col,
                                                         const int* __restrict__ eFace, const double* __restrict__ eSf,
                                                         const double* __restrict__ eA, const double* __restrict__ D,
                                                         const double* __restrict__ sigmaf, const double* __restrict__ V,
                                                         const double* __restrict__ hist, double* __restrict__ source, int N, int ld,
                                                         int ldF, long long nE, int nSlices, double rgx, double rgy, double rgz,
                                                         const double* __restrict__ faceT /* null: Sf & sigmaf */) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nSlices; s += nWarps) {
        const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5;
        const int row = s * 32 + lane;
        const int r = row < N ? row : 0;
        const double DP[3] = {D[r], D[(size_t)ld + r], D[2 * (size_t)ld + r]};
        double acc[3] = {0, 0, 0};
        for (int k = 0; k < width; k++) {
            const long long e = (long long)base + 32 * k + lane;
            const int ef = eFace[e];
            if (ef == 0) continue;
            const int f = (ef > 0 ? ef : -ef) - 1, cc = col[e];
            const double a = eA[e];
            if (faceT) {        // total-Lagrangian: the face traction vector, owner orientation
                const double sgn = ef > 0 ? 1.0 : -1.0;
                acc[0] += sgn * faceT[f] - a * (D[cc] - DP[0]);
                acc[1] += sgn * faceT[(size_t)ldF + f] - a * (D[(size_t)ld + cc] - DP[1]);
                acc[2] += sgn * faceT[2 * (size_t)ldF + f] - a * (D[2 * (size_t)ld + cc] - DP[2]);
                continue;
            }
            const double S[3] = {eSf[e], eSf[nE + e], eSf[2 * nE + e]};
            double sg[6];
#pragma unroll
            for (int q = 0; q < 6; q++) sg[q] = sigmaf[(size_t)q * ldF + f];
            acc[0] += S[0] * sg[0] + S[1] * sg[1] + S[2] * sg[2] - a * (D[cc] - DP[0]);
            acc[1] += S[0] * sg[1] + S[1] * sg[3] + S[2] * sg[4] - a * (D[(size_t)ld + cc] - DP[1]);
            acc[2] += S[0] * sg[2] + S[1] * sg[4] + S[2] * sg[5] - a * (D[2 * (size_t)ld + cc] - DP[2]);
        }
        if (row < N) {
            const double v = V[row], rg[3] = {rgx, rgy, rgz};
#pragma unroll
            for (int q = 0; q < 3; q++) source[(size_t)q * ld + row] = acc[q] + v * rg[q] + (hist ? v * hist[(size_t)q * ld + row] : 0.0);
        }
    }
}

__global__ void k_vol_to_point_dev(const int* __restrict__ ptPtr, const int* __restrict__ ptCol, const double* __restrict__ ptW,
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

}  // namespace

void s4f_uns_destroy(s4fgpu_ctx* c) { delete c->uns; c->uns = nullptr; }

// face->vertex CSR, owner/neighbour, the face of every row entry, and 3/sum(St & Ct) per cell (geometry only)
int s4f_uns_setup(s4fgpu_ctx* c) {
    if (c->nPoints == 0) { c->err = "unsLinearGeometry needs the mesh points (s4fgpu_set_points)"; return 1; }
    if (c->nRanks > 1 && c->extPtr.size() != (size_t)c->nPoints + 1) {
        c->err = "unsLinearGeometry on a decomposed mesh: call s4fgpu_set_points before s4fgpu_set_geometry";
        return 1;
    }
    if (!c->unsFinite() && c->law.kind != S4F_LAW_LINEAR_ELASTIC) { c->err = "unsLinearGeometry: linearElastic is the law available on the faces"; return 1; }
    if (c->unsFinite() && c->law.kind != S4F_LAW_NEO_HOOKEAN_ELASTIC) {
        c->err = "unsNonLinearGeometryTotalLagrangian / UpdatedLagrangian: neoHookeanElastic is the law available on the faces"; return 1;
    }
    const int N = c->N, F = c->F, B = c->B;
    if (!c->uns) c->uns = new S4fUns();
    S4fUns& u = *c->uns;
    u.nF = F + B; u.ldF = ((F + B + 31) / 32) * 32;
    S4F_CHECK_CUDA(c, u.fPtr.upload(c->hFvPtr)); S4F_CHECK_CUDA(c, u.fVerts.upload(c->hFv));
    std::vector<int> own(std::max(F, 1), 0), nei(std::max(F, 1), 0);
    for (int f = 0; f < F; f++) { own[f] = c->own[f]; nei[f] = c->nei[f]; }
    S4F_CHECK_CUDA(c, u.fOwn.upload(own)); S4F_CHECK_CUDA(c, u.fNei.upload(nei));
    S4F_CHECK_CUDA(c, u.pts.upload(c->hPoints));
    {
        std::vector<int> pf(std::max(c->G, 1), 0);
        for (int b = 0; b < B; b++) if (c->ghostOfFace[b] >= 0) pf[c->ghostOfFace[b] - N] = b;
        S4F_CHECK_CUDA(c, u.procFace.upload(pf));
    }
    // rows in the order of s4f_build_rows: lower neighbours, upper neighbours, boundary faces
    std::vector<int> cnt(N, 0);
    for (int f = 0; f < F; f++) { cnt[c->own[f]]++; cnt[c->nei[f]]++; }
    for (int b = 0; b < B; b++) cnt[c->faceCells[b]]++;
    std::vector<long long> rowPtr(N + 1, 0);
    for (int i = 0; i < N; i++) rowPtr[i + 1] = rowPtr[i] + cnt[i];
    std::vector<int> rFace(rowPtr[N]);
    {
        std::vector<long long> cur(rowPtr.begin(), rowPtr.end() - 1);
        for (int f = 0; f < F; f++) rFace[cur[c->nei[f]]++] = -(f + 1);
        for (int f = 0; f < F; f++) rFace[cur[c->own[f]]++] = f + 1;
        for (int b = 0; b < B; b++) rFace[cur[c->faceCells[b]]++] = F + b + 1;
    }
    std::vector<int> sp(c->nSlices + 1, 0);
    for (int s = 0; s < c->nSlices; s++) {
        int w = 0;
        for (int r = s * 32; r < std::min(N, s * 32 + 32); r++) w = std::max(w, cnt[r]);
        sp[s + 1] = sp[s] + 32 * w;
    }
    std::vector<int> ef((size_t)std::max<long long>(c->nEntries, 1), 0);
    for (int s = 0; s < c->nSlices; s++) {
        const int w = (sp[s + 1] - sp[s]) / 32;
        for (int lane = 0; lane < 32; lane++) {
            const int P = s * 32 + lane;
            if (P >= N) continue;
            for (int k = 0; k < cnt[P] && k < w; k++) ef[(size_t)sp[s] + 32 * (size_t)k + lane] = rFace[rowPtr[P] + k];
        }
    }
    S4F_CHECK_CUDA(c, u.eFace.upload(ef));
    // 3 / sum St & Ct (fvcGradf.C:508-566, :670); the unmirrored faces of empty patches add the cell volume per empty direction
    std::vector<double> V3(N, 0.0);
    const double* X = c->hPoints.data();
    for (int f = 0; f < F + B; f++) {
        const int a = c->hFvPtr[f], m = c->hFvPtr[f + 1] - a;
        double Vf = 0;
        auto P = [&](int i) { return &X[3 * (size_t)c->hFv[a + i]]; };
        if (m == 3) {
            const double e1[3] = {P(1)[0] - P(0)[0], P(1)[1] - P(0)[1], P(1)[2] - P(0)[2]}, e2[3] = {P(2)[0] - P(0)[0], P(2)[1] - P(0)[1], P(2)[2] - P(0)[2]};
            const double S[3] = {0.5 * (e1[1] * e2[2] - e1[2] * e2[1]), 0.5 * (e1[2] * e2[0] - e1[0] * e2[2]), 0.5 * (e1[0] * e2[1] - e1[1] * e2[0])};
            for (int q = 0; q < 3; q++) Vf += S[q] * (P(0)[q] + P(1)[q] + P(2)[q]) / 3.0;
        } else {
            double cp[3] = {0, 0, 0};
            for (int i = 0; i < m; i++) for (int q = 0; q < 3; q++) cp[q] += P(i)[q];
            for (int q = 0; q < 3; q++) cp[q] /= m;
            for (int i = 0; i < m; i++) {
                const double* pa = P(i); const double* pb = P((i + 1) % m);
                const double ra[3] = {pa[0] - cp[0], pa[1] - cp[1], pa[2] - cp[2]}, rb[3] = {pb[0] - cp[0], pb[1] - cp[1], pb[2] - cp[2]};
                const double St[3] = {0.5 * (ra[1] * rb[2] - ra[2] * rb[1]), 0.5 * (ra[2] * rb[0] - ra[0] * rb[2]), 0.5 * (ra[0] * rb[1] - ra[1] * rb[0])};
                for (int q = 0; q < 3; q++) Vf += St[q] * (cp[q] + pa[q] + pb[q]) / 3.0;
            }
        }
        V3[f < F ? c->own[f] : c->faceCells[f - F]] += Vf;
        if (f < F) V3[c->nei[f]] -= Vf;
    }
    const int nEmpty = (c->solD[0] ? 0 : 1) + (c->solD[1] ? 0 : 1) + (c->solD[2] ? 0 : 1);
    std::vector<double> r(c->ld, 1.0);
    for (int i = 0; i < N; i++) r[i] = 3.0 / (V3[i] + nEmpty * c->hV[i]);
    S4F_CHECK_CUDA(c, u.rV3.upload(r));
    const size_t lf = u.ldF;
    if (u.fT.n != 9 * lf) {
        S4F_CHECK_CUDA(c, u.fT.alloc(9 * lf)); S4F_CHECK_CUDA(c, u.fG.alloc(9 * lf));
        S4F_CHECK_CUDA(c, u.gradDf.alloc(9 * lf)); S4F_CHECK_CUDA(c, u.sigmaf.alloc(6 * lf));
    }
    if (c->nonOrth && u.gLS.n != 9 * (size_t)c->ld) S4F_CHECK_CUDA(c, u.gLS.alloc(9 * (size_t)c->ld));
    bool keptState = false;
    if (c->unsFinite()) { if (u.faceT.n != 3 * lf) S4F_CHECK_CUDA(c, u.faceT.alloc(3 * lf)); else keptState = true; } else u.faceT.release();
    if (c->unsUL() && u.Ff.n != 9 * lf) {          // identity at the start; kept across mesh motion (face fields follow the topology)
        S4F_CHECK_CUDA(c, u.Ff.alloc(9 * lf)); S4F_CHECK_CUDA(c, u.FfOld.alloc(9 * lf));
        for (DevBuf<double>* b : {&u.Ff, &u.FfOld}) for (int d : {0, 4, 8}) {
            k_uns_fill<<<(unsigned)((lf + 255) / 256), 256, 0, c->stream>>>(b->p + (size_t)d * lf, 1.0, (long long)lf);
            c->launches++;
        }
    }
    if (keptState) {      // a geometry refresh (mesh motion): the face tractions with the new area vectors
        k_uns_flux_refresh<<<(u.nF + 127) / 128, 128, 0, c->stream>>>(c->faceEntry.p, c->eSf.p, c->bSf.p, u.gradDf.p, u.sigmaf.p, u.faceT.p, F, B, u.ldF,
                                                                    c->nEntries);
        c->launches++;
        S4F_CHECK_CUDA(c, cudaGetLastError());
    }
    c->unsValid = true;
    return 0;
}

// mechanical().interpolate(D, pointD, false); mechanical().grad(D, pointD, gradD, gradDf); mechanical().correct(sigmaf)
int s4f_uns_gradients(s4fgpu_ctx* c) {
    if (!c->unsValid) { int rc = s4f_uns_setup(c); if (rc) return rc; }
    S4fUns& u = *c->uns;
    const int N = c->N, F = c->F, B = c->B, bOff = c->bOff(), ld = c->ld, nF = F + B;
    const int gridR = s4f_grid(c->numSMs, (long long)c->nSlices * 32, 4);
    { int rx = s4f_point_ghost_exchange(c, c->D.p, 3); if (rx) return rx; }       // decomposed: the other ranks' values at shared points
    k_vol_to_point_dev<<<(c->nPoints + 127) / 128, 128, 0, c->stream>>>(c->ptPtr.p, c->ptCol.p, c->ptW.p, c->ptN.p, c->D.p, c->ptOut.p, c->nPoints, ld);
    k_uns_face_pre<<<(nF + 127) / 128, 128, 0, c->stream>>>(u.fPtr.p, u.fVerts.p, u.pts.p, c->ptOut.p, c->faceEntry.p, c->eSf.p, c->bSf.p, u.fT.p,
                                                           u.fG.p, F, B, u.ldF, c->nEntries);
    c->launches += 2;
    if (B > 0) {       // the boundary conditions' snGrad() with the registered (previous) grad(D)
        int rc = s4f_bc_sngrad_store(c); if (rc) return rc;
    }
    k_uns_cell_grad<<<gridR, S4F_BLOCK, 0, c->stream>>>(c->slicePtr.p, u.eFace.p, u.fG.p, u.rV3.p, c->gradD.p, N, ld, u.ldF, c->nSlices);
    c->launches++;
    if (B > 0) {
        k_uns_grad_boundary<<<(B + 127) / 128, 128, 0, c->stream>>>(c->bKind.p, c->bN.p, c->bSn.p, u.fT.p, c->gradD.p, F, B, bOff, ld, u.ldF);
        c->launches++;
        int rc = s4f_bc_sngrad_store(c); if (rc) return rc;            // again, with the gradD just assigned: fvc::snGrad(D) on the patches
    }
    if (c->nonOrth) {
        int rc = s4f_grad_calculated_interior(c, c->D.p, u.gLS.p); if (rc) return rc;
        if ((rc = s4f_halo_exchange(c, u.gLS.p, 9))) return rc;
    }
    if (c->G > 0) {
        k_uns_proc_sngrad<<<(c->G + 127) / 128, 128, 0, c->stream>>>(u.procFace.p, c->procEntry.p, c->bFaceCell.p, c->eSf.p, c->eDn.p, c->eW.p,
                                                                    c->nonOrth ? c->eCorr.p : nullptr, c->D.p, u.gLS.p, c->bSn.p, c->G, N, B, ld,
                                                                    c->nEntries);
        c->launches++;
    }
    S6u s0; for (int q = 0; q < 6; q++) s0.v[q] = c->law.sigma0[q];
    k_uns_face_stress<<<(nF + 127) / 128, 128, 0, c->stream>>>(u.fOwn.p, u.fNei.p, c->faceEntry.p, c->eSf.p, c->eDn.p, c->eW.p,
                                                              c->nonOrth ? c->eCorr.p : nullptr, c->bN.p, c->bSn.p, c->D.p, u.gLS.p, u.fT.p,
                                                              u.gradDf.p, c->unsGradientsOnly ? nullptr : u.sigmaf.p /* constructor
                                                              (unsLinGeomSolid.C:86-89): the gradients, sigmaf keeps its initial value */,
                                                              F, B, ld, u.ldF, c->nEntries, c->law.mu, c->law.lambda, s0,
                                                              c->bSf.p, c->unsFinite() ? u.faceT.p : nullptr, c->law.K,
                                                              c->unsUL() ? u.FfOld.p : nullptr, c->unsUL() ? u.Ff.p : nullptr);
    c->launches++;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return s4f_halo_exchange(c, c->gradD.p, 9);
}

int s4f_uns_bc_update(s4fgpu_ctx* c) {
    if (!c->unsValid) { int rc = s4f_uns_setup(c); if (rc) return rc; }
    S4fUns& u = *c->uns;
    if (c->B == 0) return 0;
    k_bc_update_uns<<<(c->B + 127) / 128, 128, 0, c->stream>>>(c->bKind.p, c->bN.p, c->bcValue.p, c->bcPressure.p, c->impK.p, u.sigmaf.p, u.gradDf.p,
                                                              c->tracGrad.p, c->D.p, c->F, c->B, c->bOff(), c->ld, u.ldF,
                                                              c->unsUL() ? 2 : (c->unsTL() ? 1 : 0), c->incremental() ? c->Dold.p : nullptr);
    c->launches++;
    return 0;
}

int s4f_uns_source(s4fgpu_ctx* c) {
    if (!c->unsValid) { int rc = s4f_uns_setup(c); if (rc) return rc; }
    S4fUns& u = *c->uns;
    // updated Lagrangian: the explicit inertia terms exist for every scheme (fvc::d2dt2(rho_, D.oldTime())); rho()*g() with the
    // reference density is added here, not inside the history (unsNonLinGeomUpdatedLagSolid.C:263)
    const double* hist = (c->ctl.d2dt2Scheme == S4F_D2DT2_STEADY_STATE && !c->UL()) ? nullptr : c->d2Hist.p;
    const double rs = c->law.rho;
    k_source_uns<<<s4f_grid(c->numSMs, (long long)c->nSlices * 32, 4), S4F_BLOCK, 0, c->stream>>>(
        c->slicePtr.p, c->col.p, u.eFace.p, c->eSf.p, c->eA.p, c->D.p, u.sigmaf.p, c->V.p, hist, c->source.p, c->N, c->ld, u.ldF, c->nEntries,
        c->nSlices, rs * c->ctl.g[0], rs * c->ctl.g[1], rs * c->ctl.g[2], c->unsFinite() ? u.faceT.p : nullptr);
    c->launches++;
    return 0;
}

// time-step roll of the face deformation gradient: Ff.oldTime() = Ff
int s4f_uns_new_timestep(s4fgpu_ctx* c) {
    if (!c->unsUL() || !c->uns || c->uns->Ff.n == 0) return 0;
    S4F_CHECK_CUDA(c, cudaMemcpyAsync(c->uns->FfOld.p, c->uns->Ff.p, c->uns->Ff.n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

// host copies of the face fields (AoS, [F+B])
int s4f_uns_download(s4fgpu_ctx* c, int field, double* host) {
    if (!c->uns || !c->unsValid) { c->err = "face fields exist for the unsLinearGeometry model only"; return 1; }
    S4fUns& u = *c->uns;
    const int nc = field == S4F_FIELD_SIGMA_F ? 6 : 9;
    const double* src = field == S4F_FIELD_SIGMA_F ? u.sigmaf.p : u.gradDf.p;
    std::vector<double> h((size_t)nc * u.ldF);
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    S4F_CHECK_CUDA(c, cudaMemcpy(h.data(), src, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int f = 0; f < u.nF; f++) for (int q = 0; q < nc; q++) host[(size_t)nc * f + q] = h[(size_t)q * u.ldF + f];
    return 0;
}
