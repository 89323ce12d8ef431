// s4f_fv.cu -- finite-volume face loops as atomic-free cell-centric gathers over the SELL-32 rows:
//   * assembly of fvm::laplacian(impKf, D) (coefficients once, they only change with impKf),
//   * the explicit right-hand side: - fvc::laplacian(impKf, D) + fvc::div(sigma) + rho g + Rhie-Chow,
//   * fvc::grad(D) (least squares / Gauss linear) with the boundary normal-gradient correction,
//   * the D boundary conditions (solidTraction, fixedDisplacement, solidSymmetry),
//   * field relaxation and the residual reductions of solidModel::converged().
// Reference lines are cited at each kernel; the CPU restatement of the same operators in the
// reference's own LDU face-loop form is oracle/s4f_oracle.cpp.
#include <cmath>
#include <cstdlib>

#include "s4f_ctx.h"
#include "s4f_dev.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// fvm::laplacian(impKf, D): coefficient per row entry a = impKf_f * magSf_f * nonOrthDeltaCoeffs_f
// ([OF-ext] gaussLaplacianScheme::fvmLaplacianUncorrected: upper = a, diag = -sum a; the momentum
// equation "A == B" flips the sign: off-diagonal -a, diagonal +sum a), impKf = linear interpolate of
// impK (mechanicalModel.C:409-415).  Also the Rhie-Chow face coefficient gamma_f
// (momentumStabilisation.C:112-150, :198-206).   d2dt2Coeff*V added to the diagonal.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(S4F_BLOCK) k_assemble_laplacian(
    const int* __restrict__ slicePtr, const int* __restrict__ col, const double* __restrict__ eW, const double* __restrict__ eDn,
    const double* __restrict__ eSf, const double* __restrict__ eCorr /* may be null */, const double* __restrict__ impK,
    const double* __restrict__ V, double* __restrict__ eA, double* __restrict__ eGam, double* __restrict__ eU, double* __restrict__ eC0,
    double* __restrict__ eVc /* null when orthogonal */, double* __restrict__ rowK, double* __restrict__ diag0, int N, int bOff, int ld,
    long long nE, int nSlices, double stabScale, int stabOn, double d2dt2Coeff, const double* __restrict__ rhoF /* UL: density field */) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nSlices; s += nWarps) {
        const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5;
        const int row = s * 32 + lane;
        const double kP = (row < N) ? impK[row] : 0.0;
        double sum = 0, sU[3] = {0, 0, 0}, sV[3] = {0, 0, 0};
        for (int k = 0; k < width; k++) {
            const long long e = (long long)base + 32 * k + lane;
            const int cc = col[e];
            const double w = eW[e], dn = eDn[e], w1 = 1.0 - w;
            const double kN = impK[cc];
            const double kf = w * kP + w1 * kN;
            const double a = kf * dn;
            eA[e] = a;
            sum += a;
            double gf = 0.0;
            if (stabOn && cc < bOff) {   // zero on non-coupled boundary faces
                gf = w * (stabScale * kP) + w1 * (stabScale * kN);
                if (fabs(kP - kN) > S4F_SMALL) gf = 0.01 * 0.5 * (kP + kN);   // material interface
            }
            eGam[e] = gf;
            // factored right-hand-side coefficients (see k_source_g): the neighbour value of M = T - gamma grad(D) enters
            // with u = (1-w) Sf, of D with c0 = gamma_f magSf delta - a, of grad(D) with vc = gamma (1-w) corr (non-orthogonal
            // meshes); the row's own values with the sums U = sum w Sf and Vc = sum gamma w corr over its entries.
            const double c0 = gf * dn - a;
            eC0[e] = c0;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const double Sq = eSf[(size_t)q * nE + e];
                eU[(size_t)q * nE + e] = w1 * Sq;
                sU[q] += w * Sq;
                if (eCorr) {
                    const double cq = eCorr[(size_t)q * nE + e];
                    eVc[(size_t)q * nE + e] = gf * w1 * cq;
                    sV[q] += gf * w * cq;
                }
            }
        }
        if (row < N) {
            diag0[row] = sum + d2dt2Coeff * (rhoF ? rhoF[row] : 1.0) * V[row];
#pragma unroll
            for (int q = 0; q < 3; q++) { rowK[(size_t)q * ld + row] = sU[q]; rowK[(size_t)(3 + q) * ld + row] = sV[q]; }
        }
    }
}

__global__ void k_diag_copy(const double* __restrict__ diag0, double* __restrict__ diagC, int N, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double d = diag0[i];
    diagC[i] = d; diagC[(size_t)ld + i] = d; diagC[2 * (size_t)ld + i] = d;
}

__global__ void k_diag_recip(const double* __restrict__ diagC, double* __restrict__ rD, int N, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
#pragma unroll
    for (int q = 0; q < 3; q++) rD[(size_t)q * ld + i] = 1.0 / diagC[(size_t)q * ld + i];
}

// addBoundaryDiag: internalCoeffs = impKf_b*magSf_b*deltaCoeffs_b for fixedValue, times |n_c| for the
// symmetry plane ([OF-ext] basicSymmetry snGradTransformDiag), zero for fixedGradient.
__global__ void k_diag_boundary(const int* __restrict__ bcCells, const int* __restrict__ bcPtr, const int* __restrict__ bcFaces,
                                const int* __restrict__ bKind, const double* __restrict__ bN, const double* __restrict__ bDelta,
                                const double* __restrict__ bMagSf, const double* __restrict__ impK, double* __restrict__ diagC,
                                int nBCells, int B, int bOff, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nBCells) return;
    const int P = bcCells[i];
    double add[3] = {0, 0, 0};
    for (int j = bcPtr[i]; j < bcPtr[i + 1]; j++) {
        const int b = bcFaces[j];
        const int kind = bKind[b];
        const double gm = impK[bOff + b] * bMagSf[b] * bDelta[b];
        if (kind == S4F_BC_FIXED_DISPLACEMENT) { add[0] += gm; add[1] += gm; add[2] += gm; }
        else if (kind == S4F_BC_SOLID_SYMMETRY) {
            add[0] += gm * fabs(bN[b]); add[1] += gm * fabs(bN[(size_t)B + b]); add[2] += gm * fabs(bN[2 * (size_t)B + b]);
        }
    }
    diagC[P] += add[0]; diagC[(size_t)ld + P] += add[1]; diagC[2 * (size_t)ld + P] += add[2];
}

// ------------------------------------------------------------------------------------------------
// updateCoeffs() of the patches (triggered by the fvMatrix constructor):
//  solidTraction  gradient() = tractionBoundarySnGrad: linGeomTotalDispSolid.C:235-271
//                 ((t - n p) - (n & (sigma_b - impK gradD_b)))/impK ;  TL form with the deformed normal
//                 nonLinGeomTotalLagTotalDispSolid.C:284-328
//  fixedDisplacement  value = totalDisp: fixedDisplacementFvPatchVectorField.C:258-294
// ------------------------------------------------------------------------------------------------
__global__ void k_bc_update(const int* __restrict__ bKind, const double* __restrict__ bN, const double* __restrict__ bcValue,
                            const double* __restrict__ bcPressure, const double* __restrict__ impK, const double* __restrict__ sigma,
                            const double* __restrict__ gradD, const double* __restrict__ Finv, double* __restrict__ tracGrad,
                            double* __restrict__ D, const double* __restrict__ DoldIncr /* null unless the field is DD */, int B,
                            int bOff, int ld, int TL) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int kind = bKind[b];
    const size_t j = (size_t)bOff + b;
    if (kind == S4F_BC_FIXED_DISPLACEMENT) {
#pragma unroll
        for (int c = 0; c < 3; c++)      // DD field: disp -= Dold.boundaryField()  (fixedDisplacement...C:279-287)
            D[(size_t)c * ld + j] = bcValue[(size_t)c * B + b] - (DoldIncr ? DoldIncr[(size_t)c * ld + j] : 0.0);
    } else if (kind == S4F_BC_SOLID_TRACTION) {
        double n[3], t[3], g[9], s[6];
#pragma unroll
        for (int c = 0; c < 3; c++) { n[c] = bN[(size_t)c * B + b]; t[c] = bcValue[(size_t)c * B + b]; }
#pragma unroll
        for (int q = 0; q < 9; q++) g[q] = gradD[(size_t)q * ld + j];
#pragma unroll
        for (int q = 0; q < 6; q++) s[q] = sigma[(size_t)q * ld + j];
        const double p = bcPressure[b], k = impK[j], rk = 1.0 / k;
        double out[3];
        if (!TL) {
            double M[9]; s_to_t(s, M);
#pragma unroll
            for (int q = 0; q < 9; q++) M[q] -= k * g[q];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double nM = n[0] * M[c] + n[1] * M[3 + c] + n[2] * M[6 + c];
                out[c] = ((t[c] - n[c] * p) - nM) * rk;
            }
        } else {
            double Fi[9];
#pragma unroll
            for (int q = 0; q < 9; q++) Fi[q] = Finv[(size_t)q * ld + j];
            double nc[3];   // Finv.T() & n
#pragma unroll
            for (int c = 0; c < 3; c++) nc[c] = Fi[c] * n[0] + Fi[3 + c] * n[1] + Fi[6 + c] * n[2];
            const double m = sqrt(nc[0] * nc[0] + nc[1] * nc[1] + nc[2] * nc[2]);
            nc[0] /= m; nc[1] /= m; nc[2] /= m;
            const double ns[3] = {s[0] * nc[0] + s[1] * nc[1] + s[2] * nc[2], s[1] * nc[0] + s[3] * nc[1] + s[4] * nc[2],
                                  s[2] * nc[0] + s[4] * nc[1] + s[5] * nc[2]};
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double ng = n[0] * g[c] + n[1] * g[3 + c] + n[2] * g[6 + c];
                out[c] = ((t[c] - nc[c] * p) - ns[c] + k * ng) * rk;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) tracGrad[(size_t)c * B + b] = out[c];
    }
}

// snGrad() of boundary face b with the registered grad(D) at the face cell:
//  solidTraction: gradient();  fixedDisplacement: (D_b - (D_P + k & gradD_P)) deltaCoeffs (fixedDisplacement...C:297-326)
//  solidSymmetry: (transform(I - 2nn, DP) - DP) deltaCoeffs/2  (solidSymmetry...C:148-196)
__device__ __forceinline__ void bc_sngrad(int kind, int b, int P, int B, int bOff, int ld, const double* __restrict__ bN,
                                          const double* __restrict__ bK, const double* __restrict__ bDelta,
                                          const double* __restrict__ tracGrad, const double* __restrict__ D,
                                          const double* __restrict__ gradD, double* sn, double* kgOut) {
    if (kind == S4F_BC_SOLID_TRACTION) {
#pragma unroll
        for (int c = 0; c < 3; c++) sn[c] = tracGrad[(size_t)c * B + b];
        if (kgOut) {
            double k[3] = {bK[b], bK[(size_t)B + b], bK[2 * (size_t)B + b]};
#pragma unroll
            for (int c = 0; c < 3; c++) kgOut[c] = k[0] * gradD[(size_t)c * ld + P] + k[1] * gradD[(size_t)(3 + c) * ld + P] + k[2] * gradD[(size_t)(6 + c) * ld + P];
        }
        return;
    }
    const double k[3] = {bK[b], bK[(size_t)B + b], bK[2 * (size_t)B + b]};
    const double delta = bDelta[b];
    double kg[3], DP[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        kg[c] = k[0] * gradD[(size_t)c * ld + P] + k[1] * gradD[(size_t)(3 + c) * ld + P] + k[2] * gradD[(size_t)(6 + c) * ld + P];
        DP[c] = D[(size_t)c * ld + P] + kg[c];
        if (kgOut) kgOut[c] = kg[c];
    }
    if (kind == S4F_BC_FIXED_DISPLACEMENT) {
#pragma unroll
        for (int c = 0; c < 3; c++) sn[c] = (D[(size_t)c * ld + bOff + b] - DP[c]) * delta;
    } else {
        const double n[3] = {bN[b], bN[(size_t)B + b], bN[2 * (size_t)B + b]};
        const double nDP = n[0] * DP[0] + n[1] * DP[1] + n[2] * DP[2];
#pragma unroll
        for (int c = 0; c < 3; c++) sn[c] = ((DP[c] - 2.0 * n[c] * nDP) - DP[c]) * (delta / 2.0);
    }
}

// snGrad of every boundary face stored for the gradient's boundary correction (uses the OLD grad(D):
// gradD = fvc::grad(D) is assigned after the evaluation)
__global__ void k_bc_sngrad_store(const int* __restrict__ bFaceCell, const int* __restrict__ bKind, const double* __restrict__ bN,
                                  const double* __restrict__ bK, const double* __restrict__ bDelta, const double* __restrict__ tracGrad,
                                  const double* __restrict__ D, const double* __restrict__ gradD, double* __restrict__ bSn, int B,
                                  int bOff, int ld) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int kind = bKind[b];
    if (kind == S4F_BC_PROCESSOR) return;
    double sn[3];
    bc_sngrad(kind, b, bFaceCell[b], B, bOff, ld, bN, bK, bDelta, tracGrad, D, gradD, sn, nullptr);
#pragma unroll
    for (int c = 0; c < 3; c++) bSn[(size_t)c * B + b] = sn[c];
}

// evaluate(): solidTraction D_b = D_P + k & gradD_P + gradient()/deltaCoeffs (solidTraction...C:398-463);
// solidSymmetry D_b = (DP + transform(I-2nn, DP))/2 (solidSymmetry...C:200-260)
__global__ void k_bc_evaluate(const int* __restrict__ bFaceCell, const int* __restrict__ bKind, const double* __restrict__ bN,
                              const double* __restrict__ bK, const double* __restrict__ bDelta, const double* __restrict__ tracGrad,
                              double* __restrict__ D, const double* __restrict__ gradD, int B, int bOff, int ld) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int kind = bKind[b];
    if (kind != S4F_BC_SOLID_TRACTION && kind != S4F_BC_SOLID_SYMMETRY) return;
    const int P = bFaceCell[b];
    const double k[3] = {bK[b], bK[(size_t)B + b], bK[2 * (size_t)B + b]};
    double DP[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double kg = k[0] * gradD[(size_t)c * ld + P] + k[1] * gradD[(size_t)(3 + c) * ld + P] + k[2] * gradD[(size_t)(6 + c) * ld + P];
        DP[c] = D[(size_t)c * ld + P] + kg;
    }
    if (kind == S4F_BC_SOLID_TRACTION) {
        const double delta = bDelta[b];
#pragma unroll
        for (int c = 0; c < 3; c++) D[(size_t)c * ld + bOff + b] = DP[c] + tracGrad[(size_t)c * B + b] / delta;
    } else {
        const double n[3] = {bN[b], bN[(size_t)B + b], bN[2 * (size_t)B + b]};
        const double nDP = n[0] * DP[0] + n[1] * DP[1] + n[2] * DP[2];
#pragma unroll
        for (int c = 0; c < 3; c++) D[(size_t)c * ld + bOff + b] = (DP[c] + (DP[c] - 2.0 * n[c] * nDP)) / 2.0;
    }
}

// ------------------------------------------------------------------------------------------------
// Explicit right-hand side, one gather per cell over its faces:
//   - V fvc::laplacian(impKf,D) [compact part; the non-orthogonal parts of fvm and fvc cancel]
//   + V fvc::div(T)            [OF-ext] gaussDivScheme + linear:  Sf & (w T_P + (1-w) T_N); boundary Sf_b & T_b
//   + V rho g + d2dt2 old-time terms
//   + V RhieChow = sum_f gamma_f [ magSf (delta (D_N-D_P) + corr & gradD_f) - Sf & gradD_f ]   (momentumStabilisation.C:210-217)
// T is sigma (linear geometry) or J Finv & sigma / relJ relFinv & sigma (total / updated Lagrangian).
//
// Factored form.  With linear face interpolation w X_P + (1-w) X_N and the uniform Rhie-Chow coefficient gamma of a
// single-law case, the kernel that produces the stress also writes ONE combined tensor per cell,
//      M = T - gamma grad(D)   (cells, ghost cells);   M_b = T_b on boundary slots (gamma_f = 0 there),
// and the sum over the faces of a cell separates into a part that only needs NEIGHBOUR values and per-entry
// coefficients fixed at assembly, and a part in the cell's own values with per-row sums:
//      rhs_q(P) = sum_e [ u_e & M_N(:,q) + vc_e & gradD_N(:,q) + c0_e (D_N,q - D_P,q) ]
//                 + U & M_P(:,q) + Vc & gradD_P(:,q) + V (rho g_q + hist_q),
//      u = (1-w) Sf,  vc = gamma (1-w) corr,  c0 = gamma magSf delta - a,  U = sum w Sf,  Vc = sum gamma w corr.
// k_source_g is the general kernel (non-orthogonal meshes: plateHole, the notched bar, every moved updated-Lagrangian
// mesh): it streams col, u, vc, c0 (60 B per entry) and gathers M, grad(D), D (21 values).  Round 1's version gave every
// displacement component its own warp (each re-streaming the coefficients and gathering T, grad(D) and gamma_f
// separately: 0.41 of the HBM peak); here one lane owns a row for all three components, as in k_source_m.
// On orthogonal meshes vc = Vc = 0 and the gather of grad(D) disappears: k_source_m below.
// ------------------------------------------------------------------------------------------------
template <int MINB>
__global__ void __launch_bounds__(S4F_BLOCK, MINB) k_source_g(
    const int* __restrict__ slicePtr, const int* __restrict__ col, const double* __restrict__ eU, const double* __restrict__ eVc,
    const double* __restrict__ eC0, const double* __restrict__ rowK, const double* __restrict__ D, const double* __restrict__ M,
    const double* __restrict__ gradD, const double* __restrict__ V, const double* __restrict__ hist, double* __restrict__ source, int N,
    int ld, long long nE, int nSlices, double rhoGx, double rhoGy, double rhoGz) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    const double* __restrict__ eU1 = eU + nE;
    const double* __restrict__ eU2 = eU + 2 * nE;
    const double* __restrict__ eV1 = eVc + nE;
    const double* __restrict__ eV2 = eVc + 2 * nE;
    for (int s = warp; s < nSlices; s += nWarps) {
        const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5;
        const int row = s * 32 + lane;
        const int r = (row < N) ? row : 0;
        const double DP[3] = {D[r], D[(size_t)ld + r], D[2 * (size_t)ld + r]};
        double acc[3] = {0, 0, 0};
        // software pipeline over the entries: the streamed coefficients of entry k+1 are loaded while entry k's 21 gathers fly
        int cc = __ldcs(col + base + lane);
        double u0 = 0, u1 = 0, u2 = 0, v0 = 0, v1 = 0, v2 = 0, c0 = 0;
        if (width > 0) {
            const long long e = (long long)base + lane;
            u0 = __ldcs(eU + e); u1 = __ldcs(eU1 + e); u2 = __ldcs(eU2 + e);
            v0 = __ldcs(eVc + e); v1 = __ldcs(eV1 + e); v2 = __ldcs(eV2 + e); c0 = __ldcs(eC0 + e);
        }
        for (int k = 0; k < width; k++) {
            const int n = cc;
            const double a0 = u0, a1 = u1, a2 = u2, b0 = v0, b1 = v1, b2 = v2, cz = c0;
            double m[9], g[9], d[3];
            // faces without a correction vector (the orthogonal part of a hex-dominant mesh) need no grad(D): warp-uniform skip
            const bool needG = __any_sync(0xffffffffu, (b0 != 0.0) | (b1 != 0.0) | (b2 != 0.0));
#pragma unroll
            for (int q = 0; q < 9; q++) m[q] = M[(size_t)q * ld + n];
            if (needG) {
#pragma unroll
                for (int q = 0; q < 9; q++) g[q] = gradD[(size_t)q * ld + n];
            } else {
#pragma unroll
                for (int q = 0; q < 9; q++) g[q] = 0.0;
            }
#pragma unroll
            for (int q = 0; q < 3; q++) d[q] = D[(size_t)q * ld + n];
            if (k + 1 < width) {
                const long long e = (long long)base + 32 * (k + 1) + lane;
                cc = __ldcs(col + e);
                u0 = __ldcs(eU + e); u1 = __ldcs(eU1 + e); u2 = __ldcs(eU2 + e);
                v0 = __ldcs(eVc + e); v1 = __ldcs(eV1 + e); v2 = __ldcs(eV2 + e); c0 = __ldcs(eC0 + e);
            }
#pragma unroll
            for (int q = 0; q < 3; q++)
                acc[q] += a0 * m[q] + a1 * m[3 + q] + a2 * m[6 + q] + b0 * g[q] + b1 * g[3 + q] + b2 * g[6 + q] + cz * (d[q] - DP[q]);
        }
        if (row < N) {
            const double U0 = rowK[row], U1 = rowK[(size_t)ld + row], U2 = rowK[2 * (size_t)ld + row], v = V[row];
            const double W0 = rowK[3 * (size_t)ld + row], W1 = rowK[4 * (size_t)ld + row], W2 = rowK[5 * (size_t)ld + row];
            const double rg[3] = {rhoGx, rhoGy, rhoGz};
#pragma unroll
            for (int q = 0; q < 3; q++) {
                double sv = acc[q] + U0 * M[(size_t)q * ld + row] + U1 * M[(size_t)(3 + q) * ld + row] + U2 * M[(size_t)(6 + q) * ld + row]
                          + W0 * gradD[(size_t)q * ld + row] + W1 * gradD[(size_t)(3 + q) * ld + row] + W2 * gradD[(size_t)(6 + q) * ld + row];
                sv += v * rg[q];
                if (hist) sv += v * hist[(size_t)q * ld + row];
                source[(size_t)q * ld + row] = sv;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Right-hand side on ORTHOGONAL meshes (no correction vectors): the factored form above without the grad(D) terms,
//      rhs_q(P) = sum_e [ u_e & M_N(:,q) + c0_e (D_N,q - D_P,q) ] + U & M_P(:,q) + V (rho g_q + hist_q).
// M is written by the kernel that produces the stress (law kernels / the flux-tensor kernel), so the right-hand side
// streams 5 values per entry (col, u, c0) and gathers 12 (M, D).  One row per lane, a slice per warp, entries in pairs
// with the streamed loads issued ahead of the gathers.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(S4F_BLOCK, 3) k_source_m(
    const int* __restrict__ slicePtr, const int* __restrict__ col, const double* __restrict__ eU, const double* __restrict__ eC0,
    const double* __restrict__ rowK, const double* __restrict__ D, const double* __restrict__ M, const double* __restrict__ V,
    const double* __restrict__ hist, double* __restrict__ source, int N, int ld, long long nE, int nSlices, double rhoGx, double rhoGy,
    double rhoGz) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    const double* __restrict__ eU1 = eU + nE;
    const double* __restrict__ eU2 = eU + 2 * nE;
    constexpr int G = 2;
    for (int s = warp; s < nSlices; s += nWarps) {
        const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5;
        const int row = s * 32 + lane;
        const int r = (row < N) ? row : 0;
        const double DP[3] = {D[r], D[(size_t)ld + r], D[2 * (size_t)ld + r]};
        double acc[3] = {0, 0, 0};
        for (int k0 = 0; k0 < width; k0 += G) {
            int cc[G]; double u0[G], u1[G], u2[G], c0[G];
#pragma unroll
            for (int k = 0; k < G; k++) {
                const bool ok = k0 + k < width;
                const long long e = (long long)base + 32 * (ok ? k0 + k : k0) + lane;
                cc[k] = __ldcs(col + e);            // streamed once: evict-first, keep L2 for the gathered M / D lines
                u0[k] = ok ? __ldcs(eU + e) : 0.0; u1[k] = ok ? __ldcs(eU1 + e) : 0.0; u2[k] = ok ? __ldcs(eU2 + e) : 0.0;
                c0[k] = ok ? __ldcs(eC0 + e) : 0.0;
            }
            double m[G][9], d[G][3];
#pragma unroll
            for (int k = 0; k < G; k++) {
                const int n = cc[k];
#pragma unroll
                for (int q = 0; q < 9; q++) m[k][q] = M[(size_t)q * ld + n];
#pragma unroll
                for (int q = 0; q < 3; q++) d[k][q] = D[(size_t)q * ld + n];
            }
#pragma unroll
            for (int k = 0; k < G; k++)
#pragma unroll
                for (int q = 0; q < 3; q++)
                    acc[q] += u0[k] * m[k][q] + u1[k] * m[k][3 + q] + u2[k] * m[k][6 + q] + c0[k] * (d[k][q] - DP[q]);
        }
        if (row < N) {
            const double U0 = rowK[row], U1 = rowK[(size_t)ld + row], U2 = rowK[2 * (size_t)ld + row], v = V[row];
            const double rg[3] = {rhoGx, rhoGy, rhoGz};
#pragma unroll
            for (int q = 0; q < 3; q++) {
                double sv = acc[q] + U0 * M[(size_t)q * ld + row] + U1 * M[(size_t)(3 + q) * ld + row] + U2 * M[(size_t)(6 + q) * ld + row];
                sv += v * rg[q];
                if (hist) sv += v * hist[(size_t)q * ld + row];
                source[(size_t)q * ld + row] = sv;
            }
        }
    }
}

// M = sigma - gamma grad(D) for the linear-geometry model when the stress was not produced by a law kernel
// (uploaded / initial fields)
__global__ void __launch_bounds__(S4F_BLOCK) k_make_m(const double* __restrict__ sigma, const double* __restrict__ gradD,
                                                      double* __restrict__ M, int N, int bOff, int B, int ld, double gamma) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N + B; t += gridDim.x * blockDim.x) {
        const int i = (t < N) ? t : bOff + (t - N);
        double s6[6], s9[9];
#pragma unroll
        for (int q = 0; q < 6; q++) s6[q] = sigma[(size_t)q * ld + i];
        s_to_t(s6, s9);
        const double g = (t < N) ? gamma : 0.0;
#pragma unroll
        for (int q = 0; q < 9; q++) M[(size_t)q * ld + i] = s9[q] - g * gradD[(size_t)q * ld + i];
    }
}

// boundary part of the laplacian pair: - impKf_b magSf_b snGrad_b (explicit) + boundaryCoeffs
// (addBoundarySource), boundaryCoeffs = impKf_b magSf_b gradientBoundaryCoeffs_b with
//   fixedGradient: gradient();  fixedDisplacement: deltaCoeffs (D_b - k & gradD_P)  (fixedDisplacement...C:328-356)
//   symmetry [OF-ext] transformFvPatchField: snGrad() + deltaCoeffs |n_c| D_P,c
__global__ void k_source_boundary(const int* __restrict__ bcCells, const int* __restrict__ bcPtr, const int* __restrict__ bcFaces,
                                  const int* __restrict__ bKind, const double* __restrict__ bN, const double* __restrict__ bK,
                                  const double* __restrict__ bDelta, const double* __restrict__ bMagSf, const double* __restrict__ impK,
                                  const double* __restrict__ tracGrad, const double* __restrict__ D, const double* __restrict__ gradD,
                                  double* __restrict__ source, int nBCells, int B, int bOff, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nBCells) return;
    const int P = bcCells[i];
    double add[3] = {0, 0, 0};
    for (int j = bcPtr[i]; j < bcPtr[i + 1]; j++) {
        const int b = bcFaces[j];
        const int kind = bKind[b];
        const double gm = impK[bOff + b] * bMagSf[b];
        double sn[3], kg[3];
        bc_sngrad(kind, b, P, B, bOff, ld, bN, bK, bDelta, tracGrad, D, gradD, sn, kg);
        double gbc[3];
        if (kind == S4F_BC_SOLID_TRACTION) { gbc[0] = sn[0]; gbc[1] = sn[1]; gbc[2] = sn[2]; }
        else if (kind == S4F_BC_FIXED_DISPLACEMENT) {
            const double delta = bDelta[b];
#pragma unroll
            for (int c = 0; c < 3; c++) gbc[c] = delta * (D[(size_t)c * ld + bOff + b] - kg[c]);
        } else {
            const double delta = bDelta[b];
#pragma unroll
            for (int c = 0; c < 3; c++) gbc[c] = sn[c] + delta * fabs(bN[(size_t)c * B + b]) * D[(size_t)c * ld + P];
        }
#pragma unroll
        for (int c = 0; c < 3; c++) add[c] += -gm * sn[c] + gm * gbc[c];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) source[(size_t)c * ld + P] += add[c];
}

// ------------------------------------------------------------------------------------------------
// fvc::grad(D).  Least squares: grad_P = sum_e ls_e (D_e - D_P)  (extendedLeastSquaresGrad.C:103-152 in
// gather form; boundary faces are row entries whose column is the boundary-value slot).
// Gauss linear: grad_P = (1/V) sum_e Sf_e (w D_P + (1-w) D_e).     mechanicalModel::grad, mechanicalModel.C:571-582
// ------------------------------------------------------------------------------------------------
template <bool GAUSS>
__global__ void __launch_bounds__(S4F_BLOCK, GAUSS ? 2 : 3) k_grad(const int* __restrict__ slicePtr, const int* __restrict__ col,
                                                       const double* __restrict__ eVec /* eLs or eSf */, const double* __restrict__ eW,
                                                       const double* __restrict__ D, const double* __restrict__ rV,
                                                       double* __restrict__ gradD, int N, int ld, long long nE, int nSlices) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    constexpr int G = 4;    // entries whose index / vector loads are issued together, before the gathers
    for (int s = warp; s < nSlices; s += nWarps) {
        const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5;
        const int row = s * 32 + lane;
        const int r = (row < N) ? row : 0;
        const double DP[3] = {D[r], D[(size_t)ld + r], D[2 * (size_t)ld + r]};
        double g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int k0 = 0; k0 < width; k0 += G) {
            int cc[G]; double v[G][3]; double wv[G];
#pragma unroll
            for (int k = 0; k < G; k++) {
                const bool ok = k0 + k < width;
                const long long e = (long long)base + 32 * (ok ? k0 + k : k0) + lane;
                cc[k] = col[e];
                v[k][0] = ok ? eVec[e] : 0.0; v[k][1] = ok ? eVec[nE + e] : 0.0; v[k][2] = ok ? eVec[2 * nE + e] : 0.0;
                if (GAUSS) wv[k] = eW[e];
            }
#pragma unroll
            for (int k = 0; k < G; k++) {
                double d[3];
                if (GAUSS) {
#pragma unroll
                    for (int c = 0; c < 3; c++) d[c] = wv[k] * DP[c] + (1.0 - wv[k]) * D[(size_t)c * ld + cc[k]];
                } else {
#pragma unroll
                    for (int c = 0; c < 3; c++) d[c] = D[(size_t)c * ld + cc[k]] - DP[c];
                }
#pragma unroll
                for (int i = 0; i < 3; i++)
#pragma unroll
                    for (int j = 0; j < 3; j++) g[3 * i + j] += v[k][i] * d[j];
            }
        }
        if (row < N) {
            const double sc = GAUSS ? rV[row] : 1.0;
#pragma unroll
            for (int q = 0; q < 9; q++) gradD[(size_t)q * ld + row] = g[q] * sc;
        }
    }
}

// gaussGrad::correctBoundaryConditions: grad_b = grad_P + n (snGrad_b - n & grad_P)
__global__ void k_grad_boundary(const int* __restrict__ bFaceCell, const int* __restrict__ bKind, const double* __restrict__ bN,
                                const double* __restrict__ bSn, double* __restrict__ gradD, int B, int bOff, int ld) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (bKind[b] == S4F_BC_PROCESSOR) return;
    const int P = bFaceCell[b];
    const double n[3] = {bN[b], bN[(size_t)B + b], bN[2 * (size_t)B + b]};
    double g[9];
#pragma unroll
    for (int q = 0; q < 9; q++) g[q] = gradD[(size_t)q * ld + P];
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

// ------------------------------------------------------------------------------------------------
// relaxField (fixed): D = D.prevIter + alpha (D - D.prevIter) on cells and boundary values
// (solidModel.C:823-906 -> [OF-ext] GeometricField::relax), fused with the three gMax reductions of
// converged(): solidModelTemplates.C:42-103.
// ------------------------------------------------------------------------------------------------
struct FinOuter {
    OuterScalars* S;
    __device__ void operator()(const double* tot) const { S->maxDelta = tot[0]; S->maxIncr = tot[1]; S->maxMag = tot[2]; }
};
__global__ void __launch_bounds__(S4F_BLOCK) k_relax_residual(double* __restrict__ D, const double* __restrict__ Dprev,
                                                              const double* __restrict__ Dold, int N, int bOff, int B, int ld,
                                                              double alpha, OuterScalars* S, RedCtx red) {
    double v[3] = {0, 0, 0};
    const int total = N + B;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int i = (t < N) ? t : bOff + (t - N);
        double d[3], dp[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            dp[c] = Dprev[(size_t)c * ld + i];
            d[c] = D[(size_t)c * ld + i];
            if (alpha != 1.0) { d[c] = dp[c] + alpha * (d[c] - dp[c]); D[(size_t)c * ld + i] = d[c]; }
        }
        if (t < N) {
            const double o[3] = {Dold[i], Dold[(size_t)ld + i], Dold[2 * (size_t)ld + i]};
            const double a = sqrt((d[0] - dp[0]) * (d[0] - dp[0]) + (d[1] - dp[1]) * (d[1] - dp[1]) + (d[2] - dp[2]) * (d[2] - dp[2]));
            const double bb = sqrt((d[0] - o[0]) * (d[0] - o[0]) + (d[1] - o[1]) * (d[1] - o[1]) + (d[2] - o[2]) * (d[2] - o[2]));
            const double m = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            v[0] = fmax(v[0], a); v[1] = fmax(v[1], bb); v[2] = fmax(v[2], m);
        }
    }
    grid_reduce<3, OpMax>(v, red, FinOuter{S});
}

// Aitken relaxation (solidModel.C:842-897): cell-wise alpha, incl. boundary values
__global__ void __launch_bounds__(S4F_BLOCK) k_relax_aitken(double* __restrict__ D, const double* __restrict__ Dprev,
                                                            const double* __restrict__ Dold, double* __restrict__ res,
                                                            double* __restrict__ resPrev, double* __restrict__ aAlpha, int N, int bOff,
                                                            int B, int ld, int first, double alpha0, OuterScalars* S, RedCtx red) {
    double v[3] = {0, 0, 0};
    const int total = N + B;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int i = (t < N) ? t : bOff + (t - N);
        double d[3], dp[3], rn[3], rp[3];
        double num = 0, den = 0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const size_t j = (size_t)c * ld + i;
            dp[c] = Dprev[j]; d[c] = D[j];
            rp[c] = res[j];               // aitkenResidual_.storePrevIter()
            rn[c] = dp[c] - d[c];
            const double dl = rp[c] - rn[c];
            num += rp[c] * dl; den += dl * dl;
        }
        double a;
        if (first) a = alpha0;
        else { a = aAlpha[i] * num / (den + S4F_SMALL); a = fmax(0.0, fmin(2.0, a)); }
        aAlpha[i] = a;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const size_t j = (size_t)c * ld + i;
            resPrev[j] = rp[c]; res[j] = rn[c];
            d[c] -= a * rn[c];
            D[j] = d[c];
        }
        if (t < N) {
            const double o[3] = {Dold[i], Dold[(size_t)ld + i], Dold[2 * (size_t)ld + i]};
            const double x = sqrt((d[0] - dp[0]) * (d[0] - dp[0]) + (d[1] - dp[1]) * (d[1] - dp[1]) + (d[2] - dp[2]) * (d[2] - dp[2]));
            const double bb = sqrt((d[0] - o[0]) * (d[0] - o[0]) + (d[1] - o[1]) * (d[1] - o[1]) + (d[2] - o[2]) * (d[2] - o[2]));
            const double m = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            v[0] = fmax(v[0], x); v[1] = fmax(v[1], bb); v[2] = fmax(v[2], m);
        }
    }
    grid_reduce<3, OpMax>(v, red, FinOuter{S});
}

}  // namespace

// ================================================================================================
// host drivers
// ================================================================================================

// rho*fvm::d2dt2(D) as  diag += k.diag V  and  source += V (h0 D.o + h1 D.oo + h2 D.ooo + h3 D.oooo):
//  Euler     [OF-ext] EulerD2dt2Scheme::fvmD2dt2 (variable deltaT form)
//  backward  numerics/backwardD2dt2Scheme/backwardD2dt2Scheme.C:309-395 (coefficients :350-356; deltaT0_(vf) = GREAT
//            during the first time step :48-68) around [OF-ext] backwardDdtScheme::fvmDdt / fvcDdt; see the oracle.
struct D2dt2Coeffs { double diag, h[4]; };
static D2dt2Coeffs d2dt2_coeffs(const s4fgpu_ctx* c) {
    D2dt2Coeffs k{0, {0, 0, 0, 0}};
    const double dt = c->ctl.deltaT, dt0 = c->ctl.deltaT0 > 0 ? c->ctl.deltaT0 : dt, rho = c->law.rho;
    if (c->ctl.d2dt2Scheme == S4F_D2DT2_EULER) {
        const double coefft = (dt + dt0) / (2 * dt), coefft00 = (dt + dt0) / (2 * dt0), rDeltaT2 = 4.0 / ((dt + dt0) * (dt + dt0));
        k.diag = coefft * rDeltaT2 * rho;
        k.h[0] = rDeltaT2 * rho * (coefft + coefft00); k.h[1] = -rDeltaT2 * rho * coefft00;
    } else if (c->ctl.d2dt2Scheme == S4F_D2DT2_BACKWARD) {
        const bool first = c->timeIndex <= 1;
        const double cc = first ? 1.0 : 1.0 + dt / (dt + dt0), c00 = first ? 0.0 : dt * dt / (dt0 * (dt + dt0)), c0 = cc + c00;
        const double b = 1.0 + dt / (dt + dt0), b00 = dt * dt / (dt0 * (dt + dt0)), b0 = b + b00;
        const double r = rho / (dt * dt);
        k.diag = r * cc * b;
        k.h[0] = r * (cc * b0 + c0 * b);
        k.h[1] = r * (-cc * b00 - c0 * b0 - c00 * b);
        k.h[2] = r * (c0 * b00 + c00 * b0);
        k.h[3] = -r * c00 * b00;
    }
    return k;
}

namespace {
__global__ void k_d2dt2_hist(const double* __restrict__ D1, const double* __restrict__ D2, const double* __restrict__ D3,
                             const double* __restrict__ D4, double* __restrict__ hist, long long n, double h0, double h1, double h2,
                             double h3) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) hist[i] = h0 * D1[i] + h1 * D2[i] + h2 * D3[i] + h3 * D4[i];
}
}  // namespace

// Updated-Lagrangian inertia fvm::d2dt2(rho_, DD()) + fvc::d2dt2(rho_, D().oldTime()) with the density field
// (nonLinGeomUpdatedLagSolid.C:173-176; backwardD2dt2Scheme.C:149-222, :391-470; [OF-ext] EulerD2dt2Scheme rho-field
// overloads): diagonal coefficient (times rho_P V_P) and the explicit part per unit volume incl. rho_P g.  The
// restatement with its coefficient table is spelled out in the oracle (ulD2dt2).
struct UlCoeffs { double diag; double m[4]; double mo[3]; double moo[3]; double c[3]; double co[3]; double coo[3]; double g[3]; int mode; };
static UlCoeffs ul_coeffs(const s4fgpu_ctx* c) {
    UlCoeffs k{}; k.mode = c->ctl.d2dt2Scheme;
    for (int q = 0; q < 3; q++) k.g[q] = c->unsUL() ? 0.0 : c->ctl.g[q];      // the uns model adds rho()*g() itself (s4f_uns_source)
    const double dt = c->ctl.deltaT, dt0 = c->ctl.deltaT0 > 0 ? c->ctl.deltaT0 : dt;
    if (k.mode == S4F_D2DT2_EULER) {
        const double cf = (dt + dt0) / (2 * dt), cf00 = (dt + dt0) / (2 * dt0), r2 = 4.0 / ((dt + dt0) * (dt + dt0));
        k.diag = cf * r2;
        k.m[0] = r2 * (cf + cf00); k.m[1] = -r2 * cf00;                   // rho * (.. DD.o .. DD.oo)
        k.c[0] = -r2 * cf; k.c[1] = r2 * (cf + cf00); k.c[2] = -r2 * cf00;   // rho * (.. D.o, D.oo, D.ooo)
    } else if (k.mode == S4F_D2DT2_BACKWARD) {
        const bool firstM = c->timeIndex <= 1, firstC = c->timeIndex <= 2;
        const double kb = 1.0 + dt / (dt + dt0), kb00 = dt * dt / (dt0 * (dt + dt0)), kb0 = kb + kb00;
        const double cm = firstM ? 1.0 : kb, cm00 = firstM ? 0.0 : kb00, cm0 = cm + cm00;
        const double cc = firstC ? 1.0 : kb, cc00 = firstC ? 0.0 : kb00, cc0 = cc + cc00;
        const double r = 1.0 / (dt * dt);
        k.diag = r * cm * kb;
        k.m[0] = r * cm * kb0; k.m[1] = -r * cm * kb00;                                      // rho    * (DD.o, DD.oo)
        k.mo[0] = r * cm0 * kb; k.mo[1] = -r * cm0 * kb0; k.mo[2] = r * cm0 * kb00;          // rho.o  * (DD.o, DD.oo, DD.ooo)
        k.moo[0] = -r * cm00 * kb; k.moo[1] = r * cm00 * kb0; k.moo[2] = -r * cm00 * kb00;   // rho.oo * (DD.oo, DD.ooo, DD.oooo)
        k.c[0] = -r * cc * kb; k.c[1] = r * cc * kb0; k.c[2] = -r * cc * kb00;               // rho    * (D.o, D.oo, D.ooo)
        k.co[0] = r * cc0 * kb; k.co[1] = -r * cc0 * kb0; k.co[2] = r * cc0 * kb00;          // rho.o  * (D.oo, D.ooo, D.oooo)
        k.coo[0] = -r * cc00 * kb; k.coo[1] = r * cc00 * kb0; k.coo[2] = -r * cc00 * kb00;   // rho.oo * (D.ooo, D.oooo, D.ooooo)
    }
    return k;
}
namespace {
struct UlPtrs { const double *rho, *rhoO, *rhoOO, *DDo, *DDoo, *DDooo, *DDoooo, *Do, *Doo, *Dooo, *Doooo, *Dooooo; };
__global__ void k_d2dt2_hist_ul(UlPtrs p, UlCoeffs k, double* __restrict__ hist, int n, int ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r = p.rho[i], ro = p.rhoO[i], roo = p.rhoOO[i];
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const size_t j = (size_t)q * ld + i;
        double h = r * k.g[q];
        if (k.mode != S4F_D2DT2_STEADY_STATE) {
            const double a1 = p.DDo[j], a2 = p.DDoo[j], d1 = p.Do[j], d2 = p.Doo[j], d3 = p.Dooo[j];
            h += r * (k.m[0] * a1 + k.m[1] * a2) + r * (k.c[0] * d1 + k.c[1] * d2 + k.c[2] * d3);
            if (k.mode == S4F_D2DT2_BACKWARD) {
                const double a3 = p.DDooo[j], a4 = p.DDoooo[j], d4 = p.Doooo[j], d5 = p.Dooooo[j];
                h += ro * (k.mo[0] * a1 + k.mo[1] * a2 + k.mo[2] * a3) + roo * (k.moo[0] * a2 + k.moo[1] * a3 + k.moo[2] * a4)
                   + ro * (k.co[0] * d2 + k.co[1] * d3 + k.co[2] * d4) + roo * (k.coo[0] * d3 + k.coo[1] * d4 + k.coo[2] * d5);
            }
        }
        hist[j] = h;
    }
}
}  // namespace

int s4f_d2dt2_history(s4fgpu_ctx* c) {
    if (c->UL()) {
        if (c->histValid) return 0;
        const UlCoeffs k = ul_coeffs(c);
        UlPtrs p{c->rhoF.p, c->rhoO.p, c->rhoOO.p, c->DDo.p, c->DDoo.p, c->DDooo.p, c->DDoooo.p, c->Dold.p, c->DoldOld.p, c->Dooo.p, c->Doooo.p, c->Dooooo.p};
        k_d2dt2_hist_ul<<<(c->N + 255) / 256, 256, 0, c->stream>>>(p, k, c->d2Hist.p, c->N, c->ld);
        c->launches++;
        c->histValid = true;
        return 0;
    }
    if (c->ctl.d2dt2Scheme == S4F_D2DT2_STEADY_STATE || c->histValid) return 0;
    const D2dt2Coeffs k = d2dt2_coeffs(c);
    const bool deep = c->ctl.d2dt2Scheme == S4F_D2DT2_BACKWARD && c->timeIndex > 1;   // before: D.ooo = D.oooo = copies of D.oo
    const long long n = 3 * (long long)c->ld;
    k_d2dt2_hist<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->Dold.p, c->DoldOld.p, deep ? c->Dooo.p : c->DoldOld.p,
                                                                     deep ? c->Doooo.p : c->DoldOld.p, c->d2Hist.p, n, k.h[0], k.h[1],
                                                                     k.h[2], k.h[3]);
    c->launches++;
    c->histValid = true;
    return 0;
}

int s4f_upload_bc(s4fgpu_ctx* c) {
    std::vector<int> kinds(std::max(c->B, 1), S4F_BC_SOLID_TRACTION);
    for (int p = 0; p < c->nPatches; p++)
        for (int i = 0; i < c->pSize[p]; i++) kinds[c->pStart[p] + i] = (c->pKind[p] == S4F_PATCH_PROCESSOR) ? S4F_BC_PROCESSOR : c->bcKind[p];
    S4F_CHECK_CUDA(c, c->bKind.upload(kinds));
    return 0;
}

int s4f_assemble_matrix(s4fgpu_ctx* c) {
    const double dcoef = c->UL() ? ul_coeffs(c).diag : d2dt2_coeffs(c).diag;
    const int gridM = s4f_grid(c->numSMs, (long long)c->nSlices * 32);
    k_assemble_laplacian<<<gridM, S4F_BLOCK, 0, c->stream>>>(c->slicePtr.p, c->col.p, c->eW.p, c->eDn.p, c->eSf.p, c->nonOrth ? c->eCorr.p : nullptr,
                                                             c->impK.p, c->V.p, c->eA.p, c->eGam.p, c->eU.p, c->eC0.p, c->nonOrth ? c->eVc.p : nullptr,
                                                             c->rowK.p, c->diag0.p, c->N, c->bOff(), c->ld, c->nEntries, c->nSlices,
                                                             c->ctl.stabScaleFactor, c->ctl.stabilisation == S4F_STAB_RHIE_CHOW ? 1 : 0, dcoef,
                                                             c->UL() ? c->rhoF.p : nullptr);
    k_diag_copy<<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->diag0.p, c->diagC.p, c->N, c->ld);
    c->launches += 2;
    if (c->nBCells > 0) {
        k_diag_boundary<<<(c->nBCells + 127) / 128, 128, 0, c->stream>>>(c->bcCells.p, c->bcPtr.p, c->bcFaces.p, c->bKind.p, c->bN.p, c->bDelta.p,
                                                                       c->bMagSf.p, c->impK.p, c->diagC.p, c->nBCells, c->B, c->bOff(), c->ld);
        c->launches++;
    }
    k_diag_recip<<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->diagC.p, c->rDiagC.p, c->N, c->ld);
    c->launches++;
    c->matrixValid = true;
    // new coefficients on the same graph: an existing GAMG hierarchy keeps its aggregates and re-sums its levels on the device
    c->amgRefresh = (c->amg != nullptr && !getenv("S4F_NO_AMG_REFRESH"));      // the variable: A/B aid of the tests
    c->amgValid = false; c->dicValid = false;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return 0;
}

int s4f_bc_update_coeffs(s4fgpu_ctx* c) {
    if (c->B == 0) return 0;
    if (c->unsModel()) return s4f_uns_bc_update(c);
    const int TL = c->finiteStrain() ? 1 : 0;
    k_bc_update<<<(c->B + 127) / 128, 128, 0, c->stream>>>(c->bKind.p, c->bN.p, c->bcValue.p, c->bcPressure.p, c->impK.p, c->sigma.p, c->gradD.p,
                                                          c->Finv.p, c->tracGrad.p, c->D.p, c->incremental() ? c->Dold.p : nullptr, c->B,
                                                          c->bOff(), c->ld, TL);
    c->launches++;
    return 0;
}

int s4f_bc_evaluate(s4fgpu_ctx* c) {
    if (c->B == 0) return 0;
    k_bc_evaluate<<<(c->B + 127) / 128, 128, 0, c->stream>>>(c->bFaceCell.p, c->bKind.p, c->bN.p, c->bK.p, c->bDelta.p, c->tracGrad.p, c->D.p,
                                                            c->gradD.p, c->B, c->bOff(), c->ld);
    c->launches++;
    return 0;
}

int s4f_make_m(s4fgpu_ctx* c) {
    k_make_m<<<s4f_grid(c->numSMs, c->N + c->B), S4F_BLOCK, 0, c->stream>>>(c->sigma.p, c->gradD.p, c->T9.p, c->N, c->bOff(), c->B, c->ld, c->gamma0());
    c->launches++;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    c->mValid = true;
    return s4f_halo_exchange(c, c->T9.p, 9);
}

int s4f_bc_sngrad_store(s4fgpu_ctx* c) {
    if (c->B == 0) return 0;
    k_bc_sngrad_store<<<(c->B + 127) / 128, 128, 0, c->stream>>>(c->bFaceCell.p, c->bKind.p, c->bN.p, c->bK.p, c->bDelta.p, c->tracGrad.p, c->D.p,
                                                                c->gradD.p, c->bSn.p, c->B, c->bOff(), c->ld);
    c->launches++;
    return 0;
}

int s4f_assemble_source(s4fgpu_ctx* c) {
    int rh = s4f_d2dt2_history(c); if (rh) return rh;
    if (c->unsModel()) {          // unsLinGeomSolid.C:124-131: no stabilisation term, the divergence of the face stress
        int rc = s4f_uns_source(c); if (rc) return rc;
    } else {
        if (!c->mValid) {      // M = T - gamma grad(D) was not left behind by a law kernel (uploaded / initial fields)
            int rc = (c->ctl.solidModel == S4F_MODEL_LIN_GEOM_TOTAL_DISP) ? s4f_make_m(c) : s4f_kinematics(c);
            if (rc) return rc;
        }
        const double rs = c->UL() ? 0.0 : c->law.rho;          // updated Lagrangian: rho_*g() of the density field is part of d2Hist
        const double* hist = (c->ctl.d2dt2Scheme == S4F_D2DT2_STEADY_STATE && !c->UL()) ? nullptr : c->d2Hist.p;
        if (!c->nonOrth)
            k_source_m<<<s4f_grid(c->numSMs, (long long)c->nSlices * 32, 6), S4F_BLOCK, 0, c->stream>>>(
                c->slicePtr.p, c->col.p, c->eU.p, c->eC0.p, c->rowK.p, c->D.p, c->T9.p, c->V.p, hist, c->source.p, c->N, c->ld, c->nEntries,
                c->nSlices, rs * c->ctl.g[0], rs * c->ctl.g[1], rs * c->ctl.g[2]);
        else {
            const char* ev = getenv("S4F_SRCG_MINB");      // tuning aid: 2 = 98 registers, no spill; 3 = 80 registers; 6 = 80 registers, finer grid
            const int minb = ev ? atoi(ev) : 2;
            if (minb == 2)
                k_source_g<2><<<s4f_grid(c->numSMs, (long long)c->nSlices * 32, 2), S4F_BLOCK, 0, c->stream>>>(
                    c->slicePtr.p, c->col.p, c->eU.p, c->eVc.p, c->eC0.p, c->rowK.p, c->D.p, c->T9.p, c->gradD.p, c->V.p, hist, c->source.p, c->N,
                    c->ld, c->nEntries, c->nSlices, rs * c->ctl.g[0], rs * c->ctl.g[1], rs * c->ctl.g[2]);
            else
                k_source_g<3><<<s4f_grid(c->numSMs, (long long)c->nSlices * 32, minb == 3 ? 3 : 6), S4F_BLOCK, 0, c->stream>>>(
                    c->slicePtr.p, c->col.p, c->eU.p, c->eVc.p, c->eC0.p, c->rowK.p, c->D.p, c->T9.p, c->gradD.p, c->V.p, hist, c->source.p, c->N,
                    c->ld, c->nEntries, c->nSlices, rs * c->ctl.g[0], rs * c->ctl.g[1], rs * c->ctl.g[2]);
        }
        c->launches++;
    }
    if (c->nBCells > 0) {
        k_source_boundary<<<(c->nBCells + 127) / 128, 128, 0, c->stream>>>(c->bcCells.p, c->bcPtr.p, c->bcFaces.p, c->bKind.p, c->bN.p, c->bK.p,
                                                                         c->bDelta.p, c->bMagSf.p, c->impK.p, c->tracGrad.p, c->D.p, c->gradD.p,
                                                                         c->source.p, c->nBCells, c->B, c->bOff(), c->ld);
        c->launches++;
    }
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return 0;
}

int s4f_grad(s4fgpu_ctx* c) {
    if (c->unsModel()) return s4f_uns_gradients(c);
    const int gridM = s4f_grid(c->numSMs, (long long)c->nSlices * 32, 3);
    if (c->B > 0) {
        k_bc_sngrad_store<<<(c->B + 127) / 128, 128, 0, c->stream>>>(c->bFaceCell.p, c->bKind.p, c->bN.p, c->bK.p, c->bDelta.p, c->tracGrad.p, c->D.p,
                                                                    c->gradD.p, c->bSn.p, c->B, c->bOff(), c->ld);
        c->launches++;
    }
    if (c->ctl.gradScheme == S4F_GRAD_GAUSS_LINEAR)
        k_grad<true><<<gridM, S4F_BLOCK, 0, c->stream>>>(c->slicePtr.p, c->col.p, c->eSf.p, c->eW.p, c->D.p, c->rV.p, c->gradD.p, c->N, c->ld, c->nEntries, c->nSlices);
    else {
        if (c->pointCellsGrad() && !c->gValid) { int rc = s4f_build_point_stencil(c); if (rc) return rc; }
        if (c->pointCellsGrad()) { int rc = s4f_point_ghost_exchange(c, c->D.p, 3); if (rc) return rc; }
        k_grad<false><<<gridM, S4F_BLOCK, 0, c->stream>>>(c->gradSlicePtr(), c->gradCol(), c->gradLs(), c->eW.p, c->D.p, c->rV.p, c->gradD.p, c->N, c->ld, c->gradNE(), c->nSlices);
    }
    c->launches++;
    if (c->B > 0) {
        k_grad_boundary<<<(c->B + 127) / 128, 128, 0, c->stream>>>(c->bFaceCell.p, c->bKind.p, c->bN.p, c->bSn.p, c->gradD.p, c->B, c->bOff(), c->ld);
        c->launches++;
    }
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return s4f_halo_exchange(c, c->gradD.p, 9);
}

// fvc::grad of a temporary with calculated patches (gradD() = fvc::grad(D().oldTime() + DD()), nonLinGeomUpdatedLagSolid.C:243):
// snGrad_b = deltaCoeffs (X_b - X_P), then the usual boundary correction
namespace {
__global__ void k_sngrad_calculated(const int* __restrict__ bFaceCell, const int* __restrict__ bKind, const double* __restrict__ bDelta,
                                    const double* __restrict__ X, double* __restrict__ bSn, int B, int bOff, int ld) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || bKind[b] == S4F_BC_PROCESSOR) return;
    const int P = bFaceCell[b];
#pragma unroll
    for (int c = 0; c < 3; c++) bSn[(size_t)c * B + b] = bDelta[b] * (X[(size_t)c * ld + bOff + b] - X[(size_t)c * ld + P]);
}
}  // namespace
int s4f_grad_calculated_interior(s4fgpu_ctx* c, const double* X, double* gradOut) {
    const int gridM = s4f_grid(c->numSMs, (long long)c->nSlices * 32, 3);
    if (c->ctl.gradScheme == S4F_GRAD_GAUSS_LINEAR)
        k_grad<true><<<gridM, S4F_BLOCK, 0, c->stream>>>(c->slicePtr.p, c->col.p, c->eSf.p, c->eW.p, X, c->rV.p, gradOut, c->N, c->ld, c->nEntries, c->nSlices);
    else {
        if (c->pointCellsGrad() && !c->gValid) { int rc = s4f_build_point_stencil(c); if (rc) return rc; }
        if (c->pointCellsGrad()) { int rc = s4f_point_ghost_exchange(c, const_cast<double*>(X), 3); if (rc) return rc; }
        k_grad<false><<<gridM, S4F_BLOCK, 0, c->stream>>>(c->gradSlicePtr(), c->gradCol(), c->gradLs(), c->eW.p, X, c->rV.p, gradOut, c->N, c->ld, c->gradNE(), c->nSlices);
    }
    c->launches++;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return 0;
}

int s4f_grad_calculated(s4fgpu_ctx* c, const double* X, double* gradOut) {
    const int gridM = s4f_grid(c->numSMs, (long long)c->nSlices * 32, 3);
    if (c->B > 0) {
        k_sngrad_calculated<<<(c->B + 127) / 128, 128, 0, c->stream>>>(c->bFaceCell.p, c->bKind.p, c->bDelta.p, X, c->bSn.p, c->B, c->bOff(), c->ld);
        c->launches++;
    }
    if (c->ctl.gradScheme == S4F_GRAD_GAUSS_LINEAR)
        k_grad<true><<<gridM, S4F_BLOCK, 0, c->stream>>>(c->slicePtr.p, c->col.p, c->eSf.p, c->eW.p, X, c->rV.p, gradOut, c->N, c->ld, c->nEntries, c->nSlices);
    else {
        if (c->pointCellsGrad() && !c->gValid) { int rc = s4f_build_point_stencil(c); if (rc) return rc; }
        if (c->pointCellsGrad()) { int rc = s4f_point_ghost_exchange(c, const_cast<double*>(X), 3); if (rc) return rc; }
        k_grad<false><<<gridM, S4F_BLOCK, 0, c->stream>>>(c->gradSlicePtr(), c->gradCol(), c->gradLs(), c->eW.p, X, c->rV.p, gradOut, c->N, c->ld, c->gradNE(), c->nSlices);
    }
    c->launches++;
    if (c->B > 0) {
        k_grad_boundary<<<(c->B + 127) / 128, 128, 0, c->stream>>>(c->bFaceCell.p, c->bKind.p, c->bN.p, c->bSn.p, gradOut, c->B, c->bOff(), c->ld);
        c->launches++;
    }
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return 0;
}

namespace {
__global__ void k_sum2(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ b, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}
}  // namespace

// incremental models: D = D.oldTime() + DD; gradD = gradD.oldTime() + gradDD  (nonLinGeomTotalLagSolid.C:190-196)
int s4f_update_totals(s4fgpu_ctx* c, bool disp, bool grad) {
    if (!c->incremental()) return 0;
    const long long ld = c->ld;
    if (disp) { k_sum2<<<(unsigned)((3 * ld + 255) / 256), 256, 0, c->stream>>>(c->Dtot.p, c->Dold.p, c->D.p, 3 * ld); c->launches++; }
    if (grad) { k_sum2<<<(unsigned)((9 * ld + 255) / 256), 256, 0, c->stream>>>(c->gradDtot.p, c->gradDold.p, c->gradD.p, 9 * ld); c->launches++; }
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return 0;
}

int s4f_relax_and_residual(s4fgpu_ctx* c, int iCorr) {
    const int grid = s4f_grid(c->numSMs, c->N + c->B);
    if (c->ctl.relaxationMethod == S4F_RELAX_AITKEN) {
        const size_t ld = c->ld;
        if (c->aitRes.n != 3 * ld) { S4F_CHECK_CUDA(c, c->aitRes.alloc(3 * ld)); S4F_CHECK_CUDA(c, c->aitResPrev.alloc(3 * ld)); S4F_CHECK_CUDA(c, c->aitAlpha.alloc(ld)); }
        k_relax_aitken<<<grid, S4F_BLOCK, 0, c->stream>>>(c->D.p, c->Dprev.p, c->Dold.p, c->aitRes.p, c->aitResPrev.p, c->aitAlpha.p, c->N, c->bOff(), c->B,
                                                         c->ld, iCorr == 0, c->ctl.fieldRelaxD, c->outS.p, c->red());
    } else {
        k_relax_residual<<<grid, S4F_BLOCK, 0, c->stream>>>(c->D.p, c->Dprev.p, c->Dold.p, c->N, c->bOff(), c->B, c->ld, c->ctl.fieldRelaxD,
                                                           c->outS.p, c->red());
    }
    c->launches++;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return s4f_halo_exchange(c, c->D.p, 3);
}

// ---- timing of the face-loop kernels (bench.py roofline) ----------------------------------------
int s4f_time_fv_kernels(s4fgpu_ctx* c, int kernel, int reps, int flushL2, double* msOut, double* bytesOut) {
    if ((kernel == S4F_KERNEL_GAMG_VCYCLE || kernel == S4F_KERNEL_GAMG_STEP0) && !c->amgValid) {
        int rc = c->amgRefresh ? s4f_amg_refresh(c) : s4f_amg_setup(c); if (rc) return rc;
        c->amgValid = true; c->amgRefresh = false;
    }
    if (flushL2 && c->flushBuf.n < (size_t)48 * 1024 * 1024) S4F_CHECK_CUDA(c, c->flushBuf.alloc((size_t)48 * 1024 * 1024));
    cudaEvent_t e0, e1;
    S4F_CHECK_CUDA(c, cudaEventCreate(&e0)); S4F_CHECK_CUDA(c, cudaEventCreate(&e1));
    double total = 0;
    const int warm = getenv("S4F_TIME_WARMUP") ? atoi(getenv("S4F_TIME_WARMUP")) : 3;      // ncu runs set 0: one launch per kernel
    for (int r = -warm; r < reps; r++) {
        if (flushL2) S4F_CHECK_CUDA(c, cudaMemsetAsync(c->flushBuf.p, 0, c->flushBuf.n * sizeof(double), c->stream));
        S4F_CHECK_CUDA(c, cudaEventRecord(e0, c->stream));
        int rc = 0;
        if (kernel == S4F_KERNEL_GRAD) rc = s4f_grad(c);
        else if (kernel == S4F_KERNEL_LAW) rc = s4f_law_correct(c);
        else if (kernel == S4F_KERNEL_GAMG_VCYCLE) rc = s4f_amg_apply(c, c->rA.p, c->wA.p);
        else if (kernel == S4F_KERNEL_GAMG_STEP0) rc = s4f_amg_step0(c, c->rA.p, nullptr);
        else rc = s4f_assemble_source(c);
        if (rc) return rc;
        S4F_CHECK_CUDA(c, cudaEventRecord(e1, c->stream));
        S4F_CHECK_CUDA(c, cudaEventSynchronize(e1));
        float ms; S4F_CHECK_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 0) total += ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *msOut = total / reps;
    const double N = c->N, nnz = (double)c->nnzOff + (c->B - c->G);   // row entries incl. boundary faces
    if (kernel == S4F_KERNEL_GRAD) *bytesOut = 24 * N + nnz * (4 + 24) + 72 * N + 0.125 * N;                 // D, (col, ls), gradD out
    else if (kernel == S4F_KERNEL_LAW) *bytesOut = s4f_law_bytes(c);
    else if (kernel == S4F_KERNEL_GAMG_STEP0) { int rc = s4f_amg_step0(c, c->rA.p, bytesOut); if (rc) return rc; }
    else if (kernel == S4F_KERNEL_GAMG_VCYCLE) { int nl, sz[16]; double st; int rc = s4f_amg_info(c, &nl, sz, 16, bytesOut, &st); if (rc) return rc; }
    else if (!c->nonOrth) *bytesOut = (24 + 72) * N + nnz * (4 + 24 + 8) + (24 + 8 + 24 + 0.125) * N;          // D, M | col,u,c0 | U, V, out
    else *bytesOut = (24 + 72 + 72) * N + nnz * (4 + 24 + 24 + 8) + (48 + 8 + 24 + 0.125) * N;                  // D, M, gradD | col,u,vc,c0 | U,Vc, V, out
    return 0;
}
