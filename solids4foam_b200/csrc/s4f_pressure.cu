// s4f_pressure.cu -- mechanicalLaw::updateSigmaHyd with solvePressureEqn (ML/mechanicalLaw/mechanicalLaw.C:1366-1476):
// the hydrostatic stress of the law is smoothed by a pressure equation assembled and solved on the device,
//     fvm::Sp(1, p) - fvm::laplacian(rDAf, p) == pExplicit - fvc::div(rDAf*(interpolate(grad p) & Sf)),        :1446-1453
//     rDAf = pressureSmoothingScaleFactor * interpolate(impK/DEqnA),   DEqnA = DEqn.A()                        :1432-1440
// over the same SELL-32 rows as the momentum matrix (coefficients eP = rDAf magSf delta); p has zeroGradient patches
// (:452-458), so boundary faces carry no coefficient.  The scalar system rides the fused three-component PCG with
// components 1 and 2 idle.  Afterwards grad p = fvc::grad(p) (:1467) and sigma gets (p - pExplicit)/J on its diagonal
// (linearElastic.C:337-340, neoHookeanElastic.C:295-302, neoHookeanElasticMisesPlastic.C:1215-1222).
// The CPU restatement in LDU form is oracle/s4f_oracle.cpp: updateSigmaHydSmoothed.
#include <utility>

#include "s4f_ctx.h"
#include "s4f_dev.cuh"

namespace {

// r = impK/AD, AD = (diag + component average of the boundary diagonal)/V  ([OF-ext] fvMatrix::A()); patch value of AD = cell value
__global__ void k_p_ratio(const double* __restrict__ impK, const double* __restrict__ diagC, const double* __restrict__ V,
                          const int* __restrict__ bFaceCell, double* __restrict__ r, int N, int bOff, int B, int ld) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N + B) return;
    const int i = (t < N) ? t : bOff + (t - N);
    const int P = (t < N) ? t : bFaceCell[t - N];
    const double AD = (diagC[P] + diagC[(size_t)ld + P] + diagC[2 * (size_t)ld + P]) / 3.0 / V[P];
    r[i] = impK[i] / AD;
}

// coefficients, diagonal and source of the pressure equation, one row per lane
__global__ void __launch_bounds__(S4F_BLOCK) k_p_assemble(const int* __restrict__ slicePtr, const int* __restrict__ col, const double* __restrict__ eW,
                                                         const double* __restrict__ eDn, const double* __restrict__ eSf,
                                                         const double* __restrict__ eCorr /* null when orthogonal */, const double* __restrict__ r,
                                                         const double* __restrict__ gradP, const double* __restrict__ pExp, const double* __restrict__ V,
                                                         double* __restrict__ eP, double* __restrict__ pDiag, double* __restrict__ pRDiag,
                                                         double* __restrict__ pB, int N, int bOff, int ld, long long nE, int nSlices, double scale) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nSlices; s += nWarps) {
        const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5;
        const int row = s * 32 + lane;
        const int rr = row < N ? row : 0;
        const double rP = r[rr];
        const double gP[3] = {gradP[rr], gradP[(size_t)ld + rr], gradP[2 * (size_t)ld + rr]};
        double sum = 0, src = 0;
        for (int k = 0; k < width; k++) {
            const long long e = (long long)base + 32 * k + lane;
            const int cc = col[e];
            const double w = eW[e], w1 = 1.0 - w;
            const double S[3] = {eSf[e], eSf[nE + e], eSf[2 * nE + e]};
            const double gN[3] = {gradP[cc], gradP[(size_t)ld + cc], gradP[2 * (size_t)ld + cc]};
            double a = 0;
            if (cc < bOff) {          // internal and processor faces
                const double gam = scale * (w * rP + w1 * r[cc]);
                a = gam * eDn[e];
                const double gf[3] = {w * gP[0] + w1 * gN[0], w * gP[1] + w1 * gN[1], w * gP[2] + w1 * gN[2]};
                double t = -(S[0] * gf[0] + S[1] * gf[1] + S[2] * gf[2]);
                if (eCorr) t += eCorr[e] * gf[0] + eCorr[nE + e] * gf[1] + eCorr[2 * nE + e] * gf[2];
                src += gam * t;
            } else {                  // boundary face: only the explicit div term, with the patch values
                src -= scale * r[cc] * (S[0] * gN[0] + S[1] * gN[1] + S[2] * gN[2]);
            }
            eP[e] = a;
            sum += a;
        }
        if (row < N) {
            const double v = V[row], d = v + sum;
#pragma unroll
            for (int q = 0; q < 3; q++) { pDiag[(size_t)q * ld + row] = d; pRDiag[(size_t)q * ld + row] = 1.0 / d; }
            pB[row] = src + v * pExp[row];
        }
    }
}

__global__ void k_p_boundary(const int* __restrict__ bFaceCell, const int* __restrict__ bKind, double* __restrict__ p, int B, int bOff) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || bKind[b] == S4F_BC_PROCESSOR) return;
    p[bOff + b] = p[bFaceCell[b]];                       // zeroGradient
}

// fvc::grad(p): least squares / Gauss linear (the gather of k_grad for one scalar)
template <bool GAUSS>
__global__ void __launch_bounds__(S4F_BLOCK) k_grad_scalar(const int* __restrict__ slicePtr, const int* __restrict__ col, const double* __restrict__ eVec,
                                                          const double* __restrict__ eW, const double* __restrict__ p, const double* __restrict__ rV,
                                                          double* __restrict__ gradP, int N, int ld, long long nE, int nSlices) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nSlices; s += nWarps) {
        const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5;
        const int row = s * 32 + lane;
        const double pP = p[row < N ? row : 0];
        double g[3] = {0, 0, 0};
        for (int k = 0; k < width; k++) {
            const long long e = (long long)base + 32 * k + lane;
            const double pn = p[col[e]];
            const double d = GAUSS ? (eW[e] * pP + (1.0 - eW[e]) * pn) : (pn - pP);
            g[0] += eVec[e] * d; g[1] += eVec[nE + e] * d; g[2] += eVec[2 * nE + e] * d;
        }
        if (row < N) {
            const double sc = GAUSS ? rV[row] : 1.0;
#pragma unroll
            for (int q = 0; q < 3; q++) gradP[(size_t)q * ld + row] = g[q] * sc;
        }
    }
}

// gaussGrad::correctBoundaryConditions with snGrad = 0: grad_b = grad_P - n (n & grad_P)
__global__ void k_grad_scalar_boundary(const int* __restrict__ bFaceCell, const int* __restrict__ bKind, const double* __restrict__ bN,
                                       double* __restrict__ gradP, int B, int bOff, int ld) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || bKind[b] == S4F_BC_PROCESSOR) return;
    const int P = bFaceCell[b];
    const double n[3] = {bN[b], bN[(size_t)B + b], bN[2 * (size_t)B + b]};
    const double g[3] = {gradP[P], gradP[(size_t)ld + P], gradP[2 * (size_t)ld + P]};
    const double ng = n[0] * g[0] + n[1] * g[1] + n[2] * g[2];
#pragma unroll
    for (int q = 0; q < 3; q++) gradP[(size_t)q * ld + bOff + b] = g[q] - n[q] * ng;
}

// sigma += (p - pExplicit)/J I   (J = 1 for the small-strain laws)
__global__ void k_sigma_hyd_fix(double* __restrict__ sigma, const double* __restrict__ p, const double* __restrict__ pExp,
                                const double* __restrict__ J /* null: 1 */, int N, int bOff, int B, int ld) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N + B) return;
    const int i = (t < N) ? t : bOff + (t - N);
    const double d = (p[i] - pExp[i]) / (J ? J[i] : 1.0);
    sigma[i] += d; sigma[(size_t)3 * ld + i] += d; sigma[(size_t)5 * ld + i] += d;
}

__global__ void k_p_relax(double* __restrict__ p, const double* __restrict__ x, double alpha, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) p[i] = (alpha == 1.0) ? x[i] : p[i] + alpha * (x[i] - p[i]);
}

}  // namespace

int s4f_pressure_smooth(s4fgpu_ctx* c) {
    const int N = c->N, B = c->B, bOff = c->bOff(), ld = c->ld;
    const int gridR = s4f_grid(c->numSMs, (long long)c->nSlices * 32, 4);
    if (!c->matrixValid) { int rc = s4f_assemble_matrix(c); if (rc) return rc; }     // DEqnA needs the momentum diagonal
    k_p_ratio<<<(N + B + 255) / 256, 256, 0, c->stream>>>(c->impK.p, c->diagC.p, c->V.p, c->bFaceCell.p, c->pRatio.p, N, bOff, B, ld);
    c->launches++;
    int rc = s4f_halo_exchange(c, c->pRatio.p, 1); if (rc) return rc;
    S4F_CHECK_CUDA(c, cudaMemsetAsync(c->pB.p, 0, 3 * (size_t)ld * sizeof(double), c->stream));
    k_p_assemble<<<gridR, S4F_BLOCK, 0, c->stream>>>(c->slicePtr.p, c->col.p, c->eW.p, c->eDn.p, c->eSf.p, c->nonOrth ? c->eCorr.p : nullptr, c->pRatio.p,
                                                    c->gradP.p, c->pExp.p, c->V.p, c->eP.p, c->pDiag.p, c->pRDiag.p, c->pB.p, N, bOff, ld, c->nEntries,
                                                    c->nSlices, c->law.pressureSmoothingScaleFactor);
    c->launches++;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    // initial guess: the previous sigmaHyd (the field persists between calls)
    S4F_CHECK_CUDA(c, cudaMemsetAsync(c->pX.p, 0, 3 * (size_t)ld * sizeof(double), c->stream));
    S4F_CHECK_CUDA(c, cudaMemcpyAsync(c->pX.p, c->sigmaHyd.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    {   // sigmaHydEqn.solve(): the fused solver on the pressure matrix (Jacobi-preconditioned PCG; the momentum GAMG hierarchy
        // belongs to another matrix)
        if ((rc = s4f_finish_solve(c))) return rc;      // the momentum solve's statistics are read before they are set aside
        const s4fgpu_stats keep = c->last; const long long keepInner = c->totalInner;
        const int pre = c->ctl.preconditioner, sol = c->ctl.solver;
        std::swap(c->eA.p, c->eP.p); std::swap(c->diagC.p, c->pDiag.p); std::swap(c->rDiagC.p, c->pRDiag.p);
        c->ctl.preconditioner = S4F_PRECOND_DIAGONAL; c->ctl.solver = S4F_SOLVER_PCG;
        // fvSolution "solvers sigmaHyd": its own tolerances when the case gives them
        const double keepTol = c->ctl.tolerance, keepRel = c->ctl.relTol; const int keepMax = c->ctl.maxIter;
        if (c->law.sigmaHydTolerance > 0) { c->ctl.tolerance = c->law.sigmaHydTolerance; c->ctl.relTol = c->law.sigmaHydRelTol; if (c->law.sigmaHydMaxIter > 0) c->ctl.maxIter = c->law.sigmaHydMaxIter; }
        // the scalar equation rides component 0 whatever the mesh's empty directions are (a 2-D case whose empty direction
        // is x would otherwise skip it)
        const int keepD[3] = {c->solD[0], c->solD[1], c->solD[2]};
        c->solD[0] = 1; c->solD[1] = 0; c->solD[2] = 0;
        rc = s4f_solve_segregated(c, c->pX.p, c->pB.p);
        for (int q = 0; q < 3; q++) c->solD[q] = keepD[q];
        c->ctl.preconditioner = pre; c->ctl.solver = sol;
        c->ctl.tolerance = keepTol; c->ctl.relTol = keepRel; c->ctl.maxIter = keepMax;
        std::swap(c->eA.p, c->eP.p); std::swap(c->diagC.p, c->pDiag.p); std::swap(c->rDiagC.p, c->pRDiag.p);
        c->lastP = c->last; c->last = keep; c->totalInner = keepInner;
        if (rc) return rc;
    }
    {   // sigmaHyd.relax(): sigmaHyd = sigmaHyd.prevIter + alpha (solution - sigmaHyd.prevIter); alpha = 1 is a copy
        const double al = (c->law.sigmaHydRelax > 0) ? c->law.sigmaHydRelax : 1.0;
        k_p_relax<<<(N + 255) / 256, 256, 0, c->stream>>>(c->sigmaHyd.p, c->pX.p, al, N);
        c->launches++;
    }
    if (B > 0) { k_p_boundary<<<(B + 127) / 128, 128, 0, c->stream>>>(c->bFaceCell.p, c->bKind.p, c->sigmaHyd.p, B, bOff); c->launches++; }
    if ((rc = s4f_halo_exchange(c, c->sigmaHyd.p, 1))) return rc;
    if (c->ctl.gradScheme == S4F_GRAD_GAUSS_LINEAR)
        k_grad_scalar<true><<<gridR, S4F_BLOCK, 0, c->stream>>>(c->slicePtr.p, c->col.p, c->eSf.p, c->eW.p, c->sigmaHyd.p, c->rV.p, c->gradP.p, N, ld, c->nEntries, c->nSlices);
    else {
        if (c->pointCellsGrad() && !c->gValid) { int rcg = s4f_build_point_stencil(c); if (rcg) return rcg; }
        if (c->pointCellsGrad() && (rc = s4f_point_ghost_exchange(c, c->sigmaHyd.p, 1))) return rc;
        k_grad_scalar<false><<<gridR, S4F_BLOCK, 0, c->stream>>>(c->gradSlicePtr(), c->gradCol(), c->gradLs(), c->eW.p, c->sigmaHyd.p, c->rV.p, c->gradP.p, N, ld, c->gradNE(), c->nSlices);
    }
    c->launches++;
    if (B > 0) { k_grad_scalar_boundary<<<(B + 127) / 128, 128, 0, c->stream>>>(c->bFaceCell.p, c->bKind.p, c->bN.p, c->gradP.p, B, bOff, ld); c->launches++; }
    if ((rc = s4f_halo_exchange(c, c->gradP.p, 3))) return rc;
    const bool lin = (c->law.kind == S4F_LAW_LINEAR_ELASTIC || c->law.kind == S4F_LAW_LINEAR_ELASTIC_MISES_PLASTIC);
    k_sigma_hyd_fix<<<(N + B + 255) / 256, 256, 0, c->stream>>>(c->sigma.p, c->sigmaHyd.p, c->pExp.p, lin ? nullptr : c->lawJ.p, N, bOff, B, ld);
    c->launches++;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return 0;
}
