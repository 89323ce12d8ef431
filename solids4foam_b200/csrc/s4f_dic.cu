// s4f_dic.cu -- [OF-ext] DICPreconditioner / FDICPreconditioner (the preconditioner of 93 of the reference's tutorials) on the
// device, EXACTLY: the sequential face sweeps
//     rD = diag;  for f: rD[u_f] -= upper_f^2 / rD[l_f];  rD = 1/rD
//     w = rD r;   for f ascending:  w[u_f] -= rD[u_f] upper_f w[l_f];   for f descending:  w[l_f] -= rD[l_f] upper_f w[u_f]
// are triangular recurrences over the cells (faces are ordered by owner, owner < neighbour), so they are evaluated level by
// level: level(i) = 1 + max level of the lower neighbours; all cells of a level are independent.  The results equal the
// sequential sweeps up to the summation order inside a cell, so PCG takes the iteration counts of the reference's own solver
// (parity tests).  A hex mesh of nx x ny x nz cells has nx+ny+nz-2 levels: about 1000 small launches per sweep at 8 M cells --
// this is the parity preconditioner; GAMG (s4f_amg.cu) is the fast one.  Across ranks it is block Jacobi, as in OpenFOAM.
#include <algorithm>

#include "s4f_ctx.h"
#include "s4f_dev.cuh"

struct S4fDic {
    std::vector<int> lvlPtr;             // host: cells of level L are cells[lvlPtr[L] .. lvlPtr[L+1])
    DevBuf<int> cells;
    DevBuf<double> rD;                   // 3*ld: reciprocal D of the incomplete factorisation, per component
};

namespace {

// MODE 0: raw[i] = diag[i] - sum_lower a^2 / raw[l];  MODE 1: z[i] = rD[i] (r[i] + sum_lower a z[l]);
// MODE 2: z[i] += rD[i] sum_upper a z[u]          (a = -upper > 0 is the stored coefficient)
template <int MODE>
__global__ void k_dic_level(const int* __restrict__ cells, int first, int count, const int* __restrict__ slicePtr, const int* __restrict__ col,
                            const double* __restrict__ eA, const double* __restrict__ diagC, double* __restrict__ rD,
                            const double* __restrict__ r, double* __restrict__ z, int N, int ld) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int i = cells[first + t];
    const int s = i >> 5, lane = i & 31;
    const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5;
    double acc[3] = {0, 0, 0};
    for (int k = 0; k < width; k++) {
        const long long e = (long long)base + 32 * k + lane;
        const int c = col[e];
        const bool take = (MODE == 2) ? (c > i && c < N) : (c < i);
        if (!take) continue;
        const double a = eA[e];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            if (MODE == 0) acc[q] += a * a / rD[(size_t)q * ld + c];
            else acc[q] += a * z[(size_t)q * ld + c];
        }
    }
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const size_t j = (size_t)q * ld + i;
        if (MODE == 0) rD[j] = diagC[j] - acc[q];
        else if (MODE == 1) z[j] = rD[j] * (r[j] + acc[q]);
        else z[j] += rD[j] * acc[q];
    }
}

__global__ void k_dic_recip(double* __restrict__ rD, int N, int ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
#pragma unroll
    for (int q = 0; q < 3; q++) rD[(size_t)q * ld + i] = 1.0 / rD[(size_t)q * ld + i];
}

}  // namespace

void s4f_dic_destroy(s4fgpu_ctx* c) { delete c->dic; c->dic = nullptr; }

int s4f_dic_setup(s4fgpu_ctx* c) {
    if (!c->dic) c->dic = new S4fDic();
    S4fDic& d = *c->dic;
    const int N = c->N, F = c->F;
    std::vector<int> level(N, 0);
    int nL = 1;
    for (int f = 0; f < F; f++) {            // faces ordered by owner: level[owner] is final when its faces come up
        const int l = level[c->own[f]] + 1;
        if (l > level[c->nei[f]]) { level[c->nei[f]] = l; nL = std::max(nL, l + 1); }
    }
    d.lvlPtr.assign(nL + 1, 0);
    for (int i = 0; i < N; i++) d.lvlPtr[level[i] + 1]++;
    for (int L = 0; L < nL; L++) d.lvlPtr[L + 1] += d.lvlPtr[L];
    std::vector<int> cells(N), cur(d.lvlPtr.begin(), d.lvlPtr.end() - 1);
    for (int i = 0; i < N; i++) cells[cur[level[i]]++] = i;
    S4F_CHECK_CUDA(c, d.cells.upload(cells));
    if (d.rD.n != 3 * (size_t)c->ld) S4F_CHECK_CUDA(c, d.rD.alloc(3 * (size_t)c->ld));
    for (int L = 0; L < nL; L++) {
        const int n = d.lvlPtr[L + 1] - d.lvlPtr[L];
        k_dic_level<0><<<(n + 127) / 128, 128, 0, c->stream>>>(d.cells.p, d.lvlPtr[L], n, c->slicePtr.p, c->col.p, c->eA.p, c->diagC.p, d.rD.p,
                                                             nullptr, nullptr, N, c->ld);
    }
    k_dic_recip<<<(N + 255) / 256, 256, 0, c->stream>>>(d.rD.p, N, c->ld);
    c->launches += nL + 1;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    c->dicValid = true;
    return 0;
}

int s4f_dic_apply(s4fgpu_ctx* c, const double* r3, double* z3) {
    if (!c->dicValid) { int rc = s4f_dic_setup(c); if (rc) return rc; }
    S4fDic& d = *c->dic;
    const int nL = (int)d.lvlPtr.size() - 1, N = c->N;
    for (int L = 0; L < nL; L++) {
        const int n = d.lvlPtr[L + 1] - d.lvlPtr[L];
        k_dic_level<1><<<(n + 127) / 128, 128, 0, c->stream>>>(d.cells.p, d.lvlPtr[L], n, c->slicePtr.p, c->col.p, c->eA.p, c->diagC.p, d.rD.p, r3, z3,
                                                             N, c->ld);
    }
    for (int L = nL - 1; L >= 0; L--) {
        const int n = d.lvlPtr[L + 1] - d.lvlPtr[L];
        k_dic_level<2><<<(n + 127) / 128, 128, 0, c->stream>>>(d.cells.p, d.lvlPtr[L], n, c->slicePtr.p, c->col.p, c->eA.p, c->diagC.p, d.rD.p, r3, z3,
                                                             N, c->ld);
    }
    c->launches += 2 * nL;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return 0;
}
