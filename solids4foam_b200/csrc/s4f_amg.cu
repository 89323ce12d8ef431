// s4f_amg.cu -- GAMG-type preconditioner of the displacement PCG solve.
//
// Reference behaviour: fvSolution "solver PCG; preconditioner GAMG" / "solver GAMG" ([OF-ext] GAMGSolver,
// GAMGPreconditioner, pairGAMGAgglomeration "faceAreaPair"): cells are agglomerated pair-wise by their
// strongest face coefficient, coarse matrices are the sums of the fine coefficients (Galerkin product
// with piece-wise constant restriction/prolongation), a V-cycle of smoothing sweeps is applied.  This is
// the B200 restatement of that algorithm, not OpenFOAM's code:
//   * set-up (once per matrix, host): three pair-wise passes per level -> aggregates of <= 8 cells, so
//     the level sizes fall 8x and the whole hierarchy adds only ~1/7 of the fine-level traffic;
//   * every level is a SELL-32 cell-centric row store like the fine level (atomic-free gathers), the
//     three displacement components share each matrix read (they differ only in the diagonal);
//   * smoother: Chebyshev-Jacobi polynomial (symmetric, so the V-cycle is a valid PCG preconditioner,
//     and a pure SpMV chain -- no sequential sweep as in GaussSeidel/DIC);
//   * coarse-grid correction scaled by a fixed factor (OpenFOAM's scaleCorrection computes it from two
//     global dot products per level; a constant keeps the preconditioner linear and costs nothing);
//   * coarsest level (<= 512 cells): dense inverse applied by one small kernel;
//   * optional fp32 V-cycle: the preconditioner only has to be an SPD approximation, PCG stays fp64.
// Multi-rank: the hierarchy is rank-local (couplings across processor patches are dropped in the
// preconditioner only, i.e. block-Jacobi over ranks, like DIC in OpenFOAM); Amul in PCG is exact.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "s4f_ctx.h"
#include "s4f_dev.cuh"

namespace {

// ================================================================================================
// host: agglomeration
// ================================================================================================
struct HostLevel {
    int n = 0;
    std::vector<int> own, nei;          // faces, upper-triangular order
    std::vector<double> a;              // positive face coefficient (= -upper)
    std::vector<double> diag[3];
    std::vector<int> parent;            // aggregate of each cell in the next (coarser) level
};

// one pair-wise pass: greedy matching of every still-unmatched cell with its strongest unmatched
// neighbour ([OF-ext] pairGAMGAgglomeration::agglomerate, restated); returns the number of aggregates
int pairwise_pass(const HostLevel& L, std::vector<int>& agg) {
    const int n = L.n;
    const size_t F = L.own.size();
    std::vector<int> ptr(n + 1, 0);
    for (size_t f = 0; f < F; f++) { ptr[L.own[f] + 1]++; ptr[L.nei[f] + 1]++; }
    for (int i = 0; i < n; i++) ptr[i + 1] += ptr[i];
    std::vector<int> adj(2 * F);
    std::vector<float> wgt(2 * F);
    {
        std::vector<int> cur(ptr.begin(), ptr.end() - 1);
        for (size_t f = 0; f < F; f++) {   // lower neighbours first, then upper: the fine-level row order
            int e = cur[L.nei[f]]++; adj[e] = L.own[f]; wgt[e] = (float)L.a[f];
        }
        for (size_t f = 0; f < F; f++) {
            int e = cur[L.own[f]]++; adj[e] = L.nei[f]; wgt[e] = (float)L.a[f];
        }
    }
    agg.assign(n, -1);
    int nc = 0;
    for (int i = 0; i < n; i++) {
        if (agg[i] >= 0) continue;
        int best = -1; float bw = 0.f;
        for (int e = ptr[i]; e < ptr[i + 1]; e++) {
            const int j = adj[e];
            if (agg[j] < 0 && j != i && wgt[e] > bw * 1.0000001f) { bw = wgt[e]; best = j; }
        }
        agg[i] = nc;
        if (best >= 0) agg[best] = nc;
        nc++;
    }
    return nc;
}

// Galerkin coarse level for piece-wise constant transfer: coarse face = sum of the fine faces between
// two aggregates, coarse diagonal = sum of fine diagonals - 2 * (faces inside the aggregate)
void galerkin(const HostLevel& L, const std::vector<int>& agg, int nc, HostLevel& C) {
    C.n = nc;
    for (int q = 0; q < 3; q++) C.diag[q].assign(nc, 0.0);
    for (int i = 0; i < L.n; i++) for (int q = 0; q < 3; q++) C.diag[q][agg[i]] += L.diag[q][i];
    const size_t F = L.own.size();
    std::vector<int> cnt(nc + 1, 0);
    for (size_t f = 0; f < F; f++) {
        const int a = agg[L.own[f]], b = agg[L.nei[f]];
        if (a != b) cnt[std::min(a, b) + 1]++;
    }
    for (int i = 0; i < nc; i++) cnt[i + 1] += cnt[i];
    std::vector<int> hi(cnt[nc]);
    std::vector<double> w(cnt[nc]);
    {
        std::vector<int> cur(cnt.begin(), cnt.end() - 1);
        for (size_t f = 0; f < F; f++) {
            const int a = agg[L.own[f]], b = agg[L.nei[f]];
            if (a == b) { for (int q = 0; q < 3; q++) C.diag[q][a] -= 2.0 * L.a[f]; continue; }
            const int e = cur[std::min(a, b)]++;
            hi[e] = std::max(a, b); w[e] = L.a[f];
        }
    }
    C.own.clear(); C.nei.clear(); C.a.clear();
    C.own.reserve(cnt[nc] / 2 + 16); C.nei.reserve(cnt[nc] / 2 + 16); C.a.reserve(cnt[nc] / 2 + 16);
    std::vector<std::pair<int, double>> row;
    for (int i = 0; i < nc; i++) {
        row.clear();
        for (int e = cnt[i]; e < cnt[i + 1]; e++) row.emplace_back(hi[e], w[e]);
        std::sort(row.begin(), row.end(), [](const std::pair<int, double>& x, const std::pair<int, double>& y) { return x.first < y.first; });
        for (size_t k = 0; k < row.size();) {
            size_t m = k; double s = 0;
            while (m < row.size() && row[m].first == row[k].first) s += row[m++].second;
            C.own.push_back(i); C.nei.push_back(row[k].first); C.a.push_back(s);
            k = m;
        }
    }
}

// dense inverse of the coarsest matrix (SPD) by Cholesky, per component
bool dense_inverse(const HostLevel& L, int q, std::vector<double>& inv) {
    const int n = L.n;
    std::vector<double> A((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) A[(size_t)i * n + i] = L.diag[q][i];
    for (size_t f = 0; f < L.own.size(); f++) {
        A[(size_t)L.own[f] * n + L.nei[f]] -= L.a[f];
        A[(size_t)L.nei[f] * n + L.own[f]] -= L.a[f];
    }
    // a pure-Neumann component (no fixed patch) has a singular matrix: regularise like a tiny spring
    double dmax = 0; for (int i = 0; i < n; i++) dmax = std::max(dmax, A[(size_t)i * n + i]);
    for (int i = 0; i < n; i++) A[(size_t)i * n + i] += 1e-10 * dmax;
    // Cholesky A = L L^T in place (lower)
    for (int j = 0; j < n; j++) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(d > 0)) return false;
        d = std::sqrt(d); A[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[(size_t)i * n + j];
            for (int k = 0; k < j; k++) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
            A[(size_t)i * n + j] = s / d;
        }
    }
    inv.assign((size_t)n * n, 0.0);
    std::vector<double> y(n);
    for (int c = 0; c < n; c++) {        // solve L L^T x = e_c
        for (int i = 0; i < n; i++) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) s -= A[(size_t)i * n + k] * y[k];
            y[i] = s / A[(size_t)i * n + i];
        }
        for (int i = n - 1; i >= 0; i--) {
            double s = y[i];
            for (int k = i + 1; k < n; k++) s -= A[(size_t)k * n + i] * inv[(size_t)k * n + c];
            inv[(size_t)i * n + c] = s / A[(size_t)i * n + i];
        }
    }
    return true;
}

// ================================================================================================
// device kernels (T = float | double for the V-cycle arithmetic; TB = type of the right-hand side
// seen by a level: the fp64 PCG residual on level 0, T below; TO = type of the smoother output)
// ================================================================================================
struct Cheb { double c1, c2; };

// x = d = (1/theta) rD b      (first smoothing step from a zero initial guess)
template <class T, class TB>
__global__ void __launch_bounds__(S4F_BLOCK) k_amg_first(const T* __restrict__ rD, const TB* __restrict__ b, T* __restrict__ x,
                                                         T* __restrict__ d, int n, int ld, int ldb, T invTheta, const int* __restrict__ act) {
    const int a[3] = {act[0], act[1], act[2]};      // components whose PCG solve is still running (device-side flags)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int q = 0; q < 3; q++) {
            if (!a[q]) continue;
            const T v = invTheta * rD[(size_t)q * ld + i] * (T)b[(size_t)q * ldb + i];
            x[(size_t)q * ld + i] = v; d[(size_t)q * ld + i] = v;
        }
    }
}

// one Chebyshev-Jacobi step:  r = b - A x;  d' = c1 d + c2 rD r;  x' = x + d'   (out of place in x)
// MODE 0: as written; MODE 1: residual only (xo = r, nothing else written).
// Row per thread, entries in groups of eight with all index/coefficient loads issued before the gathers
// (the mapping of the PCG SpMV k_amul3, see there).
#define S4F_AMG_BLOCK 256
template <class T, class TB, class TO, int MODE>
__global__ void __launch_bounds__(S4F_AMG_BLOCK, 4) k_amg_step(const int* __restrict__ slicePtr, const int* __restrict__ col,
                                                               const T* __restrict__ a, const T* __restrict__ dg, const T* __restrict__ rD,
                                                               const TB* __restrict__ b, const T* __restrict__ x, T* __restrict__ d,
                                                               TO* __restrict__ xo, int n, int ld, int ldb, int ldo, int nSlices, T c1, T c2,
                                                               const int* __restrict__ act) {
    // a converged component of the fused PCG skips its vector traffic; the matrix stream is shared by the others
    const bool a0 = act[0] != 0, a1 = act[1] != 0, a2 = act[2] != 0;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nSlices; s += nWarps) {
        const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5;
        const int row = s * 32 + lane;
        T s0 = 0, s1 = 0, s2 = 0;
        for (int k0 = 0; k0 < width; k0 += 8) {
            int cc[8]; T e[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const bool ok = k0 + k < width;
                const int idx = base + 32 * (ok ? k0 + k : k0) + lane;
                cc[k] = col[idx];
                e[k] = ok ? a[idx] : (T)0;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (a0) s0 += e[k] * x[cc[k]];
                if (a1) s1 += e[k] * x[cc[k] + ld];
                if (a2) s2 += e[k] * x[cc[k] + 2 * ld];
            }
        }
        if (row < n) {
            const T acc[3] = {s0, s1, s2};
            const bool aq[3] = {a0, a1, a2};
#pragma unroll
            for (int q = 0; q < 3; q++) {
                if (!aq[q]) continue;
                const int j = q * ld + row;
                const T xv = x[j];
                const T r = (T)b[(size_t)q * ldb + row] - (dg[j] * xv - acc[q]);
                if (MODE == 1) { xo[(size_t)q * ldo + row] = (TO)r; }
                else {
                    const T dn = (c1 != (T)0 ? c1 * d[j] : (T)0) + c2 * rD[j] * r;
                    d[j] = dn;
                    xo[(size_t)q * ldo + row] = (TO)(xv + dn);
                }
            }
        }
    }
}

// b_c[I] = sum over the children of I of t[child]
template <class T>
__global__ void k_amg_restrict(const int* __restrict__ childPtr, const int* __restrict__ child, const T* __restrict__ t,
                               T* __restrict__ bc, int nc, int ld, int ldc, const int* __restrict__ act) {
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nc) return;
    const bool a0 = act[0] != 0, a1 = act[1] != 0, a2 = act[2] != 0;
    T s0 = 0, s1 = 0, s2 = 0;
    for (int e = childPtr[I]; e < childPtr[I + 1]; e++) {
        const int i = child[e];
        if (a0) s0 += t[i];
        if (a1) s1 += t[(size_t)ld + i];
        if (a2) s2 += t[2 * (size_t)ld + i];
    }
    bc[I] = s0; bc[(size_t)ldc + I] = s1; bc[2 * (size_t)ldc + I] = s2;
}

// x[i] += omega * e[parent[i]]
template <class T>
__global__ void __launch_bounds__(S4F_BLOCK) k_amg_prolong(const int* __restrict__ parent, const T* __restrict__ e, T* __restrict__ x,
                                                           int n, int ld, int ldc, T omega, const int* __restrict__ act) {
    const int a[3] = {act[0], act[1], act[2]};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int I = parent[i];
#pragma unroll
        for (int q = 0; q < 3; q++) if (a[q]) x[(size_t)q * ld + i] += omega * e[(size_t)q * ldc + I];
    }
}

// coarsest level: x = Ainv_q b, one warp per (row, component)
template <class T, class TB, class TO>
__global__ void k_amg_dense(const T* __restrict__ inv, const TB* __restrict__ b, TO* __restrict__ x, int n, int ldb, int ldo) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= 3 * n) return;
    const int q = gw / n, i = gw % n;
    const T* row = inv + ((size_t)q * n + i) * n;
    T s = 0;
    for (int k = lane; k < n; k += 32) s += row[k] * (T)b[(size_t)q * ldb + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) x[(size_t)q * ldo + i] = (TO)s;
}

// halo of a level-0 work vector (type T): boundary-cell values -> staging, staging -> ghost slots [N, N+G)
template <class T>
__global__ void k_amg_pack(const T* __restrict__ f, const int* __restrict__ sendCells, T* __restrict__ buf, int G, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * G) return;
    const int q = i / G, g = i % G;
    buf[i] = f[(size_t)q * ld + sendCells[g]];
}
template <class T>
__global__ void k_amg_unpack(T* __restrict__ f, const T* __restrict__ buf, int G, int N, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * G) return;
    const int q = i / G, g = i % G;
    f[(size_t)q * ld + N + g] = buf[i];
}
// all-gathered level-1 right-hand sides ([rank][3][maxLoc]) -> the replicated level-1 vector
template <class T>
__global__ void k_amg_scatter_gathered(const T* __restrict__ recv, T* __restrict__ b, const int* __restrict__ off, const int* __restrict__ cnt,
                                       int world, int maxLoc, int ldc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (r >= world || i >= cnt[r]) return;
#pragma unroll
    for (int q = 0; q < 3; q++) b[(size_t)q * ldc + off[r] + i] = recv[((size_t)r * 3 + q) * maxLoc + i];
}

template <class T>
__global__ void k_amg_convert(const double* __restrict__ in, T* __restrict__ out, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (T)in[i];
}
template <class T>
__global__ void k_amg_diag(const double* __restrict__ diagC, T* __restrict__ dg, T* __restrict__ rD, int n, int ldIn, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const double v = diagC[(size_t)q * ldIn + i];
        dg[(size_t)q * ld + i] = (T)v; rD[(size_t)q * ld + i] = (T)(1.0 / v);
    }
}
__global__ void k_gather_upper(const int* __restrict__ faceEntry, const double* __restrict__ eA, double* __restrict__ upper, int F) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < F) upper[f] = -eA[faceEntry[f]];
}


// ---- K-cycle on level 1 (gamgCycle 2): two flexible-CG steps on the coarse problem A_1 x = b, each preconditioned by the
// V-cycle from level 1 down (Notay's aggregation multigrid): the coarse correction is scaled by the Krylov step instead of a
// fixed factor.  Per component q:  c1 = V(b), v1 = A c1, rho1 = c1.v1, alpha1 = c1.b;  r = b - (alpha1/rho1) v1;
// c2 = V(r), v2 = A c2, gamma = c2.v1, beta = c2.v2, alpha2 = c2.r, rho2 = beta - gamma^2/rho1;
// x = (alpha1/rho1 - gamma alpha2/(rho1 rho2)) c1 + (alpha2/rho2) c2.
struct KcScalars { double rho1[3], alpha1[3], gamma[3], beta[3], alpha2[3]; };

// v = A c (row gather over the level's SELL rows) with the dot products the step needs, accumulated by atomics:
// STEP 1: rho1 += c.v, alpha1 += c.b;   STEP 2: gamma += c.v1, beta += c.v, alpha2 += c.r
template <class T, int STEP>
__global__ void __launch_bounds__(S4F_AMG_BLOCK, 4) k_kc_amul(const int* __restrict__ slicePtr, const int* __restrict__ col,
                                                              const T* __restrict__ a, const T* __restrict__ dg, const T* __restrict__ cvec,
                                                              const T* __restrict__ bvec /* b (step 1) or r (step 2) */,
                                                              const T* __restrict__ v1, T* __restrict__ v, int n, int ld, int nSlices,
                                                              KcScalars* S, const int* __restrict__ act) {
    const bool aq[3] = {act[0] != 0, act[1] != 0, act[2] != 0};
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    double d0[3] = {0, 0, 0}, d1[3] = {0, 0, 0}, d2[3] = {0, 0, 0};
    for (int s = warp; s < nSlices; s += nWarps) {
        const int base = slicePtr[s], width = (slicePtr[s + 1] - base) >> 5;
        const int row = s * 32 + lane;
        T acc[3] = {0, 0, 0};
        for (int k = 0; k < width; k++) {
            const int idx = base + 32 * k + lane;
            const int cc = col[idx];
            const T e = a[idx];
#pragma unroll
            for (int q = 0; q < 3; q++) if (aq[q]) acc[q] += e * cvec[cc + q * ld];
        }
        if (row < n) {
#pragma unroll
            for (int q = 0; q < 3; q++) {
                if (!aq[q]) continue;
                const int j = q * ld + row;
                const T cv = cvec[j];
                const T vv = dg[j] * cv - acc[q];
                v[j] = vv;
                d0[q] += (double)cv * (double)vv;
                d1[q] += (double)cv * (double)bvec[j];
                if (STEP == 2) d2[q] += (double)cv * (double)v1[j];
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 3; q++) {
        if (!aq[q]) continue;
        double x0 = d0[q], x1 = d1[q], x2 = d2[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { x0 += __shfl_xor_sync(0xffffffffu, x0, o); x1 += __shfl_xor_sync(0xffffffffu, x1, o); x2 += __shfl_xor_sync(0xffffffffu, x2, o); }
        if (lane == 0) {
            if (STEP == 1) { atomicAdd(&S->rho1[q], x0); atomicAdd(&S->alpha1[q], x1); }
            else { atomicAdd(&S->beta[q], x0); atomicAdd(&S->alpha2[q], x1); atomicAdd(&S->gamma[q], x2); }
        }
    }
}

// r = b - (alpha1/rho1) v1
template <class T>
__global__ void k_kc_resid(const T* __restrict__ b, const T* __restrict__ v1, T* __restrict__ r, int n, int ld, const KcScalars* S,
                           const int* __restrict__ act) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        if (!act[q]) continue;
        const double rho = S->rho1[q];
        const T sc = (T)(fabs(rho) > 1e-300 ? S->alpha1[q] / rho : 0.0);
        r[q * ld + i] = b[q * ld + i] - sc * v1[q * ld + i];
    }
}

// x = (alpha1/rho1 - gamma alpha2/(rho1 rho2)) c1 + (alpha2/rho2) c2
template <class T>
__global__ void k_kc_final(const T* __restrict__ c1, const T* __restrict__ c2, T* __restrict__ x, int n, int ld, const KcScalars* S,
                           const int* __restrict__ act) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        if (!act[q]) continue;
        const double rho1 = S->rho1[q];
        double k1 = 0, k2 = 0;
        if (fabs(rho1) > 1e-300) {
            k1 = S->alpha1[q] / rho1;
            const double rho2 = S->beta[q] - S->gamma[q] * S->gamma[q] / rho1;
            if (fabs(rho2) > 1e-300 * fabs(S->beta[q]) && fabs(rho2) > 1e-300) { k2 = S->alpha2[q] / rho2; k1 -= S->gamma[q] * k2 / rho1; }
        }
        x[q * ld + i] = (T)k1 * c1[q * ld + i] + (T)k2 * c2[q * ld + i];
    }
}

// ================================================================================================
// hierarchy
// ================================================================================================
template <class T>
struct Level {
    int n = 0, ld = 0, nSlices = 0;
    const int* slicePtr = nullptr; const int* col = nullptr; const T* a = nullptr;   // level 0 aliases the fine rows
    DevBuf<int> slicePtrB, colB; DevBuf<T> aB;
    DevBuf<T> dg, rD;                   // 3*ld
    DevBuf<int> parent;                 // [n] -> next level
    DevBuf<int> childPtr, child;        // children lists of THIS level's cells in the finer level
    DevBuf<T> b, x, x2, d, t;           // 3*ld work vectors
};

}  // namespace

struct S4fAmg {
    virtual ~S4fAmg() {}
    virtual int apply(s4fgpu_ctx* c, const double* r3, double* z3, const int* act) = 0;
    virtual int step0(s4fgpu_ctx* c, const double* r3, const int* act) = 0;   // one fine-level smoothing step alone (timing)
    double step0Bytes = 0;
    std::vector<int> sizes;
    double bytesPerApply = 0;
    double setupSeconds = 0;
};

namespace {

template <class T>
struct Hierarchy : S4fAmg {
    std::vector<std::unique_ptr<Level<T>>> lv;
    DevBuf<T> denseInv;                 // 3 * nC * nC
    int deg = 2, cycle = 0;
    DevBuf<T> kc1, kv1, kr, kc2;        // K-cycle work vectors on level 1 (3*ld each)
    DevBuf<KcScalars> kS;
    double omegaK = 1.0;                // scaling of the K-cycle's coarse correction on the fine level
    const int* act = nullptr;           // device int[3]: components to work on (the fused PCG's active flags, or all ones)
    double omega = 2.2;
    double theta = 0, delta = 0;
    // multi-rank: level 0 is this rank's part of the mesh (ghost columns, halo exchange per SpMV); levels >= 1
    // are the GLOBAL coarse levels, replicated on every rank and fed by an all-gather of the restricted residual
    bool dist = false;
    int n1Local = 0, n1Off = 0, maxLoc = 0;
    DevBuf<T> hsend, hrecv, gsend, grecv;
    DevBuf<int> rankOff, rankCnt;

    static ncclDataType_t nccl_t() { return sizeof(T) == 4 ? ncclFloat : ncclDouble; }
    int halo0(s4fgpu_ctx* c, T* x) {
        if (!dist || c->G == 0) return 0;
        const int G = c->G;
        k_amg_pack<T><<<(3 * G + 255) / 256, 256, 0, c->stream>>>(x, c->sendCells.p, hsend.p, G, c->ld);
        S4F_CHECK_NCCL(c, ncclGroupStart());
        for (const auto& nb : c->nbrs)
            for (int q = 0; q < 3; q++) {
                S4F_CHECK_NCCL(c, ncclSend(hsend.p + (size_t)q * G + nb.sendOff, nb.count, nccl_t(), nb.rank, c->comm, c->stream));
                S4F_CHECK_NCCL(c, ncclRecv(hrecv.p + (size_t)q * G + nb.sendOff, nb.count, nccl_t(), nb.rank, c->comm, c->stream));
            }
        S4F_CHECK_NCCL(c, ncclGroupEnd());
        k_amg_unpack<T><<<(3 * G + 255) / 256, 256, 0, c->stream>>>(x, hrecv.p, G, c->N, c->ld);
        c->launches += 2;
        return 0;
    }

    int build_level_rows(s4fgpu_ctx* c, Level<T>& L, const HostLevel& H) {
        const int n = H.n;
        L.n = n; L.ld = ((n + 31) / 32) * 32; L.nSlices = (n + 31) / 32;
        std::vector<int> cnt(n, 0);
        const size_t F = H.own.size();
        for (size_t f = 0; f < F; f++) { cnt[H.own[f]]++; cnt[H.nei[f]]++; }
        std::vector<long long> rowPtr(n + 1, 0);
        for (int i = 0; i < n; i++) rowPtr[i + 1] = rowPtr[i] + cnt[i];
        std::vector<int> rc(rowPtr[n]); std::vector<double> rv(rowPtr[n]);
        {
            std::vector<long long> cur(rowPtr.begin(), rowPtr.end() - 1);
            for (size_t f = 0; f < F; f++) { long long e = cur[H.nei[f]]++; rc[e] = H.own[f]; rv[e] = H.a[f]; }
            for (size_t f = 0; f < F; f++) { long long e = cur[H.own[f]]++; rc[e] = H.nei[f]; rv[e] = H.a[f]; }
        }
        std::vector<int> sp(L.nSlices + 1, 0);
        for (int s = 0; s < L.nSlices; s++) {
            int w = 0;
            for (int r = s * 32; r < std::min(n, s * 32 + 32); r++) w = std::max(w, cnt[r]);
            sp[s + 1] = sp[s] + 32 * w;
        }
        const size_t nE = sp[L.nSlices];
        std::vector<int> hc(std::max<size_t>(nE, 1), 0); std::vector<T> ha(std::max<size_t>(nE, 1), (T)0);
        for (int s = 0; s < L.nSlices; s++) {
            const int w = (sp[s + 1] - sp[s]) / 32;
            for (int lane = 0; lane < 32; lane++) {
                const int P = s * 32 + lane;
                for (int k = 0; k < w; k++) {
                    const size_t E = (size_t)sp[s] + 32 * k + lane;
                    if (P >= n) { hc[E] = 0; continue; }
                    if (k >= cnt[P]) { hc[E] = P; continue; }
                    hc[E] = rc[rowPtr[P] + k]; ha[E] = (T)rv[rowPtr[P] + k];
                }
            }
        }
        S4F_CHECK_CUDA(c, L.slicePtrB.upload(sp)); S4F_CHECK_CUDA(c, L.colB.upload(hc)); S4F_CHECK_CUDA(c, L.aB.upload(ha));
        L.slicePtr = L.slicePtrB.p; L.col = L.colB.p; L.a = L.aB.p;
        std::vector<T> hd(3 * (size_t)L.ld, (T)1), hr(3 * (size_t)L.ld, (T)1);
        for (int q = 0; q < 3; q++) for (int i = 0; i < n; i++) { hd[(size_t)q * L.ld + i] = (T)H.diag[q][i]; hr[(size_t)q * L.ld + i] = (T)(1.0 / H.diag[q][i]); }
        S4F_CHECK_CUDA(c, L.dg.upload(hd)); S4F_CHECK_CUDA(c, L.rD.upload(hr));
        return 0;
    }
    int alloc_work(s4fgpu_ctx* c, Level<T>& L) {
        const size_t m = 3 * (size_t)L.ld;
        S4F_CHECK_CUDA(c, L.b.alloc(m)); S4F_CHECK_CUDA(c, L.x.alloc(m)); S4F_CHECK_CUDA(c, L.x2.alloc(m));
        S4F_CHECK_CUDA(c, L.d.alloc(m)); S4F_CHECK_CUDA(c, L.t.alloc(m));
        return 0;
    }
    // parent[i] = coarse cell of fine cell i (global numbering); the children lists cover the coarse cells
    // [off, off+nc) that this rank's fine cells feed (off = 0, nc = all on replicated / serial levels)
    int set_transfer(s4fgpu_ctx* c, Level<T>& fine, Level<T>& coarse, const std::vector<int>& parent, int nc, int off = 0) {
        S4F_CHECK_CUDA(c, fine.parent.upload(parent));
        std::vector<int> ptr(nc + 1, 0), ch(parent.size());
        for (size_t i = 0; i < parent.size(); i++) ptr[parent[i] - off + 1]++;
        for (int i = 0; i < nc; i++) ptr[i + 1] += ptr[i];
        std::vector<int> cur(ptr.begin(), ptr.end() - 1);
        for (size_t i = 0; i < parent.size(); i++) ch[cur[parent[i] - off]++] = (int)i;
        S4F_CHECK_CUDA(c, coarse.childPtr.upload(ptr)); S4F_CHECK_CUDA(c, coarse.child.upload(ch));
        return 0;
    }

    static int step_grid(const s4fgpu_ctx* c, const Level<T>& L) { return s4f_grid(c->numSMs, (long long)L.nSlices * 32, 4); }
    // algorithmic bytes of one application (for the roofline report): every array read / written once
    double bytes_per_apply(int fineLdUnused) const {
        (void)fineLdUnused;
        double tot = 0, below1 = 0;
        const double sT = sizeof(T);
        for (size_t l = 0; l + 1 < lv.size(); l++) {
            const Level<T>& L = *lv[l];
            const double n = L.n, nz = nnz[l], sB = (l == 0) ? 8.0 : sT, sO = (l == 0) ? 8.0 : sT, nc = lv[l + 1]->n;
            const double mat = nz * (4 + sT) + n * 0.125;
            double t = 0;
            t += 3 * n * (sT + sB + 2 * sT);                                         // first
            t += (deg - 1) * (mat + 3 * n * (sB + 5 * sT + sT));                      // pre steps
            t += mat + 3 * n * (sB + 3 * sT);                                        // residual
            t += 3 * n * sT + 3 * nc * sT + 4 * n;                                   // restrict
            t += 4 * n + 6 * n * sT + 3 * nc * sT;                                   // prolong
            t += mat + 3 * n * (sB + 4 * sT + sT);                                   // post step 0 (no d read)
            t += (deg - 1) * (mat + 3 * n * (sB + 5 * sT)) + (deg > 1 ? 3 * n * sO : 0);   // post steps
            tot += t;
            if (l >= 1) below1 += t;
        }
        if (cycle == 2 && lv.size() > 2 && !dist) {      // K-cycle: a second V-cycle from level 1, two A_1 products with their dots, r, x, two copies
            const double n1 = lv[1]->n, mat1 = nnz[1] * (4 + sT) + n1 * 0.125;
            tot += below1 + 2 * (mat1 + 3 * n1 * 4 * sT) + 3 * n1 * 3 * sT + 3 * n1 * 3 * sT + 2 * 3 * n1 * 2 * sT;
        }
        return tot;
    }
    std::vector<double> nnz;

    // ---- smoothing on one level ---------------------------------------------------------------
    template <class TB, class TO>
    void step(s4fgpu_ctx* c, Level<T>& L, const TB* b, int ldb, const T* xin, TO* xout, int ldo, double c1, double c2) {
        const int grid = step_grid(c, L);
        if (dist && &L == lv[0].get()) halo0(c, const_cast<T*>(xin));
        k_amg_step<T, TB, TO, 0><<<grid, S4F_AMG_BLOCK, 0, c->stream>>>(L.slicePtr, L.col, L.a, L.dg.p, L.rD.p, b, xin, L.d.p, xout, L.n, L.ld, ldb, ldo,
                                                                     L.nSlices, (T)c1, (T)c2, act);
        c->launches++;
    }
    // Chebyshev-Jacobi of degree `deg`; fromZero: x0 = 0.  Result ends in *xres (L.x or L.x2), or, when
    // `out` is given (level 0), the last step writes the fp64 output directly.
    template <class TB>
    T* smooth(s4fgpu_ctx* c, Level<T>& L, const TB* b, int ldb, bool fromZero, T* xcur, double* out, int ldo) {
        const double sigma = theta / delta;
        double rho = 1.0 / sigma;
        T* other = (xcur == L.x.p) ? L.x2.p : L.x.p;
        int k0 = 0;
        if (fromZero) {
            const int grid = s4f_grid(c->numSMs, L.n);
            k_amg_first<T, TB><<<grid, S4F_BLOCK, 0, c->stream>>>(L.rD.p, b, xcur, L.d.p, L.n, L.ld, ldb, (T)(1.0 / theta), act);
            c->launches++;
            k0 = 1;
        }
        for (int k = k0; k < deg; k++) {
            double c1, c2;
            if (k == 0) { c1 = 0.0; c2 = 1.0 / theta; }
            else { const double rhon = 1.0 / (2.0 * sigma - rho); c1 = rhon * rho; c2 = 2.0 * rhon / delta; rho = rhon; }
            const bool last = (k == deg - 1);
            if (last && out) { step<TB, double>(c, L, b, ldb, xcur, out, ldo, c1, c2); return nullptr; }
            step<TB, T>(c, L, b, ldb, xcur, other, L.ld, c1, c2);
            std::swap(xcur, other);
        }
        return xcur;
    }

    // t = b - A x on level l, restricted into the right-hand side of level l+1
    template <class TB>
    int residual_restrict(s4fgpu_ctx* c, size_t l, const TB* b, int ldb, T* x) {
        Level<T>& L = *lv[l];
        Level<T>& C = *lv[l + 1];
        const int grid = step_grid(c, L);
        const bool d0 = dist && l == 0;
        if (d0) { int rc = halo0(c, x); if (rc) return rc; }
        k_amg_step<T, TB, T, 1><<<grid, S4F_AMG_BLOCK, 0, c->stream>>>(L.slicePtr, L.col, L.a, L.dg.p, L.rD.p, b, x, L.d.p, L.t.p, L.n, L.ld, ldb, L.ld,
                                                                       L.nSlices, (T)0, (T)0, act);
        c->launches++;
        if (!d0) {
            k_amg_restrict<T><<<(C.n + 127) / 128, 128, 0, c->stream>>>(C.childPtr.p, C.child.p, L.t.p, C.b.p, C.n, L.ld, C.ld, act);
            c->launches++;
            return 0;
        }
        // this rank's aggregates -> packed [3][maxLoc]; all-gather over NVLink; scatter into the replicated vector
        k_amg_restrict<T><<<(n1Local + 127) / 128, 128, 0, c->stream>>>(C.childPtr.p, C.child.p, L.t.p, gsend.p, n1Local, L.ld, maxLoc, act);
        S4F_CHECK_NCCL(c, ncclAllGather(gsend.p, grecv.p, 3 * (size_t)maxLoc, nccl_t(), c->comm, c->stream));
        dim3 g2((maxLoc + 255) / 256, c->nRanks);
        k_amg_scatter_gathered<T><<<g2, 256, 0, c->stream>>>(grecv.p, C.b.p, rankOff.p, rankCnt.p, c->nRanks, maxLoc, C.ld);
        c->launches += 2;
        return 0;
    }

    template <class TB>
    int cycle_level(s4fgpu_ctx* c, size_t l, const TB* b, int ldb, double* out, int ldo) {
        Level<T>& L = *lv[l];
        if (l + 1 == lv.size()) {     // coarsest: dense inverse
            const int warps = 3 * L.n, blocks = (warps * 32 + 255) / 256;
            if (out) k_amg_dense<T, TB, double><<<blocks, 256, 0, c->stream>>>(denseInv.p, b, out, L.n, ldb, ldo);
            else k_amg_dense<T, TB, T><<<blocks, 256, 0, c->stream>>>(denseInv.p, b, L.x.p, L.n, ldb, L.ld);
            c->launches++;
            return 0;
        }
        Level<T>& C = *lv[l + 1];
        T* x = smooth<TB>(c, L, b, ldb, true, L.x.p, nullptr, 0);                 // pre-smoothing from zero
        { int rr = residual_restrict<TB>(c, l, b, ldb, x); if (rr) return rr; }
        int rc;
        const bool kcycle = (cycle == 2 && l == 0 && lv.size() > 2 && !dist);
        if (kcycle) { rc = kcycle_level1(c); if (rc) return rc; }
        else { rc = cycle_level<T>(c, l + 1, C.b.p, C.ld, nullptr, 0); if (rc) return rc; }
        {
            const int grid = s4f_grid(c->numSMs, L.n);
            k_amg_prolong<T><<<grid, S4F_BLOCK, 0, c->stream>>>(L.parent.p, C.x.p, x, L.n, L.ld, C.ld, (T)(kcycle ? omegaK : omega), act);
            c->launches++;
        }
        if (cycle == 1 && l + 2 < lv.size()) {   // W-cycle: a second coarse correction on the updated residual
            rc = residual_restrict<TB>(c, l, b, ldb, x); if (rc) return rc;
            rc = cycle_level<T>(c, l + 1, C.b.p, C.ld, nullptr, 0); if (rc) return rc;
            const int gridp = s4f_grid(c->numSMs, L.n);
            k_amg_prolong<T><<<gridp, S4F_BLOCK, 0, c->stream>>>(L.parent.p, C.x.p, x, L.n, L.ld, C.ld, (T)omega, act);
            c->launches++;
        }
        T* xr = smooth<TB>(c, L, b, ldb, false, x, out, ldo);                      // post-smoothing
        if (!out && xr != L.x.p) {   // callers read the level result from L.x
            S4F_CHECK_CUDA(c, cudaMemcpyAsync(L.x.p, xr, 3 * (size_t)L.ld * sizeof(T), cudaMemcpyDeviceToDevice, c->stream));
        }
        return 0;
    }

    // two FCG steps on A_1 x = C.b preconditioned by the V-cycle from level 1; result in lv[1]->x
    int kcycle_level1(s4fgpu_ctx* c) {
        Level<T>& C = *lv[1];
        const size_t m = 3 * (size_t)C.ld;
        if (kc1.n != m) {
            S4F_CHECK_CUDA(c, kc1.alloc(m)); S4F_CHECK_CUDA(c, kv1.alloc(m)); S4F_CHECK_CUDA(c, kr.alloc(m)); S4F_CHECK_CUDA(c, kc2.alloc(m));
            S4F_CHECK_CUDA(c, kS.alloc(1));
        }
        const int grid = step_grid(c, C), gv = (C.n + 255) / 256;
        S4F_CHECK_CUDA(c, cudaMemsetAsync(kS.p, 0, sizeof(KcScalars), c->stream));
        int rc = cycle_level<T>(c, 1, C.b.p, C.ld, nullptr, 0); if (rc) return rc;                       // c1 = V(b)
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(kc1.p, C.x.p, m * sizeof(T), cudaMemcpyDeviceToDevice, c->stream));
        k_kc_amul<T, 1><<<grid, S4F_AMG_BLOCK, 0, c->stream>>>(C.slicePtr, C.col, C.a, C.dg.p, kc1.p, C.b.p, nullptr, kv1.p, C.n, C.ld, C.nSlices, kS.p, act);
        k_kc_resid<T><<<gv, 256, 0, c->stream>>>(C.b.p, kv1.p, kr.p, C.n, C.ld, kS.p, act);
        c->launches += 2;
        rc = cycle_level<T>(c, 1, kr.p, C.ld, nullptr, 0); if (rc) return rc;                            // c2 = V(r)
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(kc2.p, C.x.p, m * sizeof(T), cudaMemcpyDeviceToDevice, c->stream));
        k_kc_amul<T, 2><<<grid, S4F_AMG_BLOCK, 0, c->stream>>>(C.slicePtr, C.col, C.a, C.dg.p, kc2.p, kr.p, kv1.p, C.t.p, C.n, C.ld, C.nSlices, kS.p, act);
        k_kc_final<T><<<gv, 256, 0, c->stream>>>(kc1.p, kc2.p, C.x.p, C.n, C.ld, kS.p, act);
        c->launches += 2;
        return 0;
    }

    // the fine-level Chebyshev-Jacobi step kernel on its own (no halo): the dominant kernel of the V-cycle
    int step0(s4fgpu_ctx* c, const double* r3, const int* actIn) override {
        act = actIn;
        Level<T>& L = *lv[0];
        if (lv.size() < 2) return 0;
        const int grid = step_grid(c, L);
        k_amg_step<T, double, T, 0><<<grid, S4F_AMG_BLOCK, 0, c->stream>>>(L.slicePtr, L.col, L.a, L.dg.p, L.rD.p, r3, L.x.p, L.d.p, L.x2.p, L.n, L.ld,
                                                                            c->ld, L.ld, L.nSlices, (T)0.3, (T)0.5, act);
        c->launches++;
        return 0;
    }

    int apply(s4fgpu_ctx* c, const double* r3, double* z3, const int* actIn) override {
        act = actIn;
        int rc = cycle_level<double>(c, 0, r3, c->ld, z3, c->ld);
        if (rc) return rc;
        S4F_CHECK_CUDA(c, cudaGetLastError());
        return 0;
    }
};

struct DistInfo { bool on = false; int n1Local = 0, n1Off = 0, maxLoc = 0; std::vector<int> off, cnt; };

template <class T>
int build(s4fgpu_ctx* c, std::vector<HostLevel>& H, const DistInfo& di) {
    auto* A = new Hierarchy<T>();
    std::unique_ptr<S4fAmg> guard(A);
    A->dist = di.on; A->n1Local = di.n1Local; A->n1Off = di.n1Off; A->maxLoc = di.maxLoc;
    if (di.on) {
        const size_t G = std::max(c->G, 1);
        S4F_CHECK_CUDA(c, A->hsend.alloc(3 * G)); S4F_CHECK_CUDA(c, A->hrecv.alloc(3 * G));
        S4F_CHECK_CUDA(c, A->gsend.alloc(3 * (size_t)std::max(di.maxLoc, 1)));
        S4F_CHECK_CUDA(c, A->grecv.alloc(3 * (size_t)std::max(di.maxLoc, 1) * c->nRanks));
        S4F_CHECK_CUDA(c, A->rankOff.upload(di.off)); S4F_CHECK_CUDA(c, A->rankCnt.upload(di.cnt));
    }
    A->deg = c->ctl.gamgSmootherDegree > 0 ? c->ctl.gamgSmootherDegree : 3;
    A->cycle = c->ctl.gamgCycle;
    A->omega = c->ctl.gamgOverCorrection > 0 ? c->ctl.gamgOverCorrection : 2.2;
    A->omegaK = A->omega;            // sweep: 1.0 -> 19.0, 1.5 -> 15.5, 2.2 -> 13.4, 2.6 -> 16.2 ms per outer iteration at 8 M cells
    if (const char* e = getenv("S4F_GAMG_OMEGA_K")) A->omegaK = atof(e);      // tuning aid (profiles/microbench/gamg_sweep.py)
    const double ratio = c->ctl.gamgSmootherRatio > 0 ? c->ctl.gamgSmootherRatio : 0.3;
    const double lmax = 2.0, lmin = ratio * lmax;    // Gershgorin bound of D^-1 A for the M-matrices of every level
    A->theta = 0.5 * (lmax + lmin); A->delta = 0.5 * (lmax - lmin);
    for (size_t l = 0; l < H.size(); l++) {
        A->lv.emplace_back(new Level<T>());
        Level<T>& L = *A->lv.back();
        int rc;
        if (l == 0) {
            L.n = c->N; L.ld = c->ld; L.nSlices = c->nSlices;
            L.slicePtr = c->slicePtr.p; L.col = c->col.p;
            if (sizeof(T) == sizeof(double)) L.a = reinterpret_cast<const T*>(c->eA.p);
            else {
                S4F_CHECK_CUDA(c, L.aB.alloc((size_t)c->nEntries, false));
                k_amg_convert<T><<<(unsigned)((c->nEntries + 255) / 256), 256, 0, c->stream>>>(c->eA.p, L.aB.p, c->nEntries);
                L.a = L.aB.p;
            }
            S4F_CHECK_CUDA(c, L.dg.alloc(3 * (size_t)L.ld)); S4F_CHECK_CUDA(c, L.rD.alloc(3 * (size_t)L.ld));
            k_amg_diag<T><<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->diagC.p, L.dg.p, L.rD.p, c->N, c->ld, L.ld);
            c->launches += 2;
        } else {
            if ((rc = A->build_level_rows(c, L, H[l]))) return rc;
        }
        if ((rc = A->alloc_work(c, L))) return rc;
        if (l == 1 && di.on) { if ((rc = A->set_transfer(c, *A->lv[0], L, H[0].parent, di.n1Local, di.n1Off))) return rc; }
        else if (l > 0) { if ((rc = A->set_transfer(c, *A->lv[l - 1], L, H[l - 1].parent, H[l].n))) return rc; }
        A->sizes.push_back(H[l].n);
        A->nnz.push_back(l == 0 ? (double)c->nnzOff : 2.0 * (double)H[l].own.size());
    }
    // dense inverse on the coarsest level
    const HostLevel& HC = H.back();
    std::vector<T> inv(3 * (size_t)HC.n * HC.n);
    for (int q = 0; q < 3; q++) {
        std::vector<double> iq;
        if (!dense_inverse(HC, q, iq)) { c->err = "GAMG: coarsest-level matrix is not positive definite"; return 1; }
        for (size_t i = 0; i < iq.size(); i++) inv[(size_t)q * HC.n * HC.n + i] = (T)iq[i];
    }
    S4F_CHECK_CUDA(c, A->denseInv.upload(inv));
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    A->bytesPerApply = A->bytes_per_apply(c->ld);
    A->step0Bytes = A->nnz[0] * (4 + sizeof(T)) + 0.125 * c->N + 3.0 * c->N * (8 + 5 * sizeof(T) + sizeof(T));   // col,a | b(fp64) x d dg rD in, d x' out
    c->amg = guard.release();
    return 0;
}

}  // namespace

int s4f_download_upper(s4fgpu_ctx* c, double* hostUpper) {
    if (c->F == 0) return 0;
    DevBuf<double> up;
    S4F_CHECK_CUDA(c, up.alloc((size_t)c->F, false));
    k_gather_upper<<<(c->F + 255) / 256, 256, 0, c->stream>>>(c->faceEntry.p, c->eA.p, up.p, c->F);
    c->launches++;
    S4F_CHECK_CUDA(c, cudaMemcpyAsync(hostUpper, up.p, (size_t)c->F * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

void s4f_amg_destroy(s4fgpu_ctx* c) {
    delete c->amg;
    c->amg = nullptr;
}

namespace {

// three pair-wise passes -> aggregates of up to 8 cells; `total` maps the cells of src to the coarse cells of out
int coarsen3(const HostLevel& srcIn, int stopBelow, std::vector<int>& total, HostLevel& out) {
    const HostLevel* src = &srcIn;
    total.resize(src->n);
    for (int i = 0; i < src->n; i++) total[i] = i;
    HostLevel tmpA, tmpB;
    int nc = src->n;
    for (int pass = 0; pass < 3; pass++) {
        std::vector<int> agg;
        const int ncNew = pairwise_pass(*src, agg);
        HostLevel& dst = (pass % 2 == 0) ? tmpA : tmpB;
        galerkin(*src, agg, ncNew, dst);
        for (size_t i = 0; i < total.size(); i++) total[i] = agg[total[i]];
        src = &dst; nc = ncNew;
        if (nc <= stopBelow) break;
    }
    out = *src;
    return nc;
}

// equal-sized host blocks, all-gathered through device staging (set-up only)
int allgather_host(s4fgpu_ctx* c, const void* send, size_t bytes, void* recv) {
    DevBuf<char> ds, dr;
    S4F_CHECK_CUDA(c, ds.alloc(std::max<size_t>(bytes, 1), false)); S4F_CHECK_CUDA(c, dr.alloc(std::max<size_t>(bytes, 1) * c->nRanks, false));
    if (bytes) S4F_CHECK_CUDA(c, cudaMemcpyAsync(ds.p, send, bytes, cudaMemcpyHostToDevice, c->stream));
    S4F_CHECK_NCCL(c, ncclAllGather(ds.p, dr.p, std::max<size_t>(bytes, 1), ncclChar, c->comm, c->stream));
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    if (bytes) S4F_CHECK_CUDA(c, cudaMemcpy(recv, dr.p, bytes * c->nRanks, cudaMemcpyDeviceToHost));
    return 0;
}

// one int per processor-patch face, sent to / received from the rank across the face (patch order)
int exchange_ghost_ints(s4fgpu_ctx* c, const std::vector<int>& send, std::vector<int>& recv) {
    const int G = c->G;
    recv.assign(G, 0);
    if (G == 0) return 0;
    DevBuf<int> ds, dr;
    S4F_CHECK_CUDA(c, ds.upload(send)); S4F_CHECK_CUDA(c, dr.alloc(G));
    S4F_CHECK_NCCL(c, ncclGroupStart());
    for (const auto& nb : c->nbrs) {
        S4F_CHECK_NCCL(c, ncclSend(ds.p + nb.sendOff, nb.count, ncclInt, nb.rank, c->comm, c->stream));
        S4F_CHECK_NCCL(c, ncclRecv(dr.p + nb.sendOff, nb.count, ncclInt, nb.rank, c->comm, c->stream));
    }
    S4F_CHECK_NCCL(c, ncclGroupEnd());
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    S4F_CHECK_CUDA(c, cudaMemcpy(recv.data(), dr.p, G * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

__global__ void k_gather_entries(const int* __restrict__ idx, const double* __restrict__ eA, double* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = eA[idx[i]];
}

// Multi-rank: aggregate this rank's cells, learn the aggregates of the ghost cells, form this rank's rows of
// the global level-1 matrix (incl. the couplings across processor patches) and all-gather the pieces, so that
// every rank holds the same global level 1 and continues the coarsening identically.
int build_global_level1(s4fgpu_ctx* c, HostLevel& L0, HostLevel& H1, DistInfo& di) {
    const int N = c->N, G = c->G, world = c->nRanks, rank = c->rank;
    std::vector<int> parentLoc; HostLevel dummy;
    const int n1Local = coarsen3(L0, 0, parentLoc, dummy);
    std::vector<int> cnt(world);
    int rc = allgather_host(c, &n1Local, sizeof(int), cnt.data()); if (rc) return rc;
    std::vector<int> off(world + 1, 0);
    for (int r = 0; r < world; r++) off[r + 1] = off[r] + cnt[r];
    const int n1Global = off[world], myOff = off[rank];
    // aggregates of the ghost cells + coefficients of the processor faces
    std::vector<int> sendP(std::max(G, 1), 0), ghostParent, sendCellsH(std::max(G, 1), 0), entryH(std::max(G, 1), 0);
    for (const auto& nb : c->nbrs)
        for (int i = 0; i < nb.count; i++) {
            const int cell = c->faceCells[c->pStart[nb.patch] + i];
            sendCellsH[nb.sendOff + i] = cell;
            sendP[nb.sendOff + i] = myOff + parentLoc[cell];
        }
    sendP.resize(G); 
    if ((rc = exchange_ghost_ints(c, sendP, ghostParent))) return rc;
    std::vector<double> aPf(std::max(G, 1), 0.0);
    if (G > 0) {
        DevBuf<double> d; S4F_CHECK_CUDA(c, d.alloc(G));
        k_gather_entries<<<(G + 255) / 256, 256, 0, c->stream>>>(c->procEntry.p, c->eA.p, d.p, G);
        c->launches++;
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(aPf.data(), d.p, G * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    // augmented level: local cells + ghost cells, internal faces + processor faces
    HostLevel aug; aug.n = N + G;
    aug.own = L0.own; aug.nei = L0.nei; aug.a = L0.a;
    for (int g = 0; g < G; g++) { aug.own.push_back(sendCellsH[g]); aug.nei.push_back(N + g); aug.a.push_back(aPf[g]); }
    for (int q = 0; q < 3; q++) { aug.diag[q] = L0.diag[q]; aug.diag[q].resize(N + G, 0.0); }
    std::vector<int> agg(N + G);
    for (int i = 0; i < N; i++) agg[i] = myOff + parentLoc[i];
    for (int g = 0; g < G; g++) agg[N + g] = ghostParent[g];
    HostLevel part; galerkin(aug, agg, n1Global, part);
    // the faces this rank contributes: owner (lower id) among its own aggregates
    std::vector<int> fo, fn; std::vector<double> fa;
    for (size_t f = 0; f < part.own.size(); f++)
        if (part.own[f] >= myOff && part.own[f] < myOff + n1Local) { fo.push_back(part.own[f]); fn.push_back(part.nei[f]); fa.push_back(part.a[f]); }
    int nF = (int)fo.size();
    std::vector<int> nFs(world);
    if ((rc = allgather_host(c, &nF, sizeof(int), nFs.data()))) return rc;
    int maxF = 1, maxLoc = 1;
    for (int r = 0; r < world; r++) { maxF = std::max(maxF, nFs[r]); maxLoc = std::max(maxLoc, cnt[r]); }
    std::vector<int> si(2 * (size_t)maxF, 0), ri(2 * (size_t)maxF * world);
    std::vector<double> sd((size_t)maxF + 3 * (size_t)maxLoc, 0.0), rd(((size_t)maxF + 3 * (size_t)maxLoc) * world);
    for (int f = 0; f < nF; f++) { si[f] = fo[f]; si[maxF + f] = fn[f]; sd[f] = fa[f]; }
    for (int q = 0; q < 3; q++) for (int i = 0; i < n1Local; i++) sd[(size_t)maxF + (size_t)q * maxLoc + i] = part.diag[q][myOff + i];
    if ((rc = allgather_host(c, si.data(), si.size() * sizeof(int), ri.data()))) return rc;
    if ((rc = allgather_host(c, sd.data(), sd.size() * sizeof(double), rd.data()))) return rc;
    H1 = HostLevel(); H1.n = n1Global;
    for (int q = 0; q < 3; q++) H1.diag[q].assign(n1Global, 0.0);
    for (int r = 0; r < world; r++) {
        const int* pi = ri.data() + (size_t)r * 2 * maxF;
        const double* pd = rd.data() + (size_t)r * ((size_t)maxF + 3 * (size_t)maxLoc);
        for (int f = 0; f < nFs[r]; f++) { H1.own.push_back(pi[f]); H1.nei.push_back(pi[maxF + f]); H1.a.push_back(pd[f]); }
        for (int q = 0; q < 3; q++) for (int i = 0; i < cnt[r]; i++) H1.diag[q][off[r] + i] = pd[(size_t)maxF + (size_t)q * maxLoc + i];
    }
    L0.parent.assign(agg.begin(), agg.begin() + N);
    di.on = true; di.n1Local = n1Local; di.n1Off = myOff; di.maxLoc = maxLoc;
    di.off.assign(off.begin(), off.begin() + world); di.cnt = cnt;
    return 0;
}

}  // namespace

// Build the hierarchy from the assembled fine matrix (upper() and the per-component diagonals).
int s4f_amg_setup(s4fgpu_ctx* c) {
    s4f_amg_destroy(c);
    const auto t0 = std::chrono::steady_clock::now();
    const int N = c->N, F = c->F;
    std::vector<HostLevel> H(1);
    {
        HostLevel& L0 = H[0];
        L0.n = N; L0.own = c->own; L0.nei = c->nei; L0.a.resize(F);
        int rc = s4f_download_upper(c, L0.a.data()); if (rc) return rc;
        for (int f = 0; f < F; f++) L0.a[f] = -L0.a[f];
        std::vector<double> d(3 * (size_t)c->ld);
        S4F_CHECK_CUDA(c, cudaMemcpy(d.data(), c->diagC.p, d.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int q = 0; q < 3; q++) L0.diag[q].assign(d.begin() + (size_t)q * c->ld, d.begin() + (size_t)q * c->ld + N);
    }
    const int coarsest = 512;
    DistInfo di;
    if (c->nRanks > 1) {
        HostLevel H1;
        int rc = build_global_level1(c, H[0], H1, di); if (rc) return rc;
        H.push_back(std::move(H1));
    }
    while (H.back().n > coarsest && H.size() < 12) {
        std::vector<int> total; HostLevel next;
        const int nc = coarsen3(H.back(), coarsest / 4, total, next);
        if (nc >= H.back().n) break;        // no coarsening possible (no faces)
        H.back().parent = total;
        H.push_back(std::move(next));
    }
    int rc;
    if (c->ctl.gamgSinglePrecision) rc = build<float>(c, H, di);
    else rc = build<double>(c, H, di);
    if (!rc) c->amg->setupSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

int s4f_amg_info(s4fgpu_ctx* c, int* nLevels, int* sizes, int maxLevels, double* bytesPerApply, double* setupSeconds) {
    if (!c->amg) { c->err = "GAMG hierarchy missing"; return 1; }
    *nLevels = (int)c->amg->sizes.size();
    for (int i = 0; i < *nLevels && i < maxLevels; i++) sizes[i] = c->amg->sizes[i];
    *bytesPerApply = c->amg->bytesPerApply; *setupSeconds = c->amg->setupSeconds;
    return 0;
}

int s4f_amg_step0(s4fgpu_ctx* c, const double* r3, double* bytes) {
    if (!c->amg) { c->err = "GAMG hierarchy missing"; return 1; }
    if (bytes) *bytes = c->amg->step0Bytes;
    return c->amg->step0(c, r3, c->amgAct ? c->amgAct : c->ones3.p);
}

int s4f_amg_apply(s4fgpu_ctx* c, const double* r3, double* z3) {
    if (!c->amg) { c->err = "GAMG hierarchy missing"; return 1; }
    return c->amg->apply(c, r3, z3, c->amgAct ? c->amgAct : c->ones3.p);
}
