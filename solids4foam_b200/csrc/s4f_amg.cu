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
// Multi-rank (OpenFOAM: GAMG over processor interfaces + processorAgglomeration): every rank agglomerates its own cells;
// the couplings across processor patches are kept on every level as interface entries with ghost columns, so levels
// stay DISTRIBUTED (each rank smooths its own aggregates, one peer-memory halo exchange per smoothing step,
// s4f_comm.cu) until the global level size falls below `replicateBelow` cells; from there the level is gathered to all
// ranks and the rest of the hierarchy is replicated (identical arithmetic on every rank, no communication).
#include <algorithm>
#include <chrono>
#include <cooperative_groups.h>
#include <type_traits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "s4f_amg_setup.h"
#include "s4f_comm.h"
#include "s4f_dev.cuh"

namespace cg = cooperative_groups;
namespace {

// ================================================================================================
// host: agglomeration
// ================================================================================================
// couplings of this rank's cells to cells of one neighbour rank, in an order both sides agree on
struct Iface {
    int rank = -1; std::vector<int> cell, remote; std::vector<double> a;
    std::vector<int> remoteParent;      // aggregate (local index on the other rank) of remote[f] in the next level (coefficient refresh)
};
struct HostLevel {
    int n = 0;
    std::vector<int> own, nei;          // faces, upper-triangular order
    std::vector<double> a;              // positive face coefficient (= -upper)
    std::vector<double> diag[3];
    std::vector<int> parent;            // aggregate of each cell in the next (coarser) level: its index in that level's vectors
    bool dist = false;                  // one part per rank (n = this rank's cells); false: the whole level on every rank
    std::vector<Iface> ifc;             // dist: couplings across rank boundaries
    int childOff = 0, childCnt = 0;     // the range of the next level's cells that this rank's cells feed
    std::vector<int> rankOff;           // gathered next level: first global cell of every rank
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

// x = (1/theta) b/diag      (first smoothing step from a zero initial guess)
template <class T, class TB>
__global__ void __launch_bounds__(S4F_BLOCK) k_amg_first(const T* __restrict__ dg, const TB* __restrict__ b, T* __restrict__ x,
                                                         int n, int ld, int ldb, T invTheta, const int* __restrict__ act) {
    const int a[3] = {act[0], act[1], act[2]};      // components whose PCG solve is still running (device-side flags)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int q = 0; q < 3; q++) {
            if (!a[q]) continue;
            x[(size_t)q * ld + i] = invTheta * (T)b[(size_t)q * ldb + i] / dg[(size_t)q * ld + i];
        }
    }
}

// one Chebyshev-Jacobi step in its three-term form:  r = b - A x;  x' = x + c1 (x - xprev) + c2 r/diag   (out of place)
// Per component the step reads b, x, xprev, diag and writes x' (40 B per row in fp64; the direction-vector form of round 1
// read and wrote a fourth vector and a stored 1/diag: 56 B).  PREV 0: no previous iterate (c1 = 0, first post-smoothing
// step); 1: xprev given; 2: the previous iterate is the zero initial guess (second pre-smoothing step).
// MODE 0: as written; MODE 1: residual only (xo = r).
// Row per thread, entries in groups of eight with all index/coefficient loads issued before the gathers
// (the mapping of the PCG SpMV k_amul3, see there).
#define S4F_AMG_BLOCK 256
template <class T, class TB, class TO, int MODE>
__global__ void __launch_bounds__(S4F_AMG_BLOCK, 4) k_amg_step(const int* __restrict__ slicePtr, const int* __restrict__ col,
                                                               const T* __restrict__ a, const T* __restrict__ dg,
                                                               const TB* __restrict__ b, const T* __restrict__ x, const T* __restrict__ xprev,
                                                               TO* __restrict__ xo, int n, int ld, int ldb, int ldo, int nSlices, T c1, T c2,
                                                               int prevMode, const int* __restrict__ act) {
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
                const T xv = x[j], d = dg[j];
                const T r = (T)b[(size_t)q * ldb + row] - (d * xv - acc[q]);
                if (MODE == 1) { xo[(size_t)q * ldo + row] = (TO)r; }
                else {
                    T xn = xv + c2 * r / d;
                    if (prevMode == 1) xn += c1 * (xv - xprev[j]);
                    else if (prevMode == 2) xn += c1 * xv;
                    xo[(size_t)q * ldo + row] = (TO)xn;
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

template <class T>
__global__ void k_amg_convert(const double* __restrict__ in, T* __restrict__ out, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (T)in[i];
}
template <class T>
__global__ void k_amg_diag(const double* __restrict__ diagC, T* __restrict__ dg, int n, int ldIn, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int q = 0; q < 3; q++) dg[(size_t)q * ld + i] = (T)diagC[(size_t)q * ldIn + i];
}
__global__ void k_gather_upper(const int* __restrict__ faceEntry, const double* __restrict__ eA, double* __restrict__ upper, int F) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < F) upper[f] = -eA[faceEntry[f]];
}


// Galerkin sums of one coarse level from the (new) coefficients of the finer one, aggregates unchanged: coarse row I collects, over
// its children i and their entries (i, j, a), a into the coarse entry of J = parent[j], or -- J = I -- out of the diagonal:
// diag_c[I] = sum_i diag[i] - sum_{j in I} a_ij.  One thread per coarse row; set-up work, not on the iteration path.
#define S4F_AMG_MAXW 64
template <class T>
__global__ void k_amg_galerkin(const int* __restrict__ spF, const int* __restrict__ colF, const T* __restrict__ aF, const T* __restrict__ dgF,
                               const int* __restrict__ parent, int nFine, int ldF, const int* __restrict__ spC, const int* __restrict__ colC,
                               T* __restrict__ aC, T* __restrict__ dgC, const int* __restrict__ childPtr, const int* __restrict__ child, int nC,
                               int ldC, int* __restrict__ fail, const int* __restrict__ ghostParent /* decomposed fine level, else null */,
                               int nGhostFine, int rowOff) {
    // nC rows starting at rowOff: all rows, or -- next level gathered from the ranks -- the rows of this rank's aggregates
    const int Iloc = blockIdx.x * blockDim.x + threadIdx.x;
    if (Iloc >= nC) return;
    const int I = rowOff + Iloc;
    const int sC = I >> 5, lC = I & 31, baseC = spC[sC], wC = (spC[sC + 1] - baseC) >> 5;
    if (wC > S4F_AMG_MAXW) { *fail = 1; return; }
    int cc[S4F_AMG_MAXW]; double acc[S4F_AMG_MAXW];
    for (int k = 0; k < wC; k++) { cc[k] = colC[baseC + 32 * k + lC]; acc[k] = 0.0; }
    double dsub = 0, dsum[3] = {0, 0, 0};
    for (int ce = childPtr[Iloc]; ce < childPtr[Iloc + 1]; ce++) {
        const int i = child[ce];
        const int sF = i >> 5, lF = i & 31, baseF = spF[sF], wF = (spF[sF + 1] - baseF) >> 5;
#pragma unroll
        for (int q = 0; q < 3; q++) dsum[q] += (double)dgF[(size_t)q * ldF + i];
        for (int k = 0; k < wF; k++) {
            const double a = (double)aF[baseF + 32 * k + lF];
            if (a == 0.0) continue;
            const int j = colF[baseF + 32 * k + lF];
            if (j >= nFine && !(ghostParent && j - nFine < nGhostFine)) continue;      // boundary slots
            const int J = j < nFine ? parent[j] : ghostParent[j - nFine];                // across a processor patch: the other rank's aggregate
            if (J == I) { dsub += a; continue; }
            int k2 = 0;
            while (k2 < wC && cc[k2] != J) k2++;
            if (k2 == wC) { *fail = 2; return; }
            acc[k2] += a;
        }
    }
    for (int k = 0; k < wC; k++) aC[baseC + 32 * k + lC] = (T)acc[k];
#pragma unroll
    for (int q = 0; q < 3; q++) dgC[(size_t)q * ldC + I] = (T)(dsum[q] - dsub);
}

// Refresh of a gathered level: every rank has re-summed the rows of ITS aggregates; the rows travel to all ranks three SELL
// entry positions at a time through the gather plan of the level above (the one that carries the restricted right-hand side).
template <class T>
__global__ void k_amg_rows_pack(const int* __restrict__ sp, const T* __restrict__ a, int rowOff, int nLocal, int k0, T* __restrict__ out, int stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nLocal) return;
    const int I = rowOff + i, base = sp[I >> 5], w = (sp[(I >> 5) + 1] - base) >> 5;
#pragma unroll
    for (int q = 0; q < 3; q++) out[(size_t)q * stride + i] = (k0 + q < w) ? a[base + 32 * (k0 + q) + (I & 31)] : (T)0;
}
template <class T>
__global__ void k_amg_rows_unpack(const int* __restrict__ sp, T* __restrict__ a, int n, int k0, const T* __restrict__ in, int ld) {
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= n) return;
    const int base = sp[I >> 5], w = (sp[(I >> 5) + 1] - base) >> 5;
#pragma unroll
    for (int q = 0; q < 3; q++) if (k0 + q < w) a[base + 32 * (k0 + q) + (I & 31)] = in[(size_t)q * ld + I];
}
template <class T>
__global__ void k_amg_diag_pack(const T* __restrict__ dg, int rowOff, int nLocal, int ld, T* __restrict__ out, int stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nLocal) return;
#pragma unroll
    for (int q = 0; q < 3; q++) out[(size_t)q * stride + i] = dg[(size_t)q * ld + rowOff + i];
}

// ---- K-cycle on level 1 (gamgCycle 2): two flexible-CG steps on the coarse problem A_1 x = b, each preconditioned by the
// V-cycle from level 1 down (Notay's aggregation multigrid): the coarse correction is scaled by the Krylov step instead of a
// fixed factor.  Per component q:  c1 = V(b), v1 = A c1, rho1 = c1.v1, alpha1 = c1.b;  r = b - (alpha1/rho1) v1;
// c2 = V(r), v2 = A c2, gamma = c2.v1, beta = c2.v2, alpha2 = c2.r, rho2 = beta - gamma^2/rho1;
// x = (alpha1/rho1 - gamma alpha2/(rho1 rho2)) c1 + (alpha2/rho2) c2.
struct KcScalars { double rho1[3], alpha1[3], gamma[3], beta[3], alpha2[3]; };

template <int STEP>
struct FinKc {
    KcScalars* S;
    __device__ void operator()(const double* tot) const {
#pragma unroll
        for (int q = 0; q < 3; q++) {
            if (STEP == 1) { S->rho1[q] = tot[q]; S->alpha1[q] = tot[3 + q]; }
            else { S->beta[q] = tot[q]; S->alpha2[q] = tot[3 + q]; S->gamma[q] = tot[6 + q]; }
        }
    }
};

// v = A c (row gather over the level's SELL rows) with the dot products the step needs (deterministic grid reduction,
// all-reduced over the ranks when the level is distributed):
// STEP 1: rho1 = c.v, alpha1 = c.b;   STEP 2: beta = c.v, alpha2 = c.r, gamma = c.v1
template <class T, int STEP>
__global__ void __launch_bounds__(S4F_AMG_BLOCK, 4) k_kc_amul(const int* __restrict__ slicePtr, const int* __restrict__ col,
                                                              const T* __restrict__ a, const T* __restrict__ dg, const T* __restrict__ cvec,
                                                              const T* __restrict__ bvec /* b (step 1) or r (step 2) */,
                                                              const T* __restrict__ v1, T* __restrict__ v, int n, int ld, int nSlices,
                                                              KcScalars* S, const int* __restrict__ act, RedCtx red) {
    const bool aq[3] = {act[0] != 0, act[1] != 0, act[2] != 0};
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nWarps = (gridDim.x * blockDim.x) >> 5;
    constexpr int NV = (STEP == 1) ? 6 : 9;
    double d[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) d[i] = 0;
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
                d[q] += (double)cv * (double)vv;
                d[3 + q] += (double)cv * (double)bvec[j];
                if constexpr (STEP == 2) d[6 + q] += (double)cv * (double)v1[j];
            }
        }
    }
    grid_reduce<NV, OpSum>(d, red, FinKc<STEP>{S});
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

// device copy of the active components (a kernel, not a memcpy node: the cycle is replayed inside a CUDA graph loop)
template <class T>
__global__ void __launch_bounds__(S4F_BLOCK) k_amg_copy(const T* __restrict__ src, T* __restrict__ dst, int n, int ld, const int* __restrict__ act) {
    const int a[3] = {act[0], act[1], act[2]};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int q = 0; q < 3; q++) if (a[q]) dst[(size_t)q * ld + i] = src[(size_t)q * ld + i];
}


// ---- the small levels of the hierarchy as ONE kernel -----------------------------------------------------------------------
// An experiment kept as an option (S4F_AMG_TAIL_MAX, off by default -- it is slower, see record_tail): below ~20 k rows a
// level's kernels look launch-floor-bound: 9 launches per visit, visited twice per K-cycle.  The V-cycle from such a level
// down can be recorded once, at set-up, as a list of operations (the same calls that launch the kernels above, in recording
// mode), and replayed by one thread-block cluster of 8 x 1024 threads that steps through the list with a hardware cluster
// barrier between operations instead of a kernel boundary.  Same arithmetic, same summation order inside a row.
enum { TAIL_FIRST = 0, TAIL_STEP = 1, TAIL_RESID = 2, TAIL_RESTRICT = 3, TAIL_PROLONG = 4, TAIL_DENSE = 5, TAIL_COPY = 6 };
template <class T>
struct TailOp {
    int kind, n, ld, ldc, nSlices, prevMode;
    const int *slicePtr, *col, *idxA, *idxB;          // idxA/idxB: childPtr/child (restrict) or parent (prolong)
    const T *a, *dg, *b, *x, *xprev;
    T* xo;
    T c1, c2;
};
#define S4F_TAIL_CTAS 8
#define S4F_TAIL_THREADS 1024
template <class T>
__global__ void __cluster_dims__(S4F_TAIL_CTAS, 1, 1) __launch_bounds__(S4F_TAIL_THREADS, 1)
k_amg_tail(const TailOp<T>* __restrict__ ops, int nOps, const int* __restrict__ act) {
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = cluster.block_rank() * blockDim.x + threadIdx.x, nT = cluster.num_blocks() * blockDim.x;
    const bool aq[3] = {act[0] != 0, act[1] != 0, act[2] != 0};
    for (int o = 0; o < nOps; o++) {
        const TailOp<T> op = ops[o];
        const int n = op.n, ld = op.ld;
        if (op.kind == TAIL_FIRST) {
            for (int i = tid; i < n; i += nT)
#pragma unroll
                for (int q = 0; q < 3; q++) if (aq[q]) op.xo[(size_t)q * ld + i] = op.c2 * __ldcg(&op.b[(size_t)q * ld + i]) / op.dg[(size_t)q * ld + i];
        } else if (op.kind == TAIL_STEP || op.kind == TAIL_RESID) {
            for (int row = tid; row < op.nSlices * 32; row += nT) {
                const int sl = row >> 5, lane = row & 31, base = op.slicePtr[sl], width = (op.slicePtr[sl + 1] - base) >> 5;
                T s0 = 0, s1 = 0, s2 = 0;
                for (int k0 = 0; k0 < width; k0 += 4) {
                    int cc[4]; T e[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const bool ok = k0 + k < width;
                        const int idx = base + 32 * (ok ? k0 + k : k0) + lane;
                        cc[k] = op.col[idx];
                        e[k] = ok ? op.a[idx] : (T)0;
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (aq[0]) s0 += e[k] * __ldcg(&op.x[cc[k]]);
                        if (aq[1]) s1 += e[k] * __ldcg(&op.x[cc[k] + ld]);
                        if (aq[2]) s2 += e[k] * __ldcg(&op.x[cc[k] + 2 * ld]);
                    }
                }
                if (row < n) {
                    const T acc[3] = {s0, s1, s2};
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        if (!aq[q]) continue;
                        const int j = q * ld + row;
                        const T xv = __ldcg(&op.x[j]), d = op.dg[j];
                        const T r = __ldcg(&op.b[j]) - (d * xv - acc[q]);
                        if (op.kind == TAIL_RESID) op.xo[j] = r;
                        else {
                            T xn = xv + op.c2 * r / d;
                            if (op.prevMode == 1) xn += op.c1 * (xv - __ldcg(&op.xprev[j]));
                            else if (op.prevMode == 2) xn += op.c1 * xv;
                            op.xo[j] = xn;
                        }
                    }
                }
            }
        } else if (op.kind == TAIL_RESTRICT) {          // n = coarse rows, ld = fine ld, ldc = coarse ld
            for (int I = tid; I < n; I += nT) {
                T s0 = 0, s1 = 0, s2 = 0;
                for (int e = op.idxA[I]; e < op.idxA[I + 1]; e++) {
                    const int i = op.idxB[e];
                    if (aq[0]) s0 += __ldcg(&op.x[i]);
                    if (aq[1]) s1 += __ldcg(&op.x[(size_t)ld + i]);
                    if (aq[2]) s2 += __ldcg(&op.x[2 * (size_t)ld + i]);
                }
                op.xo[I] = s0; op.xo[(size_t)op.ldc + I] = s1; op.xo[2 * (size_t)op.ldc + I] = s2;
            }
        } else if (op.kind == TAIL_PROLONG) {           // x (fine, in xo) += c1 * e[parent]; e = op.x with ldc
            for (int i = tid; i < n; i += nT) {
                const int I = op.idxA[i];
#pragma unroll
                for (int q = 0; q < 3; q++) if (aq[q]) op.xo[(size_t)q * ld + i] = __ldcg(&op.xo[(size_t)q * ld + i]) + op.c1 * __ldcg(&op.x[(size_t)q * op.ldc + I]);
            }
        } else if (op.kind == TAIL_DENSE) {             // a = inverse [3][n][n]; one warp per (row, component)
            const int lane = tid & 31;
            for (int gw = tid >> 5; gw < 3 * n; gw += nT >> 5) {
                const int q = gw / n, i = gw % n;
                const T* row = op.a + ((size_t)q * n + i) * n;
                T sum = 0;
                for (int k = lane; k < n; k += 32) sum += row[k] * __ldcg(&op.b[(size_t)q * ld + k]);
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
                if (lane == 0) op.xo[(size_t)q * ld + i] = sum;
            }
        } else {                                        // TAIL_COPY
            for (int i = tid; i < n; i += nT)
#pragma unroll
                for (int q = 0; q < 3; q++) if (aq[q]) op.xo[(size_t)q * ld + i] = __ldcg(&op.x[(size_t)q * ld + i]);
        }
        cluster.sync();
    }
}

// ================================================================================================
// hierarchy
// ================================================================================================
template <class T>
struct Level {
    int n = 0, nGhost = 0, ld = 0, nSlices = 0;
    const int* slicePtr = nullptr; const int* col = nullptr; const T* a = nullptr;   // level 0 aliases the fine rows
    DevBuf<int> slicePtrB, colB; DevBuf<T> aB;
    DevBuf<T> dg;                       // 3*ld: per-component diagonal
    DevBuf<int> parent;                 // [n] -> index in the next level's vectors
    DevBuf<int> childPtr, child;        // children lists of THIS level's cells in the finer level
    DevBuf<T> b, x, x2, x3, t;          // 3*ld work vectors (x, x2, x3: the three rotating iterates of the smoother)
    bool dist = false;                  // one part per rank; ghost columns [n, n+nGhost) filled by `halo`
    S4fHaloPlan* halo = nullptr; bool ownHalo = false;
    // transition to the replicated part of the hierarchy: this (distributed) level restricts into `gsend`, which is
    // gathered to every rank as the right-hand side of the next level
    S4fGatherPlan* gather = nullptr; DevBuf<T> gsend; int gLocal = 0, gStride = 0;
    // coefficient refresh of a decomposed hierarchy: coarse column of every ghost column of this (distributed) level, the
    // first row of the next level this rank computes, and the widest SELL slice of this level's rows
    DevBuf<int> ghostParent; int rowOff = 0, maxWidth = 0;
    double nnz = 0;
    ~Level() { if (ownHalo) s4f_halo_plan_destroy(halo); s4f_gather_plan_destroy(gather); }
};

}  // namespace

struct S4fAmg {
    virtual ~S4fAmg() {}
    virtual int apply(s4fgpu_ctx* c, const double* r3, double* z3, const int* act) = 0;
    virtual int step0(s4fgpu_ctx* c, const double* r3, const int* act) = 0;   // one fine-level smoothing step alone (timing)
    virtual int refresh(s4fgpu_ctx* c) = 0;   // new fine-matrix coefficients, same aggregates
    double step0Bytes = 0;
    std::vector<int> sizes;
    std::vector<int> distributed;       // per level: 1 = one part per rank (sizes[] is then this rank's part)
    double bytesPerApply = 0;
    double setupSeconds = 0;
    double refreshSeconds = 0;
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
    // the V-cycle from level `tailLevel` down as one kernel (k_amg_tail): recorded once, after the levels are built
    int tailLevel = -1;
    bool recording = false;
    std::vector<TailOp<T>> tailHost;
    DevBuf<TailOp<T>> tailOps;

    int halo(s4fgpu_ctx* c, Level<T>& L, T* x) {
        if (!L.dist || !L.halo) return 0;
        return s4f_halo_run<T>(c, L.halo, x, L.ld, 3, L.n);
    }

    // rows of a coarse level: interior neighbours (lower, upper), then the couplings across rank boundaries (ghost columns)
    int build_level_rows(s4fgpu_ctx* c, Level<T>& L, const HostLevel& H) {
        const int n = H.n;
        L.n = n; L.nSlices = (n + 31) / 32; L.dist = H.dist;
        // ghosts: per neighbour the distinct remote cells in ascending order; send list: the distinct local cells
        std::vector<int> nbrRank, nbrCount, sendCells;
        std::vector<std::vector<int>> ghostIds(H.ifc.size());
        int nGhost = 0;
        std::vector<int> ghostBase(H.ifc.size(), 0);
        for (size_t k = 0; k < H.ifc.size(); k++) {
            const Iface& I = H.ifc[k];
            std::vector<int> g(I.remote), sc(I.cell);
            std::sort(g.begin(), g.end()); g.erase(std::unique(g.begin(), g.end()), g.end());
            std::sort(sc.begin(), sc.end()); sc.erase(std::unique(sc.begin(), sc.end()), sc.end());
            ghostIds[k] = g; ghostBase[k] = nGhost; nGhost += (int)g.size();
            // what I send to this neighbour are my distinct cells; what I receive are its distinct cells: both sides sort
            // by the owner's local index, so the orders agree (the two counts differ in general)
            nbrRank.push_back(I.rank); nbrCount.push_back((int)sc.size());
            sendCells.insert(sendCells.end(), sc.begin(), sc.end());
        }
        L.nGhost = nGhost;
        L.ld = ((n + nGhost + 31) / 32) * 32; if (L.ld == 0) L.ld = 32;
        std::vector<int> cnt(n, 0);
        const size_t F = H.own.size();
        for (size_t f = 0; f < F; f++) { cnt[H.own[f]]++; cnt[H.nei[f]]++; }
        for (const Iface& I : H.ifc) for (int cl : I.cell) cnt[cl]++;
        std::vector<long long> rowPtr(n + 1, 0);
        for (int i = 0; i < n; i++) rowPtr[i + 1] = rowPtr[i] + cnt[i];
        std::vector<int> rc(rowPtr[n]); std::vector<double> rv(rowPtr[n]);
        {
            std::vector<long long> cur(rowPtr.begin(), rowPtr.end() - 1);
            for (size_t f = 0; f < F; f++) { long long e = cur[H.nei[f]]++; rc[e] = H.own[f]; rv[e] = H.a[f]; }
            for (size_t f = 0; f < F; f++) { long long e = cur[H.own[f]]++; rc[e] = H.nei[f]; rv[e] = H.a[f]; }
            for (size_t k = 0; k < H.ifc.size(); k++) {
                const Iface& I = H.ifc[k];
                for (size_t f = 0; f < I.cell.size(); f++) {
                    const int g = (int)(std::lower_bound(ghostIds[k].begin(), ghostIds[k].end(), I.remote[f]) - ghostIds[k].begin());
                    long long e = cur[I.cell[f]]++; rc[e] = n + ghostBase[k] + g; rv[e] = I.a[f];
                }
            }
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
        L.nnz = (double)rowPtr[n];
        L.maxWidth = 0;
        for (int sl = 0; sl < L.nSlices; sl++) L.maxWidth = std::max(L.maxWidth, (sp[sl + 1] - sp[sl]) / 32);
        S4F_CHECK_CUDA(c, L.slicePtrB.upload(sp)); S4F_CHECK_CUDA(c, L.colB.upload(hc)); S4F_CHECK_CUDA(c, L.aB.upload(ha));
        L.slicePtr = L.slicePtrB.p; L.col = L.colB.p; L.a = L.aB.p;
        std::vector<T> hd(3 * (size_t)L.ld, (T)1);
        for (int q = 0; q < 3; q++) for (int i = 0; i < n; i++) hd[(size_t)q * L.ld + i] = (T)H.diag[q][i];
        S4F_CHECK_CUDA(c, L.dg.upload(hd));
        if (H.dist) {
            // receive counts = distinct remote cells per neighbour; the plan moves nbrCount values out and expects the same
            // number in: the two sides' lists are mirror images (my send list to r = r's ghost list of me)
            std::vector<int> recvCount;
            for (size_t k = 0; k < H.ifc.size(); k++) recvCount.push_back((int)ghostIds[k].size());
            int rcx = s4f_halo_plan_create_asym(c, nbrRank, nbrCount, recvCount, sendCells, 3, &L.halo); if (rcx) return rcx;
            L.ownHalo = true;
        }
        return 0;
    }
    int alloc_work(s4fgpu_ctx* c, Level<T>& L) {
        const size_t m = 3 * (size_t)L.ld;
        S4F_CHECK_CUDA(c, L.b.alloc(m)); S4F_CHECK_CUDA(c, L.x.alloc(m)); S4F_CHECK_CUDA(c, L.x2.alloc(m));
        S4F_CHECK_CUDA(c, L.x3.alloc(m)); S4F_CHECK_CUDA(c, L.t.alloc(m));
        return 0;
    }
    // parent[i] = index of the coarse cell of fine cell i in the next level's vectors; the children lists cover the coarse
    // cells [off, off+nc) that this rank's fine cells feed (all of them on replicated / serial levels)
    int set_transfer(s4fgpu_ctx* c, Level<T>& fine, Level<T>& coarse, const std::vector<int>& parent, int nc, int off) {
        S4F_CHECK_CUDA(c, fine.parent.upload(parent.empty() ? std::vector<int>(1, 0) : parent));
        std::vector<int> ptr(nc + 1, 0), ch(std::max<size_t>(parent.size(), 1), 0);
        for (size_t i = 0; i < parent.size(); i++) ptr[parent[i] - off + 1]++;
        for (int i = 0; i < nc; i++) ptr[i + 1] += ptr[i];
        std::vector<int> cur(ptr.begin(), ptr.end() - 1);
        for (size_t i = 0; i < parent.size(); i++) ch[cur[parent[i] - off]++] = (int)i;
        S4F_CHECK_CUDA(c, coarse.childPtr.upload(ptr)); S4F_CHECK_CUDA(c, coarse.child.upload(ch));
        return 0;
    }

    static int step_grid(const s4fgpu_ctx* c, const Level<T>& L) { return s4f_grid(c->numSMs, (long long)L.nSlices * 32, 4); }
    // algorithmic bytes of one application on this rank (for the roofline report): every array read / written once
    double bytes_per_apply() const {
        double tot = 0, below1 = 0;
        const double sT = sizeof(T);
        for (size_t l = 0; l + 1 < lv.size(); l++) {
            const Level<T>& L = *lv[l];
            const double n = L.n, nz = L.nnz, sB = (l == 0) ? 8.0 : sT, sO = (l == 0) ? 8.0 : sT, nc = lv[l + 1]->n;
            const double mat = nz * (4 + sT) + n * 0.125;
            double t = 0;
            t += 3 * n * (sT + sB + sT);                                             // first: diag, b in, x out
            t += mat + 3 * n * (sB + 2 * sT + sT);                                   // second pre step (previous iterate = 0)
            t += (deg > 2 ? deg - 2 : 0) * (mat + 3 * n * (sB + 3 * sT + sT));                      // further pre steps
            t += mat + 3 * n * (sB + 2 * sT + sT);                                   // residual
            t += 3 * n * sT + 3 * nc * sT + 4 * n;                                   // restrict
            t += 4 * n + 6 * n * sT + 3 * nc * sT;                                   // prolong
            t += mat + 3 * n * (sB + 2 * sT + sT);                                   // post step 0 (no previous iterate)
            t += (deg - 1) * (mat + 3 * n * (sB + 3 * sT + sT)) + (deg > 1 && l == 0 ? 3 * n * (sO - sT) : 0);   // post steps
            tot += t;
            if (l >= 1) below1 += t;
        }
        if (cycle == 2 && lv.size() > 2) {      // K-cycle: a second V-cycle from level 1, two A_1 products with their dots, r, x, two copies
            const double n1 = lv[1]->n, mat1 = lv[1]->nnz * (4 + sT) + n1 * 0.125;
            tot += below1 + 2 * (mat1 + 3 * n1 * 4 * sT) + 3 * n1 * 3 * sT + 3 * n1 * 3 * sT + 2 * 3 * n1 * 2 * sT;
        }
        return tot;
    }

    // ---- smoothing on one level ---------------------------------------------------------------
    // The iterates rotate through the level's three vectors x, x2, x3.  `Rot` tracks where the current and the previous
    // iterate live; the starting slot is chosen so that the result of the whole cycle on the level ends in L.x.
    struct Rot { T* buf[3]; int cur; int prev; };     // prev: -1 none, -2 the zero initial guess, else a slot
    Rot rot_start(Level<T>& L) const {
        Rot r; r.buf[0] = L.x.p; r.buf[1] = L.x2.p; r.buf[2] = L.x3.p;
        // pre-smoothing writes deg slots, post-smoothing deg more (W-cycle: the extra prolongation is in place)
        int s = (1 - 2 * deg) % 3; if (s < 0) s += 3;
        r.cur = s; r.prev = -1;
        return r;
    }
    template <class TB, class TO>
    int step(s4fgpu_ctx* c, Level<T>& L, const TB* b, int ldb, const T* xin, const T* xprev, int prevMode, TO* xout, int ldo, double c1, double c2) {
        const int grid = step_grid(c, L);
        if (recording) {
            if constexpr (std::is_same<TB, T>::value && std::is_same<TO, T>::value) {
                TailOp<T> op{}; op.kind = TAIL_STEP; op.n = L.n; op.ld = L.ld; op.nSlices = L.nSlices; op.prevMode = prevMode;
                op.slicePtr = L.slicePtr; op.col = L.col; op.a = L.a; op.dg = L.dg.p; op.b = b; op.x = xin; op.xprev = xprev; op.xo = xout;
                op.c1 = (T)c1; op.c2 = (T)c2;
                tailHost.push_back(op);
            }
            return 0;
        }
        int rc = halo(c, L, const_cast<T*>(xin)); if (rc) return rc;
        k_amg_step<T, TB, TO, 0><<<grid, S4F_AMG_BLOCK, 0, c->stream>>>(L.slicePtr, L.col, L.a, L.dg.p, b, xin, xprev, xout, L.n, L.ld, ldb, ldo,
                                                                     L.nSlices, (T)c1, (T)c2, prevMode, act);
        c->launches++;
        return 0;
    }
    // Chebyshev-Jacobi of degree `deg`; fromZero: x0 = 0.  The result is R.buf[R.cur], or, when `out` is given (level 0),
    // the last step writes the fp64 output directly.
    template <class TB>
    int smooth(s4fgpu_ctx* c, Level<T>& L, const TB* b, int ldb, bool fromZero, Rot& R, double* out, int ldo) {
        const double sigma = theta / delta;
        double rho = 1.0 / sigma;
        int k0 = 0, rc;
        if (fromZero) {
            const int grid = s4f_grid(c->numSMs, L.n);
            if (recording) {
                if constexpr (std::is_same<TB, T>::value) {
                    TailOp<T> op{}; op.kind = TAIL_FIRST; op.n = L.n; op.ld = L.ld; op.dg = L.dg.p; op.b = b; op.xo = R.buf[R.cur]; op.c2 = (T)(1.0 / theta);
                    tailHost.push_back(op);
                }
            } else {
                k_amg_first<T, TB><<<grid, S4F_BLOCK, 0, c->stream>>>(L.dg.p, b, R.buf[R.cur], L.n, L.ld, ldb, (T)(1.0 / theta), act);
                c->launches++;
            }
            R.prev = -2;
            k0 = 1;
        } else R.prev = -1;
        for (int k = k0; k < deg; k++) {
            double c1, c2;
            if (k == 0) { c1 = 0.0; c2 = 1.0 / theta; }
            else { const double rhon = 1.0 / (2.0 * sigma - rho); c1 = rhon * rho; c2 = 2.0 * rhon / delta; rho = rhon; }
            const bool last = (k == deg - 1);
            const T* xprev = R.prev >= 0 ? R.buf[R.prev] : nullptr;
            const int prevMode = (k == 0) ? 0 : (R.prev >= 0 ? 1 : (R.prev == -2 ? 2 : 0));
            if (last && out) return step<TB, double>(c, L, b, ldb, R.buf[R.cur], xprev, prevMode, out, ldo, c1, c2);
            const int nxt = (R.cur + 1) % 3;          // never the previous iterate's slot (that is cur - 1)
            if ((rc = step<TB, T>(c, L, b, ldb, R.buf[R.cur], xprev, prevMode, R.buf[nxt], L.ld, c1, c2))) return rc;
            R.prev = R.cur; R.cur = nxt;
        }
        return 0;
    }

    // t = b - A x on level l, restricted into the right-hand side of level l+1
    template <class TB>
    int residual_restrict(s4fgpu_ctx* c, size_t l, const TB* b, int ldb, T* x) {
        Level<T>& L = *lv[l];
        Level<T>& C = *lv[l + 1];
        const int grid = step_grid(c, L);
        if (recording) {
            if constexpr (std::is_same<TB, T>::value) {
                TailOp<T> op{}; op.kind = TAIL_RESID; op.n = L.n; op.ld = L.ld; op.nSlices = L.nSlices;
                op.slicePtr = L.slicePtr; op.col = L.col; op.a = L.a; op.dg = L.dg.p; op.b = b; op.x = x; op.xo = L.t.p;
                tailHost.push_back(op);
                TailOp<T> rs{}; rs.kind = TAIL_RESTRICT; rs.n = C.n; rs.ld = L.ld; rs.ldc = C.ld; rs.idxA = C.childPtr.p; rs.idxB = C.child.p;
                rs.x = L.t.p; rs.xo = C.b.p;
                tailHost.push_back(rs);
            }
            return 0;
        }
        int rc = halo(c, L, x); if (rc) return rc;
        k_amg_step<T, TB, T, 1><<<grid, S4F_AMG_BLOCK, 0, c->stream>>>(L.slicePtr, L.col, L.a, L.dg.p, b, x, nullptr, L.t.p, L.n, L.ld, ldb, L.ld,
                                                                       L.nSlices, (T)0, (T)0, 0, act);
        c->launches++;
        if (!L.gather) {
            k_amg_restrict<T><<<(C.n + 127) / 128, 128, 0, c->stream>>>(C.childPtr.p, C.child.p, L.t.p, C.b.p, C.n, L.ld, C.ld, act);
            c->launches++;
            return 0;
        }
        // this rank's aggregates -> packed [3][gStride]; pushed to every rank over NVLink into the replicated right-hand side
        k_amg_restrict<T><<<(L.gLocal + 127) / 128, 128, 0, c->stream>>>(C.childPtr.p, C.child.p, L.t.p, L.gsend.p, L.gLocal, L.ld, L.gStride, act);
        c->launches++;
        return s4f_gather_run<T>(c, L.gather, L.gsend.p, L.gStride, C.b.p, C.ld, act);
    }

    int copy(s4fgpu_ctx* c, Level<T>& L, const T* src, T* dst) {
        if (recording) {
            TailOp<T> op{}; op.kind = TAIL_COPY; op.n = L.n; op.ld = L.ld; op.x = src; op.xo = dst;
            tailHost.push_back(op);
            return 0;
        }
        k_amg_copy<T><<<s4f_grid(c->numSMs, L.n), S4F_BLOCK, 0, c->stream>>>(src, dst, L.n, L.ld, act);
        c->launches++;
        return 0;
    }

    template <class TB>
    int cycle_level(s4fgpu_ctx* c, size_t l, const TB* b, int ldb, double* out, int ldo) {
        Level<T>& L = *lv[l];
        if (!recording && (int)l == tailLevel && !out) {
            if constexpr (std::is_same<TB, T>::value) {
                if (b == L.b.p) {       // the recorded program reads this level's own right-hand side
                    k_amg_tail<T><<<S4F_TAIL_CTAS, S4F_TAIL_THREADS, 0, c->stream>>>(tailOps.p, (int)tailHost.size(), act);
                    c->launches++;
                    return 0;
                }
            }
        }
        if (l + 1 == lv.size()) {     // coarsest: dense inverse
            if (recording) {
                if constexpr (std::is_same<TB, T>::value) {
                    TailOp<T> op{}; op.kind = TAIL_DENSE; op.n = L.n; op.ld = L.ld; op.a = denseInv.p; op.b = b; op.xo = L.x.p;
                    tailHost.push_back(op);
                }
                return 0;
            }
            const int warps = 3 * L.n, blocks = (warps * 32 + 255) / 256;
            if (out) k_amg_dense<T, TB, double><<<blocks, 256, 0, c->stream>>>(denseInv.p, b, out, L.n, ldb, ldo);
            else k_amg_dense<T, TB, T><<<blocks, 256, 0, c->stream>>>(denseInv.p, b, L.x.p, L.n, ldb, L.ld);
            c->launches++;
            return 0;
        }
        Level<T>& C = *lv[l + 1];
        int rc;
        Rot R = rot_start(L);
        if ((rc = smooth<TB>(c, L, b, ldb, true, R, nullptr, 0))) return rc;               // pre-smoothing from zero
        T* x = R.buf[R.cur];
        if ((rc = residual_restrict<TB>(c, l, b, ldb, x))) return rc;
        const bool kcycle = (cycle == 2 && l == 0 && lv.size() > 2);
        if (kcycle) { rc = kcycle_level1(c); if (rc) return rc; }
        else { rc = cycle_level<T>(c, l + 1, C.b.p, C.ld, nullptr, 0); if (rc) return rc; }
        if (recording) {
            TailOp<T> op{}; op.kind = TAIL_PROLONG; op.n = L.n; op.ld = L.ld; op.ldc = C.ld; op.idxA = L.parent.p; op.x = C.x.p; op.xo = x; op.c1 = (T)omega;
            tailHost.push_back(op);
        } else {
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
        if ((rc = smooth<TB>(c, L, b, ldb, false, R, out, ldo))) return rc;                // post-smoothing
        if (!out && R.buf[R.cur] != L.x.p) copy(c, L, R.buf[R.cur], L.x.p);   // callers read the level result from L.x (deg 1 only)
        return 0;
    }

    // two FCG steps on A_1 x = C.b preconditioned by the V-cycle from level 1; result in lv[1]->x
    int kcycle_level1(s4fgpu_ctx* c) {
        Level<T>& C = *lv[1];
        const int grid = step_grid(c, C), gv = (C.n + 255) / 256;
        // a distributed level 1 all-reduces its dot products; on a replicated one every rank computes the same sums
        const RedCtx red = C.dist ? c->red() : RedCtx{c->partials.p, c->ticket.p, nullptr};
        int rc = cycle_level<T>(c, 1, C.b.p, C.ld, nullptr, 0); if (rc) return rc;                       // c1 = V(b)
        copy(c, C, C.x.p, kc1.p);
        if ((rc = halo(c, C, kc1.p))) return rc;
        k_kc_amul<T, 1><<<grid, S4F_AMG_BLOCK, 0, c->stream>>>(C.slicePtr, C.col, C.a, C.dg.p, kc1.p, C.b.p, nullptr, kv1.p, C.n, C.ld, C.nSlices, kS.p, act, red);
        k_kc_resid<T><<<gv, 256, 0, c->stream>>>(C.b.p, kv1.p, kr.p, C.n, C.ld, kS.p, act);
        c->launches += 2;
        rc = cycle_level<T>(c, 1, kr.p, C.ld, nullptr, 0); if (rc) return rc;                            // c2 = V(r)
        copy(c, C, C.x.p, kc2.p);
        if ((rc = halo(c, C, kc2.p))) return rc;
        k_kc_amul<T, 2><<<grid, S4F_AMG_BLOCK, 0, c->stream>>>(C.slicePtr, C.col, C.a, C.dg.p, kc2.p, kr.p, kv1.p, C.t.p, C.n, C.ld, C.nSlices, kS.p, act, red);
        k_kc_final<T><<<gv, 256, 0, c->stream>>>(kc1.p, kc2.p, C.x.p, C.n, C.ld, kS.p, act);
        c->launches += 2;
        return 0;
    }

    // The fine matrix changed its coefficients but not its graph (mesh motion, a new time step size): keep the aggregates,
    // re-sum every coarse level on the device, invert the coarsest matrix again.  Decomposed: couplings across processor
    // patches go to the other rank's aggregate (ghostParent); where the hierarchy is gathered, every rank re-sums the rows of
    // its own aggregates and the rows are exchanged.
    int refresh(s4fgpu_ctx* c) override {
        Level<T>& L0 = *lv[0];
        if (sizeof(T) != sizeof(double)) {
            k_amg_convert<T><<<(unsigned)((c->nEntries + 255) / 256), 256, 0, c->stream>>>(c->eA.p, L0.aB.p, c->nEntries);
            c->launches++;
        }
        k_amg_diag<T><<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->diagC.p, L0.dg.p, c->N, c->ld, L0.ld);
        c->launches++;
        DevBuf<int> fail;
        S4F_CHECK_CUDA(c, fail.alloc(1));
        for (size_t l = 1; l < lv.size(); l++) {
            Level<T>& Fn = *lv[l - 1];
            Level<T>& C = *lv[l];
            const int nRows = Fn.gather ? Fn.gLocal : C.n, rowOff = Fn.gather ? Fn.rowOff : 0;
            if (Fn.dist && Fn.nGhost > 0 && Fn.ghostParent.n == 0) { c->err = "GAMG refresh: hierarchy without interface parents"; return 1; }
            if (nRows > 0) {
                k_amg_galerkin<T><<<(nRows + 127) / 128, 128, 0, c->stream>>>(Fn.slicePtr, Fn.col, Fn.a, Fn.dg.p, Fn.parent.p, Fn.n, Fn.ld, C.slicePtrB.p,
                                                                           C.colB.p, C.aB.p, C.dg.p, C.childPtr.p, C.child.p, nRows, C.ld, fail.p,
                                                                           Fn.dist && Fn.nGhost > 0 ? Fn.ghostParent.p : nullptr, Fn.nGhost, rowOff);
                c->launches++;
            }
            S4F_CHECK_CUDA(c, cudaGetLastError());
            if (Fn.gather) {        // rows of all ranks -> every rank (C.t is free here: 3 * ld work values)
                const int gp = (Fn.gLocal + 127) / 128, gu = (C.n + 127) / 128;
                for (int k0 = 0; k0 < C.maxWidth; k0 += 3) {
                    if (Fn.gLocal > 0) k_amg_rows_pack<T><<<gp, 128, 0, c->stream>>>(C.slicePtrB.p, C.aB.p, rowOff, Fn.gLocal, k0, Fn.gsend.p, Fn.gStride);
                    int rg = s4f_gather_run<T>(c, Fn.gather, Fn.gsend.p, Fn.gStride, C.t.p, C.ld, c->ones3.p); if (rg) return rg;
                    k_amg_rows_unpack<T><<<gu, 128, 0, c->stream>>>(C.slicePtrB.p, C.aB.p, C.n, k0, C.t.p, C.ld);
                    c->launches += 2;
                }
                if (Fn.gLocal > 0) k_amg_diag_pack<T><<<gp, 128, 0, c->stream>>>(C.dg.p, rowOff, Fn.gLocal, C.ld, Fn.gsend.p, Fn.gStride);
                int rg = s4f_gather_run<T>(c, Fn.gather, Fn.gsend.p, Fn.gStride, C.dg.p, C.ld, c->ones3.p); if (rg) return rg;
                c->launches++;
                S4F_CHECK_CUDA(c, cudaGetLastError());
            }
        }
        int hf = 0;
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(&hf, fail.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
        if (hf) { c->err = "GAMG refresh: coarse rows wider than expected"; return 1; }
        return invert_coarsest(c);
    }

    // dense inverse of the coarsest matrix from its device rows (<= 512 cells; host Cholesky)
    // choose the first level from which everything is small and local, and record its V-cycle (called when the levels, their
    // work vectors and the coarsest inverse exist; pointers stay valid across coefficient refreshes)
    int record_tail(s4fgpu_ctx* c) {
        tailLevel = -1; tailHost.clear();
        // OFF by default: measured at 8 M cells (profiles/r2_amg_tail_ab.log), ms per outer iteration / launches per outer iteration:
        // off 11.36 / 297; levels <= 2048 rows fused 11.54 / 242; <= 20000 rows 11.91 / 190; <= 200000 rows 13.89 / 134.  Inside
        // the PCG graph a small kernel costs ~1.5 us and spreads over 148 SMs; one cluster of 8 SMs stepping through the same
        // operations is latency-bound and slower, barrier or not.  S4F_AMG_TAIL_MAX=<rows> turns it on.
        long long maxRows = 0;
        if (const char* e = getenv("S4F_AMG_TAIL_MAX")) maxRows = atoll(e);
        if (cycle == 1 || maxRows <= 0 || lv.size() < 3) return 0;
        int lt = -1;
        for (int l = (int)lv.size() - 1; l >= 2; l--) {
            const Level<T>& L = *lv[l];
            if (L.dist || L.gather || L.n > maxRows) break;
            lt = l;
        }
        if (lt < 0 || lt + 1 >= (int)lv.size()) return 0;          // nothing to fuse (the coarsest level alone is one launch already)
        recording = true;
        int rc = cycle_level<T>(c, (size_t)lt, lv[lt]->b.p, lv[lt]->ld, nullptr, 0);
        recording = false;
        if (rc) return rc;
        S4F_CHECK_CUDA(c, tailOps.upload(tailHost));
        tailLevel = lt;
        return 0;
    }

    int invert_coarsest(s4fgpu_ctx* c) {
        Level<T>& LC = *lv.back();
        const int n = LC.n;
        std::vector<int> sp(LC.nSlices + 1);
        // (a mesh of <= 512 cells has the fine level as its only, dense level: its rows are the fine rows)
        S4F_CHECK_CUDA(c, cudaMemcpy(sp.data(), LC.slicePtr, sp.size() * sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<int> hc(std::max(sp.back(), 1)); std::vector<T> ha(std::max(sp.back(), 1)), hd(3 * (size_t)LC.ld);
        S4F_CHECK_CUDA(c, cudaMemcpy(hc.data(), LC.col, sp.back() * sizeof(int), cudaMemcpyDeviceToHost));
        S4F_CHECK_CUDA(c, cudaMemcpy(ha.data(), LC.a, sp.back() * sizeof(T), cudaMemcpyDeviceToHost));
        S4F_CHECK_CUDA(c, cudaMemcpy(hd.data(), LC.dg.p, hd.size() * sizeof(T), cudaMemcpyDeviceToHost));
        HostLevel HC; HC.n = n;
        for (int q = 0; q < 3; q++) { HC.diag[q].resize(n); for (int i = 0; i < n; i++) HC.diag[q][i] = (double)hd[(size_t)q * LC.ld + i]; }
        for (int i = 0; i < n; i++) {
            const int sI = i >> 5, lane = i & 31, w = (sp[sI + 1] - sp[sI]) / 32;
            for (int k = 0; k < w; k++) {
                const int j = hc[sp[sI] + 32 * k + lane]; const double a = (double)ha[sp[sI] + 32 * k + lane];
                if (j > i && j < n && a != 0.0) { HC.own.push_back(i); HC.nei.push_back(j); HC.a.push_back(a); }
            }
        }
        std::vector<T> inv(3 * (size_t)n * n);
        for (int q = 0; q < 3; q++) {
            std::vector<double> iq;
            if (!dense_inverse(HC, q, iq)) { c->err = "GAMG refresh: coarsest-level matrix is not positive definite"; return 1; }
            for (size_t i = 0; i < iq.size(); i++) inv[(size_t)q * n * n + i] = (T)iq[i];
        }
        S4F_CHECK_CUDA(c, denseInv.upload(inv));
        return 0;
    }

    // the fine-level Chebyshev-Jacobi step kernel on its own (no halo): the dominant kernel of the V-cycle
    int step0(s4fgpu_ctx* c, const double* r3, const int* actIn) override {
        act = actIn;
        Level<T>& L = *lv[0];
        if (lv.size() < 2) return 0;
        const int grid = step_grid(c, L);
        k_amg_step<T, double, T, 0><<<grid, S4F_AMG_BLOCK, 0, c->stream>>>(L.slicePtr, L.col, L.a, L.dg.p, r3, L.x.p, L.x3.p, L.x2.p, L.n, L.ld,
                                                                            c->ld, L.ld, L.nSlices, (T)0.3, (T)0.5, 1, act);
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

template <class T>
int build(s4fgpu_ctx* c, std::vector<HostLevel>& H) {
    auto* A = new Hierarchy<T>();
    std::unique_ptr<S4fAmg> guard(A);
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
            L.n = c->N; L.nGhost = c->G; L.ld = c->ld; L.nSlices = c->nSlices; L.dist = H[0].dist; L.halo = c->halo0; L.nnz = (double)c->nnzOff;
            L.slicePtr = c->slicePtr.p; L.col = c->col.p;
            if (sizeof(T) == sizeof(double)) L.a = reinterpret_cast<const T*>(c->eA.p);
            else {
                S4F_CHECK_CUDA(c, L.aB.alloc((size_t)c->nEntries, false));
                k_amg_convert<T><<<(unsigned)((c->nEntries + 255) / 256), 256, 0, c->stream>>>(c->eA.p, L.aB.p, c->nEntries);
                L.a = L.aB.p;
            }
            S4F_CHECK_CUDA(c, L.dg.alloc(3 * (size_t)L.ld));
            k_amg_diag<T><<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->diagC.p, L.dg.p, c->N, c->ld, L.ld);
            c->launches += 2;
        } else {
            if ((rc = A->build_level_rows(c, L, H[l]))) return rc;
        }
        if ((rc = A->alloc_work(c, L))) return rc;
        if (l > 0) {
            Level<T>& Fn = *A->lv[l - 1];
            const HostLevel& HF = H[l - 1];
            if ((rc = A->set_transfer(c, Fn, L, HF.parent, HF.childCnt, HF.childOff))) return rc;
            Fn.rowOff = HF.childOff;
            if (HF.dist && Fn.nGhost > 0) {     // coefficient refresh: the coarse column behind every ghost column of the finer level
                const HostLevel& HC2 = H[l];
                std::vector<std::vector<int>> gC(HC2.ifc.size()); std::vector<int> baseC(HC2.ifc.size(), 0);
                if (HC2.dist) {
                    int nG = 0;
                    for (size_t k = 0; k < HC2.ifc.size(); k++) {
                        std::vector<int> g(HC2.ifc[k].remote);
                        std::sort(g.begin(), g.end()); g.erase(std::unique(g.begin(), g.end()), g.end());
                        gC[k] = g; baseC[k] = nG; nG += (int)g.size();
                    }
                }
                std::vector<int> gp(Fn.nGhost, -1);
                int baseF = 0;
                for (size_t k = 0; k < HF.ifc.size(); k++) {
                    const Iface& I = HF.ifc[k];
                    if (I.remoteParent.size() != I.remote.size()) { c->err = "GAMG: interface without the remote aggregates"; return 1; }
                    std::vector<int> gF;
                    if (l - 1 > 0) { gF = I.remote; std::sort(gF.begin(), gF.end()); gF.erase(std::unique(gF.begin(), gF.end()), gF.end()); }
                    for (size_t f = 0; f < I.remote.size(); f++) {
                        // level 0: one ghost per processor face, in patch order; coarser levels: the distinct remote cells, ascending
                        const int g = baseF + (l - 1 == 0 ? (int)f : (int)(std::lower_bound(gF.begin(), gF.end(), I.remote[f]) - gF.begin()));
                        const int R = I.remoteParent[f];
                        gp[g] = HC2.dist ? HC2.n + baseC[k] + (int)(std::lower_bound(gC[k].begin(), gC[k].end(), R) - gC[k].begin())
                                         : HF.rankOff[I.rank] + R;
                    }
                    baseF += (l - 1 == 0) ? (int)I.remote.size() : (int)gF.size();
                }
                for (int v : gp) if (v < 0) { c->err = "GAMG: a ghost column without a coarse column"; return 1; }
                S4F_CHECK_CUDA(c, Fn.ghostParent.upload(gp));
            }
            if (HF.dist && !H[l].dist) {        // the distributed part ends here: gather to all ranks
                std::vector<int> cnt(c->nRanks);
                if ((rc = s4f_allgather_host(c, &HF.childCnt, sizeof(int), cnt.data()))) return rc;
                int mx = 1; for (int v : cnt) mx = std::max(mx, v);
                Fn.gLocal = HF.childCnt; Fn.gStride = ((mx + 31) / 32) * 32;
                S4F_CHECK_CUDA(c, Fn.gsend.alloc(3 * (size_t)Fn.gStride));
                if ((rc = s4f_gather_plan_create(c, cnt, &Fn.gather))) return rc;
            }
        }
        A->sizes.push_back(H[l].n);
        A->distributed.push_back(H[l].dist ? 1 : 0);
    }
    if (A->lv.size() > 1) {      // K-cycle work vectors on level 1
        const size_t m = 3 * (size_t)A->lv[1]->ld;
        S4F_CHECK_CUDA(c, A->kc1.alloc(m)); S4F_CHECK_CUDA(c, A->kv1.alloc(m)); S4F_CHECK_CUDA(c, A->kr.alloc(m)); S4F_CHECK_CUDA(c, A->kc2.alloc(m));
        S4F_CHECK_CUDA(c, A->kS.alloc(1));
    }
    // dense inverse on the coarsest level
    const HostLevel& HC = H.back();
    if (HC.dist) { c->err = "GAMG: the coarsest level must be replicated"; return 1; }
    std::vector<T> inv(3 * (size_t)HC.n * HC.n);
    for (int q = 0; q < 3; q++) {
        std::vector<double> iq;
        if (!dense_inverse(HC, q, iq)) { c->err = "GAMG: coarsest-level matrix is not positive definite"; return 1; }
        for (size_t i = 0; i < iq.size(); i++) inv[(size_t)q * HC.n * HC.n + i] = (T)iq[i];
    }
    S4F_CHECK_CUDA(c, A->denseInv.upload(inv));
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    { int rt = A->record_tail(c); if (rt) return rt; }
    A->bytesPerApply = A->bytes_per_apply();
    A->step0Bytes = A->lv[0]->nnz * (4 + sizeof(T)) + 0.125 * c->N + 3.0 * c->N * (8 + 3 * sizeof(T) + sizeof(T));   // col,a | b(fp64) x xprev diag in, x' out
    c->amg = guard.release();
    return 0;
}

// the hierarchy from levels built on the device (s4f_amg_setup.cu): rows are adopted (fp64) or converted (fp32)
template <class T>
int build_from_device(s4fgpu_ctx* c, std::vector<std::unique_ptr<AmgDevLevel>>& D) {
    auto* A = new Hierarchy<T>();
    std::unique_ptr<S4fAmg> guard(A);
    A->deg = c->ctl.gamgSmootherDegree > 0 ? c->ctl.gamgSmootherDegree : 3;
    A->cycle = c->ctl.gamgCycle;
    A->omega = c->ctl.gamgOverCorrection > 0 ? c->ctl.gamgOverCorrection : 2.2;
    A->omegaK = A->omega;
    if (const char* e = getenv("S4F_GAMG_OMEGA_K")) A->omegaK = atof(e);
    const double ratio = c->ctl.gamgSmootherRatio > 0 ? c->ctl.gamgSmootherRatio : 0.3;
    const double lmax = 2.0, lmin = ratio * lmax;
    A->theta = 0.5 * (lmax + lmin); A->delta = 0.5 * (lmax - lmin);
    for (size_t l = 0; l < D.size(); l++) {
        A->lv.emplace_back(new Level<T>());
        Level<T>& L = *A->lv.back();
        AmgDevLevel& S = *D[l];
        int rc;
        L.n = S.n; L.nGhost = 0; L.ld = S.ld; L.nSlices = S.nSlices; L.dist = false; L.nnz = (double)S.nnz;
        if (l == 0) {
            L.slicePtr = c->slicePtr.p; L.col = c->col.p;
            if (sizeof(T) == sizeof(double)) L.a = reinterpret_cast<const T*>(c->eA.p);
            else {
                S4F_CHECK_CUDA(c, L.aB.alloc((size_t)c->nEntries, false));
                k_amg_convert<T><<<(unsigned)((c->nEntries + 255) / 256), 256, 0, c->stream>>>(c->eA.p, L.aB.p, c->nEntries);
                L.a = L.aB.p;
            }
            S4F_CHECK_CUDA(c, L.dg.alloc(3 * (size_t)L.ld));
            k_amg_diag<T><<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->diagC.p, L.dg.p, c->N, c->ld, L.ld);
            c->launches += 2;
        } else {
            L.slicePtrB.swap(S.slicePtr); L.colB.swap(S.col);
            const size_t nE = S.a.n;
            if (sizeof(T) == sizeof(double)) {
                reinterpret_cast<DevBuf<double>&>(L.aB).swap(S.a);
                reinterpret_cast<DevBuf<double>&>(L.dg).swap(S.dg);
            } else {
                S4F_CHECK_CUDA(c, L.aB.alloc(nE, false)); S4F_CHECK_CUDA(c, L.dg.alloc(3 * (size_t)L.ld, false));
                k_amg_convert<T><<<(unsigned)((nE + 255) / 256), 256, 0, c->stream>>>(S.a.p, L.aB.p, (long long)nE);
                k_amg_convert<T><<<(unsigned)((3 * (size_t)L.ld + 255) / 256), 256, 0, c->stream>>>(S.dg.p, L.dg.p, 3 * (long long)L.ld);
                c->launches += 2;
                S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
            }
            L.slicePtr = L.slicePtrB.p; L.col = L.colB.p; L.a = L.aB.p;
            L.childPtr.swap(S.childPtr); L.child.swap(S.child);
        }
        if (S.parent.n) L.parent.swap(S.parent);
        if ((rc = A->alloc_work(c, L))) return rc;
        A->sizes.push_back(L.n);
        A->distributed.push_back(0);
    }
    if (A->lv.size() > 1) {
        const size_t m = 3 * (size_t)A->lv[1]->ld;
        S4F_CHECK_CUDA(c, A->kc1.alloc(m)); S4F_CHECK_CUDA(c, A->kv1.alloc(m)); S4F_CHECK_CUDA(c, A->kr.alloc(m)); S4F_CHECK_CUDA(c, A->kc2.alloc(m));
        S4F_CHECK_CUDA(c, A->kS.alloc(1));
    }
    if (A->lv.back()->n > 4096) { c->err = "GAMG: agglomeration stalled above the size of the dense coarsest solve"; return 1; }
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    int rc = A->invert_coarsest(c); if (rc) return rc;
    if ((rc = A->record_tail(c))) return rc;
    A->bytesPerApply = A->bytes_per_apply();
    A->step0Bytes = A->lv[0]->nnz * (4 + sizeof(T)) + 0.125 * c->N + 3.0 * c->N * (8 + 3 * sizeof(T) + sizeof(T));
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
    if (c->amg) c->graphSerial++;           // captured solves hold pointers into the hierarchy
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

__global__ void k_gather_entries(const int* __restrict__ idx, const double* __restrict__ eA, double* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = eA[idx[i]];
}

// the fine level's couplings across processor patches: cell = faceCells (patch order, which both sides share), remote =
// the neighbour's cell behind the same face, a = the assembled face coefficient
int fine_interfaces(s4fgpu_ctx* c, HostLevel& L0) {
    const int G = c->G;
    std::vector<double> aPf(std::max(G, 1), 0.0);
    if (G > 0) {
        DevBuf<double> d; S4F_CHECK_CUDA(c, d.alloc(G));
        k_gather_entries<<<(G + 255) / 256, 256, 0, c->stream>>>(c->procEntry.p, c->eA.p, d.p, G);
        c->launches++;
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(aPf.data(), d.p, G * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    std::vector<int> nbrRank; std::vector<std::vector<int>> send, recv;
    for (const auto& nb : c->nbrs) {
        Iface I; I.rank = nb.rank;
        for (int i = 0; i < nb.count; i++) { I.cell.push_back(c->faceCells[c->pStart[nb.patch] + i]); I.a.push_back(aPf[nb.sendOff + i]); }
        nbrRank.push_back(nb.rank); send.push_back(I.cell);
        L0.ifc.push_back(std::move(I));
    }
    int rc = s4f_exchange_nbr_ints(c, nbrRank, send, recv); if (rc) return rc;
    for (size_t k = 0; k < L0.ifc.size(); k++) L0.ifc[k].remote = recv[k];
    return 0;
}

// Coarsen one distributed level: every rank agglomerates its own cells (couplings across ranks never join an aggregate),
// learns the aggregates behind its interface faces and sums the interface coefficients per (own aggregate, remote
// aggregate) pair, in an order both sides reproduce (sorted by the pair as seen from the lower rank).
int coarsen_distributed(s4fgpu_ctx* c, HostLevel& L, HostLevel& C) {
    std::vector<int> parentLoc;
    const int nc = coarsen3(L, 0, parentLoc, C);
    C.dist = true;
    std::vector<int> nbrRank; std::vector<std::vector<int>> send, recv;
    for (const Iface& I : L.ifc) {
        nbrRank.push_back(I.rank);
        std::vector<int> s(I.cell.size());
        for (size_t f = 0; f < I.cell.size(); f++) s[f] = parentLoc[I.cell[f]];
        send.push_back(std::move(s));
    }
    int rc = s4f_exchange_nbr_ints(c, nbrRank, send, recv); if (rc) return rc;
    for (size_t k = 0; k < L.ifc.size(); k++) L.ifc[k].remoteParent = recv[k];
    C.ifc.clear();
    for (size_t k = 0; k < L.ifc.size(); k++) {
        const Iface& I = L.ifc[k];
        const bool lower = c->rank < I.rank;
        struct E { int p, r; double a; };
        std::vector<E> es(I.cell.size());
        for (size_t f = 0; f < I.cell.size(); f++) es[f] = E{send[k][f], recv[k][f], I.a[f]};
        std::stable_sort(es.begin(), es.end(), [lower](const E& x, const E& y) {
            const int x0 = lower ? x.p : x.r, x1 = lower ? x.r : x.p, y0 = lower ? y.p : y.r, y1 = lower ? y.r : y.p;
            return x0 != y0 ? x0 < y0 : x1 < y1;
        });
        Iface J; J.rank = I.rank;
        for (size_t f = 0; f < es.size();) {
            size_t m = f; double s = 0;
            while (m < es.size() && es[m].p == es[f].p && es[m].r == es[f].r) s += es[m++].a;
            J.cell.push_back(es[f].p); J.remote.push_back(es[f].r); J.a.push_back(s);
            f = m;
        }
        C.ifc.push_back(std::move(J));
    }
    L.parent = parentLoc; L.childOff = 0; L.childCnt = nc;
    return 0;
}

// Gather a distributed level to every rank: global cell = offset of the owning rank + local cell; interior faces of all
// ranks, then every interface face once (from its lower rank).  The fine level's parents become global indices.
int gather_level(s4fgpu_ctx* c, HostLevel& fine, const HostLevel& part, HostLevel& glob) {
    const int R = c->nRanks, me = c->rank;
    std::vector<int> cnt(R), off(R + 1, 0);
    int rc = s4f_allgather_host(c, &part.n, sizeof(int), cnt.data()); if (rc) return rc;
    for (int r = 0; r < R; r++) off[r + 1] = off[r] + cnt[r];
    std::vector<int> fi; std::vector<double> fd;
    for (size_t f = 0; f < part.own.size(); f++) { fi.push_back(off[me] + part.own[f]); fi.push_back(off[me] + part.nei[f]); fd.push_back(part.a[f]); }
    for (const Iface& I : part.ifc)
        if (me < I.rank)
            for (size_t f = 0; f < I.cell.size(); f++) { fi.push_back(off[me] + I.cell[f]); fi.push_back(off[I.rank] + I.remote[f]); fd.push_back(I.a[f]); }
    const size_t nF = fd.size();
    for (int q = 0; q < 3; q++) fd.insert(fd.end(), part.diag[q].begin(), part.diag[q].end());
    std::vector<std::vector<char>> ri, rd;
    if ((rc = s4f_allgatherv_host(c, fi.data(), fi.size() * sizeof(int), ri))) return rc;
    if ((rc = s4f_allgatherv_host(c, fd.data(), fd.size() * sizeof(double), rd))) return rc;
    (void)nF;
    glob = HostLevel(); glob.n = off[R]; glob.dist = false;
    for (int q = 0; q < 3; q++) glob.diag[q].assign(glob.n, 0.0);
    for (int r = 0; r < R; r++) {
        const int* pi = reinterpret_cast<const int*>(ri[r].data());
        const double* pd = reinterpret_cast<const double*>(rd[r].data());
        const size_t nf = ri[r].size() / (2 * sizeof(int));
        for (size_t f = 0; f < nf; f++) { glob.own.push_back(pi[2 * f]); glob.nei.push_back(pi[2 * f + 1]); glob.a.push_back(pd[f]); }
        for (int q = 0; q < 3; q++) for (int i = 0; i < cnt[r]; i++) glob.diag[q][off[r] + i] = pd[nf + (size_t)q * cnt[r] + i];
    }
    for (int& p : fine.parent) p += off[me];
    fine.childOff = off[me]; fine.childCnt = cnt[me]; fine.rankOff = off;
    return 0;
}

}  // namespace

// Build the hierarchy from the assembled fine matrix (upper() and the per-component diagonals).
int s4f_amg_setup(s4fgpu_ctx* c) {
    s4f_amg_destroy(c);
    const auto t0 = std::chrono::steady_clock::now();
    // single rank: agglomeration and Galerkin products on the device (s4f_amg_setup.cu); S4F_AMG_HOST_SETUP=1 keeps the
    // sequential host agglomeration (the one decomposed runs use per rank), for A/B comparison in the tests
    if (c->nRanks == 1 && !getenv("S4F_AMG_HOST_SETUP")) {
        std::vector<std::unique_ptr<AmgDevLevel>> D;
        int rcd = s4f_amg_device_levels(c, D, 512, 3);
        const double tAgg = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!rcd) rcd = c->ctl.gamgSinglePrecision ? build_from_device<float>(c, D) : build_from_device<double>(c, D);
        if (getenv("S4F_AMG_TIMING"))
            fprintf(stderr, "libs4fgpu: GAMG set-up: agglomeration + Galerkin %.3f s, level build + coarsest inverse %.3f s\n", tAgg,
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() - tAgg);
        if (!rcd) {
            c->amg->setupSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            c->graphSerial++;
            return 0;
        }
        fprintf(stderr, "libs4fgpu: GAMG device set-up not possible (%s); agglomerating on the host\n", c->err.c_str());
        c->err.clear(); cudaGetLastError();
        s4f_amg_destroy(c);
    }
    const int N = c->N, F = c->F;
    std::vector<HostLevel> H(1);
    int rc;
    const bool timing = getenv("S4F_AMG_TIMING") != nullptr;
    auto lap = [&](const char* what) {
        if (timing && c->rank == 0) fprintf(stderr, "libs4fgpu: GAMG host set-up: %6.3f s  after %s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), what);
    };
    {
        HostLevel& L0 = H[0];
        L0.n = N; L0.own = c->own; L0.nei = c->nei; L0.a.resize(F);
        rc = s4f_download_upper(c, L0.a.data()); if (rc) return rc;
        for (int f = 0; f < F; f++) L0.a[f] = -L0.a[f];
        std::vector<double> d(3 * (size_t)c->ld);
        S4F_CHECK_CUDA(c, cudaMemcpy(d.data(), c->diagC.p, d.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int q = 0; q < 3; q++) L0.diag[q].assign(d.begin() + (size_t)q * c->ld, d.begin() + (size_t)q * c->ld + N);
        if (c->nRanks > 1) { L0.dist = true; if ((rc = fine_interfaces(c, L0))) return rc; }
    }
    lap("fine level download + interfaces");
    const int coarsest = 512;
    long long replicateBelow = 150000;          // global cells: below this a level costs less replicated than distributed
    if (const char* e = getenv("S4F_GAMG_REPLICATE_BELOW")) replicateBelow = atoll(e);
    while (H.back().dist) {
        HostLevel part;
        if ((rc = coarsen_distributed(c, H.back(), part))) return rc;
        lap("coarsen_distributed");
        std::vector<int> cnt(c->nRanks);
        if ((rc = s4f_allgather_host(c, &part.n, sizeof(int), cnt.data()))) return rc;
        long long nGlobal = 0, fineGlobal = 0;
        for (int v : cnt) nGlobal += v;
        std::vector<int> cntF(c->nRanks);
        if ((rc = s4f_allgather_host(c, &H.back().n, sizeof(int), cntF.data()))) return rc;
        for (int v : cntF) fineGlobal += v;
        const bool stalled = nGlobal * 10 > fineGlobal * 9;
        if (nGlobal <= replicateBelow || stalled || H.size() >= 6) {
            HostLevel glob;
            if ((rc = gather_level(c, H.back(), part, glob))) return rc;
            lap("gather_level");
            H.push_back(std::move(glob));
        } else H.push_back(std::move(part));
    }
    while (H.back().n > coarsest && H.size() < 12) {
        std::vector<int> total; HostLevel next;
        const int nc = coarsen3(H.back(), coarsest / 4, total, next);
        if (nc >= H.back().n) break;        // no coarsening possible (no faces)
        H.back().parent = total; H.back().childOff = 0; H.back().childCnt = nc;
        H.push_back(std::move(next));
    }
    if (H.back().n > 4096) { c->err = "GAMG: agglomeration stalled above the size of the dense coarsest solve"; return 1; }
    lap("replicated levels");
    if (c->ctl.gamgSinglePrecision) rc = build<float>(c, H);
    else rc = build<double>(c, H);
    lap("device level build");
    if (!rc) c->amg->setupSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    c->graphSerial++;
    return rc;
}

// after mesh motion: same aggregates, new coefficients.  Falls back to a full set-up when there is nothing to refresh.
int s4f_amg_refresh(s4fgpu_ctx* c) {
    if (!c->amg) return s4f_amg_setup(c);
    const auto t0 = std::chrono::steady_clock::now();
    int rc = c->amg->refresh(c);
    if (c->nRanks > 1) {        // the rebuild is collective: every rank takes it if one must
        std::vector<int> all(c->nRanks, 0);
        const int mine = rc ? 1 : 0;
        int ra = s4f_allgather_host(c, &mine, sizeof(int), all.data()); if (ra) return ra;
        for (int v : all) if (v && !rc) { rc = 1; c->err = "another rank could not refresh"; }
    }
    if (getenv("S4F_AMG_TIMING") && !rc && c->rank == 0)
        fprintf(stderr, "libs4fgpu: GAMG coefficient refresh %.4f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    if (rc) {
        fprintf(stderr, "libs4fgpu: GAMG coefficient refresh not possible (%s); rebuilding the hierarchy\n", c->err.c_str());
        c->err.clear();
        cudaGetLastError();
        return s4f_amg_setup(c);
    }
    c->amg->refreshSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

int s4f_amg_info(s4fgpu_ctx* c, int* nLevels, int* sizes, int maxLevels, double* bytesPerApply, double* setupSeconds) {
    if (!c->amg) { c->err = "GAMG hierarchy missing"; return 1; }
    *nLevels = (int)c->amg->sizes.size();
    for (int i = 0; i < *nLevels && i < maxLevels; i++) sizes[i] = c->amg->sizes[i];
    *bytesPerApply = c->amg->bytesPerApply; *setupSeconds = c->amg->setupSeconds;
    return 0;
}

int s4f_amg_distributed_levels(s4fgpu_ctx* c) {
    int n = 0;
    if (c->amg) for (int v : c->amg->distributed) n += v;
    return n;
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
