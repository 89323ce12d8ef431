// s4f_amg_setup.cu -- the GAMG hierarchy built ON THE DEVICE (single-rank runs): pair-wise agglomeration and Galerkin
// coarse matrices, level after level, from the assembled fine matrix in its SELL-32 rows.
//
// Reference behaviour: [OF-ext] pairGAMGAgglomeration ("faceAreaPair"): every cell is paired with its most strongly
// coupled free neighbour, `mergeLevels` such passes make one solver level, coarse coefficients are the sums of the fine
// ones.  Round 1 (and the decomposed runs still, per rank) did this on one host thread: 2.2 s at 8 M cells, the time of 175
// outer iterations.  The sequential greedy sweep becomes a parallel matching by mutual picks: every free cell picks its
// best free neighbour -- strongest coefficient, ties to the LOWER neighbour index ("sweep rule", see `better`) -- and two
// cells that picked each other are matched; repeat until a round matches nothing.  The lowest-numbered end point of the
// strongest class of edges is always picked back, so every round makes progress, and along a line of equally coupled cells
// the pairs form from its low end exactly as the sequential sweep forms them: on a structured hex mesh the aggregates are
// the same 2x2x2 bricks (rounds ~ longest line / 2; a hash / parity tie rule needs O(log n) rounds but staggers the bricks
// and costs PCG iterations -- it remains as the finish rule after S4F_AMG_SWEEP_ROUNDS).  0.17 s at 8 M cells;
// tests/test_gpu_parity.py::test_device_built_gamg_hierarchy_matches_the_host_built_one compares with the host sweep.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>

#include "s4f_amg_setup.h"
#include "s4f_dev.cuh"

namespace {

#define S4F_SETUP_MAXW 64

struct Mat {                 // a matrix in SELL-32 rows on the device (fp64), per-component diagonal
    int n = 0, ld = 0, nSlices = 0; long long nnz = 0; size_t nE = 0;      // nE: padded entries
    const int* slicePtr = nullptr; const int* col = nullptr; const double* a = nullptr; const double* dg = nullptr; int ldDg = 0;
    DevBuf<int> slicePtrB, colB; DevBuf<double> aB, dgB;      // scratch, grown on demand and reused from pass to pass
    void own() { slicePtr = slicePtrB.p; col = colB.p; a = aB.p; dg = dgB.p; ldDg = ld; }
};

// cudaMalloc / cudaFree synchronise the device and cost up to milliseconds for large blocks: the temporaries of the passes
// (15 passes, a dozen arrays each) live in buffers that only grow.  Measured at 8 M cells: 0.22-0.69 s of the set-up was
// allocator time, the matching itself 0.12 s.
template <class T>
cudaError_t ensure(DevBuf<T>& b, size_t count) { return b.n >= count && b.p ? cudaSuccess : b.alloc(count, false); }
struct Scratch {
    DevBuf<int> match, pick, flag, scan, leaderOf, cnt, ent, cs, changed, agg, cursor;
    DevBuf<char> tmp;
    DevBuf<int> fail;
    Mat ping, pong;
};

__device__ __forceinline__ unsigned int hash_edge(int a, int b) {
    unsigned int x = (unsigned int)a * 0x9E3779B1u ^ ((unsigned int)b + 0x7F4A7C15u) * 0x85EBCA6Bu;
    x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12;
    return x;
}

// is edge (i, j1) with coefficient a1 better than (i, j2) with a2?  Coefficients are compared after rounding the fp32
// mantissa by `mb` bits (a_ij and a_ji are sums taken in different orders; the sequential sweep of the host set-up
// treats 1e-7 relative as a tie).  Among ties:
//   sweep rule  (mb == S4F_MATCH_SWEEP_BITS): the lower neighbour index wins.  The lowest-numbered end point of the
//       strongest class of edges is always picked back by its pick, so every round matches something, and along a line
//       of equally coupled cells the pairs form from its low end -- (0,1), then (2,3), ... -- which is the alignment the
//       sequential greedy sweep ([OF-ext] pairGAMGAgglomeration, restated in s4f_amg.cu pairwise_pass) produces: coarse
//       cells stay bricks of 2x2x2 fine cells instead of staggered ones.  Rounds ~ the longest such line / 2.
//   finish rule (any other mb): smaller index distance, then the parity of floor(min(i, j) / d), then a hash of the
//       edge: O(log n) rounds whatever the numbering; used if the sweep rule has not finished after its round budget.
#define S4F_MATCH_SWEEP_BITS 8
__device__ __forceinline__ unsigned int quant(float a, int mb) { return (__float_as_uint(a) + (1u << (mb - 1))) & ~((1u << mb) - 1u); }
__device__ __forceinline__ bool better(int i, int j1, float a1, int j2, float a2, int mb) {
    const unsigned int q1 = quant(a1, mb), q2 = quant(a2, mb);
    if (q1 != q2) return q1 > q2;
    if (mb == S4F_MATCH_SWEEP_BITS) return j1 < j2;
    const int d1 = abs(i - j1), d2 = abs(i - j2);
    if (d1 != d2) return d1 < d2;
    const int e1 = (min(i, j1) / d1) & 1, e2 = (min(i, j2) / d2) & 1;
    if (e1 != e2) return e1 < e2;
    return hash_edge(min(i, j1), max(i, j1)) > hash_edge(min(i, j2), max(i, j2));
}

__global__ void k_pick(int n, const int* __restrict__ sp, const int* __restrict__ col, const double* __restrict__ a, const int* __restrict__ match,
                       int* __restrict__ pick, int mb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int best = -1; float ba = 0.f;
    if (match[i] < 0) {
        const int s = i >> 5, lane = i & 31, base = sp[s], w = (sp[s + 1] - base) >> 5;
        for (int k = 0; k < w; k++) {
            const int j = col[base + 32 * k + lane];
            const float av = (float)a[base + 32 * k + lane];
            if (j == i || j >= n || !(av > 0.f) || match[j] >= 0) continue;
            if (best < 0 || better(i, j, av, best, ba, mb)) { best = j; ba = av; }
        }
    }
    pick[i] = best;
}

__global__ void k_pair(int n, const int* __restrict__ pick, int* __restrict__ match, int* __restrict__ changed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = pick[i];
    if (j >= 0 && pick[j] == i) { match[i] = j; *changed = 1; }
}

// leaders (the lower cell of a pair, or a single cell) get consecutive aggregate numbers
__global__ void k_leader_flag(int n, int* __restrict__ match, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (match[i] < 0) match[i] = i;
    flag[i] = (match[i] >= i) ? 1 : 0;
}
__global__ void k_assign_agg(int n, const int* __restrict__ match, const int* __restrict__ scan, int* __restrict__ agg, int* __restrict__ leaderOf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = min(i, match[i]);
    agg[i] = scan[l];
    if (l == i) leaderOf[scan[i]] = i;
}

// the coarse row of aggregate I: distinct coarse neighbours with summed coefficients (ascending), what stays inside
__device__ __forceinline__ int collect_row(int I, int n, const int* sp, const int* col, const double* a, const int* agg, const int* leaderOf,
                                           const int* match, int* cc, double* acc, double& dsub, bool& overflow) {
    int cnt = 0;
    dsub = 0;
    const int l = leaderOf[I], p = match[l];
    for (int t = 0; t < 2; t++) {
        const int i = t == 0 ? l : p;
        if (t == 1 && p == l) break;
        const int s = i >> 5, lane = i & 31, base = sp[s], w = (sp[s + 1] - base) >> 5;
        for (int k = 0; k < w; k++) {
            const int j = col[base + 32 * k + lane];
            const double av = a[base + 32 * k + lane];
            if (av == 0.0 || j >= n) continue;
            const int J = agg[j];
            if (J == I) { dsub += av; continue; }
            int q = 0;
            while (q < cnt && cc[q] != J) q++;
            if (q == cnt) {
                if (cnt == S4F_SETUP_MAXW) { overflow = true; continue; }
                cc[cnt] = J; acc[cnt] = 0.0; cnt++;
            }
            acc[q] += av;
        }
    }
    return cnt;
}

__global__ void k_coarse_count(int nc, int n, const int* __restrict__ sp, const int* __restrict__ col, const double* __restrict__ a,
                               const int* __restrict__ agg, const int* __restrict__ leaderOf, const int* __restrict__ match, int* __restrict__ cnt,
                               int* __restrict__ fail) {
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nc) return;
    int cc[S4F_SETUP_MAXW]; double acc[S4F_SETUP_MAXW]; double dsub; bool ov = false;
    cnt[I] = collect_row(I, n, sp, col, a, agg, leaderOf, match, cc, acc, dsub, ov);
    if (ov) *fail = 1;
}

// width of every slice (max row length of its 32 rows) times 32
__global__ void k_slice_entries(int nc, int nSlices, const int* __restrict__ cnt, int* __restrict__ ent) {
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= nSlices) return;
    const int r = s * 32 + lane;
    int w = r < nc ? cnt[r] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (lane == 0) ent[s] = 32 * w;
}

__global__ void k_coarse_fill(int nc, int n, const int* __restrict__ sp, const int* __restrict__ col, const double* __restrict__ a,
                              const double* __restrict__ dg, int ldF, const int* __restrict__ agg, const int* __restrict__ leaderOf,
                              const int* __restrict__ match, const int* __restrict__ spC, int* __restrict__ colC, double* __restrict__ aC,
                              double* __restrict__ dgC, int ldC, int nRowsPadded) {
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nRowsPadded) return;
    const int s = I >> 5, lane = I & 31, base = spC[s], w = (spC[s + 1] - base) >> 5;
    if (I >= nc) { for (int k = 0; k < w; k++) { colC[base + 32 * k + lane] = 0; aC[base + 32 * k + lane] = 0.0; } return; }
    int cc[S4F_SETUP_MAXW]; double acc[S4F_SETUP_MAXW]; double dsub; bool ov = false;
    const int cnt = collect_row(I, n, sp, col, a, agg, leaderOf, match, cc, acc, dsub, ov);
    for (int x = 1; x < cnt; x++) {          // ascending columns: a fixed order whatever the fine rows' entry order
        const int cj = cc[x]; const double aj = acc[x];
        int y = x - 1;
        while (y >= 0 && cc[y] > cj) { cc[y + 1] = cc[y]; acc[y + 1] = acc[y]; y--; }
        cc[y + 1] = cj; acc[y + 1] = aj;
    }
    for (int k = 0; k < w; k++) {
        colC[base + 32 * k + lane] = k < cnt ? cc[k] : I;
        aC[base + 32 * k + lane] = k < cnt ? acc[k] : 0.0;
    }
    const int l = leaderOf[I], p = match[l];
#pragma unroll
    for (int q = 0; q < 3; q++) {
        double d = dg[(size_t)q * ldF + l];
        if (p != l) d += dg[(size_t)q * ldF + p];
        dgC[(size_t)q * ldC + I] = d - dsub;
    }
}

__global__ void k_compose(int n, int* __restrict__ total, const int* __restrict__ agg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) total[i] = agg[total[i]];
}
__global__ void k_iota(int n, int* __restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = i;
}
__global__ void k_child_count(int n, const int* __restrict__ parent, int* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&cnt[parent[i]], 1);
}
__global__ void k_child_fill(int n, const int* __restrict__ parent, const int* __restrict__ ptr, int* __restrict__ cursor, int* __restrict__ child) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) child[ptr[parent[i]] + atomicAdd(&cursor[parent[i]], 1)] = i;
}
__global__ void k_child_sort(int nc, const int* __restrict__ ptr, int* __restrict__ child) {      // ascending children: a fixed summation order
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nc) return;
    const int b = ptr[I], e = ptr[I + 1];
    for (int x = b + 1; x < e; x++) {
        const int v = child[x];
        int y = x - 1;
        while (y >= b && child[y] > v) { child[y + 1] = child[y]; y--; }
        child[y + 1] = v;
    }
}

int exclusive_scan(s4fgpu_ctx* c, const int* in, int* out, int n, DevBuf<char>& tmp) {      // out has n + 1 entries: out[n] = total
    size_t bytes = 0;
    S4F_CHECK_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n + 1, c->stream));
    if (tmp.n < bytes) S4F_CHECK_CUDA(c, tmp.alloc(bytes, false));
    S4F_CHECK_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n + 1, c->stream));
    return 0;
}

int read_int(s4fgpu_ctx* c, const int* p, int* host) {
    S4F_CHECK_CUDA(c, cudaMemcpyAsync(host, p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// one pair-wise pass: matching on M, aggregates (S.agg[n]), Galerkin matrix C (in its scratch buffers)
int pair_pass(s4fgpu_ctx* c, const Mat& M, Mat& C, Scratch& S) {
    const int n = M.n, g = (n + 255) / 256;
    DevBuf<int>&match = S.match, &pick = S.pick, &flag = S.flag, &scan = S.scan, &leaderOf = S.leaderOf, &cnt = S.cnt, &ent = S.ent, &agg = S.agg;
    DevBuf<char>& tmp = S.tmp;
    S4F_CHECK_CUDA(c, ensure(match, n)); S4F_CHECK_CUDA(c, ensure(pick, n));
    S4F_CHECK_CUDA(c, ensure(flag, (size_t)n + 1)); S4F_CHECK_CUDA(c, ensure(scan, (size_t)n + 1));
    S4F_CHECK_CUDA(c, cudaMemsetAsync(match.p, 0xFF, (size_t)n * sizeof(int), c->stream));
    S4F_CHECK_CUDA(c, cudaMemsetAsync(flag.p, 0, ((size_t)n + 1) * sizeof(int), c->stream));
    // sweep rule for up to S4F_AMG_SWEEP_ROUNDS rounds (default 2048), then the finish rule; stop when a group of 4 rounds
    // matched nothing
    static const int sweepRounds = getenv("S4F_AMG_SWEEP_ROUNDS") ? atoi(getenv("S4F_AMG_SWEEP_ROUNDS")) : 2048;
    DevBuf<int>& changed = S.changed;
    S4F_CHECK_CUDA(c, ensure(changed, 1));
    S4F_CHECK_CUDA(c, cudaMemsetAsync(changed.p, 0, sizeof(int), c->stream));
    bool sweep = sweepRounds > 0;
    const auto tm0 = std::chrono::steady_clock::now();
    int roundsDone = 0;
    for (int round = 0, inPhase = 0; round < sweepRounds + 64; round++, inPhase++, roundsDone++) {
        if (inPhase > 0 && inPhase % 4 == 0) {
            int h = 0;
            S4F_CHECK_CUDA(c, cudaMemcpyAsync(&h, changed.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            S4F_CHECK_CUDA(c, cudaMemsetAsync(changed.p, 0, sizeof(int), c->stream));
            S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
            if (!h) break;                      // the sweep rule always progresses: nothing matched = nothing left to match
        }
        if (sweep && round >= sweepRounds) { sweep = false; inPhase = 0; }
        k_pick<<<g, 256, 0, c->stream>>>(n, M.slicePtr, M.col, M.a, match.p, pick.p, sweep ? S4F_MATCH_SWEEP_BITS : (inPhase < 8 ? 11 : 21));
        k_pair<<<g, 256, 0, c->stream>>>(n, pick.p, match.p, changed.p);
        c->launches += 2;
    }
    if (getenv("S4F_AMG_TIMING")) {
        cudaStreamSynchronize(c->stream);
        fprintf(stderr, "libs4fgpu:   matching n = %d: %d rounds, %.1f ms\n", n, roundsDone,
                1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - tm0).count());
    }
    k_leader_flag<<<g, 256, 0, c->stream>>>(n, match.p, flag.p);
    c->launches++;
    int rc = exclusive_scan(c, flag.p, scan.p, n, tmp); if (rc) return rc;
    int nc = 0;
    if ((rc = read_int(c, scan.p + n, &nc))) return rc;
    S4F_CHECK_CUDA(c, ensure(agg, n)); S4F_CHECK_CUDA(c, ensure(leaderOf, std::max(nc, 1)));
    k_assign_agg<<<g, 256, 0, c->stream>>>(n, match.p, scan.p, agg.p, leaderOf.p);
    c->launches++;
    // coarse rows
    C.n = nc; C.nSlices = (nc + 31) / 32; C.ld = C.nSlices * 32; if (C.ld == 0) C.ld = 32;
    S4F_CHECK_CUDA(c, ensure(cnt, (size_t)std::max(nc, 1) + 1)); S4F_CHECK_CUDA(c, ensure(ent, (size_t)C.nSlices + 1));
    S4F_CHECK_CUDA(c, cudaMemsetAsync(cnt.p, 0, ((size_t)std::max(nc, 1) + 1) * sizeof(int), c->stream));
    S4F_CHECK_CUDA(c, cudaMemsetAsync(ent.p, 0, ((size_t)C.nSlices + 1) * sizeof(int), c->stream));
    k_coarse_count<<<(nc + 127) / 128, 128, 0, c->stream>>>(nc, n, M.slicePtr, M.col, M.a, agg.p, leaderOf.p, match.p, cnt.p, S.fail.p);
    k_slice_entries<<<(C.nSlices * 32 + 255) / 256, 256, 0, c->stream>>>(nc, C.nSlices, cnt.p, ent.p);
    c->launches += 2;
    S4F_CHECK_CUDA(c, ensure(C.slicePtrB, (size_t)C.nSlices + 1));
    if ((rc = exclusive_scan(c, ent.p, C.slicePtrB.p, C.nSlices, tmp))) return rc;
    int nE = 0;
    if ((rc = read_int(c, C.slicePtrB.p + C.nSlices, &nE))) return rc;
    {   // true off-diagonal entries (for the byte counts)
        S4F_CHECK_CUDA(c, ensure(S.cs, (size_t)nc + 1));
        if ((rc = exclusive_scan(c, cnt.p, S.cs.p, nc, tmp))) return rc;
        int tot = 0; if ((rc = read_int(c, S.cs.p + nc, &tot))) return rc;
        C.nnz = tot;
    }
    C.nE = std::max(nE, 1);
    S4F_CHECK_CUDA(c, ensure(C.colB, C.nE)); S4F_CHECK_CUDA(c, ensure(C.aB, C.nE));
    S4F_CHECK_CUDA(c, ensure(C.dgB, 3 * (size_t)C.ld));
    S4F_CHECK_CUDA(c, cudaMemsetAsync(C.dgB.p, 0, 3 * (size_t)C.ld * sizeof(double), c->stream));
    C.own();
    k_coarse_fill<<<(C.ld + 127) / 128, 128, 0, c->stream>>>(nc, n, M.slicePtr, M.col, M.a, M.dg, M.ldDg, agg.p, leaderOf.p, match.p, C.slicePtrB.p,
                                                           C.colB.p, C.aB.p, C.dgB.p, C.ld, C.ld);
    c->launches++;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return 0;
}

}  // namespace

// levels[0] describes the fine matrix (only `parent` is filled), levels[l >= 1] own their rows
int s4f_amg_device_levels(s4fgpu_ctx* c, std::vector<std::unique_ptr<AmgDevLevel>>& levels, int coarsest, int mergeLevels) {
    levels.clear();
    Scratch S;
    S4F_CHECK_CUDA(c, S.fail.alloc(1));
    levels.emplace_back(new AmgDevLevel());
    levels[0]->n = c->N; levels[0]->ld = c->ld; levels[0]->nSlices = c->nSlices; levels[0]->nnz = c->nnzOff;
    Mat cur;            // the matrix being coarsened: the fine rows, then the rows of the level just made
    cur.n = c->N; cur.ld = c->ld; cur.nSlices = c->nSlices; cur.slicePtr = c->slicePtr.p; cur.col = c->col.p; cur.a = c->eA.p;
    cur.dg = c->diagC.p; cur.ldDg = c->ld; cur.nnz = c->nnzOff;
    while (cur.n > coarsest && levels.size() < 12) {
        const int nFine = cur.n;
        DevBuf<int> total;                  // becomes the level's parent map
        S4F_CHECK_CUDA(c, total.alloc(nFine, false));
        k_iota<<<(nFine + 255) / 256, 256, 0, c->stream>>>(nFine, total.p);
        c->launches++;
        const Mat* m = &cur;
        Mat* last = nullptr;
        int nc = nFine;
        for (int pass = 0; pass < mergeLevels; pass++) {
            Mat* dst = (last == &S.ping) ? &S.pong : &S.ping;
            int rc = pair_pass(c, *m, *dst, S); if (rc) return rc;
            k_compose<<<(nFine + 255) / 256, 256, 0, c->stream>>>(nFine, total.p, S.agg.p);
            c->launches++;
            nc = dst->n;
            last = dst; m = dst;
            if (nc <= coarsest / 4) break;
        }
        int hf = 0;
        { int rc = read_int(c, S.fail.p, &hf); if (rc) return rc; }
        if (hf) { c->err = "GAMG device set-up: a coarse row has more than 64 neighbours"; return 1; }
        if (nc >= nFine) break;                 // no coarsening possible (no couplings)
        AmgDevLevel& F = *levels.back();
        F.parent.swap(total);
        levels.emplace_back(new AmgDevLevel());
        AmgDevLevel& L = *levels.back();
        L.n = last->n; L.ld = last->ld; L.nSlices = last->nSlices; L.nnz = last->nnz;
        // the level keeps exact-size copies; the scratch matrices go on to the next level
        S4F_CHECK_CUDA(c, L.slicePtr.alloc((size_t)L.nSlices + 1, false)); S4F_CHECK_CUDA(c, L.col.alloc(last->nE, false));
        S4F_CHECK_CUDA(c, L.a.alloc(last->nE, false)); S4F_CHECK_CUDA(c, L.dg.alloc(3 * (size_t)L.ld, false));
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(L.slicePtr.p, last->slicePtrB.p, ((size_t)L.nSlices + 1) * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(L.col.p, last->colB.p, last->nE * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(L.a.p, last->aB.p, last->nE * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        S4F_CHECK_CUDA(c, cudaMemcpyAsync(L.dg.p, last->dgB.p, 3 * (size_t)L.ld * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        // children of the new level's cells in the finer one, ascending
        S4F_CHECK_CUDA(c, ensure(S.cnt, (size_t)L.n + 1)); S4F_CHECK_CUDA(c, ensure(S.cursor, std::max(L.n, 1)));
        S4F_CHECK_CUDA(c, cudaMemsetAsync(S.cnt.p, 0, ((size_t)L.n + 1) * sizeof(int), c->stream));
        S4F_CHECK_CUDA(c, cudaMemsetAsync(S.cursor.p, 0, (size_t)std::max(L.n, 1) * sizeof(int), c->stream));
        S4F_CHECK_CUDA(c, L.childPtr.alloc((size_t)L.n + 1, false)); S4F_CHECK_CUDA(c, L.child.alloc(std::max(nFine, 1), false));
        k_child_count<<<(nFine + 255) / 256, 256, 0, c->stream>>>(nFine, F.parent.p, S.cnt.p);
        int rc = exclusive_scan(c, S.cnt.p, L.childPtr.p, L.n, S.tmp); if (rc) return rc;
        k_child_fill<<<(nFine + 255) / 256, 256, 0, c->stream>>>(nFine, F.parent.p, L.childPtr.p, S.cursor.p, L.child.p);
        k_child_sort<<<(L.n + 127) / 128, 128, 0, c->stream>>>(L.n, L.childPtr.p, L.child.p);
        c->launches += 3;
        // continue from the new level
        cur.n = L.n; cur.ld = L.ld; cur.nSlices = L.nSlices; cur.slicePtr = L.slicePtr.p; cur.col = L.col.p; cur.a = L.a.p;
        cur.dg = L.dg.p; cur.ldDg = L.ld; cur.nnz = L.nnz;
    }
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));      // the scratch buffers go out of scope
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return 0;
}
