// s4f_comm.cu -- multi-GPU exchange layer of the solid-solver hot path: processor-patch halos, global reductions and
// the coarse-level gather of the GAMG preconditioner, all over cudaIpc-mapped peer memory (NVLink / NVSwitch).
//
// Reference behaviour replaced: OpenFOAM's processorFvPatchField::initEvaluate/evaluate and
// lduMatrix::initMatrixInterfaces/updateMatrixInterfaces (halo of the cell values next to a processor patch, once per
// Amul / smoothing sweep), and Pstream::gSum / gMax ([OF-ext]).  Round 1 did both with NCCL (pack kernel,
// one ncclSend/ncclRecv per neighbour and component, unpack kernel: ~90 us per exchange; ncclAllReduce plus a one-thread
// kernel per reduction).  Here:
//   * a halo exchange is ONE kernel: every thread writes its share of the boundary-cell values straight into the
//     neighbour's mailbox (peer stores over NVLink) as LL words {payload32, sequence number} -- an 8-byte store is atomic, so
//     no fence and no flag message are needed --, then polls its share of the incoming words and copies the values into
//     the ghost slots [n, n+G) of the field: 5 us per exchange inside a CUDA graph against 14 us for data + fence + flag
//     (profiles/r2_p2p_exchange_latency.log) and ~90 us for round 1's pack / NCCL group / unpack;
//   * a reduction needs no kernel of its own: the last block of the reducing kernel exchanges the partial sums with
//     all ranks (s4f_dev.cuh, grid_reduce) and finishes the scalar step;
//   * the gather that feeds the replicated coarse GAMG levels is the same push + flag scheme to all ranks.
// NCCL remains for set-up only (exchange of IPC handles and of the agglomeration tables).
#include <algorithm>
#include <cstring>

#include "s4f_comm.h"
#include "s4f_dev.cuh"

// equal-sized host blocks, all-gathered through device staging (set-up only)
int s4f_allgather_host(s4fgpu_ctx* c, const void* send, size_t bytes, void* recv) {
    if (c->nRanks <= 1) { if (bytes) std::memcpy(recv, send, bytes); return 0; }
    DevBuf<char> ds, dr;
    S4F_CHECK_CUDA(c, ds.alloc(std::max<size_t>(bytes, 1), false)); S4F_CHECK_CUDA(c, dr.alloc(std::max<size_t>(bytes, 1) * c->nRanks, false));
    if (bytes) S4F_CHECK_CUDA(c, cudaMemcpyAsync(ds.p, send, bytes, cudaMemcpyHostToDevice, c->stream));
    S4F_CHECK_NCCL(c, ncclAllGather(ds.p, dr.p, std::max<size_t>(bytes, 1), ncclChar, c->comm, c->stream));
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    if (bytes) S4F_CHECK_CUDA(c, cudaMemcpy(recv, dr.p, bytes * c->nRanks, cudaMemcpyDeviceToHost));
    return 0;
}

// variable-sized blocks: recv[r] = what rank r sent
int s4f_allgatherv_host(s4fgpu_ctx* c, const void* send, size_t bytes, std::vector<std::vector<char>>& recv) {
    const int R = c->nRanks;
    std::vector<unsigned long long> sizes(R);
    unsigned long long mine = bytes;
    int rc = s4f_allgather_host(c, &mine, sizeof(mine), sizes.data()); if (rc) return rc;
    size_t mx = 1;
    for (int r = 0; r < R; r++) mx = std::max<size_t>(mx, sizes[r]);
    std::vector<char> sb(mx, 0), rb(mx * R);
    if (bytes) std::memcpy(sb.data(), send, bytes);
    rc = s4f_allgather_host(c, sb.data(), mx, rb.data()); if (rc) return rc;
    recv.resize(R);
    for (int r = 0; r < R; r++) recv[r].assign(rb.begin() + (size_t)r * mx, rb.begin() + (size_t)r * mx + sizes[r]);
    return 0;
}

// one block of ints per neighbour, sent to / received from the rank across (sizes agree pairwise by construction)
int s4f_exchange_nbr_ints(s4fgpu_ctx* c, const std::vector<int>& nbrRank, const std::vector<std::vector<int>>& send,
                          std::vector<std::vector<int>>& recv) {
    const size_t nn = nbrRank.size();
    recv.assign(nn, {});
    size_t tot = 0;
    for (size_t n = 0; n < nn; n++) tot += send[n].size();
    if (tot == 0) { for (size_t n = 0; n < nn; n++) recv[n].clear(); return 0; }
    std::vector<int> flat; flat.reserve(tot);
    for (size_t n = 0; n < nn; n++) flat.insert(flat.end(), send[n].begin(), send[n].end());
    DevBuf<int> ds, dr;
    S4F_CHECK_CUDA(c, ds.upload(flat)); S4F_CHECK_CUDA(c, dr.alloc(tot));
    S4F_CHECK_NCCL(c, ncclGroupStart());
    size_t off = 0;
    for (size_t n = 0; n < nn; n++) {
        if (!send[n].empty()) {
            S4F_CHECK_NCCL(c, ncclSend(ds.p + off, send[n].size(), ncclInt, nbrRank[n], c->comm, c->stream));
            S4F_CHECK_NCCL(c, ncclRecv(dr.p + off, send[n].size(), ncclInt, nbrRank[n], c->comm, c->stream));
        }
        off += send[n].size();
    }
    S4F_CHECK_NCCL(c, ncclGroupEnd());
    S4F_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<int> back(tot);
    S4F_CHECK_CUDA(c, cudaMemcpy(back.data(), dr.p, tot * sizeof(int), cudaMemcpyDeviceToHost));
    off = 0;
    for (size_t n = 0; n < nn; n++) { recv[n].assign(back.begin() + off, back.begin() + off + send[n].size()); off += send[n].size(); }
    return 0;
}

// ---- IPC: every rank publishes one allocation; the ranks in `want` map it ------------------------------------------
struct IpcRecord { cudaIpcMemHandle_t h; unsigned long long extra[2 + 2 * S4F_MAX_NBRS]; };

static int ipc_publish(s4fgpu_ctx* c, void* base, const unsigned long long* extra, int nExtra, std::vector<IpcRecord>& all) {
    IpcRecord mine; std::memset(&mine, 0, sizeof(mine));
    S4F_CHECK_CUDA(c, cudaIpcGetMemHandle(&mine.h, base));
    for (int i = 0; i < nExtra; i++) mine.extra[i] = extra[i];
    all.resize(c->nRanks);
    return s4f_allgather_host(c, &mine, sizeof(mine), all.data());
}

static int ipc_open(s4fgpu_ctx* c, const IpcRecord& rec, void** out, std::vector<void*>& opened) {
    cudaError_t e = cudaIpcOpenMemHandle(out, rec.h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        c->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e) +
                 " (the multi-GPU path needs peer access between the GPUs of the box; there is no fallback)";
        return 2;
    }
    opened.push_back(*out);
    return 0;
}

// ---- reductions ---------------------------------------------------------------------------------------------------
int s4f_comm_setup(s4fgpu_ctx* c) {
    if (c->nRanks <= 1) return 0;
    if (c->nRanks > S4F_MAX_RANKS) { c->err = "more ranks than S4F_MAX_RANKS"; return 1; }
    const int R = c->nRanks;
    const size_t boxBytes = 2 * (size_t)R * 2 * S4F_RED_MAX * sizeof(unsigned long long);
    S4F_CHECK_CUDA(c, c->redArena.alloc(boxBytes));          // zero-filled: sequence number 0 = nothing received
    S4F_CHECK_CUDA(c, cudaDeviceSynchronize());
    std::vector<IpcRecord> all;
    int rc = ipc_publish(c, c->redArena.p, nullptr, 0, all); if (rc) return rc;
    PeerRed h; std::memset(&h, 0, sizeof(h));
    h.nRanks = R; h.rank = c->rank; h.seq = 0;
    for (int r = 0; r < R; r++) {
        void* base = c->redArena.p;
        if (r != c->rank) { rc = ipc_open(c, all[r], &base, c->ipcOpened); if (rc) return rc; }
        h.box[r] = (unsigned long long*)base;
    }
    S4F_CHECK_CUDA(c, c->redDev.alloc(1));
    S4F_CHECK_CUDA(c, cudaMemcpy(c->redDev.p, &h, sizeof(h), cudaMemcpyHostToDevice));
    // nobody may push before every rank has mapped (and zeroed) its mailbox
    int one = 1; std::vector<int> ones(R);
    return s4f_allgather_host(c, &one, sizeof(int), ones.data());
}

void s4f_comm_destroy(s4fgpu_ctx* c) {
    for (void* p : c->ipcOpened) cudaIpcCloseMemHandle(p);
    c->ipcOpened.clear();
    c->redDev.release(); c->redArena.release();
}

// ---- halo plans -----------------------------------------------------------------------------------------------------
struct S4fHaloPlan {
    HaloDev d{};
    DevBuf<char> arena;
    DevBuf<int> sendCells;
    DevBuf<unsigned int> seq;
    std::vector<void*> opened;
    int total = 0;
};

namespace {

// the whole exchange in one kernel: push my boundary-cell values as LL words, poll the neighbours' words into the ghost slots
template <class T>
__global__ void __launch_bounds__(256) k_halo_xchg(HaloDev h, T* __restrict__ f, int ld, int ncomp, int ghostBase) {
    constexpr int W = sizeof(T) / 4;                 // LL words per value
    const unsigned int k = h.seq[0] + 1u, par = k & 1u;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int n = 0; n < h.nNbr; n++) {
        const int cnt = h.scount[n];
        unsigned long long* box = h.peerBox[n] + (size_t)par * 2 * h.maxComp * cnt;
        const int* sc = h.sendCells + h.soff[n];
        for (int i = tid; i < cnt * ncomp; i += nth) {
            const int q = i / cnt, g = i - q * cnt;
            ll_put(box, (size_t)i, f[(size_t)q * ld + sc[g]], k);
        }
        (void)W;
    }
    for (int n = 0; n < h.nNbr; n++) {
        const int cnt = h.rcount[n];
        const unsigned long long* box = h.myBox[n] + (size_t)par * 2 * h.maxComp * cnt;
        T* dst = f + ghostBase + h.roff[n];
        for (int i = tid; i < cnt * ncomp; i += nth) {
            const int q = i / cnt, g = i - q * cnt;
            T v; ll_get(box, (size_t)i, k, v);
            dst[(size_t)q * ld + g] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {       // every block has read seq[0] by now: the last one out publishes the new sequence number
        const unsigned int t = atomicAdd(h.seq + 1, 1u);
        if (t == gridDim.x - 1) { h.seq[1] = 0u; __threadfence(); h.seq[0] = k; }
    }
}

}  // namespace

int s4f_halo_plan_create(s4fgpu_ctx* c, const std::vector<int>& nbrRank, const std::vector<int>& nbrCount, const std::vector<int>& sendCells,
                         int maxComp, S4fHaloPlan** out) {
    return s4f_halo_plan_create_asym(c, nbrRank, nbrCount, nbrCount, sendCells, maxComp, out);
}

int s4f_halo_plan_create_asym(s4fgpu_ctx* c, const std::vector<int>& nbrRank, const std::vector<int>& sendCount, const std::vector<int>& recvCount,
                              const std::vector<int>& sendCells, int maxComp, S4fHaloPlan** out) {
    *out = nullptr;
    const int nn = (int)nbrRank.size();
    if (nn > S4F_MAX_NBRS) { c->err = "more neighbour ranks than S4F_MAX_NBRS"; return 1; }
    auto* P = new S4fHaloPlan();
    std::unique_ptr<S4fHaloPlan> guard(P);
    HaloDev& d = P->d;
    d.nNbr = nn; d.maxComp = maxComp;
    // my arena: per neighbour [2 parities][2 * maxComp * count] LL words (a double is two words)
    unsigned long long extra[2 + 2 * S4F_MAX_NBRS]; std::memset(extra, 0, sizeof(extra));
    size_t off = 0; int tot = 0, rtot = 0;
    for (int n = 0; n < nn; n++) {
        d.scount[n] = sendCount[n]; d.soff[n] = tot; tot += sendCount[n];
        d.rcount[n] = recvCount[n]; d.roff[n] = rtot; rtot += recvCount[n];
        extra[2 + 2 * n] = (unsigned long long)nbrRank[n];
        extra[2 + 2 * n + 1] = off;
        off += 2 * (size_t)2 * maxComp * recvCount[n] * sizeof(unsigned long long);
        off = (off + 255) / 256 * 256;
    }
    extra[0] = (unsigned long long)nn; extra[1] = 0;
    P->total = std::max(tot, rtot);
    S4F_CHECK_CUDA(c, P->arena.alloc(off + 256));
    S4F_CHECK_CUDA(c, P->seq.alloc(4));
    S4F_CHECK_CUDA(c, P->sendCells.upload(sendCells.empty() ? std::vector<int>(1, 0) : sendCells));
    S4F_CHECK_CUDA(c, cudaDeviceSynchronize());
    std::vector<IpcRecord> all;
    int rc = ipc_publish(c, P->arena.p, extra, 2 + 2 * nn, all); if (rc) return rc;
    std::vector<void*> base(c->nRanks, nullptr);
    for (int n = 0; n < nn; n++) {
        const int r = nbrRank[n];
        if (r < 0 || r >= c->nRanks || r == c->rank) { c->err = "halo plan: bad neighbour rank"; return 1; }
        if (!base[r]) { rc = ipc_open(c, all[r], &base[r], P->opened); if (rc) return rc; }
        // the k-th interface between the two ranks on my side is the k-th on the other side
        int kth = 0;
        for (int m = 0; m < n; m++) if (nbrRank[m] == r) kth++;
        const IpcRecord& rec = all[r];
        int found = -1, seen = 0;
        for (int j = 0; j < (int)rec.extra[0]; j++)
            if ((int)rec.extra[2 + 2 * j] == c->rank) { if (seen == kth) { found = j; break; } seen++; }
        if (found < 0) { c->err = "halo plan: the neighbour rank does not list this rank"; return 1; }
        d.peerBox[n] = (unsigned long long*)((char*)base[r] + rec.extra[2 + 2 * found + 1]);
        d.myBox[n] = (unsigned long long*)(P->arena.p + extra[2 + 2 * n + 1]);
    }
    d.sendCells = P->sendCells.p;
    d.seq = P->seq.p;
    int one = 1; std::vector<int> ones(c->nRanks);
    rc = s4f_allgather_host(c, &one, sizeof(int), ones.data()); if (rc) return rc;
    *out = guard.release();
    return 0;
}

void s4f_halo_plan_destroy(S4fHaloPlan* P) {
    if (!P) return;
    for (void* p : P->opened) cudaIpcCloseMemHandle(p);
    delete P;
}

template <class T>
int s4f_halo_run(s4fgpu_ctx* c, S4fHaloPlan* P, T* field, int ld, int ncomp, int ghostBase) {
    if (!P || P->d.nNbr == 0) return 0;
    if (ncomp > P->d.maxComp) { c->err = "halo exchange wider than the plan's mailbox"; return 1; }
    long long work = (long long)P->total * ncomp;
    int grid = (int)std::min<long long>(64, (work + 767) / 768);
    if (grid < 1) grid = 1;
    k_halo_xchg<T><<<grid, 256, 0, c->stream>>>(P->d, field, ld, ncomp, ghostBase);
    c->launches++;
    return 0;
}
template int s4f_halo_run<double>(s4fgpu_ctx*, S4fHaloPlan*, double*, int, int, int);
template int s4f_halo_run<float>(s4fgpu_ctx*, S4fHaloPlan*, float*, int, int, int);

// ---- gather to all ranks (replicated coarse levels of the GAMG hierarchy) ------------------------------------------
struct GatherDev {
    int nRanks, rank, ldG;                 // ldG: leading dimension of the gathered [3][ldG] vector
    int off[S4F_MAX_RANKS], cnt[S4F_MAX_RANKS];
    unsigned long long* box[S4F_MAX_RANKS];   // rank r's buffer [2 parities][2 * 3 * ldG] LL words
    unsigned int* seq;
};
struct S4fGatherPlan {
    GatherDev d{};
    DevBuf<char> arena;
    DevBuf<unsigned int> seq;
    std::vector<void*> opened;
    int nGlobal = 0;
};

namespace {
template <class T>
__global__ void __launch_bounds__(256) k_gather_xchg(GatherDev g, const T* __restrict__ src, int lds, T* __restrict__ dst, int ldd,
                                                     const int* __restrict__ act) {
    const unsigned int k = g.seq[0] + 1u, par = k & 1u;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int R = g.nRanks, me = g.rank, nLoc = g.cnt[me], myOff = g.off[me];
    const bool a[3] = {act[0] != 0, act[1] != 0, act[2] != 0};
    const size_t slot = (size_t)par * 2 * 3 * g.ldG;
    for (int r = 0; r < R; r++) {          // my piece to every rank (staggered start), as LL words
        unsigned long long* box = g.box[(me + r) % R] + slot;
        for (int i = tid; i < 3 * nLoc; i += nth) {
            const int q = i / nLoc, j = i - q * nLoc;
            if (a[q]) ll_put(box, (size_t)q * g.ldG + myOff + j, src[(size_t)q * lds + j], k);
        }
    }
    const unsigned long long* mine = g.box[me] + slot;
    const int nG = g.off[R - 1] + g.cnt[R - 1];
    for (int i = tid; i < 3 * nG; i += nth) {
        const int q = i / nG, j = i - q * nG;
        if (a[q]) { T v; ll_get(mine, (size_t)q * g.ldG + j, k, v); dst[(size_t)q * ldd + j] = v; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(g.seq + 1, 1u);
        if (t == gridDim.x - 1) { g.seq[1] = 0u; __threadfence(); g.seq[0] = k; }
    }
}
}  // namespace

int s4f_gather_plan_create(s4fgpu_ctx* c, const std::vector<int>& cntPerRank, S4fGatherPlan** out) {
    *out = nullptr;
    const int R = c->nRanks;
    auto* P = new S4fGatherPlan();
    std::unique_ptr<S4fGatherPlan> guard(P);
    GatherDev& d = P->d;
    d.nRanks = R; d.rank = c->rank;
    int tot = 0;
    for (int r = 0; r < R; r++) { d.off[r] = tot; d.cnt[r] = cntPerRank[r]; tot += cntPerRank[r]; }
    P->nGlobal = tot;
    d.ldG = ((tot + 31) / 32) * 32;
    const size_t boxBytes = 2 * (size_t)2 * 3 * d.ldG * sizeof(unsigned long long);
    S4F_CHECK_CUDA(c, P->arena.alloc(boxBytes));
    S4F_CHECK_CUDA(c, P->seq.alloc(4));
    S4F_CHECK_CUDA(c, cudaDeviceSynchronize());
    std::vector<IpcRecord> all;
    int rc = ipc_publish(c, P->arena.p, nullptr, 0, all); if (rc) return rc;
    for (int r = 0; r < R; r++) {
        void* base = P->arena.p;
        if (r != c->rank) { rc = ipc_open(c, all[r], &base, P->opened); if (rc) return rc; }
        d.box[r] = (unsigned long long*)base;
    }
    d.seq = P->seq.p;
    int one = 1; std::vector<int> ones(R);
    rc = s4f_allgather_host(c, &one, sizeof(int), ones.data()); if (rc) return rc;
    *out = guard.release();
    return 0;
}

void s4f_gather_plan_destroy(S4fGatherPlan* P) {
    if (!P) return;
    for (void* p : P->opened) cudaIpcCloseMemHandle(p);
    delete P;
}

template <class T>
int s4f_gather_run(s4fgpu_ctx* c, S4fGatherPlan* P, const T* src, int lds, T* dst, int ldd, const int* act) {
    long long work = 3LL * P->nGlobal;
    int grid = (int)std::min<long long>(96, (work + 2047) / 2048);
    if (grid < 1) grid = 1;
    k_gather_xchg<T><<<grid, 256, 0, c->stream>>>(P->d, src, lds, dst, ldd, act);
    c->launches++;
    return 0;
}
template int s4f_gather_run<double>(s4fgpu_ctx*, S4fGatherPlan*, const double*, int, double*, int, const int*);
template int s4f_gather_run<float>(s4fgpu_ctx*, S4fGatherPlan*, const float*, int, float*, int, const int*);
