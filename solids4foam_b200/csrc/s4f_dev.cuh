// s4f_dev.cuh -- device helpers: warp-shuffle grid reductions with a deterministic last-block
// finish, small tensor algebra in OpenFOAM component order, launch geometry.
#pragma once
#include <cuda_runtime.h>

#define S4F_SMALL 1e-15
#define S4F_BLOCK 256

// grid of persistent blocks: a multiple of the SM count (148 on B200)
static inline int s4f_grid(int numSMs, long long work, int blocksPerSM = 8) {
    long long need = (work + S4F_BLOCK - 1) / S4F_BLOCK;
    long long g = (long long)numSMs * blocksPerSM;
    if (need < g) g = need;
    if (g < 1) g = 1;
    return (int)g;
}

struct OpSum { __device__ static double f(double a, double b) { return a + b; } __device__ static double id() { return 0.0; } };
struct OpMax { __device__ static double f(double a, double b) { return a > b ? a : b; } __device__ static double id() { return -1e300; } };

template <class Op>
__device__ __forceinline__ double warp_reduce(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = Op::f(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-level reduce of NV values; result valid in thread 0.
template <int NV, class Op>
__device__ __forceinline__ void block_reduce(double (&v)[NV]) {
    __shared__ double sh[NV][S4F_BLOCK / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double x = warp_reduce<Op>(v[i]);
        if (lane == 0) sh[i][wid] = x;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double x = (lane < (int)(blockDim.x >> 5)) ? sh[i][lane] : Op::id();
            x = warp_reduce<Op>(x);
            v[i] = x;
        }
    }
    __syncthreads();
}

// ---- multi-GPU: reductions and halos over cudaIpc-mapped peer memory (NVLink), no NCCL in the iteration ----------
// Every rank owns a mailbox that its peers write into with plain stores followed by a system-scope release of a flag
// word; messages carry a sequence number kept in device memory (CUDA-graph replays need no host-side argument), and
// slots alternate by parity: in a symmetric exchange a rank sends message k+1 only after it has seen the peer's message
// k, which the peer sent after consuming message k-1, so the slot of k-1 is free.  (s4f_comm.cu sets the pointers up.)
// LL words (after NCCL's low-latency protocol): a message travels as 8-byte words {payload32, sequence number}.  An 8-byte
// store is atomic, so the receiver needs neither a fence nor a separate flag message: it polls every word until the
// sequence number matches.  A double is two words (low / high half), a float one.
__device__ __forceinline__ void ll_store(unsigned long long* p, unsigned int payload, unsigned int k) {
    *reinterpret_cast<volatile unsigned long long*>(p) = ((unsigned long long)k << 32) | payload;
}
__device__ __forceinline__ unsigned int ll_wait(const unsigned long long* p, unsigned int k) {
    unsigned long long w;
    do { w = *reinterpret_cast<const volatile unsigned long long*>(p); } while ((unsigned int)(w >> 32) != k);
    return (unsigned int)w;
}
__device__ __forceinline__ void ll_put(unsigned long long* box, size_t i, double v, unsigned int k) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    ll_store(box + 2 * i, (unsigned int)b, k); ll_store(box + 2 * i + 1, (unsigned int)(b >> 32), k);
}
__device__ __forceinline__ void ll_put(unsigned long long* box, size_t i, float v, unsigned int k) { ll_store(box + i, __float_as_uint(v), k); }
__device__ __forceinline__ void ll_get(const unsigned long long* box, size_t i, unsigned int k, double& v) {
    const unsigned int lo = ll_wait(box + 2 * i, k), hi = ll_wait(box + 2 * i + 1, k);
    v = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
__device__ __forceinline__ void ll_get(const unsigned long long* box, size_t i, unsigned int k, float& v) { v = __uint_as_float(ll_wait(box + i, k)); }

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Grid reduce: every block deposits its NV partials; the last block to arrive (ticket) combines
// them in a fixed order (deterministic for a fixed grid), all-reduces the totals with the other ranks through the
// peer mailboxes (summed in rank order on every rank: bit-identical results everywhere, so all ranks take the same
// branches) and calls fin(tot) from thread 0.  This replaces gSum/gMax + Pstream reductions ([OF-ext]) and the
// separate ncclAllReduce + scalar-step launches of round 1.
template <int NV, class Op, class Fin>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], const RedCtx& rc, Fin fin) {
    static_assert(NV <= S4F_RED_MAX, "reduction wider than the mailbox");
    double* partials = rc.partials;
    unsigned int* ticket = rc.ticket;
    block_reduce<NV, Op>(v);
    __shared__ bool isLast;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) partials[(size_t)i * gridDim.x + blockIdx.x] = v[i];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        isLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (isLast) {
        __threadfence();
        double tot[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double a = Op::id();
            for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) a = Op::f(a, __ldcg(&partials[(size_t)i * gridDim.x + b]));
            tot[i] = a;
        }
        block_reduce<NV, Op>(tot);
        PeerRed* pr = rc.peer;
        if (pr) {
            __shared__ double shTot[NV];
            __shared__ double shIn[S4F_MAX_RANKS][NV];
            const int R = pr->nRanks, me = pr->rank;
            const unsigned int k = pr->seq + 1u, par = k & 1u;
            if (threadIdx.x == 0) {
#pragma unroll
                for (int i = 0; i < NV; i++) shTot[i] = tot[i];
            }
            __syncthreads();
            // my partial sums go to every rank's mailbox (slot `me`), then I collect the R contributions in mine
            for (int t = threadIdx.x; t < R * NV; t += blockDim.x) {
                const int r = t / NV, i = t - r * NV;
                ll_put(pr->box[r] + ((size_t)par * R + me) * 2 * S4F_RED_MAX, (size_t)i, shTot[i], k);
            }
            for (int t = threadIdx.x; t < R * NV; t += blockDim.x) {
                const int r = t / NV, i = t - r * NV;
                double v;
                ll_get(pr->box[me] + ((size_t)par * R + r) * 2 * S4F_RED_MAX, (size_t)i, k, v);
                shIn[r][i] = v;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
#pragma unroll
                for (int i = 0; i < NV; i++) {
                    double a = shIn[0][i];
                    for (int r = 1; r < R; r++) a = Op::f(a, shIn[r][i]);
                    tot[i] = a;
                }
                pr->seq = k;
            }
        }
        if (threadIdx.x == 0) { *ticket = 0u; fin(tot); }
    }
}

// halo exchange of one plan (HaloPlan in s4f_comm.cu), passed to the exchange kernel by value
struct HaloDev {
    int nNbr, maxComp;
    int scount[S4F_MAX_NBRS], soff[S4F_MAX_NBRS]; // values per component sent to neighbour n; prefix sums (into sendCells)
    int rcount[S4F_MAX_NBRS], roff[S4F_MAX_NBRS]; // values per component received from neighbour n; prefix sums (ghost order)
    unsigned long long* peerBox[S4F_MAX_NBRS];    // remote: where my message to neighbour n goes, [2 parities][2 * maxComp * scount] LL words
    unsigned long long* myBox[S4F_MAX_NBRS];      // local: messages of neighbour n, [2][2 * maxComp * rcount]
    const int* sendCells;                         // [sum count] local cells whose values go out, neighbour by neighbour
    unsigned int* seq;                            // exchanges completed; [1] is the block ticket of the kernel's tail
};

// ---- tensor algebra: tensor 9 row-major, symmTensor 6 = XX XY XZ YY YZ ZZ ----------------------
__device__ __forceinline__ void t_symm(const double* T, double* S) {
    S[0] = T[0]; S[1] = 0.5 * (T[1] + T[3]); S[2] = 0.5 * (T[2] + T[6]);
    S[3] = T[4]; S[4] = 0.5 * (T[5] + T[7]); S[5] = T[8];
}
__device__ __forceinline__ double s_tr(const double* S) { return S[0] + S[3] + S[5]; }
__device__ __forceinline__ void s_dev(const double* S, double* D) {
    const double t = s_tr(S) / 3.0;
    D[0] = S[0] - t; D[1] = S[1]; D[2] = S[2]; D[3] = S[3] - t; D[4] = S[4]; D[5] = S[5] - t;
}
__device__ __forceinline__ double s_magSqr(const double* S) {
    return S[0] * S[0] + 2 * S[1] * S[1] + 2 * S[2] * S[2] + S[3] * S[3] + 2 * S[4] * S[4] + S[5] * S[5];
}
__device__ __forceinline__ double s_det(const double* S) {
    return S[0] * S[3] * S[5] + S[1] * S[4] * S[2] + S[2] * S[1] * S[4] - S[0] * S[4] * S[4] - S[1] * S[1] * S[5] - S[2] * S[3] * S[2];
}
__device__ __forceinline__ double t_det(const double* T) {
    return T[0] * (T[4] * T[8] - T[5] * T[7]) - T[1] * (T[3] * T[8] - T[5] * T[6]) + T[2] * (T[3] * T[7] - T[4] * T[6]);
}
__device__ __forceinline__ void t_inv(const double* T, double* R) {
    const double d = t_det(T);
    R[0] = (T[4] * T[8] - T[5] * T[7]) / d; R[1] = (T[2] * T[7] - T[1] * T[8]) / d; R[2] = (T[1] * T[5] - T[2] * T[4]) / d;
    R[3] = (T[5] * T[6] - T[3] * T[8]) / d; R[4] = (T[0] * T[8] - T[2] * T[6]) / d; R[5] = (T[2] * T[3] - T[0] * T[5]) / d;
    R[6] = (T[3] * T[7] - T[4] * T[6]) / d; R[7] = (T[1] * T[6] - T[0] * T[7]) / d; R[8] = (T[0] * T[4] - T[1] * T[3]) / d;
}
__device__ __forceinline__ void t_mul(const double* A, const double* B, double* R) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) R[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ __forceinline__ void s_to_t(const double* S, double* T) {
    T[0] = S[0]; T[1] = S[1]; T[2] = S[2]; T[3] = S[1]; T[4] = S[3]; T[5] = S[4]; T[6] = S[2]; T[7] = S[4]; T[8] = S[5];
}
__device__ __forceinline__ void t_transpose(const double* A, double* R) {
    R[0] = A[0]; R[1] = A[3]; R[2] = A[6]; R[3] = A[1]; R[4] = A[4]; R[5] = A[7]; R[6] = A[2]; R[7] = A[5]; R[8] = A[8];
}
