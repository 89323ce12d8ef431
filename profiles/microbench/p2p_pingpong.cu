// p2p_pingpong.cu -- can two processes (one per GPU) exchange halo data through cudaIpc-mapped peer memory with
// device-side flags, and what does one exchange cost?  (round 2: replaces pack + ncclSend/ncclRecv + unpack)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o p2p_pingpong p2p_pingpong.cu && ./p2p_pingpong [nRanks] [values]
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("rank %d: %s -> %s\n", rank, #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Shared { cudaIpcMemHandle_t h[8]; volatile int arrived[4]; };

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// one exchange in one kernel: push my n values into the peer's staging + flag; wait for the peer's flag; unpack.
// mailbox layout (per rank): [2 parities][n doubles] then flags[2]
__global__ void k_xchg(const double* __restrict__ src, double* __restrict__ dst, double* peerBox, unsigned* peerFlag, double* myBox,
                       unsigned* myFlag, unsigned* seq, unsigned* ticket, int n) {
    const unsigned k = *seq + 1, par = k & 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) peerBox[(size_t)par * n + i] = src[i];
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned last;
    if (threadIdx.x == 0) { last = atomicAdd(ticket, 1u); if (last == gridDim.x - 1) st_release_sys(peerFlag + par, k); }
    if (threadIdx.x == 0) { while (ld_acquire_sys(myFlag + par) != k) {} }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = __ldcv(myBox + (size_t)par * n + i);
    __syncthreads();
    if (threadIdx.x == 0) { unsigned t = atomicAdd(ticket + 1, 1u); if (t == gridDim.x - 1) { ticket[0] = 0; ticket[1] = 0; *seq = k; } }
}

// LL variant (after NCCL's low-latency protocol): every double travels as two 8-byte words {lo32, flag}, {hi32, flag}; an
// 8-byte store is atomic, so the receiver needs no fence and no separate flag message: it spins on each element until both
// flags carry the sequence number.  Any grid size, no block-level synchronisation on the critical path.
__global__ void k_xchg_ll(const double* __restrict__ src, double* __restrict__ dst, uint2* peerBox, uint2* myBox, unsigned* seq, int n) {
    const unsigned k = seq[0] + 1, par = k & 1;
    uint2* out = peerBox + (size_t)par * 2 * n;
    const volatile uint2* in = myBox + (size_t)par * 2 * n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long v = __double_as_longlong(src[i]);
        out[2 * i] = make_uint2((unsigned)v, k);
        out[2 * i + 1] = make_uint2((unsigned)(v >> 32), k);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint2 a, b;
        do { a.x = in[2 * i].x; a.y = in[2 * i].y; b.x = in[2 * i + 1].x; b.y = in[2 * i + 1].y; } while (a.y != k || b.y != k);
        dst[i] = __longlong_as_double(((unsigned long long)b.x << 32) | a.x);
    }
    __syncthreads();
    if (threadIdx.x == 0) { unsigned t = atomicAdd(seq + 1, 1u); if (t == gridDim.x - 1) { seq[1] = 0; __threadfence(); seq[0] = k; } }
}

int main(int argc, char** argv) {
    const int R = argc > 1 ? atoi(argv[1]) : 2;
    const int n = argc > 2 ? atoi(argv[2]) : 30000;
    Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset((void*)sh, 0, sizeof(Shared));
    int rank = 0;
    for (int r = 1; r < R; r++) { if (fork() == 0) { rank = r; break; } }
    auto barrier = [&](int id) { __sync_fetch_and_add(&sh->arrived[id], 1); while (sh->arrived[id] < R) usleep(100); };
    int ndev = 0; CK(cudaGetDeviceCount(&ndev));
    if (ndev < R) { if (rank == 0) printf("need %d GPUs, have %d\n", R, ndev); return 0; }
    CK(cudaSetDevice(rank));
    const size_t boxBytes = 2 * (size_t)n * 8 + 64;
    char* box; CK(cudaMalloc(&box, boxBytes)); CK(cudaMemset(box, 0, boxBytes));
    CK(cudaIpcGetMemHandle(&sh->h[rank], box));
    barrier(0);
    const int peer = (rank + 1) % R;     // ring: send to next, receive from previous (R = 2: the same rank)
    const int prev = (rank + R - 1) % R;
    if (R > 2 && rank == 0) printf("ring of %d\n", R);
    char* peerBox; CK(cudaIpcOpenMemHandle((void**)&peerBox, sh->h[peer], cudaIpcMemLazyEnablePeerAccess));
    (void)prev;
    double *src, *dst; unsigned *seq, *ticket;
    CK(cudaMalloc(&src, n * 8)); CK(cudaMalloc(&dst, n * 8)); CK(cudaMalloc(&seq, 4)); CK(cudaMalloc(&ticket, 8));
    CK(cudaMemset(seq, 0, 4)); CK(cudaMemset(ticket, 0, 8));
    double* hs = (double*)malloc(n * 8); for (int i = 0; i < n; i++) hs[i] = rank * 1000000.0 + i;
    CK(cudaMemcpy(src, hs, n * 8, cudaMemcpyHostToDevice));
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CK(cudaDeviceSynchronize());
    barrier(1);
    const int blocks = 32;
    auto launch = [&]() {
        k_xchg<<<blocks, 256, 0, s>>>(src, dst, (double*)peerBox, (unsigned*)(peerBox + 2 * (size_t)n * 8), (double*)box,
                                      (unsigned*)(box + 2 * (size_t)n * 8), seq, ticket, n);
    };
    for (int i = 0; i < 20; i++) launch();
    CK(cudaStreamSynchronize(s));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    barrier(2);
    const int reps = 1000;
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < reps; i++) launch();
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(hs, dst, n * 8, cudaMemcpyDeviceToHost));
    const bool ok = hs[0] == prev * 1000000.0 && hs[n - 1] == prev * 1000000.0 + n - 1;
    // the same exchange captured in a CUDA graph of 50 nodes (launch overhead removed)
    cudaGraph_t g; cudaGraphExec_t ex;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < 50; i++) launch();
    CK(cudaStreamEndCapture(s, &g)); CK(cudaGraphInstantiate(&ex, g, 0));
    barrier(3);
    CK(cudaGraphLaunch(ex, s)); CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < 20; i++) CK(cudaGraphLaunch(ex, s));
    CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
    float msg; CK(cudaEventElapsedTime(&msg, e0, e1));
    printf("rank %d: %d doubles per exchange, stream launches %.2f us, in a graph %.2f us per exchange, data %s\n", rank, n,
           1e3 * ms / reps, 1e3 * msg / 1000, ok ? "ok" : "WRONG");
    // ---- LL protocol ----
    {
        char* llbox; CK(cudaMalloc(&llbox, 2 * (size_t)n * 16)); CK(cudaMemset(llbox, 0, 2 * (size_t)n * 16));
        unsigned* seq2; CK(cudaMalloc(&seq2, 8)); CK(cudaMemset(seq2, 0, 8));
        CK(cudaDeviceSynchronize());
        CK(cudaIpcGetMemHandle(&sh->h[rank], llbox));
        __sync_synchronize();
        // second barrier set: reuse counters beyond the first four
        __sync_fetch_and_add(&sh->arrived[3], 1); while (sh->arrived[3] < 2 * R) usleep(100);
        char* peerLL; CK(cudaIpcOpenMemHandle((void**)&peerLL, sh->h[peer], cudaIpcMemLazyEnablePeerAccess));
        __sync_fetch_and_add(&sh->arrived[2], 1); while (sh->arrived[2] < 2 * R) usleep(100);
        const int blocksLL = n >= 8192 ? 32 : (n >= 1024 ? 4 : 1);
        auto launchLL = [&]() { k_xchg_ll<<<blocksLL, 256, 0, s>>>(src, dst, (uint2*)peerLL, (uint2*)llbox, seq2, n); };
        CK(cudaMemset(dst, 0, n * 8));
        for (int i = 0; i < 20; i++) launchLL();
        CK(cudaStreamSynchronize(s));
        __sync_fetch_and_add(&sh->arrived[1], 1); while (sh->arrived[1] < 2 * R) usleep(100);
        CK(cudaEventRecord(e0, s));
        for (int i = 0; i < reps; i++) launchLL();
        CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
        float msl; CK(cudaEventElapsedTime(&msl, e0, e1));
        CK(cudaMemcpy(hs, dst, n * 8, cudaMemcpyDeviceToHost));
        const bool okl = hs[0] == prev * 1000000.0 && hs[n - 1] == prev * 1000000.0 + n - 1;
        cudaGraph_t g2; cudaGraphExec_t ex2;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < 50; i++) launchLL();
        CK(cudaStreamEndCapture(s, &g2)); CK(cudaGraphInstantiate(&ex2, g2, 0));
        __sync_fetch_and_add(&sh->arrived[0], 1); while (sh->arrived[0] < 2 * R) usleep(100);
        CK(cudaGraphLaunch(ex2, s)); CK(cudaStreamSynchronize(s));
        CK(cudaEventRecord(e0, s));
        for (int i = 0; i < 20; i++) CK(cudaGraphLaunch(ex2, s));
        CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
        float msgl; CK(cudaEventElapsedTime(&msgl, e0, e1));
        printf("rank %d: LL protocol (%d blocks): stream launches %.2f us, in a graph %.2f us per exchange, data %s\n", rank, blocksLL,
               1e3 * msl / reps, 1e3 * msgl / 1000, okl ? "ok" : "WRONG");
    }
    if (rank == 0) { int st; while (wait(&st) > 0) {} }
    return 0;
}
