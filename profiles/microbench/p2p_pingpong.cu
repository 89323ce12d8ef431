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
    if (rank == 0) { int st; while (wait(&st) > 0) {} }
    return 0;
}
