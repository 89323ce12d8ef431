#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_set(cudaGraphConditionalHandle h, int* cnt) { if (threadIdx.x==0) { int v = --(*cnt); cudaGraphSetConditional(h, v > 0); } }
__global__ void k_body(int* x) { atomicAdd(x, 1); }
int main() {
  int *cnt, *x; cudaMalloc(&cnt, 4); cudaMalloc(&x, 4);
  int h5 = 5, z = 0; cudaMemcpy(cnt, &h5, 4, cudaMemcpyHostToDevice); cudaMemcpy(x, &z, 4, cudaMemcpyHostToDevice);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t n; cudaError_t e = cudaGraphAddNode(&n, g, nullptr, 0, &p); printf("add %d\n", e);
  cudaGraph_t body = p.conditional.phGraph_out[0];
  cudaStream_t s; cudaStreamCreate(&s);
  cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeGlobal);
  k_body<<<1,32,0,s>>>(x); k_set<<<1,32,0,s>>>(h, cnt);
  cudaStreamEndCapture(s, nullptr);
  cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst %d\n", e);
  cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
  cudaMemcpy(&z, x, 4, cudaMemcpyDeviceToHost); printf("x=%d\n", z);
}
