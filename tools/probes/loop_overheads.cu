// Micro-probe (not product code): what does one iteration of a device-side loop cost on B200?
//   A. CUDA-graph WHILE conditional node, body = K tiny kernels (last one decrements the condition)
//   B. plain graph of N x K tiny kernel nodes (no host involvement between them)
//   C. one persistent cooperative kernel with a grid-wide barrier per phase (G CTAs)
//   D. stream launches of tiny kernels back to back (host-issued)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o loop_overheads loop_overheads.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;

__global__ void tiny(int* c) { if (threadIdx.x == 0 && blockIdx.x == 0) c[1]++; }
__global__ void tiny_wide(float* x, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) x[i] = x[i] * 1.0001f + 1.f; }
__global__ void cond_step(int* c, cudaGraphConditionalHandle h) {
  if (threadIdx.x == 0) { int v = --c[0]; if (v <= 0) cudaGraphSetConditional(h, 0); }
}
__global__ void persistent(float* x, int n, int iters, int phases) {
  cg::grid_group g = cg::this_grid();
  for (int it = 0; it < iters; it++)
    for (int p = 0; p < phases; p++) {
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = x[i] * 1.0001f + 1.f;
      g.sync();
    }
}
#define CK(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { printf("ERR %s line %d: %s\n", #e, __LINE__, cudaGetErrorString(_e)); return 1; } } while (0)

int main() {
  cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int* d; CK(cudaMalloc(&d, 64)); float* x; const int n = 600 * 256; CK(cudaMalloc(&x, n * 4)); cudaMemset(x, 0, n * 4);
  float ms;
  const int N = 200;
  for (int K = 1; K <= 3; K++) {  // A
    cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle h; CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h;
    p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node; CK(cudaGraphAddNode(&node, g, nullptr, 0, &p));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    CK(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    for (int k = 0; k + 1 < K; k++) tiny_wide<<<600, 256, 0, s>>>(x, n);
    cond_step<<<1, 32, 0, s>>>(d, h);
    CK(cudaStreamEndCapture(s, nullptr));
    cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
    for (int rep = 0; rep < 3; rep++) {
      int init = N; cudaMemcpyAsync(d, &init, 4, cudaMemcpyHostToDevice, s);
      cudaEventRecord(e0, s); CK(cudaGraphLaunch(ex, s)); cudaEventRecord(e1, s); CK(cudaStreamSynchronize(s));
      cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("A while-graph   K=%d kernels/body: %.2f us per iteration (%d iterations)\n", K, ms * 1e3 / N, N);
    cudaGraphExecDestroy(ex); cudaGraphDestroy(g);
  }
  for (int K = 1; K <= 3; K++) {  // B
    cudaGraph_t g; CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
    for (int i = 0; i < N; i++) { for (int k = 0; k + 1 < K; k++) tiny_wide<<<600, 256, 0, s>>>(x, n); tiny<<<1, 32, 0, s>>>(d); }
    CK(cudaStreamEndCapture(s, &g));
    cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0, s); CK(cudaGraphLaunch(ex, s)); cudaEventRecord(e1, s); CK(cudaStreamSynchronize(s));
      cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("B plain graph   K=%d kernels/iter: %.2f us per iteration\n", K, ms * 1e3 / N);
    cudaGraphExecDestroy(ex); cudaGraphDestroy(g);
  }
  for (int G : {148, 296, 592}) {  // C
    int iters = N, phases = 1;
    void* args[] = {&x, (void*)&n, &iters, &phases};
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0, s);
      cudaError_t e = cudaLaunchCooperativeKernel((void*)persistent, dim3(G), dim3(256), args, 0, s);
      cudaEventRecord(e1, s); cudaStreamSynchronize(s);
      if (e != cudaSuccess) { printf("C coop launch G=%d: %s\n", G, cudaGetErrorString(e)); break; }
      cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("C persistent    G=%d CTAs: %.2f us per phase (grid.sync)\n", G, ms * 1e3 / N);
  }
  for (int K = 1; K <= 3; K++) {  // D
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0, s);
      for (int i = 0; i < N; i++) { for (int k = 0; k + 1 < K; k++) tiny_wide<<<600, 256, 0, s>>>(x, n); tiny<<<1, 32, 0, s>>>(d); }
      cudaEventRecord(e1, s); cudaStreamSynchronize(s);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("D stream launch K=%d kernels/iter: %.2f us per iteration\n", K, ms * 1e3 / N);
  }
  return 0;
}
