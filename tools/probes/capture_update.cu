// Probe (not product code): host cost of issuing a fixed sequence of N small kernels per "sweep" on B200
//   A. direct stream launches
//   B. stream capture of the same calls -> cudaGraphExecUpdate of a resident executable graph -> one cudaGraphLaunch
// with 1 and with 3 host threads issuing concurrently (each on its own stream), i.e. the situation of the streaming
// pipeline where three stage threads share one CUDA context.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
__global__ void small(float* x, int n, float a) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) x[i] = x[i] * a + 1.f; }
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct Worker {
  cudaStream_t s; float* x; int n; cudaGraphExec_t exec = nullptr; double host_s = 0, total_s = 0;
};
static void issue(Worker& w, int N, float a) { for (int k = 0; k < N; k++) small<<<(w.n + 255) / 256, 256, 0, w.s>>>(w.x, w.n + k % 2, a); }
static void run(Worker& w, int N, int iters, bool graph) {
  cudaSetDevice(0);
  const double t0 = now();
  double host = 0;
  for (int it = 0; it < iters; it++) {
    const double h0 = now();
    if (!graph) {
      issue(w, N, 1.0f + 1e-6f * it);
    } else {
      cudaGraph_t g;
      cudaStreamBeginCapture(w.s, cudaStreamCaptureModeThreadLocal);
      issue(w, N, 1.0f + 1e-6f * it);
      cudaStreamEndCapture(w.s, &g);
      if (!w.exec) cudaGraphInstantiate(&w.exec, g, 0);
      else {
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(w.exec, g, &info) != cudaSuccess) { cudaGraphExecDestroy(w.exec); cudaGraphInstantiate(&w.exec, g, 0); }
      }
      cudaGraphLaunch(w.exec, w.s);
      cudaGraphDestroy(g);
    }
    host += now() - h0;
    cudaStreamSynchronize(w.s);  // a result round trip per "sweep", like the pipeline stages
  }
  w.host_s = host / iters;
  w.total_s = (now() - t0) / iters;
}
int main() {
  const int N = 30, iters = 300, n = 20000;
  for (int T : {1, 3}) {
    for (int mode = 0; mode < 2; mode++) {
      std::vector<Worker> ws(T);
      for (auto& w : ws) { cudaStreamCreateWithFlags(&w.s, cudaStreamNonBlocking); cudaMalloc(&w.x, (n + 8) * 4); cudaMemset(w.x, 0, (n + 8) * 4); w.n = n; }
      for (auto& w : ws) run(w, N, 5, mode == 1);  // warm-up (instantiation)
      std::vector<std::thread> th;
      const double t0 = now();
      for (auto& w : ws) th.emplace_back([&w, N, iters, mode] { run(w, N, iters, mode == 1); });
      for (auto& t : th) t.join();
      const double wall = (now() - t0) / iters;
      double host = 0; for (auto& w : ws) host += w.host_s;
      printf("threads %d  %-28s  host issue %.1f us / sweep / thread, sweep period %.1f us (%d kernels per sweep)\n", T,
             mode ? "capture + update + launch" : "direct launches", 1e6 * host / T, 1e6 * wall, N);
    }
  }
  return 0;
}
