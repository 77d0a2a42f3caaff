// Probe (not product code): gn_solve (csrc/lmstep.cuh) on host vs device over random normal-equation matrices, first
// iteration (degeneracy test active), thresholds 10 / 100 / 1e5 (the last forces the eigen-decomposition path).
// nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -ffp-contract=off -I loam_velodyne_b200/csrc -o eigen_check eigen_check.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "lmstep.cuh"
using namespace loamb;
struct Out { float x[6]; int deg; float P[36]; };
__global__ void dev_solve(const float* AtA, const float* AtB, float thr, Out* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  GnState g; g.degenerate = 0; for (int k = 0; k < 36; k++) g.P[k] = 0.f;
  float x[6];
  gn_solve(AtA + 36 * i, AtB + 6 * i, true, thr, g, x);
  for (int k = 0; k < 6; k++) out[i].x[k] = x[k];
  out[i].deg = g.degenerate;
  for (int k = 0; k < 36; k++) out[i].P[k] = g.P[k];
}
int main() {
  const int n = 2000;
  float* A = (float*)malloc(n * 36 * 4); float* B = (float*)malloc(n * 6 * 4);
  srand(7);
  for (int m = 0; m < n; m++) {
    double acc[36] = {0}, accb[6] = {0};
    const int rows = 200 + rand() % 3000;
    const double wscale = (m % 4 == 0) ? 0.02 : 1.0;  // some matrices with one weak direction
    for (int r = 0; r < rows; r++) {
      double row[6];
      for (int k = 0; k < 3; k++) row[k] = 20.0 * ((rand() / (double)RAND_MAX) - 0.5);
      for (int k = 3; k < 6; k++) row[k] = 2.0 * ((rand() / (double)RAND_MAX) - 0.5);
      row[4] *= wscale;
      const double b = 0.1 * ((rand() / (double)RAND_MAX) - 0.5);
      for (int i = 0; i < 6; i++) { accb[i] += row[i] * b; for (int j = 0; j < 6; j++) acc[i * 6 + j] += row[i] * row[j]; }
    }
    for (int k = 0; k < 36; k++) A[m * 36 + k] = (float)acc[k];
    for (int k = 0; k < 6; k++) B[m * 6 + k] = (float)accb[k];
  }
  float *dA, *dB; Out* dO;
  cudaMalloc(&dA, n * 36 * 4); cudaMalloc(&dB, n * 6 * 4); cudaMalloc(&dO, n * sizeof(Out));
  cudaMemcpy(dA, A, n * 36 * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B, n * 6 * 4, cudaMemcpyHostToDevice);
  Out* hO = (Out*)malloc(n * sizeof(Out));
  for (float thr : {10.f, 100.f, 1e5f}) {
    dev_solve<<<(n + 63) / 64, 64>>>(dA, dB, thr, dO, n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hO, dO, n * sizeof(Out), cudaMemcpyDeviceToHost);
    int bad_x = 0, bad_deg = 0, bad_p = 0, ndeg = 0, first_bad = -1;
    for (int m = 0; m < n; m++) {
      GnState g; g.degenerate = 0; memset(g.P, 0, sizeof g.P);
      float x[6];
      gn_solve(A + 36 * m, B + 6 * m, true, thr, g, x);
      ndeg += g.degenerate;
      const bool bx = memcmp(x, hO[m].x, sizeof x) != 0, bd = g.degenerate != hO[m].deg;
      const bool bp = g.degenerate && memcmp(g.P, hO[m].P, sizeof g.P) != 0;
      bad_x += bx; bad_deg += bd; bad_p += bp;
      if ((bx || bd || bp) && first_bad < 0) {
        first_bad = m;
        printf("  first mismatch m=%d: host deg %d x %g %g %g %g %g %g | dev deg %d x %g %g %g %g %g %g\n", m, g.degenerate, x[0], x[1],
               x[2], x[3], x[4], x[5], hO[m].deg, hO[m].x[0], hO[m].x[1], hO[m].x[2], hO[m].x[3], hO[m].x[4], hO[m].x[5]);
      }
    }
    printf("thr %g: %d matrices, %d degenerate on host; mismatches x %d deg %d P %d\n", thr, n, ndeg, bad_x, bad_deg, bad_p);
  }
  return 0;
}
