// Probe: which piece of the degeneracy path differs between host and device?  (sym_eigen<6>, lu_inverse<6>)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "linalg.cuh"
using namespace loamb;
struct Out { float E[6], V[36], Vinv[36]; int ok; };
__global__ void dev_parts(const float* A, Out* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[36]; for (int k = 0; k < 36; k++) a[k] = A[36 * i + k];
  Out o;
  o.ok = sym_eigen<6>(a, o.E, o.V) ? 1 : 0;
  lu_inverse<6>(o.V, o.Vinv);
  out[i] = o;
}
int main() {
  const int n = 512;
  float* A = (float*)malloc(n * 36 * 4);
  srand(7);
  for (int m = 0; m < n; m++) {
    double acc[36] = {0};
    const int rows = 200 + rand() % 3000;
    const double wscale = (m % 2 == 0) ? 0.02 : 1.0;
    for (int r = 0; r < rows; r++) {
      double row[6];
      for (int k = 0; k < 3; k++) row[k] = 20.0 * ((rand() / (double)RAND_MAX) - 0.5);
      for (int k = 3; k < 6; k++) row[k] = 2.0 * ((rand() / (double)RAND_MAX) - 0.5);
      row[4] *= wscale;
      for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) acc[i + 6 * j] += row[i] * row[j];
    }
    for (int k = 0; k < 36; k++) A[m * 36 + k] = (float)acc[k];
  }
  float* dA; Out* dO;
  cudaMalloc(&dA, n * 36 * 4); cudaMalloc(&dO, n * sizeof(Out));
  cudaMemcpy(dA, A, n * 36 * 4, cudaMemcpyHostToDevice);
  dev_parts<<<(n + 63) / 64, 64>>>(dA, dO, n);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  Out* hO = (Out*)malloc(n * sizeof(Out));
  cudaMemcpy(hO, dO, n * sizeof(Out), cudaMemcpyDeviceToHost);
  int badE = 0, badV = 0, badI = 0, shown = 0;
  for (int m = 0; m < n; m++) {
    Out h;
    h.ok = sym_eigen<6>(A + 36 * m, h.E, h.V) ? 1 : 0;
    lu_inverse<6>(h.V, h.Vinv);
    const bool bE = memcmp(h.E, hO[m].E, sizeof h.E) != 0, bV = memcmp(h.V, hO[m].V, sizeof h.V) != 0;
    const bool bI = memcmp(h.Vinv, hO[m].Vinv, sizeof h.Vinv) != 0;
    badE += bE; badV += bV; badI += bI;
    if ((bE || bV) && shown < 3) {
      shown++;
      printf("m=%d ok host %d dev %d\n  E host %g %g %g %g %g %g\n  E dev  %g %g %g %g %g %g\n", m, h.ok, hO[m].ok, h.E[0], h.E[1], h.E[2], h.E[3],
             h.E[4], h.E[5], hO[m].E[0], hO[m].E[1], hO[m].E[2], hO[m].E[3], hO[m].E[4], hO[m].E[5]);
      printf("  V col0 host %g %g %g %g %g %g\n  V col0 dev  %g %g %g %g %g %g\n", h.V[0], h.V[1], h.V[2], h.V[3], h.V[4], h.V[5], hO[m].V[0],
             hO[m].V[1], hO[m].V[2], hO[m].V[3], hO[m].V[4], hO[m].V[5]);
    }
  }
  printf("%d matrices: mismatching E %d, V %d, Vinv %d\n", n, badE, badV, badI);
  return 0;
}
