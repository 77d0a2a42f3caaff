// Probe: variants of gn_solve on the device against the host's gn_solve
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "lmstep.cuh"
namespace loamb {
__host__ __device__ __noinline__ void eig6_noinline(const float* A, float* E, float* V) { sym_eigen<6>(A, E, V); }
__host__ __device__ __noinline__ void inv6_noinline(const float* V, float* Vinv) { lu_inverse<6>(V, Vinv); }
template <int VARIANT> LOAMB_HD inline void gn_solve_v(const float* AtA_rowmajor, const float* AtB, bool first, float eig_thr, GnState& g, float x[6]) {
  float A[36], b[6];
  for (int i = 0; i < 6; i++) {
    b[i] = AtB[i];
    for (int j = 0; j < 6; j++) A[i + j * 6] = AtA_rowmajor[i * 6 + j];
  }
  float Aq[36];
  for (int i = 0; i < 36; i++) Aq[i] = A[i];
  colpiv_qr_solve<6, 6>(Aq, b, x);
  // First iteration: is any eigenvalue of AtA below eig_thr?  If AtA - c I (c = eig_thr plus a margin well above the
  // fp32 noise of a 6 x 6 eigen-decomposition) has a Cholesky factorisation, every eigenvalue exceeds c and the
  // reference's own test (:567-588 / :875-896) cannot fire: skip the eigen-decomposition (the projection is unused).
  bool surely_regular = false;
  if (first && !(VARIANT & 2)) {
    float L[36];
    float tr = 0.f;
    for (int i = 0; i < 6; i++) tr += A[i + i * 6];
    const float c = eig_thr + 1e-5f * tr;
    surely_regular = true;
    for (int j = 0; j < 6 && surely_regular; j++) {
      float d = A[j + j * 6] - c;
      for (int k = 0; k < j; k++) d -= L[j + k * 6] * L[j + k * 6];
      if (!(d > 0.f)) { surely_regular = false; break; }
      const float ld = sqrtf(d);
      L[j + j * 6] = ld;
      for (int i = j + 1; i < 6; i++) {
        float v = A[i + j * 6];
        for (int k = 0; k < j; k++) v -= L[i + k * 6] * L[j + k * 6];
        L[i + j * 6] = v / ld;
      }
    }
    if (surely_regular) g.degenerate = 0;
  }
  if (first && !surely_regular) {
    float E[6], V[36], V2[36];
    if (VARIANT & 1) eig6_noinline(A, E, V); else sym_eigen<6>(A, E, V);  // ascending eigenvalues, V column-major (column = eigenvector)
    for (int i = 0; i < 36; i++) V2[i] = V[i];
    g.degenerate = 0;
    for (int i = 0; i < 6; i++) {
      if (E[i] < eig_thr) {
        for (int j = 0; j < 6; j++) V2[i + j * 6] = 0.f;  // zero ROW i
        g.degenerate = 1;
      } else {
        break;
      }
    }
    float Vinv[36];
    if (VARIANT & 1) inv6_noinline(V, Vinv); else lu_inverse<6>(V, Vinv);
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) {
        float acc = 0.f;
        for (int k = 0; k < 6; k++) acc += Vinv[i + k * 6] * V2[k + j * 6];
        g.P[i * 6 + j] = acc;
      }
  }
  if (g.degenerate) {
    float x2[6];
    for (int i = 0; i < 6; i++) x2[i] = x[i];
    for (int i = 0; i < 6; i++) {
      float acc = 0.f;
      for (int k = 0; k < 6; k++) acc += g.P[i * 6 + k] * x2[k];
      x[i] = acc;
    }
  }
}


}
using namespace loamb;
struct Out { float x[6]; int deg; };
template <int VARIANT>
__global__ void dev_solve(const float* AtA, const float* AtB, float thr, Out* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  GnState g; g.degenerate = 0; for (int k = 0; k < 36; k++) g.P[k] = 0.f;
  float x[6];
  gn_solve_v<VARIANT>(AtA + 36 * i, AtB + 6 * i, true, thr, g, x);
  for (int k = 0; k < 6; k++) out[i].x[k] = x[k];
  out[i].deg = g.degenerate;
}
int main() {
  const int n = 1000;
  float* A = (float*)malloc(n * 36 * 4); float* B = (float*)malloc(n * 6 * 4);
  srand(7);
  for (int m = 0; m < n; m++) {
    double acc[36] = {0}, accb[6] = {0};
    const int rows = 200 + rand() % 3000;
    const double wscale = (m % 4 == 0) ? 0.02 : 1.0;
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
  for (int variant = 0; variant < 4; variant++)
    for (float thr : {10.f, 100.f, 1e5f}) {
      if (variant == 0) dev_solve<0><<<(n + 63) / 64, 64>>>(dA, dB, thr, dO, n);
      if (variant == 1) dev_solve<1><<<(n + 63) / 64, 64>>>(dA, dB, thr, dO, n);
      if (variant == 2) dev_solve<2><<<(n + 63) / 64, 64>>>(dA, dB, thr, dO, n);
      if (variant == 3) dev_solve<3><<<(n + 63) / 64, 64>>>(dA, dB, thr, dO, n);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hO, dO, n * sizeof(Out), cudaMemcpyDeviceToHost);
      int bad_x = 0, bad_deg = 0, ndeg = 0, bad_xh = 0;
      for (int m = 0; m < n; m++) {
        GnState g; g.degenerate = 0; memset(g.P, 0, sizeof g.P);
        float x[6];
        gn_solve(A + 36 * m, B + 6 * m, true, thr, g, x);  // the library's host arithmetic
        GnState g2; g2.degenerate = 0; memset(g2.P, 0, sizeof g2.P);
        float x2[6];
        if (variant == 0) gn_solve_v<0>(A + 36 * m, B + 6 * m, true, thr, g2, x2);
        if (variant == 1) gn_solve_v<1>(A + 36 * m, B + 6 * m, true, thr, g2, x2);
        if (variant == 2) gn_solve_v<2>(A + 36 * m, B + 6 * m, true, thr, g2, x2);
        if (variant == 3) gn_solve_v<3>(A + 36 * m, B + 6 * m, true, thr, g2, x2);
        ndeg += g.degenerate;
        bad_x += memcmp(x, hO[m].x, sizeof x) != 0;
        bad_deg += g.degenerate != hO[m].deg;
        bad_xh += memcmp(x, x2, sizeof x) != 0;
      }
      printf("variant %d (noinline %d, no-cholesky %d) thr %g: host degenerate %d; device mismatches x %d deg %d; host variant vs host %d\n",
             variant, variant & 1, (variant >> 1) & 1, thr, ndeg, bad_x, bad_deg, bad_xh);
    }
  return 0;
}
