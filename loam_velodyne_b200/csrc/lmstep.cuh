// The Gauss-Newton loops of the odometry and mapping stages with the pose kept ON THE DEVICE between iterations.
//
// The reference solves a 6 x 6 system per iteration on the host and the first version here mirrored that with one
// kernel + one 128-byte readback + one host solve per iteration (BasicLaserOdometry.cpp:559-622,
// BasicLaserMapping.cpp:867-922): 6-9 host round trips of ~12 us per sweep with a 7 us kernel in between.  Here a
// one-warp step kernel behind every iteration kernel solves the system (pivoted Householder QR), applies the degeneracy
// projection of the first iteration, updates the pose, tests convergence and prepares the next iteration's arguments
// (sin / cos of the new angles, the Jacobian coefficient products in the reference's order) in a small state block in
// global memory.  The host enqueues iteration + step kernels back to back; kernels behind the converged iteration
// return at once.  (Running the step inside the iteration kernel's last CTA was tried first: it put the solver's
// registers and stack into the hot kernels and was slower than the host round trip it replaced.)
// gn_solve() is one source for this path and for the host-side solver of the per-iteration API.
#pragma once

#include "linalg.cuh"

#ifndef LOAMB_HD
#if defined(__CUDACC__)
#define LOAMB_HD __host__ __device__
#else
#define LOAMB_HD
#endif
#endif

namespace loamb {

struct GnState {
  float P[36];     // row-major projection V^-1 V' of the first iteration (:567-590 / :875-898)
  int degenerate;
};

// x = argmin |A x - b| through AtA x = AtB (colPivHouseholderQr), first iteration: eigenvalues of AtA below eig_thr
// (ascending, stop at the first that is not) zero ROWS of the eigenvector matrix copy, P = V^-1 V2; degenerate -> x = P x
LOAMB_HD inline void gn_solve(const float* AtA_rowmajor, const float* AtB, bool first, float eig_thr, GnState& g, float x[6]) {
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
  if (first) {
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
    sym_eigen<6>(A, E, V);  // ascending eigenvalues, V column-major (column = eigenvector)
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
    lu_inverse<6>(V, Vinv);
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

// first words of both state blocks: what the host reads back
struct LmHeader {
  float rot[3], pos[3];
  int iter;       // index of the next iteration
  int done;       // converged or iteration cap reached
  int iters_run;  // the reference's iterCount + 1 of the last executed iteration
  int mb_seq;     // sequence number the finished loop posts into the host mailbox (0 = none)
  int pad[2];
};

#if defined(__CUDACC__)
// Loop control of the device-resident Gauss-Newton loops.  The loop itself is a CUDA-graph WHILE node whose body holds
// the iteration kernels; `handle` is its condition (0 = the kernels were launched outside a graph: nothing to set).
// When the loop ends the 9 header words go to the mapped host mailbox followed by the sequence word the host spins on
// (one PCIe round trip per LOOP instead of one per iteration).
__device__ inline void lm_loop_control(const LmHeader& h, unsigned long long handle, float* mailbox_host) {
  if (handle) cudaGraphSetConditional((cudaGraphConditionalHandle)handle, h.done ? 0u : 1u);
  if (h.done && mailbox_host && h.mb_seq) {
    const int* w = reinterpret_cast<const int*>(&h);
    volatile int* o = reinterpret_cast<volatile int*>(mailbox_host);
    for (int i = 0; i < 9; i++) o[i] = w[i];
    __threadfence_system();
    o[32] = h.mb_seq;
  }
}
#endif

}  // namespace loamb
