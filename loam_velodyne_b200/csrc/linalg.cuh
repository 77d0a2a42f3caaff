// Small fixed-size dense linear algebra shared by the CUDA kernels and the host-side solve.
//
// The reference leans on Eigen for four tiny dense problems:
//   * 3x3 symmetric eigen-decomposition per corner correspondence   (BasicLaserMapping.cpp:695)
//   * 5x3 least squares  A x = -1  per surface correspondence        (BasicLaserMapping.cpp:762-768)
//   * 6x6 solve of the normal equations, pivoted Householder QR      (BasicLaserMapping.cpp:867, BasicLaserOdometry.cpp:559)
//   * 6x6 symmetric eigen-decomposition + 6x6 inverse (degeneracy)   (BasicLaserMapping.cpp:875-898, BasicLaserOdometry.cpp:567-590)
// Eigen is not vendored by the reference; these follow the published algorithms it documents
// (column-pivoted Householder QR with LAPACK-style norm down-dating; Householder tridiagonalisation +
// implicit-shift symmetric QR; partial-pivot LU) in plain fp32, written once for host and device so the
// per-query fits done in kernels and the per-iteration solve done on the host share one arithmetic.
// All code is compiled with FMA contraction off (nvcc -fmad=false, host -ffp-contract=off).
#pragma once

#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define LB_HD __host__ __device__ __forceinline__
#else
#define LB_HD inline
#endif

namespace loamb {

LB_HD float lb_abs(float v) { return fabsf(v); }
LB_HD float lb_sqrt(float v) { return sqrtf(v); }
// hypot the way glibc evaluates it for float (exact products in double, one rounding at the end)
LB_HD float lb_hypot(float a, float b) { return (float)sqrt((double)a * (double)a + (double)b * (double)b); }

// ---------------------------------------------------------------- Householder primitives
// x[0..n) with stride 1: on return x[1..n) holds the essential part, tau/beta as in LAPACK xLARFG (sign convention
// beta = -sign(x0) * ||x||).
template <int MAXN>
LB_HD void householder_make(float* x, int n, float& tau, float& beta) {
  float tail = 0.f;
  for (int i = 1; i < n; i++) tail += x[i] * x[i];
  const float c0 = x[0];
  if (n == 1 || tail <= FLT_MIN) {
    tau = 0.f;
    beta = c0;
    for (int i = 1; i < n; i++) x[i] = 0.f;
  } else {
    beta = lb_sqrt(c0 * c0 + tail);
    if (c0 >= 0.f) beta = -beta;
    const float den = c0 - beta;
    for (int i = 1; i < n; i++) x[i] = x[i] / den;
    tau = (beta - c0) / beta;
  }
}

// column-major block A (rows x cols, leading dimension ld): A <- (I - tau v v^T) A, v = [1; ess]
LB_HD void householder_apply_left(float* A, int rows, int cols, int ld, const float* ess, float tau) {
  if (rows == 1) {
    for (int j = 0; j < cols; j++) A[j * ld] *= (1.f - tau);
  } else if (tau != 0.f) {
    for (int j = 0; j < cols; j++) {
      float* c = A + j * ld;
      float t = 0.f;
      for (int i = 1; i < rows; i++) t += ess[i - 1] * c[i];
      t += c[0];
      c[0] -= tau * t;
      for (int i = 1; i < rows; i++) c[i] -= tau * ess[i - 1] * t;
    }
  }
}

// ---------------------------------------------------------------- column-pivoted Householder QR solve
// Solves min ||A x - b|| for A (R x C, column-major, destroyed) and b (R, destroyed); x (C).
// Every loop bound is a compile-time constant and the pivot column is applied through predicated swaps with static
// indices, so after unrolling the whole factorisation lives in registers on the device (the per-query 5 x 3 plane fit
// runs ~17 k times per LM iteration; with a dynamically indexed pivot the arrays went to local memory and the fit
// dominated the kernel: profiles/r1_v3_map_iterate_grid_tpq.md).  The arithmetic (operations and their order) is
// unchanged, so host and device still agree bit for bit with each other and with the previous formulation.
template <int R, int C>
LB_HD void colpiv_qr_solve(float* A, float* b, float* x) {
  constexpr int SIZE = R < C ? R : C;
  float h[SIZE];
  float nUpd[C], nDir[C];
  int perm[C];
  float maxNorm = 0.f;
#pragma unroll
  for (int k = 0; k < C; k++) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < R; i++) s += A[i + k * R] * A[i + k * R];
    nDir[k] = lb_sqrt(s);
    nUpd[k] = nDir[k];
    if (nUpd[k] > maxNorm) maxNorm = nUpd[k];
    perm[k] = k;
  }
  const float eps = FLT_EPSILON;
  const float thr0 = maxNorm * eps;
  const float threshold_helper = thr0 * thr0 / float(R);
  const float downdate_thr = lb_sqrt(eps);
  int nonzero = SIZE;
#pragma unroll
  for (int k = 0; k < SIZE; k++) {
    int big = k;
    float bigNorm = nUpd[k];
#pragma unroll
    for (int j = k + 1; j < C; j++)
      if (nUpd[j] > bigNorm) { bigNorm = nUpd[j]; big = j; }
    if (nonzero == SIZE && bigNorm * bigNorm < threshold_helper * float(R - k)) nonzero = k;
    // bring the pivot column to position k (static indices, predicated on the pivot choice)
#pragma unroll
    for (int j = k + 1; j < C; j++) {
      if (big == j) {
#pragma unroll
        for (int i = 0; i < R; i++) { const float t = A[i + k * R]; A[i + k * R] = A[i + j * R]; A[i + j * R] = t; }
        float t = nUpd[k]; nUpd[k] = nUpd[j]; nUpd[j] = t;
        t = nDir[k]; nDir[k] = nDir[j]; nDir[j] = t;
        const int ti = perm[k]; perm[k] = perm[j]; perm[j] = ti;
      }
    }
    // Householder vector of column k below the diagonal (LAPACK xLARFG convention: beta = -sign(x0) * ||x||)
    float tau, beta;
    {
      float tail = 0.f;
#pragma unroll
      for (int i = k + 1; i < R; i++) tail += A[i + k * R] * A[i + k * R];
      const float c0 = A[k + k * R];
      if (R - k == 1 || tail <= FLT_MIN) {
        tau = 0.f;
        beta = c0;
#pragma unroll
        for (int i = k + 1; i < R; i++) A[i + k * R] = 0.f;
      } else {
        beta = lb_sqrt(c0 * c0 + tail);
        if (c0 >= 0.f) beta = -beta;
        const float den = c0 - beta;
#pragma unroll
        for (int i = k + 1; i < R; i++) A[i + k * R] = A[i + k * R] / den;
        tau = (beta - c0) / beta;
      }
    }
    h[k] = tau;
    A[k + k * R] = beta;
    // apply (I - tau v v^T) to the trailing columns
    if (R - k == 1) {
#pragma unroll
      for (int j = k + 1; j < C; j++) A[k + j * R] *= (1.f - tau);
    } else if (tau != 0.f) {
#pragma unroll
      for (int j = k + 1; j < C; j++) {
        float t = 0.f;
#pragma unroll
        for (int i = k + 1; i < R; i++) t += A[i + k * R] * A[i + j * R];
        t += A[k + j * R];
        A[k + j * R] -= tau * t;
#pragma unroll
        for (int i = k + 1; i < R; i++) A[i + j * R] -= tau * A[i + k * R] * t;
      }
    }
#pragma unroll
    for (int j = k + 1; j < C; j++) {
      if (nUpd[j] != 0.f) {
        float t = lb_abs(A[k + j * R]) / nUpd[j];
        t = (1.f + t) * (1.f - t);
        t = t < 0.f ? 0.f : t;
        const float ratio = nUpd[j] / nDir[j];
        const float t2 = t * (ratio * ratio);
        if (t2 <= downdate_thr) {
          float s = 0.f;
#pragma unroll
          for (int i = k + 1; i < R; i++) s += A[i + j * R] * A[i + j * R];
          nDir[j] = lb_sqrt(s);
          nUpd[j] = nDir[j];
        } else {
          nUpd[j] *= lb_sqrt(t);
        }
      }
    }
  }
  if (nonzero == 0) {
#pragma unroll
    for (int i = 0; i < C; i++) x[i] = 0.f;
    return;
  }
  // c = Q^T b (reflectors 0 .. nonzero-1), back substitution on the leading nonzero x nonzero triangle
#pragma unroll
  for (int k = 0; k < SIZE; k++) {
    if (k < nonzero) {
      const float tau = h[k];
      if (R - k == 1) {
        b[k] *= (1.f - tau);
      } else if (tau != 0.f) {
        float t = 0.f;
#pragma unroll
        for (int i = k + 1; i < R; i++) t += A[i + k * R] * b[i];
        t += b[k];
        b[k] -= tau * t;
#pragma unroll
        for (int i = k + 1; i < R; i++) b[i] -= tau * A[i + k * R] * t;
      }
    }
  }
#pragma unroll
  for (int i = SIZE - 1; i >= 0; i--) {
    if (i < nonzero) {
      float v = b[i];
#pragma unroll
      for (int l = i + 1; l < SIZE; l++)
        if (l < nonzero) v -= A[i + l * R] * b[l];
      b[i] = v / A[i + i * R];
    }
  }
#pragma unroll
  for (int j = 0; j < C; j++) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < SIZE; i++)
      if (i < nonzero && perm[i] == j) v = b[i];
    x[j] = v;
  }
}

// ---------------------------------------------------------------- symmetric eigen-solver
struct GivensRot {
  float c, s;
  LB_HD void make(float p, float q) {
    if (q == 0.f) {
      c = p < 0.f ? -1.f : 1.f;
      s = 0.f;
    } else if (p == 0.f) {
      c = 0.f;
      s = q < 0.f ? 1.f : -1.f;
    } else if (lb_abs(p) > lb_abs(q)) {
      float t = q / p;
      float u = lb_sqrt(1.f + t * t);
      if (p < 0.f) u = -u;
      c = 1.f / u;
      s = -t * c;
    } else {
      float t = p / q;
      float u = lb_sqrt(1.f + t * t);
      if (q < 0.f) u = -u;
      s = -1.f / u;
      c = -t * s;
    }
  }
};

template <int N>
LB_HD void tridiag_qr_step(float* diag, float* sub, int start, int end, float* Q) {
  float td = (diag[end - 1] - diag[end]) * 0.5f;
  float e = sub[end - 1];
  float mu = diag[end];
  if (td == 0.f) {
    mu -= lb_abs(e);
  } else if (e != 0.f) {
    const float e2 = e * e;
    const float h = lb_hypot(td, e);
    if (e2 == 0.f)
      mu -= e / ((td + (td > 0.f ? h : -h)) / e);
    else
      mu -= e2 / (td + (td > 0.f ? h : -h));
  }
  float x = diag[start] - mu;
  float z = sub[start];
  for (int k = start; k < end && z != 0.f; ++k) {
    GivensRot r;
    r.make(x, z);
    const float sdk = r.s * diag[k] + r.c * sub[k];
    const float dkp1 = r.s * sub[k] + r.c * diag[k + 1];
    diag[k] = r.c * (r.c * diag[k] - r.s * sub[k]) - r.s * (r.c * sub[k] - r.s * diag[k + 1]);
    diag[k + 1] = r.s * sdk + r.c * dkp1;
    sub[k] = r.c * sdk - r.s * dkp1;
    if (k > start) sub[k - 1] = r.c * sub[k - 1] - r.s * z;
    x = sub[k];
    if (k < end - 1) {
      z = -r.s * sub[k + 1];
      sub[k + 1] = r.c * sub[k + 1];
    }
    float* qk = Q + k * N;
    float* qk1 = Q + (k + 1) * N;
    for (int i = 0; i < N; i++) {
      const float xi = qk[i], yi = qk1[i];
      qk[i] = r.c * xi - r.s * yi;
      qk1[i] = r.s * xi + r.c * yi;
    }
  }
}

// A: N x N symmetric, column-major, only the lower triangle is read. evals ascending, evecs column-major
// (column j = eigenvector of evals[j]).  Returns false when the QR iteration did not converge (evals/evecs then
// hold the unsorted partial result, as the solver the reference uses would leave them).
template <int N>
LB_HD bool sym_eigen(const float* Ain, float* evals, float* evecs) {
  float m[N * N];
  float scale = 0.f;
  for (int j = 0; j < N; j++)
    for (int i = 0; i < N; i++) {
      m[i + j * N] = (i >= j) ? Ain[i + j * N] : 0.f;
      if (i >= j && lb_abs(m[i + j * N]) > scale) scale = lb_abs(m[i + j * N]);
    }
  if (scale == 0.f) scale = 1.f;
  for (int j = 0; j < N; j++)
    for (int i = j; i < N; i++) m[i + j * N] /= scale;

  float sub[N > 1 ? N - 1 : 1];
  float* diag = evals;
  if (N == 3) {
    // closed-form 3x3 tridiagonalisation
    diag[0] = m[0];
    const float v1norm2 = m[2] * m[2];
    if (v1norm2 <= FLT_MIN) {
      diag[1] = m[1 + 1 * N];
      diag[2] = m[2 + 2 * N];
      sub[0] = m[1];
      sub[1] = m[2 + 1 * N];
      for (int i = 0; i < N * N; i++) evecs[i] = 0.f;
      for (int i = 0; i < N; i++) evecs[i + i * N] = 1.f;
    } else {
      const float beta = lb_sqrt(m[1] * m[1] + v1norm2);
      const float invBeta = 1.f / beta;
      const float m01 = m[1] * invBeta;
      const float m02 = m[2] * invBeta;
      const float q = 2.f * m01 * m[2 + 1 * N] + m02 * (m[2 + 2 * N] - m[1 + 1 * N]);
      diag[1] = m[1 + 1 * N] + m02 * q;
      diag[2] = m[2 + 2 * N] - m02 * q;
      sub[0] = beta;
      sub[1] = m[2 + 1 * N] - m01 * q;
      evecs[0] = 1.f; evecs[1] = 0.f;       evecs[2] = 0.f;
      evecs[3] = 0.f; evecs[4] = m01;       evecs[5] = m02;
      evecs[6] = 0.f; evecs[7] = m02;       evecs[8] = -m01;
    }
  } else {
    float hc[N > 1 ? N - 1 : 1], p[N];
    for (int i = 0; i < N - 1; ++i) {
      const int rs = N - i - 1;
      float h, beta;
      float* v = m + (i + 1) + i * N;
      householder_make<N>(v, rs, h, beta);
      v[0] = 1.f;
      for (int r = 0; r < rs; r++) {
        float acc = 0.f;
        for (int c = 0; c < rs; c++) {
          const float a = (r >= c) ? m[(i + 1 + r) + (i + 1 + c) * N] : m[(i + 1 + c) + (i + 1 + r) * N];
          acc += a * (h * v[c]);
        }
        p[r] = acc;
      }
      float dot = 0.f;
      for (int r = 0; r < rs; r++) dot += p[r] * v[r];
      const float alpha = h * -0.5f * dot;
      for (int r = 0; r < rs; r++) p[r] += alpha * v[r];
      for (int c = 0; c < rs; c++)
        for (int r = c; r < rs; r++) m[(i + 1 + r) + (i + 1 + c) * N] -= (v[r] * p[c] + p[r] * v[c]);
      v[0] = beta;
      hc[i] = h;
    }
    for (int i = 0; i < N; i++) diag[i] = m[i + i * N];
    for (int i = 0; i < N - 1; i++) sub[i] = m[(i + 1) + i * N];
    for (int i = 0; i < N * N; i++) evecs[i] = 0.f;
    for (int i = 0; i < N; i++) evecs[i + i * N] = 1.f;
    for (int k = N - 2; k >= 0; --k) {
      const int corner = N - k - 1;
      householder_apply_left(evecs + (k + 1) + (k + 1) * N, corner, corner, N, m + (k + 2) + k * N, hc[k]);
    }
  }

  int end = N - 1, start = 0, iter = 0;
  const float precision = 2.f * FLT_EPSILON;
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (lb_abs(sub[i]) <= (lb_abs(diag[i]) + lb_abs(diag[i + 1])) * precision || lb_abs(sub[i]) <= FLT_MIN)
        sub[i] = 0.f;
    while (end > 0 && sub[end - 1] == 0.f) end--;
    if (end <= 0) break;
    iter++;
    if (iter > 30 * N) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.f) start--;
    tridiag_qr_step<N>(diag, sub, start, end, evecs);
  }
  const bool ok = iter <= 30 * N;
  if (ok) {
    for (int i = 0; i < N - 1; ++i) {
      int k = 0;
      float mn = diag[i];
      for (int j = 1; j < N - i; j++)
        if (diag[i + j] < mn) { mn = diag[i + j]; k = j; }
      if (k > 0) {
        float t = diag[i]; diag[i] = diag[k + i]; diag[k + i] = t;
        for (int r = 0; r < N; r++) {
          t = evecs[r + i * N]; evecs[r + i * N] = evecs[r + (k + i) * N]; evecs[r + (k + i) * N] = t;
        }
      }
    }
  }
  for (int i = 0; i < N; i++) diag[i] *= scale;
  return ok;
}

// ---------------------------------------------------------------- partial-pivot LU inverse (column-major in/out)
template <int N>
LB_HD void lu_inverse(const float* Ain, float* inv) {
  float lu[N * N];
  int piv[N];
  for (int i = 0; i < N * N; i++) lu[i] = Ain[i];
  for (int k = 0; k < N; k++) {
    int p = k;
    float best = lb_abs(lu[k + k * N]);
    for (int i = k + 1; i < N; i++)
      if (lb_abs(lu[i + k * N]) > best) { best = lb_abs(lu[i + k * N]); p = i; }
    piv[k] = p;
    if (p != k)
      for (int j = 0; j < N; j++) { float t = lu[k + j * N]; lu[k + j * N] = lu[p + j * N]; lu[p + j * N] = t; }
    if (lu[k + k * N] != 0.f)
      for (int i = k + 1; i < N; i++) lu[i + k * N] /= lu[k + k * N];
    for (int j = k + 1; j < N; j++)
      for (int i = k + 1; i < N; i++) lu[i + j * N] -= lu[i + k * N] * lu[k + j * N];
  }
  for (int i = 0; i < N * N; i++) inv[i] = 0.f;
  for (int i = 0; i < N; i++) inv[i + i * N] = 1.f;
  for (int k = 0; k < N; k++)
    if (piv[k] != k)
      for (int j = 0; j < N; j++) { float t = inv[k + j * N]; inv[k + j * N] = inv[piv[k] + j * N]; inv[piv[k] + j * N] = t; }
  for (int j = 0; j < N; j++) {
    for (int i = 0; i < N; i++) {
      float v = inv[i + j * N];
      for (int l = 0; l < i; l++) v -= lu[i + l * N] * inv[l + j * N];
      inv[i + j * N] = v;
    }
    for (int i = N - 1; i >= 0; i--) {
      float v = inv[i + j * N];
      for (int l = i + 1; l < N; l++) v -= lu[i + l * N] * inv[l + j * N];
      inv[i + j * N] = v / lu[i + i * N];
    }
  }
}

}  // namespace loamb
