// Warp-parallel form of the 6 x 6 Gauss-Newton step for the device-resident loops (lmstep.cuh: gn_solve is the serial
// form the host uses; BasicLaserOdometry.cpp:559-597, BasicLaserMapping.cpp:867-905).
//
// The serial step kept its matrices in local memory (416 B stack frame, every access an L1 round trip) and cost ~15 us
// per iteration on one thread -- more than the PCIe round trip it was meant to replace.  Here lanes 0..5 of a warp own
// the six columns of AtA and lane 6 owns AtB, everything in registers:
//   * column-pivoted Householder QR: pivot = shuffle arg-max of the down-dated column norms, the reflector is built by
//     the pivot lane and broadcast, every lane applies it to its own column (lane 6 carries Q^T b along);
//   * back substitution on values fetched by shuffles; the permutation is undone with predicated writes.
// Per element this is the SAME sequence of fp32 operations as colpiv_qr_solve<6, 6> (linalg.cuh), so the result equals
// the serial / host solve bit for bit (checked by tests/test_gpu_parity.py::test_warp_solver_equals_host_solver).
// The first-iteration degeneracy test (eigenvalues of AtA below a threshold) keeps the Cholesky shortcut; the rare full
// eigen-decomposition runs on lane 0 through __noinline__ wrappers (inlined into one big function, nvcc 12.9 produced a
// sym_eigen<6> that disagreed with the host's on ill-conditioned matrices -- tools/probes/gn_variants.cu).
#pragma once

#include "lmstep.cuh"

namespace loamb {

#if defined(__CUDACC__)

__device__ __noinline__ void gn_projection_noinline(const float* A_colmajor, float eig_thr, GnState* g) {
  float E[6], V[36], V2[36], Vinv[36];
  sym_eigen<6>(A_colmajor, E, V);  // ascending eigenvalues, V column-major (column = eigenvector)
  for (int i = 0; i < 36; i++) V2[i] = V[i];
  int degenerate = 0;
  for (int i = 0; i < 6; i++) {
    if (E[i] < eig_thr) {
      for (int j = 0; j < 6; j++) V2[i + j * 6] = 0.f;  // zero ROW i
      degenerate = 1;
    } else {
      break;
    }
  }
  lu_inverse<6>(V, Vinv);
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      float acc = 0.f;
      for (int k = 0; k < 6; k++) acc += Vinv[i + k * 6] * V2[k + j * 6];
      g->P[i * 6 + j] = acc;
    }
  g->degenerate = degenerate;
}

// index of AtA(i, j) in the 21-entry upper triangle (row-major) the iteration kernels reduce
__device__ __forceinline__ int tri_index(int i, int j) {
  const int r = i < j ? i : j, c = i < j ? j : i;
  return r * 6 - (r * (r - 1)) / 2 + (c - r);
}

// All 32 lanes call this with the 32 reduced sums in shared memory (s_r[0..20] = AtA upper triangle, s_r[21..26] = AtB).
// On return every lane holds x[0..5].  `first` = iteration 0 of the sweep (degeneracy test); g lives in global memory.
__device__ inline void gn_solve_warp(const float* s_r, bool first, float eig_thr, GnState* g, float x[6]) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  constexpr int R = 6;
  // my column of A (lanes 0..5) / b (lane 6); other lanes carry zeros and only take part in the shuffles
  float a[6];
#pragma unroll
  for (int i = 0; i < 6; i++) a[i] = lane < 6 ? s_r[tri_index(i, lane)] : (lane == 6 ? s_r[21 + i] : 0.f);

  // ---- first iteration: is any eigenvalue of AtA below eig_thr?  (same shortcut as gn_solve)
  if (first) {
    int need_eigen = 0;
    if (lane == 0) {
      float A[36];
#pragma unroll
      for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) A[i + j * 6] = s_r[tri_index(i, j)];
      float tr = 0.f;
#pragma unroll
      for (int i = 0; i < 6; i++) tr += A[i + i * 6];
      const float c = eig_thr + 1e-5f * tr;
      float L[36];
      bool regular = true;
#pragma unroll
      for (int j = 0; j < 6; j++) {
        if (regular) {
          float d = A[j + j * 6] - c;
#pragma unroll
          for (int k = 0; k < j; k++) d -= L[j + k * 6] * L[j + k * 6];
          if (!(d > 0.f)) {
            regular = false;
          } else {
            const float ld = sqrtf(d);
            L[j + j * 6] = ld;
#pragma unroll
            for (int i = j + 1; i < 6; i++) {
              float v = A[i + j * 6];
#pragma unroll
              for (int k = 0; k < j; k++) v -= L[i + k * 6] * L[j + k * 6];
              L[i + j * 6] = v / ld;
            }
          }
        }
      }
      if (regular) {
        g->degenerate = 0;
      } else {
        need_eigen = 1;
        gn_projection_noinline(A, eig_thr, g);
      }
    }
    need_eigen = __shfl_sync(FULL, need_eigen, 0);
    __syncwarp();
  }

  // ---- column-pivoted Householder QR, operations and their order as in colpiv_qr_solve<6, 6>
  float nDir, nUpd;
  {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < R; i++) s += a[i] * a[i];
    nDir = sqrtf(s);
    nUpd = nDir;
  }
  float maxNorm = lane < 6 ? nUpd : 0.f;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    const float other = __shfl_xor_sync(FULL, maxNorm, o);
    if (other > maxNorm) maxNorm = other;
  }
  maxNorm = __shfl_sync(FULL, maxNorm, 0);
  const float eps = FLT_EPSILON;
  const float thr0 = maxNorm * eps;
  const float threshold_helper = thr0 * thr0 / float(R);
  const float downdate_thr = sqrtf(eps);
  int nonzero = 6;
  int perm = lane;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    // pivot: first column among k..5 with the largest down-dated norm
    float bigNorm = (lane >= k && lane < 6) ? nUpd : -1.f;
    int big = lane;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const float on = __shfl_xor_sync(FULL, bigNorm, o);
      const int ob = __shfl_xor_sync(FULL, big, o);
      if (on > bigNorm || (on == bigNorm && ob < big)) { bigNorm = on; big = ob; }
    }
    bigNorm = __shfl_sync(FULL, bigNorm, 0);
    big = __shfl_sync(FULL, big, 0);
    if (nonzero == 6 && bigNorm * bigNorm < threshold_helper * float(R - k)) nonzero = k;
    // bring the pivot column to position k
    {
      const int src = lane == k ? big : (lane == big ? k : lane);
#pragma unroll
      for (int i = 0; i < 6; i++) a[i] = __shfl_sync(FULL, a[i], src);
      nUpd = __shfl_sync(FULL, nUpd, src);
      nDir = __shfl_sync(FULL, nDir, src);
      perm = __shfl_sync(FULL, perm, src);
    }
    // Householder vector of column k below the diagonal (lane k), LAPACK xLARFG convention
    float tau = 0.f;
    if (lane == k) {
      float tail = 0.f;
#pragma unroll
      for (int i = k + 1; i < R; i++) tail += a[i] * a[i];
      const float c0 = a[k];
      float beta;
      if (R - k == 1 || tail <= FLT_MIN) {
        tau = 0.f;
        beta = c0;
#pragma unroll
        for (int i = k + 1; i < R; i++) a[i] = 0.f;
      } else {
        beta = sqrtf(c0 * c0 + tail);
        if (c0 >= 0.f) beta = -beta;
        const float den = c0 - beta;
#pragma unroll
        for (int i = k + 1; i < R; i++) a[i] = a[i] / den;
        tau = (beta - c0) / beta;
      }
      a[k] = beta;
    }
    tau = __shfl_sync(FULL, tau, k);
    float v[6];
#pragma unroll
    for (int i = 0; i < 6; i++) v[i] = i > k ? __shfl_sync(FULL, a[i], k) : 0.f;
    // apply (I - tau v v^T) to the trailing columns and -- while the rank allows -- to b
    const bool trailing = lane > k && lane < 6;
    const bool rhs = lane == 6 && k < nonzero;
    if (trailing || rhs) {
      if (R - k == 1) {
        a[k] *= (1.f - tau);
      } else if (tau != 0.f) {
        float t = 0.f;
#pragma unroll
        for (int i = k + 1; i < R; i++) t += v[i] * a[i];
        t += a[k];
        a[k] -= tau * t;
#pragma unroll
        for (int i = k + 1; i < R; i++) a[i] -= tau * v[i] * t;
      }
    }
    // norm down-dating of the trailing columns
    if (trailing && nUpd != 0.f) {
      float t = fabsf(a[k]) / nUpd;
      t = (1.f + t) * (1.f - t);
      t = t < 0.f ? 0.f : t;
      const float ratio = nUpd / nDir;
      const float t2 = t * (ratio * ratio);
      if (t2 <= downdate_thr) {
        float s = 0.f;
#pragma unroll
        for (int i = k + 1; i < R; i++) s += a[i] * a[i];
        nDir = sqrtf(s);
        nUpd = nDir;
      } else {
        nUpd *= sqrtf(t);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 6; j++) x[j] = 0.f;
  if (nonzero > 0) {
    // back substitution on c = Q^T b (lane 6), every lane redundantly
    float bq[6];
#pragma unroll
    for (int i = 0; i < 6; i++) bq[i] = __shfl_sync(FULL, a[i], 6);
#pragma unroll
    for (int i = 5; i >= 0; i--) {
      const float diag = __shfl_sync(FULL, a[i], i);
      float vv = bq[i];
#pragma unroll
      for (int l = i + 1; l < 6; l++) {
        const float r_il = __shfl_sync(FULL, a[i], l);
        if (l < nonzero) vv -= r_il * bq[l];
      }
      if (i < nonzero) bq[i] = vv / diag;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const int pj = __shfl_sync(FULL, perm, i);
      if (i < nonzero) {
#pragma unroll
        for (int j = 0; j < 6; j++)
          if (pj == j) x[j] = bq[i];
      }
    }
  }
  // degenerate (decided on the first iteration of the sweep): x <- P x
  __syncwarp();
  if (__ldcg(&g->degenerate)) {
    float x2[6];
#pragma unroll
    for (int i = 0; i < 6; i++) x2[i] = x[i];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 6; k++) acc += __ldcg(&g->P[i * 6 + k]) * x2[k];
      x[i] = acc;
    }
  }
}

// rad2deg / x100 convergence measure of both loops (BasicLaserOdometry.cpp:614-621, BasicLaserMapping.cpp:910-921)
__device__ __forceinline__ void gn_deltas(const float x[6], float& deltaR, float& deltaT) {
  double r2 = 0.0, t2 = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double rd = (double)(float)((double)x[i] * 180.0 / 3.14159265358979323846);  // rad2deg returns float
    r2 += rd * rd;
    const double td = (double)(x[3 + i] * 100.f);
    t2 += td * td;
  }
  deltaR = (float)sqrt(r2);
  deltaT = (float)sqrt(t2);
}

#endif  // __CUDACC__

}  // namespace loamb
