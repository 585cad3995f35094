// Per-frame evaluation errors on device (BASELINE config 5): MPJPE (evaluate/eval.py:222) and
// PA-MPJPE (utils/pose_utils.py:10-75: centre, K = X1 X2^T, 3x3 SVD, det-sign fix, scale
// tr(RK)/var1, translate).  The reference ships the joints to the host and loops NumPy
// `np.linalg.svd` per frame in float64; here one thread handles one frame, also in float64 (B200
// keeps full-rate FP64), with a one-sided Jacobi SVD of the 3x3 cross-covariance.
#pragma once
#include "common.cuh"

namespace whmr {

constexpr int kMaxEvalJoints = 64;

// One-sided (Hestenes) Jacobi: rotates columns of A (3x3, column-major a[c][r]) until orthogonal,
// accumulating V.  On exit A = U*diag(s) column-wise.
__device__ inline void svd3_jacobi(double a[3][3], double v[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 2; ++p) {
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int r = 0; r < 3; ++r) {
          alpha += a[p][r] * a[p][r]; beta += a[q][r] * a[q][r]; gamma += a[p][r] * a[q][r];
        }
        off = fmax(off, fabs(gamma) / (sqrt(alpha * beta) + 1e-300));
        if (fabs(gamma) < 1e-300) continue;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int r = 0; r < 3; ++r) {
          const double ap = a[p][r], aq = a[q][r];
          a[p][r] = c * ap - s * aq; a[q][r] = s * ap + c * aq;
          const double vp = v[p][r], vq = v[q][r];
          v[p][r] = c * vp - s * vq; v[q][r] = s * vp + c * vq;
        }
      }
    }
    if (off < 1e-15) break;
  }
}

__global__ void __launch_bounds__(128)
joint_errors_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int n, int J,
                    float* __restrict__ mpjpe, float* __restrict__ pa_mpjpe) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* P = pred + (size_t)i * J * 3;
  const float* G = gt + (size_t)i * J * 3;
  if (mpjpe) {
    double acc = 0.0;
    for (int j = 0; j < J; ++j) {
      const double dx = (double)P[j * 3] - G[j * 3], dy = (double)P[j * 3 + 1] - G[j * 3 + 1],
                   dz = (double)P[j * 3 + 2] - G[j * 3 + 2];
      acc += sqrt(dx * dx + dy * dy + dz * dz);
    }
    mpjpe[i] = (float)(acc / J);
  }
  if (!pa_mpjpe) return;
  double mu1[3] = {0, 0, 0}, mu2[3] = {0, 0, 0};
  for (int j = 0; j < J; ++j)
    for (int c = 0; c < 3; ++c) { mu1[c] += P[j * 3 + c]; mu2[c] += G[j * 3 + c]; }
  for (int c = 0; c < 3; ++c) { mu1[c] /= J; mu2[c] /= J; }
  // K = X1 X2^T (3x3), var1 = sum X1^2
  double K[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, var1 = 0.0;
  for (int j = 0; j < J; ++j) {
    double x1[3], x2[3];
    for (int c = 0; c < 3; ++c) { x1[c] = P[j * 3 + c] - mu1[c]; x2[c] = G[j * 3 + c] - mu2[c]; }
    for (int r = 0; r < 3; ++r) {
      var1 += x1[r] * x1[r];
      for (int c = 0; c < 3; ++c) K[r][c] += x1[r] * x2[c];
    }
  }
  // SVD K = U S V^T via one-sided Jacobi on the columns of K
  double a[3][3], v[3][3];   // a[c][r] = column c of K
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) a[c][r] = K[r][c];
  svd3_jacobi(a, v);
  double s[3], U[3][3];      // U[c][r] column c
  for (int c = 0; c < 3; ++c) s[c] = sqrt(a[c][0] * a[c][0] + a[c][1] * a[c][1] + a[c][2] * a[c][2]);
  int imin = 0;
  for (int c = 1; c < 3; ++c) if (s[c] < s[imin]) imin = c;
  const int i0 = (imin + 1) % 3, i1 = (imin + 2) % 3;
  for (int c = 0; c < 3; ++c) {
    if (c == imin) continue;
    const double inv = s[c] > 0 ? 1.0 / s[c] : 0.0;
    for (int r = 0; r < 3; ++r) U[c][r] = a[c][r] * inv;
  }
  // smallest-sigma left vector from the cross product (robust for rank-deficient K); its sign is
  // absorbed by the det fix below because V's matching column keeps its own sign.
  U[imin][0] = U[i0][1] * U[i1][2] - U[i0][2] * U[i1][1];
  U[imin][1] = U[i0][2] * U[i1][0] - U[i0][0] * U[i1][2];
  U[imin][2] = U[i0][0] * U[i1][1] - U[i0][1] * U[i1][0];
  // make (U, V) a consistent SVD pair for column imin: u_min should be K v_min / s_min when s_min > 0
  {
    double kv[3] = {0, 0, 0};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) kv[r] += K[r][c] * v[imin][c];
    const double dot = kv[0] * U[imin][0] + kv[1] * U[imin][1] + kv[2] * U[imin][2];
    if (dot < 0) for (int r = 0; r < 3; ++r) U[imin][r] = -U[imin][r];
  }
  // R = V Z U^T with Z = diag(1,1,sign(det(U V^T))) on the smallest singular direction
  auto det3 = [](double m[3][3]) {   // m[c][r] columns
    return m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) +
           m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
  };
  const double sgn = (det3(U) * det3(v)) < 0 ? -1.0 : 1.0;
  double R[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k) acc += v[k][r] * (k == imin ? sgn : 1.0) * U[k][c];
      R[r][c] = acc;
    }
  double trRK = 0.0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) trRK += R[r][c] * K[c][r];
  const double scale = trRK / var1;
  double t[3];
  for (int r = 0; r < 3; ++r)
    t[r] = mu2[r] - scale * (R[r][0] * mu1[0] + R[r][1] * mu1[1] + R[r][2] * mu1[2]);
  double acc = 0.0;
  for (int j = 0; j < J; ++j) {
    double d2 = 0.0;
    for (int r = 0; r < 3; ++r) {
      const double h = scale * (R[r][0] * P[j * 3] + R[r][1] * P[j * 3 + 1] + R[r][2] * P[j * 3 + 2]) + t[r];
      const double d = h - G[j * 3 + r];
      d2 += d * d;
    }
    acc += sqrt(d2);
  }
  pa_mpjpe[i] = (float)(acc / J);
}

// Per-vertex error (evaluate/eval.py:208-209): pve[i] = mean_v ||pred[i,v] - gt[i,v]||.  One CTA per frame; the
// vertex distances are summed in a fixed order (per-thread strided partial sums, then a shared-memory tree), so the
// result is deterministic.  HBM: 2 * 12 B per vertex.
__global__ void __launch_bounds__(256)
vertex_errors_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int V, float* __restrict__ pve) {
  __shared__ float red[256];
  const size_t base = (size_t)blockIdx.x * V * 3;
  float acc = 0.f;
  for (int v = threadIdx.x; v < V; v += 256) {
    const float dx = pred[base + v * 3] - gt[base + v * 3], dy = pred[base + v * 3 + 1] - gt[base + v * 3 + 1],
                dz = pred[base + v * 3 + 2] - gt[base + v * 3 + 2];
    acc += sqrtf(dx * dx + dy * dy + dz * dz);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) pve[blockIdx.x] = red[0] / (float)V;
}

}  // namespace whmr
