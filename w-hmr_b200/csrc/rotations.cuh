// Rotation glue around the SMPL call inside Regressor.forward: each is a chain of
// ~10-25 tiny ATen launches on [B*24,3,3] in the reference; here one thread per rotation, one launch.
//   rot6d_to_rotmat               utils/geometry.py:243-257  (models/whmr.py:65)
//   unbiased_gram_schmidt         utils/geometry.py:260-272  (models/whmr.py:129-130, eval mode)
//   rotation_matrix_to_angle_axis utils/geometry.py:54-83,86-136,160-240 (kornia path; models/whmr.py:174,632)
#pragma once
#include "common.cuh"

namespace whmr {

__device__ __forceinline__ void normalize3(float& x, float& y, float& z) {   // F.normalize: v / max(||v||, 1e-12)
  const float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
  x /= n; y /= n; z /= n;
}

__global__ void __launch_bounds__(256) rot6d_to_rotmat_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = x + (size_t)i * 6;   // viewed as [3,2]: a1 = column 0, a2 = column 1
  float a1x = p[0], a1y = p[2], a1z = p[4], a2x = p[1], a2y = p[3], a2z = p[5];
  normalize3(a1x, a1y, a1z);
  const float d = a1x * a2x + a1y * a2y + a1z * a2z;
  float b2x = a2x - d * a1x, b2y = a2y - d * a1y, b2z = a2z - d * a1z;
  normalize3(b2x, b2y, b2z);
  const float b3x = a1y * b2z - a1z * b2y, b3y = a1z * b2x - a1x * b2z, b3z = a1x * b2y - a1y * b2x;
  float* o = out + (size_t)i * 9;       // columns (b1, b2, b3)
  o[0] = a1x; o[1] = b2x; o[2] = b3x;
  o[3] = a1y; o[4] = b2y; o[5] = b3y;
  o[6] = a1z; o[7] = b2z; o[8] = b3z;
}

// utils/geometry.py:14-51: the quaternion variant of batch_rodrigues (core/trainer.py:244 -- ground-truth poses of the
// training step): angle = ||theta + 1e-8||, quat = (cos(angle/2), sin(angle/2) * theta/angle), normalised, -> matrix.
// (The smplx variant inside SMPL.forward is whmr_batch_rodrigues; the two differ at ~1e-7.)
__global__ void __launch_bounds__(256) batch_rodrigues_quat_kernel(const float* __restrict__ theta, int n,
                                                                   float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float tx = theta[(size_t)i * 3 + 0], ty = theta[(size_t)i * 3 + 1], tz = theta[(size_t)i * 3 + 2];
  const float ex = tx + 1e-8f, ey = ty + 1e-8f, ez = tz + 1e-8f;
  const float angle = sqrtf(ex * ex + ey * ey + ez * ez);
  const float nx = tx / angle, ny = ty / angle, nz = tz / angle;
  const float half = angle * 0.5f;
  const float vs = sinf(half);
  float w = cosf(half), x = vs * nx, y = vs * ny, z = vs * nz;
  const float qn = sqrtf(w * w + x * x + y * y + z * z);
  w /= qn; x /= qn; y /= qn; z /= qn;
  const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  float* o = out + (size_t)i * 9;
  o[0] = w2 + x2 - y2 - z2; o[1] = 2 * xy - 2 * wz;    o[2] = 2 * wy + 2 * xz;
  o[3] = 2 * wz + 2 * xy;    o[4] = w2 - x2 + y2 - z2; o[5] = 2 * yz - 2 * wx;
  o[6] = 2 * xz - 2 * wy;    o[7] = 2 * wx + 2 * yz;    o[8] = w2 - x2 - y2 + z2;
}

// utils/geometry.py:260-272 on one row-major 3x3 (in and out may alias)
__device__ __forceinline__ void unbiased_gram_schmidt_dev(const float* p, float* o) {
  const float t1x = p[0], t1y = p[3], t1z = p[6], t2x = p[1], t2y = p[4], t2z = p[7], t3x = p[2], t3y = p[5], t3z = p[8];
  float r1x = ((t2y * t3z - t2z * t3y) + t1x) / 2.0f, r1y = ((t2z * t3x - t2x * t3z) + t1y) / 2.0f,
        r1z = ((t2x * t3y - t2y * t3x) + t1z) / 2.0f;
  normalize3(r1x, r1y, r1z);
  const float qx = ((t3y * r1z - t3z * r1y) + t2x) / 2.0f, qy = ((t3z * r1x - t3x * r1z) + t2y) / 2.0f,
              qz = ((t3x * r1y - t3y * r1x) + t2z) / 2.0f;
  const float d = qx * r1x + qy * r1y + qz * r1z;
  float r2x = qx - d * r1x, r2y = qy - d * r1y, r2z = qz - d * r1z;
  normalize3(r2x, r2y, r2z);
  const float r3x = r1y * r2z - r1z * r2y, r3y = r1z * r2x - r1x * r2z, r3z = r1x * r2y - r1y * r2x;
  o[0] = r1x; o[1] = r2x; o[2] = r3x;
  o[3] = r1y; o[4] = r2y; o[5] = r3y;
  o[6] = r1z; o[7] = r2z; o[8] = r3z;
}

__global__ void __launch_bounds__(256) unbiased_gram_schmidt_kernel(const float* __restrict__ x, int n,
                                                                    float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r[9];
  unbiased_gram_schmidt_dev(x + (size_t)i * 9, r);
#pragma unroll
  for (int k = 0; k < 9; ++k) out[(size_t)i * 9 + k] = r[k];
}

// utils/geometry.py:54-83,86-136,160-240 on one row-major 3x3 -> axis-angle
__device__ __forceinline__ void rotmat_to_axis_angle_dev(const float* p, float& ox, float& oy, float& oz) {
  // rmat_t = R^T  (utils/geometry.py:198); m[a][b] = R[b][a]
  const float m00 = p[0], m01 = p[3], m02 = p[6], m10 = p[1], m11 = p[4], m12 = p[7], m20 = p[2], m21 = p[5], m22 = p[8];
  const bool d2 = m22 < 1e-6f, d0_d1 = m00 > m11, d0_nd1 = m00 < -m11;
  float q0, q1, q2, q3, t;
  if (d2 && d0_d1) {
    t = 1 + m00 - m11 - m22; q0 = m12 - m21; q1 = t; q2 = m01 + m10; q3 = m20 + m02;
  } else if (d2) {
    t = 1 - m00 + m11 - m22; q0 = m20 - m02; q1 = m01 + m10; q2 = t; q3 = m12 + m21;
  } else if (d0_nd1) {
    t = 1 - m00 - m11 + m22; q0 = m01 - m10; q1 = m20 + m02; q2 = m12 + m21; q3 = t;
  } else {
    t = 1 + m00 + m11 + m22; q0 = t; q1 = m12 - m21; q2 = m20 - m02; q3 = m01 - m10;
  }
  const float s = 0.5f / sqrtf(t);
  q0 *= s; q1 *= s; q2 *= s; q3 *= s;
  // quaternion_to_angle_axis (:86-136)
  const float ss = q1 * q1 + q2 * q2 + q3 * q3;
  const float st = sqrtf(ss);
  const float two_theta = 2.0f * (q0 < 0.0f ? atan2f(-st, -q0) : atan2f(st, q0));
  const float k = ss > 0.0f ? two_theta / st : 2.0f;
  float ax = q1 * k, ay = q2 * k, az = q3 * k;
  if (isnan(ax)) ax = 0.0f;   // aa[torch.isnan(aa)] = 0.0 (:82)
  if (isnan(ay)) ay = 0.0f;
  if (isnan(az)) az = 0.0f;
  ox = ax; oy = ay; oz = az;
}

__global__ void __launch_bounds__(256) rotmat_to_axis_angle_kernel(const float* __restrict__ R, int n,
                                                                   float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float ax, ay, az;
  rotmat_to_axis_angle_dev(R + (size_t)i * 9, ax, ay, az);
  out[(size_t)i * 3 + 0] = ax; out[(size_t)i * 3 + 1] = ay; out[(size_t)i * 3 + 2] = az;
}

}  // namespace whmr
