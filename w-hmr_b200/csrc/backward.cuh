// Backward kernels of the custom ops (what Trainer.train_step needs to flow gradients through
// the drop-ins).  What the reference's graph actually asks for (models/whmr.py:145-173, 586-591): the projections
// see detached joints, so their gradients go to pred_cam (weak) and Tz (predicted-focal block); the sampling points
// are detached, so the sampling gradient goes to the feature maps only; SMPL vertices/joints go to betas/rotmats.
// The projection kernels also return d/d points (cheap, and needed by any other caller).
#pragma once
#include "common.cuh"
#include "sampling.cuh"

namespace whmr {

// deterministic block sum of up to 6 values (fixed tree), result valid in thread 0
template <int NV, int THREADS>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red /*[NV*THREADS]*/) {
#pragma unroll
  for (int k = 0; k < NV; ++k) red[k * THREADS + threadIdx.x] = v[k];
  __syncthreads();
  for (int s = THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
#pragma unroll
      for (int k = 0; k < NV; ++k) red[k * THREADS + threadIdx.x] += red[k * THREADS + threadIdx.x + s];
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = red[k * THREADS];
}

// ---- weak projection (utils/geometry.py:289-307): u = a*px/pz, v = b*py/pz, p = X + (tx, ty, 2f/(H s + 1e-9)) ----
// accumulates the gradient of one point w.r.t. the point and the camera; returns d/dX, d/dY, d/dZ
__device__ __forceinline__ void weak_point_bwd(float X, float Y, float Z, float s, float tx, float ty, float focal,
                                               float img_w, float img_h, float gu, float gv, float& gX, float& gY,
                                               float& gZ) {
  const float tz = 2.0f * focal / (img_h * s + 1e-9f);
  const float px = X + tx, py = Y + ty, pz = Z + tz;
  const float a = focal / (img_w * 0.5f), b = focal / (img_h * 0.5f);
  const float iz = 1.0f / pz;
  gX = gu * a * iz;
  gY = gv * b * iz;
  gZ = -(gu * a * px + gv * b * py) * iz * iz;
}

// grid = B, block 128.  g_points may be null.
__global__ void __launch_bounds__(128)
project_weak_bwd_kernel(const float* __restrict__ points, const float* __restrict__ cam, const float* __restrict__ g_out,
                        int N, float focal, float img_w, float img_h, float* __restrict__ g_points,
                        float* __restrict__ g_cam) {
  __shared__ float red[3 * 128];
  const int b = blockIdx.x;
  const float s = cam[b * 3 + 0], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2];
  float acc[3] = {0.f, 0.f, 0.f};   // d/d tx, ty, tz
  for (int n = threadIdx.x; n < N; n += 128) {
    const size_t i = (size_t)b * N + n;
    float gX, gY, gZ;
    weak_point_bwd(points[i * 3], points[i * 3 + 1], points[i * 3 + 2], s, tx, ty, focal, img_w, img_h, g_out[i * 2],
                   g_out[i * 2 + 1], gX, gY, gZ);
    if (g_points) { g_points[i * 3] = gX; g_points[i * 3 + 1] = gY; g_points[i * 3 + 2] = gZ; }
    acc[0] += gX; acc[1] += gY; acc[2] += gZ;
  }
  block_sum<3, 128>(acc, red);
  if (threadIdx.x == 0) {
    const float d = img_h * s + 1e-9f;
    g_cam[b * 3 + 0] = acc[2] * (-2.0f * focal * img_h / (d * d));   // d tz / d s
    g_cam[b * 3 + 1] = acc[0];
    g_cam[b * 3 + 2] = acc[1];
  }
}

// ---- utils/geometry.py:310-341 perspective_projection (no distortion): x = R p + t; u = f x/z + cx, v = f y/z + cy ------
// grid = B, block 128.  Every output pointer may be null; the retained z column (z/z) has zero gradient.
__global__ void __launch_bounds__(128)
perspective_projection_bwd_kernel(const float* __restrict__ points, const float* __restrict__ rotation, int rot_batch,
                                  const float* __restrict__ translation, const float* __restrict__ focal_dev,
                                  float focal_scalar, const float* __restrict__ g_out, int N, int retain_z,
                                  float* __restrict__ g_points, float* __restrict__ g_trans, float* __restrict__ g_focal,
                                  float* __restrict__ g_center) {
  __shared__ float red[6 * 128];
  const int b = blockIdx.x;
  const float* R = rotation ? rotation + (rot_batch > 1 ? (size_t)b * 9 : 0) : nullptr;
  const float f = focal_dev ? focal_dev[b] : focal_scalar;
  const float tx = translation ? translation[b * 3 + 0] : 0.f, ty = translation ? translation[b * 3 + 1] : 0.f,
              tz = translation ? translation[b * 3 + 2] : 0.f;
  const int ld = retain_z ? 3 : 2;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // d/d(tx, ty, tz), d/d focal, d/d(cx, cy)
  for (int n = threadIdx.x; n < N; n += 128) {
    const size_t i = (size_t)b * N + n;
    float x = points[i * 3 + 0], y = points[i * 3 + 1], z = points[i * 3 + 2];
    if (R) {
      const float rx = R[0] * x + R[1] * y + R[2] * z, ry = R[3] * x + R[4] * y + R[5] * z, rz = R[6] * x + R[7] * y + R[8] * z;
      x = rx; y = ry; z = rz;
    }
    x += tx; y += ty; z += tz;
    const float gu = g_out[i * ld + 0], gv = g_out[i * ld + 1];
    const float iz = 1.0f / z;
    const float gx = gu * f * iz, gy = gv * f * iz, gz = -(gu * x + gv * y) * f * iz * iz;
    if (g_points) {
      if (R) {   // g_p = R^T g_x
        g_points[i * 3 + 0] = R[0] * gx + R[3] * gy + R[6] * gz;
        g_points[i * 3 + 1] = R[1] * gx + R[4] * gy + R[7] * gz;
        g_points[i * 3 + 2] = R[2] * gx + R[5] * gy + R[8] * gz;
      } else {
        g_points[i * 3 + 0] = gx; g_points[i * 3 + 1] = gy; g_points[i * 3 + 2] = gz;
      }
    }
    acc[0] += gx; acc[1] += gy; acc[2] += gz;
    acc[3] += (gu * x + gv * y) * iz;
    acc[4] += gu; acc[5] += gv;
  }
  block_sum<6, 128>(acc, red);
  if (threadIdx.x == 0) {
    if (g_trans) { g_trans[b * 3 + 0] = acc[0]; g_trans[b * 3 + 1] = acc[1]; g_trans[b * 3 + 2] = acc[2]; }
    if (g_focal) g_focal[b] = acc[3];
    if (g_center) { g_center[b * 2 + 0] = acc[4]; g_center[b * 2 + 1] = acc[5]; }
  }
}

// ---- weak + predicted-focal block (models/whmr.py:142-173), the backward of project_full_kernel ------------------
// upstream: g_kp_weak [B,N,2] (or null), g_kp_norm [B,N,2] (or null), g_focal [B] (or null), g_cam_t [B,3] (or null)
// out: g_points [B,N,3] (or null), g_cam [B,3], g_Tz [B]
__global__ void __launch_bounds__(128)
project_full_bwd_kernel(const float* __restrict__ points, const float* __restrict__ cam,
                        const float* __restrict__ bbox_height, const float* __restrict__ center,
                        const float* __restrict__ orig_shape, const float* __restrict__ Tz, int N,
                        float wfocal, float wimg_w, float wimg_h, const float* __restrict__ g_kp_weak,
                        const float* __restrict__ g_kp_norm, const float* __restrict__ g_focal,
                        const float* __restrict__ g_cam_t, float* __restrict__ g_points, float* __restrict__ g_cam,
                        float* __restrict__ g_Tz) {
  __shared__ float red[6 * 128];
  const int b = blockIdx.x;
  const float s = cam[b * 3 + 0], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2];
  const float h = bbox_height[b], tz = Tz[b];
  const float img_h = orig_shape[b * 2 + 0], img_w = orig_shape[b * 2 + 1];
  const float focal = s * h * tz / 2.0f;
  const float ccx = img_w / 2.0f, ccy = img_h / 2.0f;
  const float sh = s * h;
  const float dx = 2.0f * (center[b * 2 + 0] - ccx), dy = 2.0f * (center[b * 2 + 1] - ccy);
  const float ctx = tx + dx / sh, cty = ty + dy / sh;
  // acc: full block d/d(ctx, cty, ctz), d/d focal; weak block d/d(tx, ty) and d/d wtz folded into [4],[5] below
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float wz_acc = 0.f;
  for (int n = threadIdx.x; n < N; n += 128) {
    const size_t i = (size_t)b * N + n;
    const float X = points[i * 3], Y = points[i * 3 + 1], Z = points[i * 3 + 2];
    float gX = 0.f, gY = 0.f, gZ = 0.f;
    if (g_kp_norm) {
      const float x = X + ctx, y = Y + cty, z = Z + tz;
      const float iz = 1.0f / z;
      const float gnx = g_kp_norm[i * 2], gny = g_kp_norm[i * 2 + 1];
      const float Fx = focal / ccx, Fy = focal / ccy;
      const float fX = gnx * Fx * iz, fY = gny * Fy * iz;
      const float fZ = -(gnx * Fx * x + gny * Fy * y) * iz * iz;
      gX += fX; gY += fY; gZ += fZ;
      acc[0] += fX; acc[1] += fY; acc[2] += fZ;
      acc[3] += gnx * x * iz / ccx + gny * y * iz / ccy;
    }
    if (g_kp_weak) {
      float wX, wY, wZ;
      weak_point_bwd(X, Y, Z, s, tx, ty, wfocal, wimg_w, wimg_h, g_kp_weak[i * 2], g_kp_weak[i * 2 + 1], wX, wY, wZ);
      gX += wX; gY += wY; gZ += wZ;
      acc[4] += wX; acc[5] += wY; wz_acc += wZ;
    }
    if (g_points) { g_points[i * 3] = gX; g_points[i * 3 + 1] = gY; g_points[i * 3 + 2] = gZ; }
  }
  block_sum<6, 128>(acc, red);
  __syncthreads();
  float w1[1] = {wz_acc};
  block_sum<1, 128>(w1, red);
  if (threadIdx.x == 0) {
    float g_ctx = acc[0], g_cty = acc[1], g_ctz = acc[2], g_f = acc[3];
    if (g_cam_t) { g_ctx += g_cam_t[b * 3 + 0]; g_cty += g_cam_t[b * 3 + 1]; g_ctz += g_cam_t[b * 3 + 2]; }
    if (g_focal) g_f += g_focal[b];
    const float d = wimg_h * s + 1e-9f;
    // focal = s h Tz / 2 ; ctx = tx + dx/(s h) ; cty = ty + dy/(s h) ; ctz = Tz
    g_cam[b * 3 + 0] = g_f * h * tz / 2.0f - (g_ctx * dx + g_cty * dy) / (s * sh) + w1[0] * (-2.0f * wfocal * wimg_h / (d * d));
    g_cam[b * 3 + 1] = g_ctx + acc[4];
    g_cam[b * 3 + 2] = g_cty + acc[5];
    g_Tz[b] = g_ctz + g_f * sh / 2.0f;
  }
}

// ---- bilinear sampling, gradient w.r.t. the feature maps ----------------------------------------------------------
// g_feat must be zero-initialised by the caller.  fp32 atomics: the summation order over points that share a pixel
// is not fixed (as in ATen's grid_sampler_2d_backward).
// NCHW: threads walk the flat [C*N] gradient of a body (coalesced reads), 4 scattered atomics each.
__global__ void __launch_bounds__(256)
sample_bilinear_bwd_nchw_kernel(const float* __restrict__ g_out, const float* __restrict__ points, int pts_bstride,
                                float* __restrict__ g_feat, int C, int H, int W, int N) {
  const int b = blockIdx.y;
  const long long total = (long long)C * N;
  const size_t plane = (size_t)H * W;
  float* fb = g_feat + (size_t)b * C * plane;
  const float* pb = points + (size_t)b * pts_bstride;
  const float* gb = g_out + (size_t)b * total;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c = (int)(i / N), n = (int)(i - (long long)c * N);
    const float2 g = *reinterpret_cast<const float2*>(pb + (size_t)n * 2);
    const Taps t = make_taps(g.x, g.y, H, W);
    const float go = gb[i];
    float* pl = fb + (size_t)c * plane;
    if (t.w00 != 0.f) atomicAdd(pl + t.o00, go * t.w00);
    if (t.w01 != 0.f) atomicAdd(pl + t.o01, go * t.w01);
    if (t.w10 != 0.f) atomicAdd(pl + t.o10, go * t.w10);
    if (t.w11 != 0.f) atomicAdd(pl + t.o11, go * t.w11);
  }
}

// NHWC: a warp takes one (body, point) and walks the channels, so each tap is a coalesced run of atomics.
__global__ void __launch_bounds__(256)
sample_bilinear_bwd_nhwc_kernel(const float* __restrict__ g_out, const float* __restrict__ points, int pts_bstride,
                                float* __restrict__ g_feat, int B, int C, int H, int W, int N) {
  const long long w = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)B * N) return;
  const int b = (int)(w / N), n = (int)(w - (long long)b * N);
  const float2 g = *reinterpret_cast<const float2*>(points + (size_t)b * pts_bstride + (size_t)n * 2);
  const Taps t = make_taps(g.x, g.y, H, W);
  float* fb = g_feat + (size_t)b * H * W * C;
  const float* gb = g_out + (size_t)b * C * N + n;
  for (int c = lane; c < C; c += 32) {
    const float go = gb[(size_t)c * N];
    if (t.w00 != 0.f) atomicAdd(fb + (size_t)t.o00 * C + c, go * t.w00);
    if (t.w01 != 0.f) atomicAdd(fb + (size_t)t.o01 * C + c, go * t.w01);
    if (t.w10 != 0.f) atomicAdd(fb + (size_t)t.o10 * C + c, go * t.w10);
    if (t.w11 != 0.f) atomicAdd(fb + (size_t)t.o11 * C + c, go * t.w11);
  }
}

}  // namespace whmr
