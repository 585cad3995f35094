// Projection kernels.  One thread per (body, point); the reference's ~12 tiny
// launches per call (zeros, 4 indexed writes into K, 2 einsums, div, slice, div) become one.
// Operation order follows the reference so fp32 rounding matches: rotate -> translate -> divide
// by z -> intrinsics -> normalise.  These kernels move ~1 KB per body: they are launch-latency
// bound at any realistic batch and exist to take launches off the critical path, not for bandwidth.
#pragma once
#include "common.cuh"

namespace whmr {

// utils/geometry.py:289-307 for one point (shared by the stand-alone kernels and the read-out finishing pass, so that
// both schedules of the loop produce the same bits)
__device__ __forceinline__ void project_weak_point(const float* __restrict__ pt, float s, float tx, float ty, float focal,
                                                   float img_w, float img_h, float* __restrict__ out2) {
  const float tz = 2.0f * focal / (img_h * s + 1e-9f);
  const float px = pt[0] + tx;
  const float py = pt[1] + ty;
  const float pz = pt[2] + tz;
  const float qx = px / pz, qy = py / pz;
  out2[0] = (focal * qx) / (img_w * 0.5f);
  out2[1] = (focal * qy) / (img_h * 0.5f);
}

// models/whmr.py:147-173 (+ utils/geometry.py:139-157) for point n of body b; n == 0 also writes focal / cam_t
__device__ __forceinline__ void project_full_point(const float* __restrict__ pt, int b, int n, const float* __restrict__ cam,
                                                   const float* __restrict__ bbox_height, const float* __restrict__ center,
                                                   const float* __restrict__ orig_shape, const float* __restrict__ Tz,
                                                   float* __restrict__ kp_norm2, float* __restrict__ kp_px2,
                                                   float* __restrict__ focal_out, float* __restrict__ cam_t_out,
                                                   float* __restrict__ kp_weak2, float wfocal, float wimg_w, float wimg_h) {
  const float s = cam[b * 3 + 0], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2];
  const float h = bbox_height[b], tz = Tz[b];
  const float img_h = orig_shape[b * 2 + 0], img_w = orig_shape[b * 2 + 1];
  const float focal = s * h * tz / 2.0f;                       // whmr.py:149
  const float ccx = img_w / 2.0f, ccy = img_h / 2.0f;          // :152-153
  const float sh = s * h;
  const float ctx = tx + 2.0f * (center[b * 2 + 0] - (img_w / 2.0f)) / sh;   // geometry.py:152-155
  const float cty = ty + 2.0f * (center[b * 2 + 1] - (img_h / 2.0f)) / sh;
  if (n == 0) {
    if (focal_out) focal_out[b] = focal;
    if (cam_t_out) { cam_t_out[b * 3 + 0] = ctx; cam_t_out[b * 3 + 1] = cty; cam_t_out[b * 3 + 2] = tz; }
  }
  const float x = pt[0] + ctx, y = pt[1] + cty, z = pt[2] + tz;
  const float qx = x / z, qy = y / z, qz = z / z;
  const float u = focal * qx + ccx * qz;
  const float v = focal * qy + ccy * qz;
  if (kp_px2) { kp_px2[0] = u; kp_px2[1] = v; }
  if (kp_norm2) { kp_norm2[0] = u / ccx - 1.0f; kp_norm2[1] = v / ccy - 1.0f; }   // :173
  if (kp_weak2) {   // utils/geometry.py:289-307 on the same points (Regressor.forward evaluates both, :142-173)
    const float wtz = 2.0f * wfocal / (wimg_h * s + 1e-9f);
    const float wx = pt[0] + tx, wy = pt[1] + ty, wz = pt[2] + wtz;
    kp_weak2[0] = (wfocal * (wx / wz)) / (wimg_w * 0.5f);
    kp_weak2[1] = (wfocal * (wy / wz)) / (wimg_h * 0.5f);
  }
}

// utils/geometry.py:289-307
__global__ void __launch_bounds__(256)
project_weak_kernel(const float* __restrict__ points, const float* __restrict__ cam, int B, int N,
                    float focal, float img_w, float img_h, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  const int b = (int)(i / N);
  project_weak_point(points + i * 3, cam[b * 3 + 0], cam[b * 3 + 1], cam[b * 3 + 2], focal, img_w, img_h, out + i * 2);
}

// utils/geometry.py:310-341
__global__ void __launch_bounds__(256)
perspective_projection_kernel(const float* __restrict__ points, const float* __restrict__ rotation,
                              int rot_batch, const float* __restrict__ translation,
                              const float* __restrict__ focal_dev, float focal_scalar,
                              const float* __restrict__ camera_center,
                              const float* __restrict__ distortion, int B, int N, int retain_z,
                              float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  const int b = (int)(i / N);
  float x = points[i * 3 + 0], y = points[i * 3 + 1], z = points[i * 3 + 2];
  if (rotation) {
    const float* R = rotation + (rot_batch > 1 ? (size_t)b * 9 : 0);
    const float rx = R[0] * x + R[1] * y + R[2] * z;
    const float ry = R[3] * x + R[4] * y + R[5] * z;
    const float rz = R[6] * x + R[7] * y + R[8] * z;
    x = rx; y = ry; z = rz;
  }
  if (translation) { x += translation[b * 3 + 0]; y += translation[b * 3 + 1]; z += translation[b * 3 + 2]; }
  if (distortion) {   // models/maf_extractor.py:212-225 (5-coefficient radial + tangential)
    const float* kc = distortion + (size_t)b * 5;
    const float px = x / z, py = y / z;
    const float r2 = px * px + py * py;
    const float dx = 2.0f * kc[2] * px * py + kc[3] * (r2 + 2.0f * px * px);
    const float dy = 2.0f * kc[3] * px * py + kc[2] * (r2 + 2.0f * py * py);
    const float rad = 1.0f + kc[0] * r2 + kc[1] * (r2 * r2) + kc[4] * (r2 * r2 * r2);
    x = rad * px + dx; y = rad * py + dy; z = 1.0f;
  }
  const float qx = x / z, qy = y / z, qz = z / z;
  const float f = focal_dev ? focal_dev[b] : focal_scalar;
  const float cx = camera_center[b * 2 + 0], cy = camera_center[b * 2 + 1];
  const float u = f * qx + cx * qz;
  const float v = f * qy + cy * qz;
  if (retain_z) {
    out[i * 3 + 0] = u; out[i * 3 + 1] = v; out[i * 3 + 2] = qz;
  } else {
    out[i * 2 + 0] = u; out[i * 2 + 1] = v;
  }
}

// models/whmr.py:147-173 (+ utils/geometry.py:139-157)
__global__ void __launch_bounds__(256)
project_full_kernel(const float* __restrict__ points, const float* __restrict__ cam,
                    const float* __restrict__ bbox_height, const float* __restrict__ center,
                    const float* __restrict__ orig_shape, const float* __restrict__ Tz, int B, int N,
                    float* __restrict__ kp_norm, float* __restrict__ kp_px,
                    float* __restrict__ focal_out, float* __restrict__ cam_t_out,
                    float* __restrict__ kp_weak, float wfocal, float wimg_w, float wimg_h) {
  pdl_wait();
  pdl_trigger();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  const int b = (int)(i / N);
  const int n = (int)(i % N);
  project_full_point(points + i * 3, b, n, cam, bbox_height, center, orig_shape, Tz, kp_norm ? kp_norm + i * 2 : nullptr,
                     kp_px ? kp_px + i * 2 : nullptr, focal_out, cam_t_out, kp_weak ? kp_weak + i * 2 : nullptr, wfocal, wimg_w,
                     wimg_h);
}

// models/maf_extractor.py:145-235 (project + get_trans + perspective_projection w/ distortion)
__global__ void __launch_bounds__(256)
project_crop_kernel(const float* __restrict__ points, const float* __restrict__ cam,
                    const float* __restrict__ center, const float* __restrict__ scale,
                    const float* __restrict__ img_focal, const float* __restrict__ img_center,
                    const float* __restrict__ distortion, int B, int N, float crop_size, float img_w,
                    float img_h, float* __restrict__ full_out, float* __restrict__ crop_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  const int b = (int)(i / N);
  const float bb = scale[b] * 200.0f;
  const float s = cam[b * 3 + 0], tx = cam[b * 3 + 1], ty = cam[b * 3 + 2];
  const float cx = center[b * 2 + 0], cy = center[b * 2 + 1];
  const float icx = img_center[b * 2 + 0], icy = img_center[b * 2 + 1];
  const float f = img_focal[b];
  const float bs = bb * s;
  float x = points[i * 3 + 0] + (tx + 2.0f * (cx - icx) / bs);
  float y = points[i * 3 + 1] + (ty + 2.0f * (cy - icy) / bs);
  float z = points[i * 3 + 2] + (2.0f * f / bs);
  if (distortion) {
    const float* kc = distortion + (size_t)b * 5;
    const float px = x / z, py = y / z;
    const float r2 = px * px + py * py;
    const float dx = 2.0f * kc[2] * px * py + kc[3] * (r2 + 2.0f * px * px);
    const float dy = 2.0f * kc[3] * px * py + kc[2] * (r2 + 2.0f * py * py);
    const float rad = 1.0f + kc[0] * r2 + kc[1] * (r2 * r2) + kc[4] * (r2 * r2 * r2);
    x = rad * px + dx; y = rad * py + dy; z = 1.0f;
  }
  const float qx = x / z, qy = y / z, qz = z / z;
  const float u = f * qx + icx * qz;
  const float v = f * qy + icy * qz;
  if (full_out) { full_out[i * 2 + 0] = u; full_out[i * 2 + 1] = v; }
  if (crop_out) {
    const float hx = img_w * 0.5f, hy = img_h * 0.5f;
    const float px = (u - (cx - bb / 2.0f)) * (crop_size / bb);
    const float py = (v - (cy - bb / 2.0f)) * (crop_size / bb);
    crop_out[i * 2 + 0] = (px - hx) / hx;
    crop_out[i * 2 + 1] = (py - hy) / hy;
  }
}

// utils/geometry.py:344-408 estimate_translation: per sample, the camera translation that best re-projects the 3-D
// joints onto the 2-D key points, a confidence-weighted least-squares problem with the 3x3 normal equations
//   rows (per joint i):  [F 0 (cx - x_i)] t = (x_i - cx) Z_i - F X_i ,   [0 F (cy - y_i)] t = (y_i - cy) Z_i - F Y_i
// weighted by sqrt(conf_i) (so conf_i in the normal equations).  The reference does this on the host, one
// np.linalg.solve per sample in float64, with a device->host->device round trip every training step
// (core/trainer.py:435); here one thread per sample accumulates and solves in float64 on the device.
// Joints [j0, N) take part (the reference slices 25: of 49 -- the ground-truth joints).  A singular system (all
// confidences zero) makes np.linalg.solve raise; here it yields non-finite values.
__global__ void __launch_bounds__(128)
estimate_translation_kernel(const float* __restrict__ S, const float* __restrict__ joints_2d, int B, int N, int j0,
                            float focal, float img_w, float img_h, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double F = (double)focal, cx = (double)img_w * 0.5, cy = (double)img_h * 0.5;
  double a00 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0, b0 = 0, b1 = 0, b2 = 0;
  for (int i = j0; i < N; ++i) {
    const float* s = S + ((size_t)b * N + i) * 3;
    const float* k = joints_2d + ((size_t)b * N + i) * 3;
    const double X = s[0], Y = s[1], Z = s[2], x = k[0], y = k[1];
    const double w = sqrt((double)k[2]);     // weight2; the normal equations see w*w
    const double qx = w * (cx - x), qy = w * (cy - y), wf = w * F;
    const double c0 = w * ((x - cx) * Z - F * X), c1 = w * ((y - cy) * Z - F * Y);
    a00 += wf * wf; a02 += wf * qx; a11 += wf * wf; a12 += wf * qy; a22 += qx * qx + qy * qy;
    b0 += wf * c0; b1 += wf * c1; b2 += qx * c0 + qy * c1;
  }
  // A = [[a00 0 a02] [0 a11 a12] [a02 a12 a22]]: eliminate the two leading unknowns
  const double m0 = a02 / a00, m1 = a12 / a11;
  const double d = a22 - m0 * a02 - m1 * a12;
  const double tz = (b2 - m0 * b0 - m1 * b1) / d;
  out[b * 3 + 0] = (float)((b0 - a02 * tz) / a00);
  out[b * 3 + 1] = (float)((b1 - a12 * tz) / a11);
  out[b * 3 + 2] = (float)tz;
}

}  // namespace whmr
