// Mesh-aligned feature sampling: F.grid_sample(im_feat, points[:, :, None, :],
// align_corners=True)[..., 0] of models/maf_extractor.py:119 -- bilinear, zero padding,
// pixel = (g + 1)/2 * (size - 1), points[...,0] <-> W.
//
// ATen's grid_sampler_2d assigns a thread per point and loops over C with stride H*W, so both its
// reads and its [B,C,N] writes are uncoalesced.  Here:
//  * NCHW (the contract layout): the output of one body is the flat array out_b[c*N + n]; threads
//    walk that flat index, so every warp store is 128 contiguous bytes whatever N is (63, 67, 431,
//    6890 ...), and the 4 taps of neighbouring lanes fall in the same channel plane (L1/L2 reuse).
//    Each (point, channel, row) costs one 32 B sector -- the floor for a sparse gather from NCHW.
//  * NHWC: lanes run over channels so every tap is a coalesced 128 B read; the [points x channels]
//    tile is transposed through shared memory so the [B,C,N] store is coalesced too.
#pragma once
#include <stdlib.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace whmr {

struct Taps {
  int o00, o01, o10, o11;      // element offsets inside a plane (clamped, always valid)
  float w00, w01, w10, w11;    // weights, zeroed for out-of-bounds taps
};

// PyTorch grid_sampler_2d semantics (align_corners=True, padding_mode=zeros, bilinear)
__device__ __forceinline__ Taps make_taps(float gx, float gy, int H, int W) {
  Taps t;
  const float ix = ((gx + 1.0f) / 2.0f) * (float)(W - 1);
  const float iy = ((gy + 1.0f) / 2.0f) * (float)(H - 1);
  // keep the float->int conversion defined for far-away / non-finite points
  const bool sane = (ix > -2.0f) && (ix < (float)W + 1.0f) && (iy > -2.0f) && (iy < (float)H + 1.0f);
  const float fx = sane ? floorf(ix) : -2.0f;
  const float fy = sane ? floorf(iy) : -2.0f;
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - fx, wx0 = (fx + 1.0f) - ix;
  const float wy1 = iy - fy, wy0 = (fy + 1.0f) - iy;
  const bool vx0 = sane && x0 >= 0 && x0 < W, vx1 = sane && x1 >= 0 && x1 < W;
  const bool vy0 = sane && y0 >= 0 && y0 < H, vy1 = sane && y1 >= 0 && y1 < H;
  const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
  const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
  t.o00 = cy0 * W + cx0; t.o01 = cy0 * W + cx1; t.o10 = cy1 * W + cx0; t.o11 = cy1 * W + cx1;
  t.w00 = (vx0 && vy0) ? wx0 * wy0 : 0.0f;   // nw
  t.w01 = (vx1 && vy0) ? wx1 * wy0 : 0.0f;   // ne
  t.w10 = (vx0 && vy1) ? wx0 * wy1 : 0.0f;   // sw
  t.w11 = (vx1 && vy1) ? wx1 * wy1 : 0.0f;   // se
  return t;
}

constexpr int kSampleItems = 4;   // flat outputs per thread (independent gathers in flight)

// Weak-perspective projection parameters for the fused MAF_Extractor.forward path (projection + sampling
// in one launch): points are then [B,N,3] mesh points and the 2-D grid coordinate is computed on the fly.
struct SampleProj {
  const float* cam;     // [B,3] (s,tx,ty)
  float focal, img_w, img_h;
  float* pts2d_out;     // [B,N,2] or null
};

// grid = (ceil(C*N / (256*kSampleItems)), B)
template <bool kProject>
__global__ void __launch_bounds__(256)
sample_bilinear_nchw_kernel(const float* __restrict__ feat, const float* __restrict__ points, int pts_bstride,
                            float* __restrict__ out, int C, int H, int W, int N, SampleProj pj) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  float cs = 0.f, ctx = 0.f, cty = 0.f, ctz = 0.f;
  if (kProject) {   // utils/geometry.py:289-307
    cs = pj.cam[b * 3 + 0]; ctx = pj.cam[b * 3 + 1]; cty = pj.cam[b * 3 + 2];
    ctz = 2.0f * pj.focal / (pj.img_h * cs + 1e-9f);
  }
  const int total = C * N;
  const size_t plane = (size_t)H * W;
  const float* fb = feat + (size_t)b * C * plane;
  const float* pb = points + (size_t)b * pts_bstride;
  float* ob = out + (size_t)b * total;
  const int base = blockIdx.x * (256 * kSampleItems) + threadIdx.x;

  float v00[kSampleItems], v01[kSampleItems], v10[kSampleItems], v11[kSampleItems];
  Taps tp[kSampleItems];
#pragma unroll
  for (int it = 0; it < kSampleItems; ++it) {
    const int i = base + it * 256;
    if (i < total) {
      const int c = i / N, n = i - c * N;
      float2 g;
      if (kProject) {
        const float* q = pb + (size_t)n * 3;
        const float px = q[0] + ctx, py = q[1] + cty, pz = q[2] + ctz;
        g.x = (pj.focal * (px / pz)) / (pj.img_w * 0.5f);
        g.y = (pj.focal * (py / pz)) / (pj.img_h * 0.5f);
        if (pj.pts2d_out && c == 0) *reinterpret_cast<float2*>(pj.pts2d_out + ((size_t)b * N + n) * 2) = g;
      } else {
        g = *reinterpret_cast<const float2*>(pb + (size_t)n * 2);
      }
      tp[it] = make_taps(g.x, g.y, H, W);
      const float* pl = fb + (size_t)c * plane;
      v00[it] = __ldg(pl + tp[it].o00); v01[it] = __ldg(pl + tp[it].o01);
      v10[it] = __ldg(pl + tp[it].o10); v11[it] = __ldg(pl + tp[it].o11);
    }
  }
#pragma unroll
  for (int it = 0; it < kSampleItems; ++it) {
    const int i = base + it * 256;
    if (i < total) {
      float acc = v00[it] * tp[it].w00;
      acc = fmaf(v01[it], tp[it].w01, acc);
      acc = fmaf(v10[it], tp[it].w10, acc);
      acc = fmaf(v11[it], tp[it].w11, acc);
      ob[i] = acc;
    }
  }
}

// Dense regime (4N >= H*W: the 431 down-sampled vertices on 14x14 / 28x28 maps, BASELINE configs[3]): nearly every
// pixel of a plane is a tap of some point, and the direct gather is bound by L1 wavefronts (32 lanes hit ~7 cache
// lines of one small plane per load), not by memory.  Here a CTA stages CG whole channel planes of one body in shared
// memory -- consecutive channels are contiguous in NCHW, so this is one coalesced 16-byte-vector stream and every byte
// of the map is read exactly once -- and gathers the taps from shared memory.  A thread keeps the taps of its point(s)
// in registers across the CG channels; for a fixed channel consecutive threads write consecutive n (coalesced).
// Same arithmetic order as the gather kernel: results are bit-identical.
// grid = (ceil(C / CG), B), block 256, dynamic smem = CG*H*W*4 (CG chosen by staged_channels: ~1120*sqrt(H*W) bytes).
template <bool kProject>
__global__ void __launch_bounds__(256)
sample_bilinear_nchw_staged_kernel(const float* __restrict__ feat, const float* __restrict__ points, int pts_bstride,
                                   float* __restrict__ out, int C, int H, int W, int N, int CG, SampleProj pj) {
  extern __shared__ __align__(16) float planes[];   // [CG][H*W]
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y, c0 = blockIdx.x * CG;
  const int cg = min(CG, C - c0);
  const int HW = H * W;
  const float* src = feat + ((size_t)b * C + c0) * HW;
  const int total = cg * HW;
  // cp.async: all 16-byte copies of a thread in flight at once (a load -> store loop keeps ~16 KB per CTA in flight)
  const uint32_t planes_s = (uint32_t)__cvta_generic_to_shared(planes);
  if ((reinterpret_cast<size_t>(src) & 15) == 0 && (total & 3) == 0) {
    for (int i = threadIdx.x; i < (total >> 2); i += 256)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(planes_s + (uint32_t)i * 16u), "l"(src + 4 * (size_t)i) : "memory");
  } else {
    for (int i = threadIdx.x; i < total; i += 256)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(planes_s + (uint32_t)i * 4u), "l"(src + i) : "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  float cs = 0.f, ctx = 0.f, cty = 0.f, ctz = 0.f;
  if (kProject) {   // utils/geometry.py:289-307
    cs = pj.cam[b * 3 + 0]; ctx = pj.cam[b * 3 + 1]; cty = pj.cam[b * 3 + 2];
    ctz = 2.0f * pj.focal / (pj.img_h * cs + 1e-9f);
  }
  const float* pb = points + (size_t)b * pts_bstride;
  for (int n = threadIdx.x; n < N; n += 256) {
    float2 g;
    if (kProject) {
      const float* q = pb + (size_t)n * 3;
      const float px = q[0] + ctx, py = q[1] + cty, pz = q[2] + ctz;
      g.x = (pj.focal * (px / pz)) / (pj.img_w * 0.5f);
      g.y = (pj.focal * (py / pz)) / (pj.img_h * 0.5f);
      if (pj.pts2d_out && blockIdx.x == 0) *reinterpret_cast<float2*>(pj.pts2d_out + ((size_t)b * N + n) * 2) = g;
    } else {
      g = *reinterpret_cast<const float2*>(pb + (size_t)n * 2);
    }
    const Taps tp = make_taps(g.x, g.y, H, W);
    float* ob = out + ((size_t)b * C + c0) * N + n;
#pragma unroll 4
    for (int c = 0; c < cg; ++c) {
      const float* pl = planes + c * HW;
      float acc = pl[tp.o00] * tp.w00;
      acc = fmaf(pl[tp.o01], tp.w01, acc);
      acc = fmaf(pl[tp.o10], tp.w10, acc);
      acc = fmaf(pl[tp.o11], tp.w11, acc);
      ob[(size_t)c * N] = acc;
    }
  }
}

// NHWC input: grid = (ceil(N/32), ceil(C/64), B), block 256 (8 warps x 4 points each).  kProject: `points` are the
// [B,N,3] mesh points and the weak projection (utils/geometry.py:289-307) is evaluated here (MAF_Extractor.forward).
template <bool kProject>
__global__ void __launch_bounds__(256)
sample_bilinear_nhwc_kernel(const float* __restrict__ feat, const float* __restrict__ points, int pts_bstride,
                            float* __restrict__ out, int C, int H, int W, int N, SampleProj pj) {
  __shared__ float tile[64][33];
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.z, c0 = blockIdx.y * 64, n0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* fb = feat + (size_t)b * H * W * C;
  float ctx = 0.f, cty = 0.f, ctz = 0.f;
  if (kProject) {
    const float cs = pj.cam[b * 3 + 0];
    ctx = pj.cam[b * 3 + 1]; cty = pj.cam[b * 3 + 2];
    ctz = 2.0f * pj.focal / (pj.img_h * cs + 1e-9f);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int pt = warp * 4 + q;
    const int n = n0 + pt;
    if (n < N) {
      float2 g;
      if (kProject) {
        const float* qp = points + (size_t)b * pts_bstride + (size_t)n * 3;
        const float px = qp[0] + ctx, py = qp[1] + cty, pz = qp[2] + ctz;
        g.x = (pj.focal * (px / pz)) / (pj.img_w * 0.5f);
        g.y = (pj.focal * (py / pz)) / (pj.img_h * 0.5f);
        if (pj.pts2d_out && blockIdx.y == 0 && lane == 0)
          *reinterpret_cast<float2*>(pj.pts2d_out + ((size_t)b * N + n) * 2) = g;
      } else {
        g = *reinterpret_cast<const float2*>(points + (size_t)b * pts_bstride + (size_t)n * 2);
      }
      const Taps t = make_taps(g.x, g.y, H, W);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + h * 32 + lane;
        float acc = 0.0f;
        if (c < C) {
          acc = __ldg(fb + (size_t)t.o00 * C + c) * t.w00;
          acc = fmaf(__ldg(fb + (size_t)t.o01 * C + c), t.w01, acc);
          acc = fmaf(__ldg(fb + (size_t)t.o10 * C + c), t.w10, acc);
          acc = fmaf(__ldg(fb + (size_t)t.o11 * C + c), t.w11, acc);
        }
        tile[h * 32 + lane][pt] = acc;
      }
    }
  }
  __syncthreads();
  // write [64 channels][32 points]: each warp handles 8 channels, lanes over points
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int cl = warp * 8 + q;
    const int c = c0 + cl, n = n0 + lane;
    if (c < C && n < N) out[((size_t)b * C + c) * N + n] = tile[cl][lane];
  }
}

// Dense regime, both layouts (the default for NHWC maps; see dense_channels for the measured choice): a CTA stages the
// WHOLE map of one body for CG channels in shared memory as tile[pixel][channel] -- NHWC input: a straight 16-byte-vector
// copy; NCHW input: coalesced plane reads, transposed on the way in (row pitch CG+1, so both the staging stores and the
// gather reads are bank-conflict free) -- and every byte of the map is read from HBM exactly once.  The gather then runs
// with LANES = CHANNELS: the four taps of a point are four conflict-free shared-memory reads of CB = min(CG,32)
// consecutive words (the plane-staged kernel has lanes = points, i.e. 32 random addresses per read: ~3.5 wavefronts
// each, which is what bounds it), and the [32 points x CB channels] block is transposed through a per-warp tile so that
// the [B,C,N] stores are 128 contiguous bytes per (channel, 32 points).  Same tap arithmetic and order as the other
// kernels: bit-identical results.
// grid = (ceil(C/CG), B), block 256; dynamic smem = H*W*pitch*4 + 8 warps x (32 points x 8 tap words + CB x 33 floats).
template <int CG, bool kNchwIn, bool kProject>
__global__ void __launch_bounds__(256)
sample_bilinear_dense_kernel(const float* __restrict__ feat, const float* __restrict__ points, int pts_bstride,
                             float* __restrict__ out, int C, int H, int W, int N, SampleProj pj) {
  constexpr int CB = CG < 32 ? CG : 32;          // channels per warp block == lanes per point
  constexpr int PPL = 32 / CB;                   // points handled per warp step
  constexpr int kPitch = kNchwIn ? CG + 1 : CG;
  extern __shared__ __align__(16) float dense_smem[];
  const int HW = H * W;
  float* tile = dense_smem;                                        // [HW][kPitch]
  float* wtaps = dense_smem + (((size_t)HW * kPitch + 3) & ~(size_t)3);   // [8 warps][32 points][8]
  float* wout = wtaps + 8 * 32 * 8;                                // [8 warps][CB][33]
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y, c0 = blockIdx.x * CG;
  const int cg = min(CG, C - c0);
  // staging with cp.async: every copy of a thread is in flight at once (a load->store loop keeps only a few KB per SM
  // in flight and runs at a quarter of the HBM rate), no registers, and the 4-byte form does the NCHW transposition.
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  if (kNchwIn) {
    const float* src = feat + ((size_t)b * C + c0) * HW;
    const int total = cg * HW;
    int c = threadIdx.x / HW, p = threadIdx.x - c * HW;
    for (int i = threadIdx.x; i < total; i += 256) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tile_s + (uint32_t)(p * kPitch + c) * 4u), "l"(src + i) : "memory");
      p += 256;
      while (p >= HW) { p -= HW; ++c; }
    }
  } else {
    const float* src = feat + (size_t)b * HW * C + c0;
    if (cg == CG && (C & 3) == 0 && (reinterpret_cast<size_t>(feat) & 15) == 0) {
      constexpr int V4 = CG / 4;
      for (int i = threadIdx.x; i < HW * V4; i += 256) {
        const int p = i / V4, j = i - p * V4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tile_s + (uint32_t)(p * kPitch + 4 * j) * 4u),
                     "l"(src + (size_t)p * C + 4 * j) : "memory");
      }
    } else {
      for (int i = threadIdx.x; i < HW * cg; i += 256) {
        const int p = i / cg, c = i - p * cg;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tile_s + (uint32_t)(p * kPitch + c) * 4u),
                     "l"(src + (size_t)p * C + c) : "memory");
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  float ctx = 0.f, cty = 0.f, ctz = 0.f;
  if (kProject) {   // utils/geometry.py:289-307
    const float cs = pj.cam[b * 3 + 0];
    ctx = pj.cam[b * 3 + 1]; cty = pj.cam[b * 3 + 2];
    ctz = 2.0f * pj.focal / (pj.img_h * cs + 1e-9f);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / CB, cl = lane - sub * CB;      // point within the step, channel within the block
  float* my_taps = wtaps + warp * (32 * 8);
  float* my_out = wout + warp * (CB * 33);
  const float* pb = points + (size_t)b * pts_bstride;
  const int n_pblocks = (N + 31) >> 5;
  constexpr int n_cblocks = CG / CB;
  for (int blk = warp; blk < n_pblocks * n_cblocks; blk += 8) {
    const int pblk = blk / n_cblocks, cblk = blk - pblk * n_cblocks;
    const int n0 = pblk * 32, cb0 = cblk * CB;
    {   // lane l: taps of point n0 + l -> the warp's tap table
      const int n = min(n0 + lane, N - 1);
      float2 g;
      if (kProject) {
        const float* q = pb + (size_t)n * 3;
        const float px = q[0] + ctx, py = q[1] + cty, pz = q[2] + ctz;
        g.x = (pj.focal * (px / pz)) / (pj.img_w * 0.5f);
        g.y = (pj.focal * (py / pz)) / (pj.img_h * 0.5f);
        if (pj.pts2d_out && blockIdx.x == 0 && cblk == 0 && n0 + lane < N)
          *reinterpret_cast<float2*>(pj.pts2d_out + ((size_t)b * N + n) * 2) = g;
      } else {
        g = *reinterpret_cast<const float2*>(pb + (size_t)n * 2);
      }
      const Taps tp = make_taps(g.x, g.y, H, W);
      __syncwarp();      // the previous block's reads of the tables are done
      *reinterpret_cast<int4*>(my_taps + lane * 8) = make_int4(tp.o00 * kPitch, tp.o01 * kPitch, tp.o10 * kPitch, tp.o11 * kPitch);
      *reinterpret_cast<float4*>(my_taps + lane * 8 + 4) = make_float4(tp.w00, tp.w01, tp.w10, tp.w11);
      __syncwarp();
    }
    const float* tc = tile + cb0 + cl;
#pragma unroll 4
    for (int s = 0; s < CB; ++s) {
      const int pt = s * PPL + sub;
      const int4 o = *reinterpret_cast<const int4*>(my_taps + pt * 8);
      const float4 w = *reinterpret_cast<const float4*>(my_taps + pt * 8 + 4);
      float acc = tc[o.x] * w.x;
      acc = fmaf(tc[o.y], w.y, acc);
      acc = fmaf(tc[o.z], w.z, acc);
      acc = fmaf(tc[o.w], w.w, acc);
      my_out[cl * 33 + pt] = acc;
    }
    __syncwarp();
    const int n = n0 + lane;
    if (n < N) {
      float* ob = out + ((size_t)b * C + c0 + cb0) * N + n;
      const int cmax = min(CB, cg - cb0);
#pragma unroll 4
      for (int c = 0; c < cmax; ++c) ob[(size_t)c * N] = my_out[c * 33 + lane];
    }
  }
}

// ---- host side ----------------------------------------------------------------------------------
// Dense-regime plan: channels per CTA of sample_bilinear_dense_kernel (64, 32 or 16; two or more CTAs per SM), or 0.
static inline size_t dense_smem_bytes(int cgp, int nchw_in, int HW) {
  const int pitch = nchw_in ? cgp + 1 : cgp;
  const int cb = cgp < 32 ? cgp : 32;
  return ((((size_t)HW * pitch + 3) & ~(size_t)3) + 8 * 32 * 8 + 8 * cb * 33) * sizeof(float);
}
static int dense_channels(int C, int H, int W, int N, int nchw_in) {
  static const int mode = getenv("WHMR_SAMPLE_DENSE") ? atoi(getenv("WHMR_SAMPLE_DENSE")) : -1;   // 0 never, 1 whenever it fits
  if (mode == 0 || (C & 15)) return 0;
  const long long HW = (long long)H * W;
  if (mode != 1 && 4LL * N < HW) return 0;
  // Measured at configs[3] (B=1024, N=431, fraction of the HBM roofline on algorithmic bytes, 14x14 / 28x28):
  //   NHWC maps: this kernel 0.47 / 0.64 (CG=32; 0.42 / 0.63 at CG=64), direct NHWC gather 0.25 / 0.41;
  //   NCHW maps: this kernel 0.42 / 0.41 (4-byte transposing copies), plane-staged kernel 0.47 / 0.59
  // so NCHW input keeps the plane-staged kernel unless WHMR_SAMPLE_DENSE=1 asks for this one.
  if (nchw_in && mode != 1) return 0;
  if (HW > 16384) return 0;
  static const int cap = getenv("WHMR_DENSE_CG") ? atoi(getenv("WHMR_DENSE_CG")) : 32;   // experiments: 64 / 32 / 16
  for (int cgp = 64; cgp >= 16; cgp >>= 1)
    if (cgp <= C && cgp <= cap && dense_smem_bytes(cgp, nchw_in, (int)HW) <= 100 * 1024) return cgp;
  return 0;
}

template <bool kNchwIn, bool kProject>
static cudaError_t launch_dense(int cgp, const float* feat, const float* points, int pts_bstride, float* out, int B, int C,
                                int H, int W, int N, SampleProj pj, cudaStream_t st) {
  const size_t smem = dense_smem_bytes(cgp, kNchwIn, H * W);
  const dim3 grid(ceil_div(C, cgp), B);
#define WHMR_DENSE(CGV)                                                                                            \
  do {                                                                                                             \
    cudaError_t e_ = ensure_dyn_smem(sample_bilinear_dense_kernel<CGV, kNchwIn, kProject>, 100 * 1024);            \
    if (e_ != cudaSuccess) return e_;                                                                              \
    launch_pdl(kPdlSample, sample_bilinear_dense_kernel<CGV, kNchwIn, kProject>, grid, dim3(256), smem, st, feat,  \
               points, pts_bstride, out, C, H, W, N, pj);                                                          \
  } while (0)
  if (cgp == 64) WHMR_DENSE(64); else if (cgp == 32) WHMR_DENSE(32); else WHMR_DENSE(16);
#undef WHMR_DENSE
  return cudaSuccess;
}


// Dense regime of the NCHW sampler (sampling.cuh): channels per CTA for the shared-memory-staged kernel, or 0 when
// the direct gather is the better kernel (sparse sampling, or planes too large to stage a useful number of).
static int staged_channels(int C, int H, int W, int N) {
  static const int mode = getenv("WHMR_SAMPLE_STAGED") ? atoi(getenv("WHMR_SAMPLE_STAGED")) : -1;   // 0 never, 1 whenever it fits
  if (mode == 0) return 0;
  const long long HW = (long long)H * W;
  // whole planes are cheaper than the gather's two 64-byte DRAM atoms per (point, channel) up to ~8N pixels per plane
  // (configs[3], 56x56, N = 431: staged 0.648 ms vs gather 0.748 ms once the staging uses cp.async)
  if (mode != 1 && 8LL * N < HW) return 0;
  // planes per CTA: small planes want many small CTAs per SM, large planes need >= 4 planes per CTA to amortise the taps.
  // Measured optimum (configs[3], fraction of the HBM roofline): 14x14 16 KB (0.57; 64 KB: 0.50), 28x28 32 KB (0.83; 64 KB:
  // 0.77), 56x56 64 KB (0.54; gather 0.47)  =>  ~1120 * sqrt(H*W) bytes.  WHMR_STAGED_KB overrides.
  static const long long budget_env = getenv("WHMR_STAGED_KB") ? atoll(getenv("WHMR_STAGED_KB")) * 1024 : 0;
  long long budget = budget_env ? budget_env : (long long)(1120.0 * std::sqrt((double)HW));
  // (cap 56 KB: at 56x56 four planes per CTA measured 0.565 of the roofline, five -- 62.7 KB -- 0.547)
  budget = std::min<long long>(std::max<long long>(std::min<long long>(budget, 56 * 1024), 4 * HW * 4), 100 * 1024);
  const int cg = (int)std::min<long long>(C, budget / (HW * 4));
  return cg >= 4 ? cg : 0;
}

template <bool kProject>
static void launch_staged(const float* feat, const float* points, int pts_bstride, float* out, int B, int C, int H, int W,
                          int N, int cg, SampleProj pj, cudaStream_t st) {
  ensure_dyn_smem(sample_bilinear_nchw_staged_kernel<kProject>, 112 * 1024);   // per device
  launch_pdl(kPdlSample, sample_bilinear_nchw_staged_kernel<kProject>, dim3(ceil_div(C, cg), B), dim3(256),
             (size_t)cg * H * W * sizeof(float), st, feat, points, pts_bstride, out, C, H, W, N, cg, pj);
}

}  // namespace whmr
