// Pose-blend contraction on the CUDA cores, exact fp32 products (WHMR_GEMM_FP32_SIMT).
//
//   offsets[b, n] = sum_k pf[b, k] * P[k, n]      b < B, n < NP = 3*VP (planar padded), k < KP
//
// This is smplx's `torch.matmul(pose_feature, posedirs)` (twin
// models/smpl_webuser/verts.py:49-50).  It exists for bring-up and to apportion error: its
// products are exact fp32 FMAs, so GPU-vs-oracle differences here come from summation order
// only.  The production path is the tcgen05 kernel in pose_blend_tc.cuh.
//
// Classic register-tiled SGEMM: CTA tile 64 bodies x 128 coords, K step 16, 256 threads, each
// thread 4 bodies x 8 coords; operands staged in shared memory with a register prefetch of the
// next K slab.  FFMA-bound by construction (3 LDS.128 per 32 FFMA).
#pragma once
#include "common.cuh"

namespace whmr {

constexpr int kSimtBM = 64, kSimtBN = 128, kSimtBK = 16;

__global__ void __launch_bounds__(256)
pose_blend_simt_kernel(const float* __restrict__ pf,  // [B,KP]
                       const float* __restrict__ P,   // [KP,NP]
                       float* __restrict__ out,       // [B,NP]
                       int B, int KP, int NP) {
  __shared__ __align__(16) float Ps[kSimtBK][kSimtBN];
  __shared__ __align__(16) float As[kSimtBK][kSimtBM + 4];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * kSimtBN;
  const int b0 = blockIdx.y * kSimtBM;
  const int tn = tid & 15;   // coords tn*4..+3 and 64+tn*4..+3
  const int tm = tid >> 4;   // bodies tm*4..+3

  // loader roles
  const int pr = tid >> 5;          // P rows pr and pr+8
  const int pc = (tid & 31) * 4;    // P col (float4)
  const int ab = tid >> 2;          // body within tile
  const int ak = (tid & 3) * 4;     // k offset (float4)
  const bool a_ok = (b0 + ab) < B;

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[i][q] = 0.0f;

  float4 p_reg0, p_reg1, a_reg;
  auto load_slab = [&](int k0) {
    p_reg0 = *reinterpret_cast<const float4*>(P + (size_t)(k0 + pr) * NP + n0 + pc);
    p_reg1 = *reinterpret_cast<const float4*>(P + (size_t)(k0 + pr + 8) * NP + n0 + pc);
    a_reg = a_ok ? *reinterpret_cast<const float4*>(pf + (size_t)(b0 + ab) * KP + k0 + ak)
                 : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto store_slab = [&]() {
    *reinterpret_cast<float4*>(&Ps[pr][pc]) = p_reg0;
    *reinterpret_cast<float4*>(&Ps[pr + 8][pc]) = p_reg1;
    As[ak + 0][ab] = a_reg.x;
    As[ak + 1][ab] = a_reg.y;
    As[ak + 2][ab] = a_reg.z;
    As[ak + 3][ab] = a_reg.w;
  };

  load_slab(0);
  for (int k0 = 0; k0 < KP; k0 += kSimtBK) {
    __syncthreads();   // previous slab fully consumed
    store_slab();
    __syncthreads();
    if (k0 + kSimtBK < KP) load_slab(k0 + kSimtBK);
#pragma unroll
    for (int k = 0; k < kSimtBK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
      const float4 q0 = *reinterpret_cast<const float4*>(&Ps[k][tn * 4]);
      const float4 q1 = *reinterpret_cast<const float4*>(&Ps[k][64 + tn * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[i][q] = fmaf(av[i], qv[q], acc[i][q]);
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + tm * 4 + i;
    if (b < B) {
      float* o = out + (size_t)b * NP + n0;
      *reinterpret_cast<float4*>(o + tn * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(o + 64 + tn * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
}

}  // namespace whmr
