// Fused per-body pose kernel: Rodrigues (optional) + rest joints from betas + 24-joint kinematic
// chain + relative (skinning) transforms + pose-blend feature.
//
// Replaces, per SMPL.forward call, smplx's batch_rodrigues (~15 launches), vertices2joints
// (dense [24,6890] einsum), the 23-step Python loop of 4x4 bmm's in batch_rigid_transform, the
// stack/pad/subtract that removes the rest pose, and the (R - I) pose-feature build
// (SURVEY K1, K3, K5).  In-tree statement of the math: models/smpl_webuser/lbs.py:27-60,
// posemapper.py:36-43, serialization.py:104-107.
//
// Mapping: one warp per body, one lane per joint (J <= 32).  The chain is evaluated level by
// level of the kinematic tree (max depth 9 for SMPL): at level d every lane pulls its parent's
// 3x4 world transform with 12 warp shuffles and the lanes whose joint sits at depth d compose
// G_j = G_parent * [R_j | J_j - J_parent].  No shared memory, no block barrier.
// Traffic per body: reads 4*(NB + 216|72) B, writes 4*(J*12 + J*3 + KP) B  (~2.5 KB).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "rotations.cuh"

namespace whmr {

struct ChainParams {
  const float* betas;   // [B,NB]
  const float* pose;    // [B,J,9] or [B,J,3]  (root_pose != null: joints 1..J-1 only, [B,J-1,9] or [B,J-1,3])
  const float* root_pose;   // [B,9] or [B,3] or null
  const float* transl;  // [B,3] or null
  int pose_is_rotmat;
  int B, J, NB, KP, max_depth;
  const float* J_template;   // [J,3]
  const float* J_shapedirs;  // [J,3,NB]
  const int* parents;        // [J]
  const int* depth;          // [J]
  float* A;                  // [B,J,12]
  float* joints;             // [B,J,3] or null
  float* A_user;             // optional second copy of A for the caller (rel_transforms), or null
  float* pf;                 // [B,KP] or null
  __nv_bfloat16* pf_split;   // [B,2,KP] hi|lo or null
  float* pf_tf32;            // [B,2,KP] hi|lo tf32-valued floats or null
  float* At;                 // [2, At_rows, 32] tf32 hi|lo of A transposed: row (b*12+e), column = joint; or null
  size_t At_part_stride;     // floats between the hi and lo parts (= At_rows*32)
  // Regressor.forward's rotation glue (models/whmr.py:129-130, 174, 190), folded in: no extra launches, no extra HBM pass
  int gram_schmidt;          // rotmat mode: R <- unbiased_gram_schmidt(R) before anything else (eval mode)
  float* rotmat_out;         // [B,J,9] the rotations actually used (the orthonormalised ones), or null
  float* pose_aa_out;        // [B,J,3] rotation_matrix_to_angle_axis of them ('pose'), or null
  float* theta_out;          // [B, 3+NB+3J] = cat(cam, betas, pose) ('theta'), or null; needs `cam`
  const float* cam;          // [B,3]
  __half* At16;              // [At_rows, 64] fp16: A transposed, hi in columns 0..31, lo in 32..63, row = (b/2)*24 + e*2 + b%2 (fused kernel); or null
};

// smplx.lbs.batch_rodrigues for one vector: angle = ||v + 1e-8||, R = I + sin*K + (1-cos)*K*K
__device__ __forceinline__ void rodrigues_smplx(float x, float y, float z, float* R) {
  const float ex = x + 1e-8f, ey = y + 1e-8f, ez = z + 1e-8f;
  const float angle = sqrtf(ex * ex + ey * ey + ez * ez);
  const float dx = x / angle, dy = y / angle, dz = z / angle;
  const float s = sinf(angle), c1 = 1.0f - cosf(angle);
  // K = [[0,-dz,dy],[dz,0,-dx],[-dy,dx,0]]
  const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
  const float xy = dx * dy, xz = dx * dz, yz = dy * dz;
  R[0] = 1.0f + c1 * (-(yy + zz));
  R[1] = s * (-dz) + c1 * xy;
  R[2] = s * dy + c1 * xz;
  R[3] = s * dz + c1 * xy;
  R[4] = 1.0f + c1 * (-(xx + zz));
  R[5] = s * (-dx) + c1 * yz;
  R[6] = s * (-dy) + c1 * xz;
  R[7] = s * dx + c1 * yz;
  R[8] = 1.0f + c1 * (-(xx + yy));
}

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

constexpr int kChainWarpsPerBlock = 4;

__global__ void __launch_bounds__(kChainWarpsPerBlock * 32) smpl_chain_kernel(ChainParams p) {
  pdl_wait();      // pose / betas come from the caller's previous kernels; the scratch may still be read by them
  pdl_trigger();
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * kChainWarpsPerBlock + (threadIdx.x >> 5);
  if (b >= p.B) return;  // warp-uniform
  const bool active = lane < p.J;
  const int j = active ? lane : 0;

  // ---- local rotation -------------------------------------------------------------------
  float R[9];
  const int pstride = p.pose_is_rotmat ? 9 : 3;
  const float* src = !p.root_pose ? p.pose + ((size_t)b * p.J + j) * pstride
                     : (j == 0 ? p.root_pose + (size_t)b * pstride : p.pose + ((size_t)b * (p.J - 1) + (j - 1)) * pstride);
  if (p.pose_is_rotmat) {
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = src[i];
    if (p.gram_schmidt) unbiased_gram_schmidt_dev(R, R);
  } else {
    rodrigues_smplx(src[0], src[1], src[2], R);
  }

  // ---- rest joint J_j = J_template_j + J_shapedirs_j . beta  (pre-contracted regressor) ----
  const float my_beta = lane < p.NB ? p.betas[(size_t)b * p.NB + lane] : 0.0f;
  // every load of this lane is issued before the first use: with a run-time trip count the loop below waited one L2
  // latency per shape coefficient (10 x ~0.3 us of a 7.5 us kernel).  Terms k >= NB multiply beta = 0 by 0: same bits.
  int par = active ? p.parents[j] : 0;
  const int dep = active ? p.depth[j] : -1;
  float Jr[3], jsd[3 * kMaxBetas];
#pragma unroll
  for (int c = 0; c < 3; ++c) Jr[c] = p.J_template[j * 3 + c];
#pragma unroll
  for (int k = 0; k < kMaxBetas; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) jsd[c * kMaxBetas + k] = k < p.NB ? __ldg(p.J_shapedirs + (j * 3 + c) * p.NB + k) : 0.0f;
#pragma unroll
  for (int k = 0; k < kMaxBetas; ++k) {
    const float bk = __shfl_sync(full, my_beta, k);
#pragma unroll
    for (int c = 0; c < 3; ++c) Jr[c] = fmaf(bk, jsd[c * kMaxBetas + k], Jr[c]);
  }

  const bool is_root = par < 0;
  if (is_root) par = j;

  // local transform [R | J_j - J_parent]; root: [R | J_0]
  float G[12];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float jp = __shfl_sync(full, Jr[c], par);
    G[c * 4 + 3] = is_root ? Jr[c] : Jr[c] - jp;
    G[c * 4 + 0] = R[c * 3 + 0];
    G[c * 4 + 1] = R[c * 3 + 1];
    G[c * 4 + 2] = R[c * 3 + 2];
  }

  // ---- kinematic chain, one tree level per step ---------------------------------------------
  for (int d = 1; d <= p.max_depth; ++d) {
    float P[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) P[i] = __shfl_sync(full, G[i], par);
    if (dep == d) {
      float N[12];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float acc = P[r * 4 + 0] * G[0 * 4 + c];
          acc = fmaf(P[r * 4 + 1], G[1 * 4 + c], acc);
          acc = fmaf(P[r * 4 + 2], G[2 * 4 + c], acc);
          if (c == 3) acc += P[r * 4 + 3];
          N[r * 4 + c] = acc;
        }
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) G[i] = N[i];
    }
  }

  // ---- skinning operand for the tensor-core kernel: A_j (with rest pose removed) transposed so that
  //      joints run along K: At[part][(b*12+e)*32 + j], zero for j >= J (lanes >= J write the padding)
  float Arow[12];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float t = G[r * 4 + 0] * Jr[0];
    t = fmaf(G[r * 4 + 1], Jr[1], t);
    t = fmaf(G[r * 4 + 2], Jr[2], t);
    Arow[r * 4 + 0] = G[r * 4 + 0]; Arow[r * 4 + 1] = G[r * 4 + 1]; Arow[r * 4 + 2] = G[r * 4 + 2];
    Arow[r * 4 + 3] = G[r * 4 + 3] - t;
  }
  if (p.At) {
    float* a0 = p.At + ((size_t)b * 12) * 32 + lane;
#pragma unroll
    for (int e = 0; e < 12; ++e) {
      const float x = active ? Arow[e] : 0.0f;
      const float hi = tf32_round(x);
      a0[(size_t)e * 32] = hi;
      a0[(size_t)e * 32 + p.At_part_stride] = tf32_round(x - hi);
    }
  }

  if (p.At16) {   // fp16 hi|lo split: 22 mantissa bits, |lo| below the fp16 normal range only costs < 6e-8 absolute
    // rows of a body PAIR interleaved, (pair, element, body-in-pair): the blended transforms of two bodies then land
    // in adjacent TMEM columns = adjacent registers of the skinning epilogue, which applies them with packed f32x2 FMAs
    __half* a0 = p.At16 + ((size_t)(b >> 1) * 24 + (b & 1)) * 64 + lane;
#pragma unroll
    for (int e = 0; e < 12; ++e) {
      const float x = active ? Arow[e] : 0.0f;
      const __half hi = __float2half_rn(x);
      a0[(size_t)e * 128] = hi;
      a0[(size_t)e * 128 + 32] = __float2half_rn(x - __half2float(hi));
    }
  }

  if (p.theta_out && lane < 3 + p.NB) {   // theta = [cam | betas | pose]; lanes 0..12 copy the head
    p.theta_out[(size_t)b * (3 + p.NB + 3 * p.J) + lane] =
        lane < 3 ? (p.cam ? p.cam[(size_t)b * 3 + lane] : 0.0f) : p.betas[(size_t)b * p.NB + (lane - 3)];
  }
  if (!active) return;

  // ---- outputs ----------------------------------------------------------------------------
  if (p.rotmat_out) {
    float* ro = p.rotmat_out + ((size_t)b * p.J + j) * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) ro[i] = R[i];
  }
  if (p.pose_aa_out || p.theta_out) {
    float ax, ay, az;
    rotmat_to_axis_angle_dev(R, ax, ay, az);
    if (p.pose_aa_out) {
      float* o = p.pose_aa_out + ((size_t)b * p.J + j) * 3;
      o[0] = ax; o[1] = ay; o[2] = az;
    }
    if (p.theta_out) {
      float* o = p.theta_out + (size_t)b * (3 + p.NB + 3 * p.J) + 3 + p.NB + j * 3;
      o[0] = ax; o[1] = ay; o[2] = az;
    }
  }
  if (p.joints) {
    float tx = 0.f, ty = 0.f, tz = 0.f;
    if (p.transl) {
      tx = p.transl[(size_t)b * 3 + 0];
      ty = p.transl[(size_t)b * 3 + 1];
      tz = p.transl[(size_t)b * 3 + 2];
    }
    float* jo = p.joints + ((size_t)b * p.J + j) * 3;
    jo[0] = G[3] + tx;
    jo[1] = G[7] + ty;
    jo[2] = G[11] + tz;
  }
  // A_j = G_j with translation  t - R_g . J_j   (rest-pose removal, lbs.py:49-58)
  float4 row[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) row[r] = make_float4(Arow[r * 4 + 0], Arow[r * 4 + 1], Arow[r * 4 + 2], Arow[r * 4 + 3]);
  float4* Ao = reinterpret_cast<float4*>(p.A + ((size_t)b * p.J + j) * 12);
  Ao[0] = row[0]; Ao[1] = row[1]; Ao[2] = row[2];
  if (p.A_user) {
    float* Au = p.A_user + ((size_t)b * p.J + j) * 12;  // caller buffer: no alignment assumption
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      Au[r * 4 + 0] = row[r].x; Au[r * 4 + 1] = row[r].y; Au[r * 4 + 2] = row[r].z; Au[r * 4 + 3] = row[r].w;
    }
  }

  // pose feature (R_j - I) for j >= 1, row-major per joint (posemapper.py:36-43), zero padded to KP
  const int nfeat = (p.J - 1) * 9;
  if (j >= 1) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float f = R[i] - ((i == 0 || i == 4 || i == 8) ? 1.0f : 0.0f);
      const size_t o = (size_t)(j - 1) * 9 + i;
      if (p.pf) p.pf[(size_t)b * p.KP + o] = f;
      if (p.pf_split) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(f);
        const __nv_bfloat16 lo = __float2bfloat16_rn(f - __bfloat162float(hi));
        p.pf_split[((size_t)b * 2 + 0) * p.KP + o] = hi;
        p.pf_split[((size_t)b * 2 + 1) * p.KP + o] = lo;
      }
      if (p.pf_tf32) {
        const float hi = tf32_round(f);
        const float lo = tf32_round(f - hi);
        p.pf_tf32[((size_t)b * 2 + 0) * p.KP + o] = hi;
        p.pf_tf32[((size_t)b * 2 + 1) * p.KP + o] = lo;
      }
    }
  }
  // K tail: the NB shape coefficients (the shape blend is folded into the same contraction:
  // rows nfeat..nfeat+NB-1 of the posedirs operand hold shapedirs), then zero padding up to KP
  for (int o = nfeat + j; o < p.KP; o += p.J) {
    const float f = (o - nfeat) < p.NB ? p.betas[(size_t)b * p.NB + (o - nfeat)] : 0.0f;
    if (p.pf) p.pf[(size_t)b * p.KP + o] = f;
    if (p.pf_split) {
      const __nv_bfloat16 hi = __float2bfloat16_rn(f);
      p.pf_split[((size_t)b * 2 + 0) * p.KP + o] = hi;
      p.pf_split[((size_t)b * 2 + 1) * p.KP + o] = __float2bfloat16_rn(f - __bfloat162float(hi));
    }
    if (p.pf_tf32) {
      const float hi = tf32_round(f);
      p.pf_tf32[((size_t)b * 2 + 0) * p.KP + o] = hi;
      p.pf_tf32[((size_t)b * 2 + 1) * p.KP + o] = tf32_round(f - hi);
    }
  }
}

// Standalone smplx-variant Rodrigues: aa [n,3] -> R [n,3,3]
__global__ void __launch_bounds__(256) rodrigues_kernel(const float* __restrict__ aa, int n,
                                                        float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float R[9];
  rodrigues_smplx(aa[(size_t)i * 3 + 0], aa[(size_t)i * 3 + 1], aa[(size_t)i * 3 + 2], R);
#pragma unroll
  for (int k = 0; k < 9; ++k) out[(size_t)i * 9 + k] = R[k];
}

}  // namespace whmr
