// Linear blend skinning with sparse weights, fused with the shape blend.
//
//   v_posed  = v_template + (shapedirs . beta + pose_offsets)  (the bracket comes out of the pose-blend
//              contraction, whose K dimension carries the 10 shape coefficients after the 207 pose terms)
//   T_v      = sum_j w_vj A_j ;  v' = T_v . [v_posed ; 1]      (SURVEY K6 + K7; lbs.py:63-79)
//
// The reference materialises T as a dense [B,6890,24]x[B,24,16] matmul (441 KB/body of
// intermediate).  Here the skinning weights are stored as an ELL table (ell_k non-zeros per
// vertex, 4 for SMPL) in coordinate-planar layout so a warp reads them coalesced, the per-vertex
// constants (template, 3x10 shape directions, joint ids, weights) live in registers across the
// CTA's whole body loop, and the only per-(vertex, body) traffic is 12 B of pose offsets in and
// 12 B of vertex out.  Stores go through a per-warp shared-memory transpose so each warp writes
// 384 contiguous bytes per body.
//
// CTA = 128 threads = 128 consecutive vertices; grid = (VP/128, ceil(B/kSkinBodies)).
// HBM bytes per body: 4*NP (offsets in, L2-resident when chunked) + 4*3V (verts out) + 4*J*12.
#pragma once
#include "common.cuh"

namespace whmr {

constexpr int kSkinBodies = 8;   // bodies per CTA (A matrices for all of them staged in smem)

struct SkinParams {
  const float* offsets;   // [B,NP] planar padded pose offsets
  const float* A;         // [B,J,12]
  const float* betas;     // [B,NB]
  const float* transl;    // [B,3] or null
  const float* v_template_p;  // [3,VP]
  const float* shapedirs_p;   // [3,NB,VP]
  const int* ell_idx;         // [ell_k,VP]
  const float* ell_w;         // [ell_k,VP]
  float* verts;               // [B,V,3]
  int B, V, VP, NP, J, NB, ell_k;
};

template <int ELLK, int NBT>   // ELLK = 0 / NBT = 0: runtime-sized generic path
__global__ void __launch_bounds__(kVertTile) skin_kernel(SkinParams p) {
  extern __shared__ __align__(16) float smem[];
  // smem: A_s [kSkinBodies][J*12] | beta_s [kSkinBodies][kMaxBetas] | tr_s [kSkinBodies][4] | stage [4][96]
  float* A_s = smem;
  float* beta_s = A_s + kSkinBodies * p.J * 12;
  float* tr_s = beta_s + kSkinBodies * kMaxBetas;
  float* stage = tr_s + kSkinBodies * 4;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int v = blockIdx.x * kVertTile + tid;      // < VP always
  const int b0 = blockIdx.y * kSkinBodies;
  const int nb_here = min(kSkinBodies, p.B - b0);
  const int ell_k = ELLK > 0 ? ELLK : p.ell_k;
  const int NB = NBT > 0 ? NBT : p.NB;

  // ---- stage A, betas, transl for this CTA's bodies -------------------------------------------
  {
    const int nA4 = nb_here * p.J * 3;   // float4 count
    const float4* src = reinterpret_cast<const float4*>(p.A + (size_t)b0 * p.J * 12);
    float4* dst = reinterpret_cast<float4*>(A_s);
    for (int i = tid; i < nA4; i += kVertTile) dst[i] = src[i];
    for (int i = tid; i < nb_here * kMaxBetas; i += kVertTile) {
      const int bi = i / kMaxBetas, k = i % kMaxBetas;
      beta_s[i] = k < NB ? p.betas[(size_t)(b0 + bi) * NB + k] : 0.0f;
    }
    for (int i = tid; i < nb_here * 4; i += kVertTile) {
      const int bi = i >> 2, c = i & 3;
      tr_s[i] = (p.transl && c < 3) ? p.transl[(size_t)(b0 + bi) * 3 + c] : 0.0f;
    }
  }

  // ---- per-vertex constants in registers -------------------------------------------------------
  float tmpl[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) tmpl[c] = p.v_template_p[c * p.VP + v];
  int jid[ELLK > 0 ? ELLK : 1];
  float jw[ELLK > 0 ? ELLK : 1];
  if (ELLK > 0) {
#pragma unroll
    for (int k = 0; k < ELLK; ++k) {
      jid[k] = p.ell_idx[(size_t)k * p.VP + v] * 12;
      jw[k] = p.ell_w[(size_t)k * p.VP + v];
    }
  }
  __syncthreads();

  const size_t out_base = (size_t)(blockIdx.x * kVertTile + warp * 32) * 3;   // float index in a body
  const int out_lim = p.V * 3;
  float* stg = stage + warp * 96;

  // prefetch offsets of the first body
  float ox = 0.f, oy = 0.f, oz = 0.f;
  if (nb_here > 0) {
    const float* o = p.offsets + (size_t)b0 * p.NP + v;
    ox = o[0]; oy = o[p.VP]; oz = o[2 * p.VP];
  }

  for (int bi = 0; bi < nb_here; ++bi) {
    const int b = b0 + bi;
    const float cx = ox, cy = oy, cz = oz;
    if (bi + 1 < nb_here) {   // software prefetch of the next body's offsets
      const float* o = p.offsets + (size_t)(b + 1) * p.NP + v;
      ox = o[0]; oy = o[p.VP]; oz = o[2 * p.VP];
    }
    // offsets already hold pose offsets + shape blend (both come out of the pose-blend contraction)
    const float px = cx + tmpl[0], py = cy + tmpl[1], pz = cz + tmpl[2];

    // blended transform T = sum_k w_k A_{j_k}
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0, t2 = t0;
    const float* Ab = A_s + bi * p.J * 12;
    auto accum = [&](int joff, float w) {
      const float4 r0 = *reinterpret_cast<const float4*>(Ab + joff);
      const float4 r1 = *reinterpret_cast<const float4*>(Ab + joff + 4);
      const float4 r2 = *reinterpret_cast<const float4*>(Ab + joff + 8);
      t0.x = fmaf(w, r0.x, t0.x); t0.y = fmaf(w, r0.y, t0.y); t0.z = fmaf(w, r0.z, t0.z); t0.w = fmaf(w, r0.w, t0.w);
      t1.x = fmaf(w, r1.x, t1.x); t1.y = fmaf(w, r1.y, t1.y); t1.z = fmaf(w, r1.z, t1.z); t1.w = fmaf(w, r1.w, t1.w);
      t2.x = fmaf(w, r2.x, t2.x); t2.y = fmaf(w, r2.y, t2.y); t2.z = fmaf(w, r2.z, t2.z); t2.w = fmaf(w, r2.w, t2.w);
    };
    if (ELLK > 0) {
#pragma unroll
      for (int k = 0; k < ELLK; ++k) accum(jid[k], jw[k]);
    } else {
      for (int k = 0; k < ell_k; ++k)
        accum(p.ell_idx[(size_t)k * p.VP + v] * 12, p.ell_w[(size_t)k * p.VP + v]);
    }
    // v' = T . [v_posed ; 1]   (+ transl, smplx body_models.py apply_trans)
    float rx = fmaf(t0.x, px, fmaf(t0.y, py, fmaf(t0.z, pz, t0.w)));
    float ry = fmaf(t1.x, px, fmaf(t1.y, py, fmaf(t1.z, pz, t1.w)));
    float rz = fmaf(t2.x, px, fmaf(t2.y, py, fmaf(t2.z, pz, t2.w)));
    rx += tr_s[bi * 4 + 0]; ry += tr_s[bi * 4 + 1]; rz += tr_s[bi * 4 + 2];

    // transpose through smem: lane l holds xyz of vertex l -> 3 coalesced 128-byte rows
    stg[lane * 3 + 0] = rx; stg[lane * 3 + 1] = ry; stg[lane * 3 + 2] = rz;
    __syncwarp();
    float* ob = p.verts + (size_t)b * p.V * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const size_t idx = out_base + r * 32 + lane;
      if (idx < (size_t)out_lim) ob[idx] = stg[r * 32 + lane];
    }
    __syncwarp();
  }
}

static inline size_t skin_smem_bytes(int J) {
  return sizeof(float) * (size_t)(kSkinBodies * J * 12 + kSkinBodies * kMaxBetas + kSkinBodies * 4 + 4 * 96);
}

}  // namespace whmr
