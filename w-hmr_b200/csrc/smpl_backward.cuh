// SMPL backward: gradients of vertices / posed chain joints w.r.t. betas and the rotation
// matrices, so the drop-in can sit inside Trainer.train_step (core/trainer.py:380-636, losses on pred_vertices,
// pred_keypoints_3d and the projected keypoints).  Chain rule over the forward of smpl_chain.cuh / skinning.cuh:
//
//   v'_v = T_v [p_v ; 1],  T_v = sum_j w_vj A_j,  p_v = t_v + (P pf)_v           (lbs.py:63-79, verts.py:42-50)
//     g_p_v = Trot_v^T g_v ;  g_A_j += w_vj * (g_v (x) [p_v ; 1])                 skin_backward_kernel
//     g_pf  = P^T g_p                                                             pose_blend_backward_kernel
//   A_j = [Rg_j | tg_j - Rg_j J_j],  Rg_j = Rg_p R_j,  tg_j = Rg_p (J_j - J_p) + tg_p,  J = Jt + Jd beta,
//   pf = [(R_j - I)_{j>=1} ; beta]                                                chain_backward_kernel
//
// CUDA-core kernels (tests: autograd through the float64 oracle).  The per-joint transform gradients are a small dense
// contraction per CTA (no shared-memory atomics), the P^T contraction is a register-tiled FFMA kernel with a split K;
// moving both onto the tensor cores is a later step.  Global fp32 atomics across CTAs => the last bits of the
// gradients depend on the execution order.
#pragma once
#include "common.cuh"
#include "readout.cuh"
#include "smpl_chain.cuh"

namespace whmr {

constexpr int kBwdBodies = 8;

struct SkinBwdParams {
  const float* g_verts;       // [B,V,3]
  const float* offsets;       // [B,NP] planar padded (pose offsets + shape blend), recomputed by the forward GEMM
  const float* A;             // [B,J,12]
  const float* v_template_p;  // [3,VP]
  const int* ell_idx;         // [ell_k,VP]
  const float* ell_w;         // [ell_k,VP]
  float* g_offsets;           // [B,NP] planar padded (pad vertices get 0)
  float* g_A;                 // [B,J,12], zero-initialised, accumulated with atomics
  int B, V, VP, NP, J, ell_k;
};

// CTA = 128 consecutive vertices x kBwdBodies bodies, 256 threads.
//  phase A (thread = vertex x body half): blended rotation from the ELL weights, g_p = Trot^T g_v -> g_offsets, and the
//           outer products gT[v] = g_v (x) [p_v ; 1] staged in shared memory, element-major;
//  phase B (warp = body, lane = 3 joints x 3 elements): g_A[j][e] += sum_v W[v][j] gT[v][e] as a small dense
//           contraction over the tile's 128 vertices against a dense [128, J] weight tile rebuilt from the ELL form --
//           every shared-memory load is a broadcast, no atomics; one global atomicAdd per (body, joint, element) and CTA.
//  (First version: 48 shared-memory atomicAdds per vertex and body, 316 us at B=256; this one: see profiles/r01_notes.md.)
constexpr int kBwdThreads = 256;
constexpr int kGtLd = kVertTile + 1;   // element stride of the staged outer products (odd: the 4 element groups of a warp hit 4 banks)
__host__ __device__ inline size_t skin_backward_smem_bytes(int J) {
  // A_s [bodies][J*12] + W_s [128][JP] + gT_s [bodies][12][128], JP = J rounded up to a multiple of 3
  const int JP = (J + 2) / 3 * 3;
  return ((size_t)kBwdBodies * J * 12 + (size_t)kVertTile * JP + (size_t)kBwdBodies * 12 * kGtLd) * sizeof(float);
}

__global__ void __launch_bounds__(kBwdThreads) skin_backward_kernel(SkinBwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int JP = (p.J + 2) / 3 * 3;
  float* A_s = smem;                                         // [kBwdBodies][J*12]
  float* W_s = A_s + kBwdBodies * p.J * 12;                   // [128][JP]
  float* gT_s = W_s + kVertTile * JP;                         // [kBwdBodies][12][kGtLd]
  const int tid = threadIdx.x;
  const int v0 = blockIdx.x * kVertTile;
  const int b0 = blockIdx.y * kBwdBodies;
  const int nb_here = min(kBwdBodies, p.B - b0);
  const int nA = nb_here * p.J * 12;
  for (int i = tid; i < nA; i += kBwdThreads) A_s[i] = p.A[(size_t)b0 * p.J * 12 + i];
  for (int i = tid; i < kVertTile * JP; i += kBwdThreads) W_s[i] = 0.f;
  __syncthreads();
  if (tid < kVertTile) {   // dense weight tile from the ELL rows (a vertex's joints are distinct: no write conflicts)
    const int v = v0 + tid;
    for (int k = 0; k < p.ell_k; ++k) {
      const float w = p.ell_w[(size_t)k * p.VP + v];
      if (w != 0.f) W_s[tid * JP + p.ell_idx[(size_t)k * p.VP + v]] += w;
    }
  }
  // ---- phase A ----
  {
    const int lv = tid & (kVertTile - 1), half = tid >> 7;     // bodies [half*4, half*4+4)
    const int v = v0 + lv;
    const bool real = v < p.V;
    const float tx = p.v_template_p[v], ty = p.v_template_p[p.VP + v], tz = p.v_template_p[2 * p.VP + v];
    int jidx[4]; float jw[4];                                  // first four ELL entries in registers (the SMPL case)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      jidx[k] = k < p.ell_k ? p.ell_idx[(size_t)k * p.VP + v] * 12 : 0;
      jw[k] = k < p.ell_k ? p.ell_w[(size_t)k * p.VP + v] : 0.f;
    }
    for (int bi = half * (kBwdBodies / 2); bi < min(nb_here, (half + 1) * (kBwdBodies / 2)); ++bi) {
      const int b = b0 + bi;
      float gx = 0.f, gy = 0.f, gz = 0.f;
      if (real) {
        const float* g = p.g_verts + ((size_t)b * p.V + v) * 3;
        gx = g[0]; gy = g[1]; gz = g[2];
      }
      const float* o = p.offsets + (size_t)b * p.NP + v;
      const float px = o[0] + tx, py = o[p.VP] + ty, pz = o[2 * p.VP] + tz;
      const float* Ab = A_s + bi * p.J * 12;
      float r[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // blended rotation part of T
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float w = jw[k];
        const float* a = Ab + jidx[k];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int c = 0; c < 3; ++c) r[rr * 3 + c] = fmaf(w, a[rr * 4 + c], r[rr * 3 + c]);
      }
      for (int k = 4; k < p.ell_k; ++k) {   // models with more than four influences per vertex
        const float w = p.ell_w[(size_t)k * p.VP + v];
        if (w == 0.f) continue;
        const float* a = Ab + p.ell_idx[(size_t)k * p.VP + v] * 12;
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int c = 0; c < 3; ++c) r[rr * 3 + c] = fmaf(w, a[rr * 4 + c], r[rr * 3 + c]);
      }
      // g_p = Trot^T g_v
      float* go = p.g_offsets + (size_t)b * p.NP + v;
      go[0] = r[0] * gx + r[3] * gy + r[6] * gz;
      go[p.VP] = r[1] * gx + r[4] * gy + r[7] * gz;
      go[2 * p.VP] = r[2] * gx + r[5] * gy + r[8] * gz;
      float* gt = gT_s + (size_t)bi * 12 * kGtLd + lv;            // [e][v]
      gt[0 * kGtLd] = gx * px; gt[1 * kGtLd] = gx * py; gt[2 * kGtLd] = gx * pz; gt[3 * kGtLd] = gx;
      gt[4 * kGtLd] = gy * px; gt[5 * kGtLd] = gy * py; gt[6 * kGtLd] = gy * pz; gt[7 * kGtLd] = gy;
      gt[8 * kGtLd] = gz * px; gt[9 * kGtLd] = gz * py; gt[10 * kGtLd] = gz * pz; gt[11 * kGtLd] = gz;
    }
  }
  __syncthreads();
  // ---- phase B ----
  {
    const int bi = tid >> 5, lane = tid & 31;
    if (bi >= nb_here) return;                                  // warp-uniform
    const int eg = (lane & 3) * 3;                              // elements eg .. eg+2
    const float* gt = gT_s + (size_t)bi * 12 * kGtLd + eg * kGtLd;
    for (int jg = (lane >> 2) * 3; jg < JP; jg += 24) {         // joints jg .. jg+2 (one pass for J <= 24)
      float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float* wv = W_s + jg;
#pragma unroll 8
      for (int v = 0; v < kVertTile; ++v) {
        const float w0 = wv[v * JP], w1 = wv[v * JP + 1], w2 = wv[v * JP + 2];
        const float t0 = gt[v], t1 = gt[kGtLd + v], t2 = gt[2 * kGtLd + v];
        acc[0] = fmaf(w0, t0, acc[0]); acc[1] = fmaf(w0, t1, acc[1]); acc[2] = fmaf(w0, t2, acc[2]);
        acc[3] = fmaf(w1, t0, acc[3]); acc[4] = fmaf(w1, t1, acc[4]); acc[5] = fmaf(w1, t2, acc[5]);
        acc[6] = fmaf(w2, t0, acc[6]); acc[7] = fmaf(w2, t1, acc[7]); acc[8] = fmaf(w2, t2, acc[8]);
      }
      float* dst = p.g_A + ((size_t)(b0 + bi) * p.J) * 12;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (jg + a >= p.J) continue;
#pragma unroll
        for (int e = 0; e < 3; ++e)
          if (acc[a * 3 + e] != 0.f) atomicAdd(dst + (jg + a) * 12 + eg + e, acc[a * 3 + e]);
      }
    }
  }
}

// g_pf[b, k] = sum_n g_off[b, n] * P[k, n]   (P = posedirs_p [KP, NP] fp32 planar; K of this product = NP = 20736)
// grid = (ceil(KP/64), ceil(B/64), split), 256 threads: a 64 x 64 output tile per CTA over one slice of n, 4 x 4 outputs
// per thread; both operands are staged transposed ([n][row], 32 n at a time) so that the inner loop is two 16-byte
// shared-memory loads per 16 FMAs.  Slices are accumulated into g_pf (zero-initialised) with atomics.
constexpr int kPbTile = 64, kPbN = 32, kPbLd = kPbTile + 4;
__global__ void __launch_bounds__(256)
pose_blend_backward_kernel(const float* __restrict__ g_off, const float* __restrict__ P, float* __restrict__ g_pf, int B,
                           int KP, int NP) {
  __shared__ __align__(16) float Gs[kPbN][kPbLd], Ps[kPbN][kPbLd];
  const int tid = threadIdx.x;
  const int tb = tid >> 4, tk = tid & 15;                 // bodies tb*4.., pose-feature columns tk*4..
  const int k0 = blockIdx.x * kPbTile, b0 = blockIdx.y * kPbTile;
  const int n_per = NP / gridDim.z;                       // a multiple of kPbN (the host picks the split)
  const int n_begin = blockIdx.z * n_per, n_end = n_begin + n_per;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int n0 = n_begin; n0 < n_end; n0 += kPbN) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256, row = idx >> 3, q = (idx & 7) * 4;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f), w = g;
      if (b0 + row < B) g = __ldg(reinterpret_cast<const float4*>(g_off + (size_t)(b0 + row) * NP + n0 + q));
      if (k0 + row < KP) w = __ldg(reinterpret_cast<const float4*>(P + (size_t)(k0 + row) * NP + n0 + q));
      Gs[q + 0][row] = g.x; Gs[q + 1][row] = g.y; Gs[q + 2][row] = g.z; Gs[q + 3][row] = g.w;
      Ps[q + 0][row] = w.x; Ps[q + 1][row] = w.y; Ps[q + 2][row] = w.z; Ps[q + 3][row] = w.w;
    }
    __syncthreads();
#pragma unroll 8
    for (int n = 0; n < kPbN; ++n) {
      const float4 a = *reinterpret_cast<const float4*>(&Gs[n][tb * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ps[n][tk * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + tb * 4 + i;
    if (b >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tk * 4 + j;
      if (k < KP) atomicAdd(g_pf + (size_t)b * KP + k, acc[i][j]);
    }
  }
}

struct ChainBwdParams {
  const float* betas;        // [B,NB]
  const float* pose;         // [B,J,9] rotation matrices
  int B, J, NB, KP, max_depth;
  const float* J_template;   // [J,3]
  const float* J_shapedirs;  // [J,3,NB]
  const int* parents;        // [J]
  const int* depth;          // [J]
  const float* g_A;          // [B,J,12]
  const float* g_joints;     // [B,J,3] or null (gradient w.r.t. the posed chain joints)
  const float* g_pf;         // [B,KP]
  float* g_pose;             // [B,J,9]
  float* g_betas;            // [B,NB]
};

// warp per body, lane per joint (mirrors smpl_chain_kernel; the forward chain is recomputed in registers)
__global__ void __launch_bounds__(kChainWarpsPerBlock * 32) chain_backward_kernel(ChainBwdParams p) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * kChainWarpsPerBlock + (threadIdx.x >> 5);
  if (b >= p.B) return;   // warp-uniform
  const bool active = lane < p.J;
  const int j = active ? lane : 0;
  float R[9];
  {
    const float* src = p.pose + ((size_t)b * p.J + j) * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = src[i];
  }
  const float my_beta = lane < p.NB ? p.betas[(size_t)b * p.NB + lane] : 0.0f;
  float Jr[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) Jr[c] = p.J_template[j * 3 + c];
  for (int k = 0; k < p.NB; ++k) {
    const float bk = __shfl_sync(full, my_beta, k);
#pragma unroll
    for (int c = 0; c < 3; ++c) Jr[c] = fmaf(bk, p.J_shapedirs[(j * 3 + c) * p.NB + k], Jr[c]);
  }
  int par = active ? p.parents[j] : 0;
  const int dep = active ? p.depth[j] : -1;
  const bool is_root = par < 0;
  if (is_root) par = j;
  // forward chain: G = [Rg | tg]
  float G[12], Jp[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    Jp[c] = __shfl_sync(full, Jr[c], par);
    G[c * 4 + 3] = is_root ? Jr[c] : Jr[c] - Jp[c];
    G[c * 4 + 0] = R[c * 3 + 0]; G[c * 4 + 1] = R[c * 3 + 1]; G[c * 4 + 2] = R[c * 3 + 2];
  }
  for (int d = 1; d <= p.max_depth; ++d) {
    float P[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) P[i] = __shfl_sync(full, G[i], par);
    if (dep == d) {
      float N[12];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float acc = P[r * 4 + 0] * G[0 * 4 + c];
          acc = fmaf(P[r * 4 + 1], G[1 * 4 + c], acc);
          acc = fmaf(P[r * 4 + 2], G[2 * 4 + c], acc);
          if (c == 3) acc += P[r * 4 + 3];
          N[r * 4 + c] = acc;
        }
#pragma unroll
      for (int i = 0; i < 12; ++i) G[i] = N[i];
    }
  }
  // parent's world rotation (for the root: unused)
  float Rp[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Rp[r * 3 + c] = __shfl_sync(full, G[r * 4 + c], par);

  // ---- seeds from A_j = [Rg_j | tg_j - Rg_j J_j] and the posed joints tg_j ----
  float gA[12];
  {
    const float* src = p.g_A + ((size_t)b * p.J + j) * 12;
#pragma unroll
    for (int i = 0; i < 12; ++i) gA[i] = active ? src[i] : 0.f;
  }
  float gRg[9], gtg[3], gJ[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    gtg[r] = gA[r * 4 + 3] + ((active && p.g_joints) ? p.g_joints[((size_t)b * p.J + j) * 3 + r] : 0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) gRg[r * 3 + c] = gA[r * 4 + c] - gA[r * 4 + 3] * Jr[c];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
    gJ[c] = -(G[0 * 4 + c] * gA[0 * 4 + 3] + G[1 * 4 + c] * gA[1 * 4 + 3] + G[2 * 4 + c] * gA[2 * 4 + 3]);

  // ---- reverse pass: children before parents (parents[j] < j) ----
  float gR[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) gR[i] = 0.f;
  for (int jj = p.J - 1; jj >= 1; --jj) {
    const int pj = p.parents[jj];
    // lane jj finalises its own local gradients and prepares what it sends to its parent
    float cR[9], ctg[3], cJ[3];
    const float dJ[3] = {Jr[0] - Jp[0], Jr[1] - Jp[1], Jr[2] - Jp[2]};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      ctg[r] = gtg[r];
#pragma unroll
      for (int c = 0; c < 3; ++c)   // (gRg R^T)[r][c] + gtg[r] * dJ[c]
        cR[r * 3 + c] = gRg[r * 3 + 0] * R[c * 3 + 0] + gRg[r * 3 + 1] * R[c * 3 + 1] + gRg[r * 3 + 2] * R[c * 3 + 2] +
                        gtg[r] * dJ[c];
    }
    float RpT_gtg[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) RpT_gtg[c] = Rp[0 * 3 + c] * gtg[0] + Rp[1 * 3 + c] * gtg[1] + Rp[2 * 3 + c] * gtg[2];
    if (lane == jj) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)   // gR = Rp^T gRg
          gR[r * 3 + c] = Rp[0 * 3 + r] * gRg[0 * 3 + c] + Rp[1 * 3 + r] * gRg[1 * 3 + c] + Rp[2 * 3 + r] * gRg[2 * 3 + c];
#pragma unroll
      for (int c = 0; c < 3; ++c) gJ[c] += RpT_gtg[c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) cJ[c] = -RpT_gtg[c];
    // broadcast from lane jj, accumulate in lane pj
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float x = __shfl_sync(full, cR[i], jj);
      if (lane == pj) gRg[i] += x;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float x = __shfl_sync(full, ctg[i], jj);
      const float y = __shfl_sync(full, cJ[i], jj);
      if (lane == pj) { gtg[i] += x; gJ[i] += y; }
    }
  }
  if (lane == 0) {   // root: Rg_0 = R_0, tg_0 = J_0
#pragma unroll
    for (int i = 0; i < 9; ++i) gR[i] = gRg[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) gJ[c] += gtg[c];
  }
  // pose feature (R_j - I), j >= 1, and the shape rows of the pose-blend contraction
  const int nfeat = (p.J - 1) * 9;
  const float* gpf = p.g_pf + (size_t)b * p.KP;
  if (active && j >= 1) {
#pragma unroll
    for (int i = 0; i < 9; ++i) gR[i] += gpf[(j - 1) * 9 + i];
  }
  if (active) {
    float* dst = p.g_pose + ((size_t)b * p.J + j) * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) dst[i] = gR[i];
  }
  // g_beta[k] = sum_j Jd[j,:,k] . gJ_j + g_pf[nfeat + k]
  for (int k = 0; k < p.NB; ++k) {
    float s = 0.f;
    if (active) {
#pragma unroll
      for (int c = 0; c < 3; ++c) s = fmaf(p.J_shapedirs[(j * 3 + c) * p.NB + k], gJ[c], s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(full, s, o);
    if (lane == 0) p.g_betas[(size_t)b * p.NB + k] = s + gpf[nfeat + k];
  }
}

// Transposed read-out: g_src[b, col] += val * (g_out[b, r] - sum of the g_out of rows that subtract row r)
//   verts part -> g_verts [B,V,3], chain-joint part -> g_joints [B,J,3]; both accumulated with atomics (the caller
//   zero-fills or pre-loads them with the direct gradients).  One thread per (body, row).
struct ReadoutBwdParams {
  ReadoutParams rp;          // tables; rp.out = g_out (group-major, same layout as the forward output)
  float* g_verts;            // [B,V,3]
  float* g_joints;           // [B,J,3] or null when no row references a chain joint
};

__device__ __forceinline__ void readout_scatter_row(const ReadoutParams& p, float* g_verts, float* g_joints, int b, int r,
                                                    float gx, float gy, float gz) {
  for (int k = p.row_ptr[r]; k < p.row_ptr[r + 1]; ++k) {
    const float w = p.vals[k];
    const int col = p.col_idx[k];
    float* dst = col < p.V ? g_verts + ((size_t)b * p.V + col) * 3 : g_joints + ((size_t)b * p.J + (col - p.V)) * 3;
    atomicAdd(dst + 0, w * gx); atomicAdd(dst + 1, w * gy); atomicAdd(dst + 2, w * gz);
  }
}

__global__ void __launch_bounds__(256) readout_backward_kernel(ReadoutBwdParams q) {
  const ReadoutParams& p = q.rp;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.B * p.R) return;
  const int b = (int)(i / p.R), r = (int)(i % p.R);
  const float* g = readout_dst(p, b, r);
  const float gx = g[0], gy = g[1], gz = g[2];
  if (gx == 0.f && gy == 0.f && gz == 0.f) return;
  readout_scatter_row(p, q.g_verts, q.g_joints, b, r, gx, gy, gz);
  const int sr = p.sub_row ? p.sub_row[r] : -1;
  if (sr >= 0) readout_scatter_row(p, q.g_verts, q.g_joints, b, sr, -gx, -gy, -gz);
}

}  // namespace whmr
