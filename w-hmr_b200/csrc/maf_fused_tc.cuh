// MAF_Extractor.sampling / .forward as ONE kernel: bilinear sampling of the feature maps at the mesh points
// (models/maf_extractor.py:119) and the `reduce_dim` Conv1d(k=1) MLP with its skip concatenations (:75-101)
// on the 5th-generation tensor cores -- the [B,C_s,N] point features never make the HBM round trip.
//
//   x  = grid_sample(im_feat, points)                          [C0]      per point (C0 = 256)
//   y0 = leaky_relu(W0 x + b0)                                 [C1=128]
//   y1 = leaky_relu(W1 [y0 ; x] + b1)                          [C2=64]
//   y2 = relu      (W2 [y1 ; x] + b2)                          [C3=32]   -> mesh_align_feat[b, c*N + n]
//
// The three contractions against x are linear, so they are ONE pass over x: Z = x . [W0 | W1[:,C1:] | W2[:,C2:]]^T
// (N = C1+C2+C3 = 224 accumulator columns in TMEM); the layer-1 and layer-2 terms on y0 / y1 are then accumulated onto
// the columns that already hold their skip part.  x is therefore streamed: a 32-channel chunk of the sampled tile is
// written to shared memory as the A operand, multiplied, and its stage is refilled two chunks later.
//
// Arithmetic: 3xTF32 (kind::tf32; a = hi + lo with both parts tf32-representable, a*w ~= lo*hi + hi*lo + hi*hi
// accumulated in fp32 in TMEM: ~2^-21 relative per product), the same scheme as pose_blend_tc.cuh.
//
// Tile = 128 consecutive points of the flattened [B*N] point list (UMMA M = 128, TMEM lane = point).  CTA = 16 worker
// warps + 1 control warp, one CTA per SM, persistent over tiles.  A tile is 14 "ops" (8 chunks of x, 4 of y0, 2 of y1),
// each = {A-operand chunk, weight chunk {hi,lo}, 12 tcgen05.mma}.
//   * The A operand lives in TENSOR MEMORY (tcgen05.mma with [a_tmem]): 2 stages x {64 hi + 64 lo} columns (a ROUND of
//     64 channels = two ops) next to the 224 accumulator columns.  The workers write it with tcgen05.st straight from registers -- no shared-memory round
//     trip, no proxy fence -- and shared memory is left entirely to the weights.
//   * The weight chunks (56 KB per op from L2, ~1.5 us to arrive against 0.7 us of MMAs) run through a 4-stage TMA ring
//     filled three ops ahead.  (Measured with two co-resident CTAs and ONE weight stage each: 27 us per tile and SM;
//     with two stages: 24 us -- the tensor pipe waited for weights half of the time.)
//   worker thread = (TMEM lane = point r = 32*(warp%4) + lane, 16-channel slice cq = warp/4 of the 64-channel round):
//     sampling rounds: 16 channels of its point (NCHW: 64 scalar taps, lanes = neighbouring points of one plane; NHWC:
//     16 LDG.128 -- a round is one exposed gather latency, so it is made as wide as the registers allow), hi/lo split,
//     tcgen05.st; round r+1 is gathered while the tensor core multiplies round r;
//     layer-1/2 rounds: tcgen05.ld of 16 accumulator columns, + bias, leaky_relu, split, tcgen05.st as the next A operand;
//     end of tile: 8 of the 32 output channels of its point (lanes = consecutive n: coalesced);
//   control thread: weight TMA, 4 K steps x 3 tcgen05.mma (M=128, N=224 | 64 | 32, K=8), tcgen05.commit.
#pragma once
#include "pose_blend_tc.cuh"
#include "sampling.cuh"
#include "skin_tc.cuh"
#include "smpl_fused_tc.cuh"

namespace whmr {

constexpr int kMafWorkerWarps = 16;
constexpr int kMafWorkers = kMafWorkerWarps * 32;
constexpr int kMafThreads = kMafWorkers + 32;    // + 1 control warp
constexpr int kMafRows = 128;                    // points per tile (UMMA M)
constexpr int kMafTmemCols = 512;                // accumulators [0, 256) + A-operand stages at 256 + 128 s: {64 hi | 64 lo}
constexpr int kMafXCol = 256;
constexpr int kMafXStage = 128;
constexpr int kMafMaxWStages = 4;
constexpr int kMafStgPitch = 20;                 // floats per row of the NHWC gather staging tile (16 + pad: conflict-free reads)
constexpr int kMafStgBytes = 16 * 32 * kMafStgPitch * 4;   // 40 KB
constexpr int kMafMaxOut = 256;                  // C1 + C2 + C3 (accumulator columns, TMA box rows)

struct MafDims { int c0, c1, c2, c3; };

// weight-ring depth: as many {hi,lo} chunks of C1+C2+C3 rows as fit beside the biases and barriers (224 rows: 4 stages)
static inline int maf_w_stages(const MafDims& d) {
  const size_t st = (size_t)2 * (d.c1 + d.c2 + d.c3) * 128;
  const size_t room = 232448 - ((size_t)kMafMaxOut * 4 + 128 + 1024 + kMafStgBytes);
  int n = (int)(room / st);
  // Shared memory and L1 are one 256 KB array: with a 4-stage ring (227 KB) the gathers run through ~24 KB of L1 and the
  // sparse NCHW levels take 2x longer (B=256, 128x96: 0.399 ms vs 0.169 ms at 2 stages; 3 stages 0.215 ms), while the
  // dense levels do not gain from the deeper ring (0.73-0.75 ms at 2, 3 and 4 stages).  Default 2; WHMR_MAF_WSTAGES overrides.
  static const int cap = getenv("WHMR_MAF_WSTAGES") ? atoi(getenv("WHMR_MAF_WSTAGES")) : 2;
  if (n > cap) n = cap;
  return n > kMafMaxWStages ? kMafMaxWStages : (n < 2 ? 2 : n);
}
static inline size_t maf_smem_bytes(const MafDims& d, bool nhwc) {   // the staging tile only exists for NHWC maps
  const int nx = d.c1 + d.c2 + d.c3;
  return maf_w_stages(d) * ((size_t)2 * nx * 128) + (size_t)kMafMaxOut * 4 + 128 + 1024 + (nhwc ? kMafStgBytes : 0);
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// rows [0,rows): hi parts; rows [rows, 2*rows): lo parts -- one TMA box each
__global__ void maf_split_weights_kernel(const float* __restrict__ w0, const float* __restrict__ b0,
                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                         const float* __restrict__ w2, const float* __restrict__ b2, MafDims d,
                                         float* __restrict__ wx, float* __restrict__ w1y, float* __restrict__ w2y,
                                         float* __restrict__ bias) {
  const int nx = d.c1 + d.c2 + d.c3;
  const int n_x = nx * d.c0, n_1 = d.c2 * d.c1, n_2 = d.c3 * d.c2;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float x;
  float* dst;
  int rows, r, k, K;
  if (i < n_x) {
    r = i / d.c0; k = i - r * d.c0; rows = nx; K = d.c0; dst = wx;
    if (r < d.c1) x = w0[(size_t)r * d.c0 + k];
    else if (r < d.c1 + d.c2) x = w1[(size_t)(r - d.c1) * (d.c1 + d.c0) + d.c1 + k];
    else x = w2[(size_t)(r - d.c1 - d.c2) * (d.c2 + d.c0) + d.c2 + k];
  } else if ((i -= n_x) < n_1) {
    r = i / d.c1; k = i - r * d.c1; rows = d.c2; K = d.c1; dst = w1y;
    x = w1[(size_t)r * (d.c1 + d.c0) + k];
  } else if ((i -= n_1) < n_2) {
    r = i / d.c2; k = i - r * d.c2; rows = d.c3; K = d.c2; dst = w2y;
    x = w2[(size_t)r * (d.c2 + d.c0) + k];
  } else if ((i -= n_2) < nx) {
    bias[i] = i < d.c1 ? (b0 ? b0[i] : 0.f) : i < d.c1 + d.c2 ? (b1 ? b1[i - d.c1] : 0.f) : (b2 ? b2[i - d.c1 - d.c2] : 0.f);
    return;
  } else {
    return;
  }
  const float hi = tf32_rna(x);
  dst[(size_t)r * K + k] = hi;
  dst[(size_t)(rows + r) * K + k] = tf32_rna(x - hi);
}

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem], kind::tf32 (A: lane = row, one 32-bit column per K element)
__device__ __forceinline__ void umma_ta_tf32(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// 8 values of one thread -> hi | lo columns of the A operand stage at `x_addr` (lane field set), columns col .. col+7
__device__ __forceinline__ void maf_store_operand8(uint32_t x_addr, int col, const float* v) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float h = tf32_rna(v[i]);
    hi[i] = __float_as_uint(h);
    lo[i] = __float_as_uint(tf32_rna(v[i] - h));
  }
  tmem_st_32x32b_x8(x_addr + (uint32_t)col, hi);
  tmem_st_32x32b_x8(x_addr + (uint32_t)(64 + col), lo);
}

// kLayout 0: feat [B,C0,H,W]; 1: feat [B,H,W,C0].  kProject: `points` are [B,N,3] mesh points and the weak projection
// (utils/geometry.py:289-307) is evaluated here (MAF_Extractor.forward, models/maf_extractor.py:126-143).
template <int kLayout, bool kProject>
__global__ void __launch_bounds__(kMafThreads, 1)
maf_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_1,
                 const __grid_constant__ CUtensorMap map_2, const float* __restrict__ feat,
                 const float* __restrict__ points, int pts_bstride, const float* __restrict__ bias_g,
                 float* __restrict__ out, float* __restrict__ pf_out, int B, int N, int H, int W, MafDims d,
                 int n_tiles, int wstages, SampleProj pj) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nx = d.c1 + d.c2 + d.c3;
  uint8_t* wring = smem;                                       // weight chunks {hi rows, lo rows}: `wstages` stages
  const uint32_t wstage_bytes = 2u * (uint32_t)nx * 128u;
  float* bias = reinterpret_cast<float*>(wring + (size_t)wstages * wstage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias + kMafMaxOut);
  uint64_t* x_full = bars;         // [2] 512 worker arrivals: operand stage written (round r: stage r&1, phase r>>1)
  uint64_t* mma_done = bars + 2;   // [2] tcgen05.commit: every MMA up to round r has retired (barrier r&1, phase r>>1)
  uint64_t* w_full = bars + 4;     // [kMafMaxWStages] TMA: weight chunk landed
  uint64_t* w_empty = w_full + kMafMaxWStages;   // [kMafMaxWStages] tcgen05.commit: the chunk's MMAs have retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_empty + kMafMaxWStages);
  float* gstage = reinterpret_cast<float*>(bars + 16);          // NHWC gather: [16 warps][32 rows][kMafStgPitch]
  (void)gstage;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&x_full[s], kMafWorkers); mbar_init(&mma_done[s], 1); }
    for (int s = 0; s < kMafMaxWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMafWorkerWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"((uint32_t)kMafTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  pdl_wait();      // the split weights / biases may come from the launch just before this one
  pdl_trigger();
  for (int i = threadIdx.x; i < nx; i += kMafThreads) bias[i] = bias_g[i];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n0 = d.c0 >> 5, n1 = d.c1 >> 5, n2 = d.c2 >> 5;   // 32-channel weight chunks (ops) per layer; a round = 2 ops
  // "round j has retired".  A waiter is never more than one phase of a barrier behind: the next completion on barrier
  // j&1 is round j+2, which cannot be issued before this waiter has moved on.
#define WHMR_MAF_WAIT_OP(bar, j) mbar_wait(&(bar)[(j) & 1], ((uint32_t)(j) >> 1) & 1u)

  if (warp == kMafWorkerWarps) {
    // ================================ control: weight TMA + MMA issue ================================
    if (elect_one()) {
      auto op_layer = [&](int k, int& layer, int& kc) {       // k-th op of a tile
        if (k < n0) { layer = 0; kc = k; } else if (k < n0 + n1) { layer = 1; kc = k - n0; } else { layer = 2; kc = k - n0 - n1; }
      };
      auto load_w = [&](uint32_t g, int layer, int kc) {
        const int rows = layer == 0 ? nx : layer == 1 ? d.c2 : d.c3;
        const CUtensorMap* map = layer == 0 ? &map_x : layer == 1 ? &map_1 : &map_2;
        const int st = (int)(g % (uint32_t)wstages);
        uint8_t* wb = wring + (size_t)st * wstage_bytes;
        mbar_arrive_expect_tx(&w_full[st], (uint32_t)rows * 256u);
        tma_load_2d(wb, map, &w_full[st], kc * 32, 0);
        tma_load_2d(wb + (size_t)rows * 128, map, &w_full[st], kc * 32, rows);
      };
      const int ops_per_tile = n0 + n1 + n2;
      const long long my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      const long long total_ops = my_tiles * ops_per_tile;
      for (int i = 0; i < wstages - 1 && i < total_ops; ++i) {   // ring prefill (ops 0 .. stages-2)
        int l, c;
        op_layer(i % ops_per_tile, l, c);
        load_w((uint32_t)i, l, c);
      }
      uint32_t g = 0;     // op (weight chunk) counter; round = g >> 1
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int k = 0; k < ops_per_tile; ++k, ++g) {
          int layer, kc;
          op_layer(k, layer, kc);
          const int rows = layer == 0 ? nx : layer == 1 ? d.c2 : d.c3;     // UMMA N
          const uint32_t d_tmem = tmem_base + (uint32_t)(layer == 0 ? 0 : layer == 1 ? d.c1 : d.c1 + d.c2);
          const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(rows >> 3) << 17) |
                                 ((uint32_t)(kMafRows >> 4) << 24);
          const int st = (int)(g % (uint32_t)wstages);
          const uint32_t rd = g >> 1;
          const uint32_t a_hi = tmem_base + (uint32_t)(kMafXCol + (rd & 1) * kMafXStage + (g & 1) * 32), a_lo = a_hi + 64;
          const uint32_t b_hi = smem_u32(wring + (size_t)st * wstage_bytes), b_lo = b_hi + (uint32_t)rows * 128u;
          if ((g & 1) == 0) WHMR_MAF_WAIT_OP(x_full, rd);
          mbar_wait(&w_full[st], (g / (uint32_t)wstages) & 1u);
          tcgen05_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t dB_hi = umma_desc_sw128(b_hi + ks * 32), dB_lo = umma_desc_sw128(b_lo + ks * 32);
            umma_ta_tf32(d_tmem, a_lo + ks * 8, dB_hi, idesc, (layer | kc | ks) != 0);   // small terms first
            umma_ta_tf32(d_tmem, a_hi + ks * 8, dB_lo, idesc, 1u);
            umma_ta_tf32(d_tmem, a_hi + ks * 8, dB_hi, idesc, 1u);
          }
          if ((long long)g + wstages < total_ops) tcgen05_commit(&w_empty[st]);   // only where the refill below will wait for it
          if (g & 1) tcgen05_commit(&mma_done[rd & 1]);
          // weight chunk of op g + stages - 1 into the ring stage op g - 1 used, once that op has retired (the next
          // completion on that barrier is op g + stages - 1 itself: not issued yet, so the parity wait is safe)
          const long long nxt = (long long)g + wstages - 1;
          if (nxt < total_ops) {
            int nl, nk;
            op_layer((int)(nxt % ops_per_tile), nl, nk);
            if (g >= 1) mbar_wait(&w_empty[(g - 1) % (uint32_t)wstages], ((g - 1) / (uint32_t)wstages) & 1u);
            load_w((uint32_t)nxt, nl, nk);
          }
        }
      }
    }
  } else {
    // ================================ workers: sample / activate / store ==============================
    const int q = warp & 3, cq = warp >> 2;     // TMEM lane quarter (hardware rule), 16-channel slice of a round
    const int r = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const long long total = (long long)B * N;
    const size_t plane = (size_t)H * W;
    uint32_t rd = 0;     // round counter
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long p = (long long)tile * kMafRows + r;
      const bool valid = p < total;
      const int b = valid ? (int)(p / N) : 0;
      const int n = valid ? (int)(p - (long long)b * N) : 0;
      Taps tp;
      {
        float2 gp;
        if (kProject) {   // utils/geometry.py:289-307
          const float cs = pj.cam[b * 3 + 0], ctx = pj.cam[b * 3 + 1], cty = pj.cam[b * 3 + 2];
          const float ctz = 2.0f * pj.focal / (pj.img_h * cs + 1e-9f);
          const float* qp = points + (size_t)b * pts_bstride + (size_t)n * 3;
          const float px = qp[0] + ctx, py = qp[1] + cty, pz = qp[2] + ctz;
          gp.x = (pj.focal * (px / pz)) / (pj.img_w * 0.5f);
          gp.y = (pj.focal * (py / pz)) / (pj.img_h * 0.5f);
          if (pj.pts2d_out && valid && cq == 0) *reinterpret_cast<float2*>(pj.pts2d_out + ((size_t)b * N + n) * 2) = gp;
        } else {
          gp = *reinterpret_cast<const float2*>(points + (size_t)b * pts_bstride + (size_t)n * 2);
        }
        tp = make_taps(gp.x, gp.y, H, W);
        if (!valid) { tp.w00 = tp.w01 = tp.w10 = tp.w11 = 0.f; tp.o00 = tp.o01 = tp.o10 = tp.o11 = 0; }
      }
      // ---- layer 0 + both skip terms: stream the sampled tile through the tensor core, 64 channels per round ----
#pragma unroll 1
      for (int k2 = 0; k2 < (n0 >> 1); ++k2, ++rd) {
        const int cb = k2 * 64 + cq * 16;
        float v[16];
        if (kLayout == 0) {
          const float* pl = feat + ((size_t)b * d.c0 + cb) * plane;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float* c = pl + (size_t)i * plane;
            const float v00 = __ldg(c + tp.o00), v01 = __ldg(c + tp.o01), v10 = __ldg(c + tp.o10), v11 = __ldg(c + tp.o11);
            float acc = v00 * tp.w00;
            acc = fmaf(v01, tp.w01, acc);
            acc = fmaf(v10, tp.w10, acc);
            v[i] = fmaf(v11, tp.w11, acc);
          }
        } else {
          // NHWC: lanes = CHANNELS for the loads.  Step s covers 8 rows of this warp's lane quarter x 4 pieces (16 B) of
          // its 16-channel slice: a warp request touches 8 lines of 64 contiguous bytes (with lane = row it is 32 lines,
          // 16 bytes each -- the L1 tag stage, one line per cycle, was the limit: ncu l1tex 60-87 %).  The taps of the
          // other rows come by shuffle from the lanes that own them; the [32 rows x 16 channels] block goes through a
          // per-warp staging tile back to lane = row.
          const int r8 = lane >> 2, pc = lane & 3;
          float* stg = gstage + warp * (32 * kMafStgPitch);
          __syncwarp();
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            const int src = s4 * 8 + r8;
            const int ob = __shfl_sync(0xffffffffu, b, src);
            const int o00 = __shfl_sync(0xffffffffu, tp.o00, src), o01 = __shfl_sync(0xffffffffu, tp.o01, src);
            const int o10 = __shfl_sync(0xffffffffu, tp.o10, src), o11 = __shfl_sync(0xffffffffu, tp.o11, src);
            const float w00 = __shfl_sync(0xffffffffu, tp.w00, src), w01 = __shfl_sync(0xffffffffu, tp.w01, src);
            const float w10 = __shfl_sync(0xffffffffu, tp.w10, src), w11 = __shfl_sync(0xffffffffu, tp.w11, src);
            const float* fb = feat + (size_t)ob * plane * d.c0 + cb + pc * 4;
            const float4 a = __ldg(reinterpret_cast<const float4*>(fb + (size_t)o00 * d.c0));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(fb + (size_t)o01 * d.c0));
            const float4 c = __ldg(reinterpret_cast<const float4*>(fb + (size_t)o10 * d.c0));
            const float4 e = __ldg(reinterpret_cast<const float4*>(fb + (size_t)o11 * d.c0));
            float4 o;
            o.x = fmaf(e.x, w11, fmaf(c.x, w10, fmaf(bb.x, w01, a.x * w00)));
            o.y = fmaf(e.y, w11, fmaf(c.y, w10, fmaf(bb.y, w01, a.y * w00)));
            o.z = fmaf(e.z, w11, fmaf(c.z, w10, fmaf(bb.z, w01, a.z * w00)));
            o.w = fmaf(e.w, w11, fmaf(c.w, w10, fmaf(bb.w, w01, a.w * w00)));
            *reinterpret_cast<float4*>(stg + src * kMafStgPitch + pc * 4) = o;
          }
          __syncwarp();
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 o = *reinterpret_cast<const float4*>(stg + lane * kMafStgPitch + i4 * 4);
            v[i4 * 4 + 0] = o.x; v[i4 * 4 + 1] = o.y; v[i4 * 4 + 2] = o.z; v[i4 * 4 + 3] = o.w;
          }
        }
        if (pf_out && valid) {   // optional [B,C0,N] point features (the reference returns them; the loop never reads them)
          float* o = pf_out + ((size_t)b * d.c0 + cb) * N + n;
#pragma unroll
          for (int i = 0; i < 16; ++i) o[(size_t)i * N] = v[i];
        }
        if (rd >= 2) { WHMR_MAF_WAIT_OP(mma_done, rd - 2); tcgen05_fence_after(); }   // the stage's previous round has retired
        const uint32_t xs = t_lane + (uint32_t)(kMafXCol + (rd & 1) * kMafXStage);
        maf_store_operand8(xs, cq * 16, v);
        maf_store_operand8(xs, cq * 16 + 8, v + 8);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tcgen05_fence_before();
        mbar_arrive(&x_full[rd & 1]);
      }
      // ---- layers 1 and 2: activation of the accumulator columns becomes the next A operand ----
#pragma unroll 1
      for (int layer = 1; layer < 3; ++layer) {
        const int nrd = (layer == 1 ? n1 : n2) >> 1;
        const int col0 = layer == 1 ? 0 : d.c1;
#pragma unroll 1
        for (int k2 = 0; k2 < nrd; ++k2, ++rd) {
          // round 0 reads what the whole previous layer produced (round rd-1); later rounds only need their stage back
          if (rd >= 2) WHMR_MAF_WAIT_OP(mma_done, rd - 2);
          if (k2 == 0) WHMR_MAF_WAIT_OP(mma_done, rd - 1);
          tcgen05_fence_after();
          const int col = col0 + k2 * 64 + cq * 16;
          uint32_t z[16];
          tmem_ld_32x32b_x16(t_lane + (uint32_t)col, z);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float y[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = __uint_as_float(z[i]) + bias[col + i];
            y[i] = a > 0.f ? a : 0.01f * a;    // F.leaky_relu default slope
          }
          const uint32_t xs = t_lane + (uint32_t)(kMafXCol + (rd & 1) * kMafXStage);
          maf_store_operand8(xs, cq * 16, y);
          maf_store_operand8(xs, cq * 16 + 8, y + 8);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tcgen05_fence_before();
          mbar_arrive(&x_full[rd & 1]);
        }
      }
      // ---- output: relu(z + b2) -> mesh_align_feat[b, c*N + n] ----
      if (rd >= 2) WHMR_MAF_WAIT_OP(mma_done, rd - 2);
      WHMR_MAF_WAIT_OP(mma_done, rd - 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int cg = cq * 8; cg < d.c3; cg += 32) {
        const int col = d.c1 + d.c2 + cg;
        uint32_t z[8];
        tmem_ld_32x32b_x8(t_lane + (uint32_t)col, z);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (valid) {
          float* o = out + (size_t)b * d.c3 * N + (size_t)cg * N + n;
#pragma unroll
          for (int i = 0; i < 8; ++i) o[(size_t)i * N] = fmaxf(__uint_as_float(z[i]) + bias[col + i], 0.f);
        }
      }
      tcgen05_fence_before();   // ordered before the x_full arrival that lets the next tile overwrite the accumulators
    }
  }
#undef WHMR_MAF_WAIT_OP

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMafWorkerWarps) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kMafTmemCols));
  }
}

// 2-D map {K, 2*rows} over a [2*rows, K] fp32 hi|lo weight matrix: box {32 floats, rows}, 128B swizzle
static inline int maf_encode_weights(void* fn, CUtensorMap* map, void* base, int K, int rows) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)2 * rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<PFN_encodeTiled>(fn)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(WHMR_E_CUDA, "cuTensorMapEncodeTiled (maf weights) failed with CUresult %d", (int)r);
  return WHMR_OK;
}

}  // namespace whmr
