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
// written to shared memory as the A operand, multiplied, and overwritten by the next chunk.
//
// Arithmetic: 3xTF32 (kind::tf32; a = hi + lo with both parts tf32-representable, a*w ~= lo*hi + hi*lo + hi*hi
// accumulated in fp32 in TMEM: ~2^-21 relative per product), the same scheme as pose_blend_tc.cuh.
//
// Tile = 128 consecutive points of the flattened [B*N] point list (UMMA M = 128, TMEM lane = point).  CTA = 8 worker
// warps + 1 control warp, 2 CTAs per SM (92 KB shared memory, 256 TMEM columns each): inside a CTA the phases of a tile
// are sequential (sample chunk -> MMA ... -> layer-1 chunks -> layer-2 chunks -> output) with one operand stage; the
// two co-resident CTAs overlap one's gathers with the other's MMAs.  Within the sampling phase the global loads of
// chunk k+1 are issued before the wait for the MMAs of chunk k.
//   worker thread (row r = 32*(warp%4) + lane, half h = warp/4): samples 16 channels of its point per chunk
//     (NCHW: 64 scalar loads, lanes = neighbouring points of one plane; NHWC: 16 LDG.128), splits hi/lo, writes two
//     16-byte pieces x 4 into the 128B-swizzled K-major operand tile (conflict-free: 8 lanes hit 8 different pieces);
//     in the layer-1/2 phases it pulls 16 accumulator columns of its TMEM lane, adds the bias, applies leaky_relu and
//     writes them back as the next A operand; at the end it writes 16 of the 32 output channels of its point
//     (lanes = consecutive n: coalesced);
//   control thread: TMA of the weight chunk {hi, lo} (rows = output channels, 128B swizzle), 4 K steps x 3
//     tcgen05.mma (M=128, N=224 | 64 | 32, K=8), tcgen05.commit.
#pragma once
#include "pose_blend_tc.cuh"
#include "sampling.cuh"
#include "skin_tc.cuh"

namespace whmr {

constexpr int kMafThreads = 288;                 // 8 worker warps + 1 control warp
constexpr int kMafRows = 128;                    // points per tile (UMMA M)
constexpr int kMafXBytes = kMafRows * 128;       // one 32-channel operand part (hi or lo)
constexpr int kMafTmemCols = 256;
constexpr int kMafMaxOut = 256;                  // C1 + C2 + C3 (accumulator columns, TMA box rows)

struct MafDims { int c0, c1, c2, c3; };

static inline size_t maf_smem_bytes(const MafDims& d) {
  const int nx = d.c1 + d.c2 + d.c3;
  return 2 * kMafXBytes + (size_t)2 * nx * 128 + (size_t)kMafMaxOut * 4 + 64 + 1024;
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

// 16 values of one thread -> hi/lo parts of the A operand tile (row r, 16-byte pieces 4h..4h+3 of the 128-byte row)
__device__ __forceinline__ void maf_store_operand(uint8_t* xh, uint8_t* xl, int r, int h, const float* v) {
  const uint32_t row_off = (uint32_t)r * 128u;
  const int rx = r & 7;
#pragma unroll
  for (int i4 = 0; i4 < 4; ++i4) {
    const uint32_t off = row_off + (uint32_t)(((h * 4 + i4) ^ rx) << 4);
    float4 hi, lo;
    hi.x = tf32_rna(v[i4 * 4 + 0]); lo.x = tf32_rna(v[i4 * 4 + 0] - hi.x);
    hi.y = tf32_rna(v[i4 * 4 + 1]); lo.y = tf32_rna(v[i4 * 4 + 1] - hi.y);
    hi.z = tf32_rna(v[i4 * 4 + 2]); lo.z = tf32_rna(v[i4 * 4 + 2] - hi.z);
    hi.w = tf32_rna(v[i4 * 4 + 3]); lo.w = tf32_rna(v[i4 * 4 + 3] - hi.w);
    *reinterpret_cast<float4*>(xh + off) = hi;
    *reinterpret_cast<float4*>(xl + off) = lo;
  }
}

// kLayout 0: feat [B,C0,H,W]; 1: feat [B,H,W,C0].  kProject: `points` are [B,N,3] mesh points and the weak projection
// (utils/geometry.py:289-307) is evaluated here (MAF_Extractor.forward, models/maf_extractor.py:126-143).
template <int kLayout, bool kProject>
__global__ void __launch_bounds__(kMafThreads, 2)
maf_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_1,
                 const __grid_constant__ CUtensorMap map_2, const float* __restrict__ feat,
                 const float* __restrict__ points, int pts_bstride, const float* __restrict__ bias_g,
                 float* __restrict__ out, float* __restrict__ pf_out, int B, int N, int H, int W, MafDims d,
                 int n_tiles, SampleProj pj) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nx = d.c1 + d.c2 + d.c3;
  uint8_t* xh = smem;
  uint8_t* xl = smem + kMafXBytes;
  uint8_t* wbuf = smem + 2 * kMafXBytes;                       // {hi rows, lo rows} of the current weight chunk
  float* bias = reinterpret_cast<float*>(wbuf + (size_t)2 * nx * 128);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias + kMafMaxOut);
  uint64_t* x_full = bars;       // 256 worker arrivals: operand chunk written
  uint64_t* w_full = bars + 1;   // TMA: weight chunk landed
  uint64_t* mma_done = bars + 2; // tcgen05.commit: every MMA issued so far has retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(x_full, 256); mbar_init(w_full, 1); mbar_init(mma_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
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

  const int n0 = d.c0 >> 5, n1 = d.c1 >> 5, n2 = d.c2 >> 5;   // 32-channel chunks per layer

  if (warp == 8) {
    // ================================ control: weight TMA + MMA issue ================================
    if (elect_one()) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int layer = 0; layer < 3; ++layer) {
          const int nch = layer == 0 ? n0 : layer == 1 ? n1 : n2;
          const int rows = layer == 0 ? nx : layer == 1 ? d.c2 : d.c3;     // UMMA N
          const CUtensorMap* map = layer == 0 ? &map_x : layer == 1 ? &map_1 : &map_2;
          const uint32_t d_tmem = tmem_base + (uint32_t)(layer == 0 ? 0 : layer == 1 ? d.c1 : d.c1 + d.c2);
          const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(rows >> 3) << 17) |
                                 ((uint32_t)(kMafRows >> 4) << 24);
          const uint32_t a_hi = smem_u32(xh), a_lo = smem_u32(xl);
          const uint32_t b_hi = smem_u32(wbuf), b_lo = b_hi + (uint32_t)rows * 128u;
#pragma unroll 1
          for (int kc = 0; kc < nch; ++kc) {
            if (g > 0) mbar_wait(mma_done, (g - 1) & 1);     // weight buffer free
            mbar_arrive_expect_tx(w_full, (uint32_t)rows * 256u);
            tma_load_2d(wbuf, map, w_full, kc * 32, 0);
            tma_load_2d(wbuf + (size_t)rows * 128, map, w_full, kc * 32, rows);
            mbar_wait(x_full, g & 1);
            mbar_wait(w_full, g & 1);
            tcgen05_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t dA_hi = umma_desc_sw128(a_hi + ks * 32), dA_lo = umma_desc_sw128(a_lo + ks * 32);
              const uint64_t dB_hi = umma_desc_sw128(b_hi + ks * 32), dB_lo = umma_desc_sw128(b_lo + ks * 32);
              umma<1>(d_tmem, dA_lo, dB_hi, idesc, (layer | kc | ks) != 0);   // small terms first
              umma<1>(d_tmem, dA_hi, dB_lo, idesc, 1u);
              umma<1>(d_tmem, dA_hi, dB_hi, idesc, 1u);
            }
            tcgen05_commit(mma_done);
            ++g;
          }
        }
      }
    }
  } else {
    // ================================ workers: sample / activate / store ==============================
    const int q = warp & 3, h = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const long long total = (long long)B * N;
    const size_t plane = (size_t)H * W;
    uint32_t g = 0;
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
          if (pj.pts2d_out && valid && h == 0) *reinterpret_cast<float2*>(pj.pts2d_out + ((size_t)b * N + n) * 2) = gp;
        } else {
          gp = *reinterpret_cast<const float2*>(points + (size_t)b * pts_bstride + (size_t)n * 2);
        }
        tp = make_taps(gp.x, gp.y, H, W);
        if (!valid) { tp.w00 = tp.w01 = tp.w10 = tp.w11 = 0.f; tp.o00 = tp.o01 = tp.o10 = tp.o11 = 0; }
      }
      // ---- layer 0 + both skip terms: stream the sampled tile through the tensor core ----
#pragma unroll 1
      for (int kc = 0; kc < n0; ++kc) {
        const int cb = kc * 32 + h * 16;
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
          const float* fb = feat + (size_t)b * plane * d.c0 + cb;
          const float4* t00 = reinterpret_cast<const float4*>(fb + (size_t)tp.o00 * d.c0);
          const float4* t01 = reinterpret_cast<const float4*>(fb + (size_t)tp.o01 * d.c0);
          const float4* t10 = reinterpret_cast<const float4*>(fb + (size_t)tp.o10 * d.c0);
          const float4* t11 = reinterpret_cast<const float4*>(fb + (size_t)tp.o11 * d.c0);
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 a = __ldg(t00 + i4), bb = __ldg(t01 + i4), c = __ldg(t10 + i4), e = __ldg(t11 + i4);
            v[i4 * 4 + 0] = fmaf(e.x, tp.w11, fmaf(c.x, tp.w10, fmaf(bb.x, tp.w01, a.x * tp.w00)));
            v[i4 * 4 + 1] = fmaf(e.y, tp.w11, fmaf(c.y, tp.w10, fmaf(bb.y, tp.w01, a.y * tp.w00)));
            v[i4 * 4 + 2] = fmaf(e.z, tp.w11, fmaf(c.z, tp.w10, fmaf(bb.z, tp.w01, a.z * tp.w00)));
            v[i4 * 4 + 3] = fmaf(e.w, tp.w11, fmaf(c.w, tp.w10, fmaf(bb.w, tp.w01, a.w * tp.w00)));
          }
        }
        if (pf_out && valid) {   // optional [B,C0,N] point features (the reference returns them; the loop never reads them)
          float* o = pf_out + ((size_t)b * d.c0 + cb) * N + n;
#pragma unroll
          for (int i = 0; i < 16; ++i) o[(size_t)i * N] = v[i];
        }
        if (g > 0) mbar_wait(mma_done, (g - 1) & 1);   // operand buffer free
        maf_store_operand(xh, xl, r, h, v);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(x_full);
        ++g;
      }
      // ---- layers 1 and 2: activation of the accumulator columns becomes the next A operand ----
#pragma unroll 1
      for (int layer = 1; layer < 3; ++layer) {
        const int nch = layer == 1 ? n1 : n2;
        const int col0 = layer == 1 ? 0 : d.c1;
#pragma unroll 1
        for (int kc = 0; kc < nch; ++kc) {
          mbar_wait(mma_done, (g - 1) & 1);
          tcgen05_fence_after();
          const int col = col0 + kc * 32 + h * 16;
          uint32_t z[16];
          tmem_ld_32x32b_x16(t_lane + (uint32_t)col, z);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float y = __uint_as_float(z[i]) + bias[col + i];
            v[i] = y > 0.f ? y : 0.01f * y;    // F.leaky_relu default slope
          }
          maf_store_operand(xh, xl, r, h, v);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          tcgen05_fence_before();
          mbar_arrive(x_full);
          ++g;
        }
      }
      // ---- output: relu(z + b2) -> mesh_align_feat[b, c*N + n] ----
      mbar_wait(mma_done, (g - 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int cg = h * 16; cg < d.c3; cg += 32) {
        const int col = d.c1 + d.c2 + cg;
        uint32_t z[16];
        tmem_ld_32x32b_x16(t_lane + (uint32_t)col, z);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (valid) {
          float* o = out + (size_t)b * d.c3 * N + (size_t)cg * N + n;
#pragma unroll
          for (int i = 0; i < 16; ++i) o[(size_t)i * N] = fmaxf(__uint_as_float(z[i]) + bias[col + i], 0.f);
        }
      }
      tcgen05_fence_before();   // ordered before the x_full arrival that lets the next tile overwrite the accumulators
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) {
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
