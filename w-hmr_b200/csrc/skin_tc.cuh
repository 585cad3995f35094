// Linear blend skinning with the blend T = W.A on the tensor cores (tcgen05 + TMEM + TMA).
//
//   T_v[b] = sum_j w_vj A_j[b]   (3x4 per vertex and body)        -> UMMA, 3xTF32 split
//   v'     = T_v[b] . [v_template + offsets[b] ; 1] (+ transl)   (K7)        -> epilogue FMAs
//
// Why: the CUDA-core formulation needs 4 x 48 B of A per (vertex, body) out of shared memory and is
// bound by shared-memory wavefronts (profiles/r01_ncu_full_summary_v1.txt: l1tex 91 %, DRAM 13 %).  As a
// GEMM the same blend is  D[128 vertices, 12*GB] = Wtile[128, J<=32] . At[12*GB, J]^T  : the weight
// tile stays in shared memory for a whole vertex tile, each A matrix is read once per tile of 128
// vertices, and the result lands in TMEM where every epilogue thread (= vertex) reads its own 12
// numbers per body.  Dense in J, so any weight sparsity (incl. fully dense weights) costs the same.
// fp32 accuracy: w and A are split x = hi + lo into tf32-exact halves, 3 MMAs (hi.hi + hi.lo + lo.hi).
//
// CTA = 576 threads: warp 0 TMA producer, warp 1 TMEM alloc + MMA issue, warps 2-17 epilogue
// (4 warps per TMEM lane quarter, 4 bodies each).  Persistent over (vertex tile x 16-body group) items,
// vertex-major, so a CTA reloads the weight tile only when its vertex tile changes.
// HBM/L2 traffic per body: 4*NP (offsets, L2-resident) + 4*3V (out) + 54 * 3 KB of At from L2.
#pragma once
#include <cuda/std/type_traits>

#include "pose_blend_tc.cuh"
#include "readout.cuh"

namespace whmr {

constexpr int kSkinGB = 16;                 // bodies per MMA tile
constexpr int kSkinN = kSkinGB * 12;        // 192 accumulator columns per tile
constexpr int kSkinStages = 2;               // At ring (released by the MMA's commit)
constexpr int kSkinOffStages = 3;            // pose-offset ring (released by the epilogue warps)
constexpr int kSkinOffBytes = kSkinGB * 3 * kTcM * 4;   // 24 KB: [16 bodies][3 planes][128 vertices] fp32
constexpr int kSkinEpiWarps = 16;            // 4 per TMEM lane quarter, 4 bodies each
constexpr int kSkinThreads = (2 + kSkinEpiWarps) * 32;
constexpr int kSkinWPart = kTcM * 128;      // 16 KB: 128 vertices x 32 joints (tf32)
constexpr int kSkinAtPart = kSkinN * 128;   // 24 KB: 192 rows x 32 joints
constexpr int kSkinStageBytes = 2 * kSkinAtPart;
constexpr int kSkinSmem = 2 * kSkinWPart + kSkinStages * kSkinStageBytes + kSkinOffStages * kSkinOffBytes +
                          kSkinEpiWarps * 192 * 4 + 256 + 1024;
constexpr int kSkinTmemStage = 256;         // column stride between the two accumulator stages

struct SkinTcParams {
  const float* offsets;       // [nb, NP] planar padded: pose offsets + shape blend
  const float* v_template_p;  // [3, VP]
  const float* transl;        // [nb,3] or null
  float* verts;               // [nb, V, 3]
  // read-outs fused into the epilogue: per 32-vertex group a list of EmitEntry (readout.cuh); emit.grp_ptr == null: off
  EmitTable emit;
  float* ro_out;              // group-major read-out buffer of the WHOLE batch
  int ro_B, ro_b0;            // total batch of the read-out buffer, first body of this chunk
  int nb, V, VP, NP, n_groups, n_items;
  int ksteps;                 // ceil(J/8) tf32 K steps (3 for SMPL's 24 joints)
};

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

__global__ void __launch_bounds__(kSkinThreads, 1)
skin_tc_kernel(const __grid_constant__ CUtensorMap tmapW, const __grid_constant__ CUtensorMap tmapAt,
               const __grid_constant__ CUtensorMap tmapOff, SkinTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // align by OFFSET (not by integer round-trip) so the compiler keeps the shared address space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* w_smem = smem;                                   // [2][16 KB]
  uint8_t* at_smem = smem + 2 * kSkinWPart;                 // [stages][2][24 KB]
  float* off_smem = reinterpret_cast<float*>(at_smem + kSkinStages * kSkinStageBytes);    // [off stages][16][3][128]
  float* stage_out = off_smem + kSkinOffStages * (kSkinOffBytes / 4);                     // [8 warps][2][96]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + kSkinEpiWarps * 192);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + 1;
  uint64_t* at_full = bars + 2;                  // [stages]
  uint64_t* at_empty = at_full + kSkinStages;    // [stages]
  uint64_t* off_full = at_empty + kSkinStages;   // [off stages]
  uint64_t* off_empty = off_full + kSkinOffStages;
  uint64_t* tmem_full = off_empty + kSkinOffStages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_begin = (int)(((long long)blockIdx.x * p.n_items) / gridDim.x);
  const int t_end = (int)(((long long)(blockIdx.x + 1) * p.n_items) / gridDim.x);

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(w_empty, 1);
    for (int s = 0; s < kSkinStages; ++s) { mbar_init(&at_full[s], 1); mbar_init(&at_empty[s], 1); }
    for (int s = 0; s < kSkinOffStages; ++s) { mbar_init(&off_full[s], 1); mbar_init(&off_empty[s], kSkinEpiWarps); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], kSkinEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0, w_par = 1;
      int ostage = 0; uint32_t ophase = 0;
      int cur_vt = -1;
      for (int t = t_begin; t < t_end; ++t) {
        const int vt = t / p.n_groups, g = t % p.n_groups;
        if (vt != cur_vt) {
          mbar_wait(w_empty, w_par);   // MMAs on the previous weight tile have retired
          w_par ^= 1;
          mbar_arrive_expect_tx(w_full, 2 * kSkinWPart);
          tma_load_3d(w_smem, &tmapW, w_full, 0, vt * kTcM, 0);
          tma_load_3d(w_smem + kSkinWPart, &tmapW, w_full, 0, vt * kTcM, 1);
          cur_vt = vt;
        }
        mbar_wait(&at_empty[stage], phase ^ 1);
        uint8_t* st = at_smem + stage * kSkinStageBytes;
        mbar_arrive_expect_tx(&at_full[stage], kSkinStageBytes);
        tma_load_3d(st, &tmapAt, &at_full[stage], 0, g * kSkinN, 0);
        tma_load_3d(st + kSkinAtPart, &tmapAt, &at_full[stage], 0, g * kSkinN, 1);
        if (++stage == kSkinStages) { stage = 0; phase ^= 1; }
        // pose offsets of the item: [16 bodies][3 planes][128 vertices]
        mbar_wait(&off_empty[ostage], ophase ^ 1);
        mbar_arrive_expect_tx(&off_full[ostage], kSkinOffBytes);
        tma_load_3d(off_smem + ostage * (kSkinOffBytes / 4), &tmapOff, &off_full[ostage], vt * kTcM, 0, g * kSkinGB);
        if (++ostage == kSkinOffStages) { ostage = 0; ophase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kSkinN >> 3) << 17) |
                                 ((uint32_t)(kTcM >> 4) << 24);   // tf32 x tf32 -> f32, M=128, N=192
      int stage = 0; uint32_t phase = 0, w_phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int cur_vt = -1;
      const uint32_t w_hi = smem_u32(w_smem), w_lo = w_hi + kSkinWPart;
      for (int t = t_begin; t < t_end; ++t) {
        const int vt = t / p.n_groups;
        if (vt != cur_vt) { mbar_wait(w_full, w_phase); w_phase ^= 1; cur_vt = vt; }
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        mbar_wait(&at_full[stage], phase);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kSkinTmemStage);
        const uint32_t a_hi = smem_u32(at_smem + stage * kSkinStageBytes), a_lo = a_hi + kSkinAtPart;
        for (int ks = 0; ks < p.ksteps; ++ks) {   // 24 joints = 3 K steps of 8 tf32; the padding step is skipped
          const uint64_t dW_hi = umma_desc_sw128(w_hi + ks * 32), dW_lo = umma_desc_sw128(w_lo + ks * 32);
          const uint64_t dA_hi = umma_desc_sw128(a_hi + ks * 32), dA_lo = umma_desc_sw128(a_lo + ks * 32);
          umma<1>(d_tmem, dW_lo, dA_hi, idesc, ks != 0);
          umma<1>(d_tmem, dW_hi, dA_lo, idesc, 1u);
          umma<1>(d_tmem, dW_hi, dA_hi, idesc, 1u);
        }
        tcgen05_commit(&at_empty[stage]);
        tcgen05_commit(&tmem_full[acc]);
        const int next_vt = (t + 1 < t_end) ? (t + 1) / p.n_groups : -1;
        if (next_vt != vt) tcgen05_commit(w_empty);
        if (++stage == kSkinStages) { stage = 0; phase ^= 1; }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int q = warp & 3;            // TMEM lane quarter
    const int hb = (warp - 2) >> 2;    // which 4 bodies of the 16-body group
    float* stg = stage_out + (warp - 2) * 192;   // two 96-float transpose buffers, alternated per body
    int acc = 0; uint32_t acc_phase = 0;
    int cur_vt = -1;
    float tx = 0.f, ty = 0.f, tz = 0.f;
    int v = 0;
    // read-out entries of this warp's 32 vertices: lane l owns entries e0+l and e0+32+l (registers); further
    // entries (rare) are walked from memory.  Per entry: source lane, weight, destination base/stride (elements).
    int e0 = 0, n_e = 0;
    int lvA = 0, lvB = 0, strA = 0, strB = 0;
    long long baseA = 0, baseB = 0;
    float wA = 0.f, wB = 0.f;
    float* bufA = nullptr; float* bufB = nullptr;
    auto load_entry = [&](int e, int& lv, float& w, long long& base, int& stride, float*& buf) {
      const EmitEntry en = p.emit.entries[e];
      lv = (en.lv_kind & 0xff) * 3;
      w = en.w;
      if (en.lv_kind >> 8) {     // regressor term -> partial[b_local][d]
        base = 3LL * en.d; stride = 3 * p.emit.n_partial; buf = p.emit.partial;
      } else {                   // one-hot row -> final output slot
        base = 3LL * ((long long)p.ro_B * en.a + (long long)p.ro_b0 * en.c + en.d); stride = 3 * en.c; buf = p.ro_out;
      }
    };
    const int V3 = p.V * 3;
    const bool has_transl = p.transl != nullptr;
    int parity = 0;
    int ostage = 0; uint32_t ophase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int vt = t / p.n_groups, g = t % p.n_groups;
      if (vt != cur_vt) {
        cur_vt = vt;
        v = vt * kTcM + q * 32 + lane;                        // < VP
        tx = p.v_template_p[v]; ty = p.v_template_p[p.VP + v]; tz = p.v_template_p[2 * p.VP + v];
        n_e = 0;
        if (p.emit.grp_ptr) {
          const int g32 = vt * (kTcM / 32) + q;
          e0 = p.emit.grp_ptr[g32];
          n_e = p.emit.grp_ptr[g32 + 1] - e0;
          if (lane < n_e) load_entry(e0 + lane, lvA, wA, baseA, strA, bufA);
          if (lane + 32 < n_e) load_entry(e0 + 32 + lane, lvB, wB, baseB, strB, bufB);
        }
      }
      const int body_base = g * kSkinGB + hb * 4;
      const int n_valid = min(4, p.nb - body_base);             // bodies this warp really has (may be <= 0)
      const int out_col = (vt * kTcM + q * 32) * 3 + lane;      // float index inside a body row
      // pose offsets of this item were put in shared memory by the TMA producer
      const float* offs = off_smem + ostage * (kSkinOffBytes / 4) + (hb * 4) * (3 * kTcM) + q * 32 + lane;
      mbar_wait(&off_full[ostage], ophase);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kSkinTmemStage + hb * 48);
      float* outp = p.verts + (size_t)max(body_base, 0) * V3 + out_col;
      // guarded (ragged last group / last vertex tile / transl) and unguarded instantiations of the same body
      auto run = [&](auto guard_tag) {
        constexpr bool G = decltype(guard_tag)::value;
        {
          constexpr int half = 0;
          uint32_t T[48];
          tmem_ld_32x32b_x32(taddr, T);
          tmem_ld_32x32b_x16(taddr + 32, T + 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int bi = half * 4 + i;
            if (G && bi >= n_valid) continue;                   // warp-uniform
            const float px = offs[bi * 3 * kTcM] + tx, py = offs[bi * 3 * kTcM + kTcM] + ty,
                        pz = offs[bi * 3 * kTcM + 2 * kTcM] + tz;
#define WHMR_T(k) __uint_as_float(T[i * 12 + (k)])
            float rx = fmaf(WHMR_T(0), px, fmaf(WHMR_T(1), py, fmaf(WHMR_T(2), pz, WHMR_T(3))));
            float ry = fmaf(WHMR_T(4), px, fmaf(WHMR_T(5), py, fmaf(WHMR_T(6), pz, WHMR_T(7))));
            float rz = fmaf(WHMR_T(8), px, fmaf(WHMR_T(9), py, fmaf(WHMR_T(10), pz, WHMR_T(11))));
#undef WHMR_T
            if (G && has_transl) {
              const float* tr = p.transl + (size_t)(body_base + bi) * 3;
              rx += tr[0]; ry += tr[1]; rz += tr[2];
            }
            // transpose through smem (double-buffered: one __syncwarp per body): lane l holds xyz of
            // vertex l -> 3 coalesced 128-byte rows per warp and body
            float* sb = stg + parity * 96;
            parity ^= 1;
            sb[lane * 3 + 0] = rx; sb[lane * 3 + 1] = ry; sb[lane * 3 + 2] = rz;
            __syncwarp();
            float* ob = outp + (size_t)bi * V3;
            if (G) {
#pragma unroll
              for (int r = 0; r < 3; ++r)
                if (out_col + r * 32 < V3) ob[r * 32] = sb[r * 32 + lane];
            } else {
              ob[0] = sb[lane]; ob[32] = sb[32 + lane]; ob[64] = sb[64 + lane];
            }
            // fused read-outs: every table entry that references one of this warp's 32 vertices takes its
            // value from the staged tile (one-hot rows: straight to the output; regressor terms: w*v to partial)
            if (n_e > 0) {
              const int bl = body_base + bi;
              if (lane < n_e) {
                float* o = bufA + baseA + (long long)bl * strA;
                o[0] = wA * sb[lvA]; o[1] = wA * sb[lvA + 1]; o[2] = wA * sb[lvA + 2];
              }
              if (lane + 32 < n_e) {
                float* o = bufB + baseB + (long long)bl * strB;
                o[0] = wB * sb[lvB]; o[1] = wB * sb[lvB + 1]; o[2] = wB * sb[lvB + 2];
              }
              for (int e = 64 + lane; e < n_e; e += 32) {
                int lv, st; long long ba; float w; float* bf;
                load_entry(e0 + e, lv, w, ba, st, bf);
                float* o = bf + ba + (long long)bl * st;
                o[0] = w * sb[lv]; o[1] = w * sb[lv + 1]; o[2] = w * sb[lv + 2];
              }
            }
          }
        }
      };
      if (n_valid > 0) {
        const bool fast = n_valid == 4 && !has_transl && (vt + 1) * kTcM <= p.V;
        if (fast) run(cuda::std::false_type{}); else run(cuda::std::true_type{});
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&tmem_empty[acc]); mbar_arrive(&off_empty[ostage]); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (++ostage == kSkinOffStages) { ostage = 0; ophase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

}  // namespace whmr
