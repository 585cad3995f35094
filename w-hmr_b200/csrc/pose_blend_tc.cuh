// Pose-blend contraction on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   offsets[b, n] = sum_k pf[b, k] * P[k, n]        (smplx: torch.matmul(pose_feature, posedirs))
//
// fp32 accuracy from low-precision MMAs by operand splitting: x = hi + lo with hi, lo exactly
// representable in the MMA input type, and   x*y ~= hi_x*hi_y + hi_x*lo_y + lo_x*hi_y   accumulated in
// fp32 in TMEM.  Two arithmetic modes share this kernel:
//   WHMR_GEMM_TC_BF16X3 : kind::f16 with bf16 inputs, relative error ~2^-16 per product, half the
//                         tensor-pipe cycles of the tf32 variant (default; the pose offsets are <= ~5 cm,
//                         so 2^-16 relative is < 1e-6 m -- measured in tests/test_parity_gpu.py);
//   WHMR_GEMM_TC_3XTF32 : kind::tf32, relative error ~2^-21 per product (the north star's "3xTF32").
//
// GEMM orientation: M = vertex coordinates (128 rows of the planar-padded posedirs, the constant
// operand "A"), N = bodies (NB = 128 or 256 rows of the split pose feature, operand "B"), so the
// accumulator tile in TMEM has one coordinate per lane and one body per column: an epilogue warp
// reads a lane-row with tcgen05.ld and every store instruction writes 32 consecutive coordinates of
// one body (128 contiguous bytes) straight from registers -- no shared-memory staging.  This is also
// the orientation a fused skinning epilogue needs (thread = vertex).
//
// Structure (persistent, warp-specialised, 192 threads, 1 CTA/SM):
//   warp 0  : TMA producer -- per K chunk (128 B of K) one stage = {A_hi, A_lo, B_hi, B_lo} tiles,
//             128B-swizzled, landing on an mbarrier (cp.async.bulk.tensor.3d);
//   warp 1  : TMEM allocator + MMA issuer -- per 32-byte K step three tcgen05.mma into the same TMEM
//             accumulator; tcgen05.commit releases the smem stage / publishes the accumulator;
//   warps 2-5: epilogue -- tcgen05.ld 32x32b.x32, coalesced global stores, hands the accumulator back.
//   Two TMEM accumulator stages (2*NB columns) overlap the epilogue of tile i with the MMAs of i+1.
// Work items (coordinate tile x body tile) are split contiguously over the CTAs, coordinate-major,
// so consecutive items of a CTA re-read the same posedirs tile from L2.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <vector>

#include "common.cuh"

namespace whmr {

struct DeviceArena;

constexpr int kTcBodyTile = 256;   // workspace rows are padded to this
constexpr int kTcM = 128;          // coordinates per tile (UMMA M)
constexpr int kTcThreads = 192;

struct TcPlan {
  int ready = 0;
  void* A_bf16 = nullptr;   // [NP,2,KP] bf16 hi|lo
  void* A_tf32 = nullptr;   // [NP,2,KP] fp32 holding tf32-representable hi|lo
  CUtensorMap tmapA_bf16, tmapA_tf32;
  CUtensorMap tmapA_both[2];   // [kind] hi + lo of a 128-row tile in ONE box (tc_encode_both_parts), valid iff both_ok
  int both_ok = 0;
  void* encode_fn = nullptr;   // cuTensorMapEncodeTiled
  int num_sms = 148;
  // tensor-core skinning: dense tf32 hi|lo weights [2, VP, 32 joints]
  void* W_tf32 = nullptr;
  CUtensorMap tmapW;
  // fused kernel: fp16 hi|lo weights [VP, 64]
  void* W_f16 = nullptr;
  CUtensorMap tmapW16;
};

// ---------------------------------------------------------------------------------------------
// device helpers (raw PTX; names follow the PTX ISA)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ int* g_trap_buf = nullptr;   // whmr_debug_set_trap_buffer: pinned host memory or null (single translation unit)
__device__ __noinline__ void mbar_timeout(uint64_t* bar, uint32_t parity) {
  int* t = g_trap_buf;
  if (t && atomicCAS(t, 0, 1) == 0) {
    t[1] = (int)blockIdx.x; t[2] = (int)threadIdx.x; t[3] = (int)(smem_u32(bar)); t[4] = (int)parity;
    __threadfence_system();
  }
  printf("whmr: mbarrier wait timed out (block %d thread %d barrier@%u parity %u)\n", (int)blockIdx.x, (int)threadIdx.x,
         smem_u32(bar), parity);
  __trap();
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (clock64() - t0 < 4000000000LL)   // ~2 s at 1.9 GHz
    if (mbar_try_wait(bar, parity)) return;
  mbar_timeout(bar, parity);
}
// Same with a sleep between polls: for single-thread roles that wait long (an idle poller still takes issue slots
// and mbarrier-unit bandwidth from the epilogue warps of its scheduler).
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, unsigned ns) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (clock64() - t0 < 4000000000LL) {
    if (ns) __nanosleep(ns);
    if (mbar_try_wait(bar, parity)) return;
  }
  mbar_timeout(bar, parity);
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);   // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                       // leading byte offset (ignored)  [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;               // stride byte offset  [32,46)
  d |= static_cast<uint64_t>(1) << 46;                       // descriptor version 1 (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B
  return d;
}
// kColl: collector usage of the A operand (PTX tcgen05.mma .collector::a::{fill,use,lastuse}).  Two consecutive MMAs on
// the SAME A tile (the hi part of a split operand meets the lo and the hi part of the other one): the first keeps the
// tile in the collector buffer (fill), the second takes it from there (lastuse) instead of reading 4 KB of shared
// memory again.  0: default (discard).
template <int kKind, int kColl = 0>   // kind 0: kind::f16 (bf16/fp16 in), 1: kind::tf32
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
#define WHMR_UMMA_ASM(KIND, COLL)                                                            \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                           \
               "tcgen05.mma.cta_group::1.kind::" KIND COLL " [%0], %1, %2, %3, p;\n\t}"      \
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory")
  if (kKind == 0) {
    if (kColl == 1) WHMR_UMMA_ASM("f16", ".collector::a::fill");
    else if (kColl == 2) WHMR_UMMA_ASM("f16", ".collector::a::use");
    else if (kColl == 3) WHMR_UMMA_ASM("f16", ".collector::a::lastuse");
    else WHMR_UMMA_ASM("f16", "");
  } else {
    if (kColl == 1) WHMR_UMMA_ASM("tf32", ".collector::a::fill");
    else if (kColl == 2) WHMR_UMMA_ASM("tf32", ".collector::a::use");
    else if (kColl == 3) WHMR_UMMA_ASM("tf32", ".collector::a::lastuse");
    else WHMR_UMMA_ASM("tf32", "");
  }
#undef WHMR_UMMA_ASM
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// One lane of the (converged) warp, chosen by elect.sync: ptxas then knows the guarded region runs on a single
// thread and feeds UTCHMMA / UTMALDG their uniform-register operands directly.  With `if (lane == 0)` it wraps
// every such instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~13 extra instructions per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int NB>
struct TcSmem {
  static constexpr int kABytes = kTcM * 128;          // one 128-row x 128-byte operand tile
  static constexpr int kBBytes = NB * 128;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kStages = NB == 256 ? 2 : 3;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kOutTile = 32 * kTcM * 4;      // epilogue staging: [32 bodies][128 coords] fp32, x2 buffers
  static constexpr int kTotal = kStages * kStageBytes + 2 * kOutTile + kBarrierBytes + 1024;   // + alignment slack
};

// kKind 0: bf16 operands (2 B), 1: tf32 operands (4 B).  K is walked in 128-byte chunks (one TMA box
// per operand tile) of four 32-byte UMMA K steps; `ksteps` = total valid K steps (KP*elem/32).
template <int kKind, int NB>
__global__ void __launch_bounds__(kTcThreads, 1)
pose_blend_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                     const __grid_constant__ CUtensorMap tmapOut, int nb, int NP, int n_body_tiles, int n_items, int ksteps,
                     long long* __restrict__ dbg, int dbg_mode) {
  using S = TcSmem<NB>;
  constexpr int kElem = kKind == 0 ? 2 : 4;
  constexpr int kChunkElems = 128 / kElem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  float* out_smem = reinterpret_cast<float*>(smem + S::kStages * S::kStageBytes);   // [2][32][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes + 2 * S::kOutTile);
  uint64_t* full = bars;                       // [kStages]
  uint64_t* empty = bars + S::kStages;         // [kStages]
  uint64_t* tmem_full = bars + 2 * S::kStages; // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kch = (ksteps + 3) >> 2;
  const int t_begin = (int)(((long long)blockIdx.x * n_items) / gridDim.x);
  const int t_end = (int)(((long long)(blockIdx.x + 1) * n_items) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation: whole warp, 2 accumulator stages of NB fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"((uint32_t)(2 * NB)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      long long dbg_wait = 0;
      const long long k0 = dbg ? clock64() : 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int coord0 = (t / n_body_tiles) * kTcM;
        const int body0 = (t % n_body_tiles) * NB;
        for (int kc = 0; kc < kch; ++kc) {
          const long long w0 = dbg ? clock64() : 0;
          mbar_wait(&empty[stage], phase ^ 1);
          if (dbg) dbg_wait += clock64() - w0;
          uint8_t* st = smem + stage * S::kStageBytes;
          mbar_arrive_expect_tx(&full[stage], S::kStageBytes);
          tma_load_3d(st, &tmapA, &full[stage], kc * kChunkElems, 0, coord0);
          tma_load_3d(st + S::kABytes, &tmapA, &full[stage], kc * kChunkElems, 1, coord0);
          tma_load_3d(st + 2 * S::kABytes, &tmapB, &full[stage], kc * kChunkElems, 0, body0);
          tma_load_3d(st + 2 * S::kABytes + S::kBBytes, &tmapB, &full[stage], kc * kChunkElems, 1, body0);
          if (++stage == S::kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (dbg) { dbg[blockIdx.x * 8 + 0] = dbg_wait; dbg[blockIdx.x * 8 + 1] = clock64() - k0; }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (elect_one()) {
      // instruction descriptor: D=f32, A/B format, K-major both, N>>3 at [17,23), M>>4 at [24,29)
      constexpr uint32_t fmt = kKind == 0 ? 1u /*BF16*/ : 2u /*TF32*/;
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NB >> 3) << 17) |
                                 ((uint32_t)(kTcM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      long long dbg_full = 0, dbg_tm = 0;
      const long long k0 = dbg ? clock64() : 0;
      for (int t = t_begin; t < t_end; ++t) {
        long long w0 = dbg ? clock64() : 0;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        if (dbg) dbg_tm += clock64() - w0;
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NB);
        for (int kc = 0; kc < kch; ++kc) {
          w0 = dbg ? clock64() : 0;
          mbar_wait(&full[stage], phase);
          if (dbg) dbg_full += clock64() - w0;
          tcgen05_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t a_lo = a_hi + S::kABytes;
          const uint32_t b_hi = a_hi + 2 * S::kABytes;
          const uint32_t b_lo = b_hi + S::kBBytes;
          const int nks = min(4, ksteps - kc * 4);
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t dA_hi = umma_desc_sw128(a_hi + ks * 32), dA_lo = umma_desc_sw128(a_lo + ks * 32);
            const uint64_t dB_hi = umma_desc_sw128(b_hi + ks * 32), dB_lo = umma_desc_sw128(b_lo + ks * 32);
            umma<kKind>(d_tmem, dA_lo, dB_hi, idesc, (kc | ks) != 0);   // small terms first
            umma<kKind>(d_tmem, dA_hi, dB_lo, idesc, 1u);
            umma<kKind>(d_tmem, dA_hi, dB_hi, idesc, 1u);
          }
          tcgen05_commit(&empty[stage]);   // smem stage reusable once these MMAs retire
          if (++stage == S::kStages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(&tmem_full[acc]);   // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (dbg) { dbg[blockIdx.x * 8 + 2] = dbg_full; dbg[blockIdx.x * 8 + 3] = dbg_tm; dbg[blockIdx.x * 8 + 4] = clock64() - k0; }
    }
  } else {
    // ===================================== epilogue =========================================
    // TMEM -> registers -> shared staging tile [32 bodies][128 coords] -> TMA store.  (Per-thread global
    // stores from 4 warps topped out at ~2 TB/s here; the bulk store keeps the LSU out of the way.)
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const bool issuer = (warp == 2 && lane == 0);
    int acc = 0; uint32_t acc_phase = 0;
    int buf = 0;
    long long dbg_w = 0;
    const long long k0 = dbg ? clock64() : 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int coord0 = (t / n_body_tiles) * kTcM;
      const int body0 = (t % n_body_tiles) * NB;
      const long long w0 = dbg ? clock64() : 0;
      mbar_wait(&tmem_full[acc], acc_phase);
      if (dbg) dbg_w += clock64() - w0;
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NB);
#pragma unroll 1
      for (int c0 = 0; c0 < NB; c0 += 32) {
        if (body0 + c0 >= nb) break;   // warp-uniform (and uniform across the 4 epilogue warps)
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        // the staging buffer we are about to overwrite was handed to a bulk store two blocks ago
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* tile = out_smem + buf * (32 * kTcM) + q * 32 + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) tile[j * kTcM] = __uint_as_float(v[j]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer && dbg_mode == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmapOut)), "r"(smem_u32(out_smem + buf * (32 * kTcM))),
                         "r"(coord0), "r"(body0 + c0) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        buf ^= 1;
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (dbg && warp == 2 && lane == 0) { dbg[blockIdx.x * 8 + 5] = dbg_w; dbg[blockIdx.x * 8 + 6] = clock64() - k0; }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * NB)));
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(r >> 16);
}
static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline float f32_to_tf32_rna(float f) {   // cvt.rna.tf32.f32
  uint32_t u;
  memcpy(&u, &f, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D map {32 floats of K, rows, 2 parts} over a part-major [2, part_rows, 32] tf32 operand (128 B rows)
static inline int tc_encode_rows32(void* fn, CUtensorMap* map, void* base, size_t rows, size_t part_stride_bytes,
                                   int box_rows) {
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {128, (cuuint64_t)part_stride_bytes};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill)>(fn)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(WHMR_E_CUDA, "cuTensorMapEncodeTiled (rows32) failed with CUresult %d", (int)r);
  return WHMR_OK;
}

// 2-D map {64 halfs, rows} over a [rows, 64] fp16 hi|lo operand (128-byte rows), 128B swizzle
static inline int tc_encode_rows64h(void* fn, CUtensorMap* map, void* base, size_t rows, int box_rows) {
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill)>(fn)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(WHMR_E_CUDA, "cuTensorMapEncodeTiled (rows64h) failed with CUresult %d", (int)r);
  return WHMR_OK;
}

// 3-D map over a [rows, 2, KP] hi|lo operand: box = (128 bytes of K, 1 part, box_rows rows), 128B swizzle
static inline int tc_encode(void* fn, CUtensorMap* map, int kind, void* base, int KP, int rows, int box_rows) {
  const int elem = kind == 0 ? 2 : 4;
  cuuint64_t dims[3] = {(cuuint64_t)KP, 2, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)KP * elem, (cuuint64_t)2 * KP * elem};
  cuuint32_t box[3] = {(cuuint32_t)(128 / elem), 1, (cuuint32_t)box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<PFN_encodeTiled>(fn)(
      map, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box,
      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(WHMR_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return WHMR_OK;
}

// The same [rows, 2, KP] hi|lo operand with the PART as the outermost box dimension: dims {KP, rows, 2}, strides {row, part},
// box {128 bytes of K, box_rows, 2} -> shared-memory image [part][row][128 B] = the hi tile followed by the lo tile, by ONE copy.
static inline int tc_encode_both_parts(void* fn, CUtensorMap* map, int kind, void* base, int KP, int rows, int box_rows) {
  const int elem = kind == 0 ? 2 : 4;
  cuuint64_t dims[3] = {(cuuint64_t)KP, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)2 * KP * elem, (cuuint64_t)KP * elem};
  cuuint32_t box[3] = {(cuuint32_t)(128 / elem), (cuuint32_t)box_rows, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<PFN_encodeTiled>(fn)(
      map, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box,
      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? WHMR_OK : WHMR_E_CUDA;    // the caller falls back to the per-part maps
}

template <typename Arena>
static inline int tc_plan_create(const SmplDevice& d, const std::vector<float>& posedirs_p /*[KP,NP]*/, Arena& arena,
                                 TcPlan* plan) {
  // driver entry point without linking libcuda
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return set_error(WHMR_E_CUDA, "cuTensorMapEncodeTiled unavailable: %s", cudaGetErrorString(e));
  plan->encode_fn = fn;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&plan->num_sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t n = (size_t)d.NP * 2 * d.KP;
  std::vector<uint16_t> hb(n);
  std::vector<float> ht(n);
  for (int row = 0; row < d.NP; ++row)
    for (int k = 0; k < d.KP; ++k) {
      const float x = posedirs_p[(size_t)k * d.NP + row];
      const uint16_t bh = f32_to_bf16_rn(x);
      const uint16_t bl = f32_to_bf16_rn(x - bf16_to_f32(bh));
      hb[((size_t)row * 2 + 0) * d.KP + k] = bh;
      hb[((size_t)row * 2 + 1) * d.KP + k] = bl;
      const float th = f32_to_tf32_rna(x);
      ht[((size_t)row * 2 + 0) * d.KP + k] = th;
      ht[((size_t)row * 2 + 1) * d.KP + k] = f32_to_tf32_rna(x - th);
    }
  uint16_t* db = nullptr;
  float* dt = nullptr;
  e = arena.upload(hb, &db);
  if (e == cudaSuccess) e = arena.upload(ht, &dt);
  if (e != cudaSuccess) return set_error(WHMR_E_CUDA, "posedirs split upload failed: %s", cudaGetErrorString(e));
  plan->A_bf16 = db;
  plan->A_tf32 = dt;
  int rc = tc_encode(fn, &plan->tmapA_bf16, 0, db, d.KP, d.NP, kTcM);
  if (rc) return rc;
  rc = tc_encode(fn, &plan->tmapA_tf32, 1, dt, d.KP, d.NP, kTcM);
  if (rc) return rc;
  plan->both_ok = tc_encode_both_parts(fn, &plan->tmapA_both[0], 0, db, d.KP, d.NP, kTcM) == WHMR_OK &&
                  tc_encode_both_parts(fn, &plan->tmapA_both[1], 1, dt, d.KP, d.NP, kTcM) == WHMR_OK;
  cudaFuncSetAttribute(pose_blend_tc_kernel<0, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<256>::kTotal);
  cudaFuncSetAttribute(pose_blend_tc_kernel<1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<256>::kTotal);
  cudaFuncSetAttribute(pose_blend_tc_kernel<0, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<128>::kTotal);
  cudaFuncSetAttribute(pose_blend_tc_kernel<1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<128>::kTotal);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(WHMR_E_CUDA, "cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
  plan->ready = 1;
  return WHMR_OK;
}

// offsets for bodies [b0, b0+nb) of the split pose feature `pf_split` ([B,2,KP], bf16 or tf32) -> out[0..nb)
static inline int tc_pose_blend_launch(const TcPlan& plan, const SmplDevice& d, int gemm_mode, const void* pf_split,
                                       int B, int b0, int nb, float* out, cudaStream_t st) {
  (void)B;
  if (!plan.ready) return set_error(WHMR_E_INVALID, "tensor-core plan not initialised");
  const int kind = gemm_mode == WHMR_GEMM_TC_BF16X3 ? 0 : 1;
  const int elem = kind == 0 ? 2 : 4;
  const int ksteps = d.KP * elem / 32;
  const int n_coord_tiles = d.NP / kTcM;
  // body-tile width: 256 halves the operand traffic per MMA; 128 gives more, smaller work items when
  // the batch is too small to fill the machine with 256-wide tiles
  const bool wide = (long long)n_coord_tiles * ceil_div(nb, 256) >= 2LL * plan.num_sms;
  const int NB = wide ? 256 : 128;
  const int n_body_tiles = ceil_div(nb, NB);
  const int n_items = n_coord_tiles * n_body_tiles;
  CUtensorMap tmapB;
  char* base = const_cast<char*>(static_cast<const char*>(pf_split)) + (size_t)b0 * 2 * d.KP * elem;
  // extent rounded up to whole tiles: the workspace is padded to kTcBodyTile rows and a column
  // (body) of the accumulator depends only on its own row, so the pad rows are computed and dropped
  int rc = tc_encode(plan.encode_fn, &tmapB, kind, base, d.KP, n_body_tiles * NB, NB);
  if (rc) return rc;
  const int grid = std::min(plan.num_sms, n_items);
  const CUtensorMap& tmapA = kind == 0 ? plan.tmapA_bf16 : plan.tmapA_tf32;
  // output: 2-D map {NP coords, nb bodies} over the planar pose-offset buffer, box {128, 32}; rows >= nb are clipped
  CUtensorMap tmapOut;
  {
    cuuint64_t dims[2] = {(cuuint64_t)d.NP, (cuuint64_t)nb};
    cuuint64_t strides[1] = {(cuuint64_t)d.NP * 4};
    cuuint32_t box[2] = {(cuuint32_t)kTcM, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<PFN_encodeTiled>(plan.encode_fn)(
        &tmapOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(WHMR_E_CUDA, "cuTensorMapEncodeTiled (out) failed with CUresult %d", (int)r);
  }
  long long* dbg = nullptr;
  static const int dbg_mode = getenv("WHMR_TC_DEBUG_MODE") ? atoi(getenv("WHMR_TC_DEBUG_MODE")) : 0;
  static const bool dbg_on = getenv("WHMR_TC_DEBUG") != nullptr;
  if (dbg_on) { cudaMalloc(&dbg, sizeof(long long) * 8 * grid); cudaMemsetAsync(dbg, 0, sizeof(long long) * 8 * grid, st); }
  if (kind == 0 && NB == 256)
    pose_blend_tc_kernel<0, 256><<<grid, kTcThreads, TcSmem<256>::kTotal, st>>>(tmapA, tmapB, tmapOut, nb, d.NP, n_body_tiles, n_items, ksteps, dbg, dbg_mode);
  else if (kind == 0)
    pose_blend_tc_kernel<0, 128><<<grid, kTcThreads, TcSmem<128>::kTotal, st>>>(tmapA, tmapB, tmapOut, nb, d.NP, n_body_tiles, n_items, ksteps, dbg, dbg_mode);
  else if (NB == 256)
    pose_blend_tc_kernel<1, 256><<<grid, kTcThreads, TcSmem<256>::kTotal, st>>>(tmapA, tmapB, tmapOut, nb, d.NP, n_body_tiles, n_items, ksteps, dbg, dbg_mode);
  else
    pose_blend_tc_kernel<1, 128><<<grid, kTcThreads, TcSmem<128>::kTotal, st>>>(tmapA, tmapB, tmapOut, nb, d.NP, n_body_tiles, n_items, ksteps, dbg, dbg_mode);
  WHMR_LAUNCHED("pose_blend_tc_kernel");
  if (dbg) {   // WHMR_TC_DEBUG: per-role wait/total cycles, averaged over CTAs (debug only: synchronises)
    cudaStreamSynchronize(st);
    std::vector<long long> hd((size_t)8 * grid);
    cudaMemcpy(hd.data(), dbg, hd.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(dbg);
    double a[8] = {0};
    for (int c = 0; c < grid; ++c) for (int k = 0; k < 8; ++k) a[k] += (double)hd[(size_t)c * 8 + k] / grid;
    fprintf(stderr, "[whmr tc dbg] NB=%d nb=%d items=%d grid=%d | producer wait_empty %.0f of %.0f | mma wait_full %.0f "
            "wait_tmem %.0f of %.0f | epi wait_full %.0f of %.0f cycles\n", NB, nb, n_items, grid, a[0], a[1], a[2], a[3], a[4],
            a[5], a[6]);
  }
  return WHMR_OK;
}

}  // namespace whmr
