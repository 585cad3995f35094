// One kernel for the vertex half of the SMPL forward: pose+shape blend (tcgen05, bf16x3), blend of the skinning
// transforms (tcgen05, fp16 hi/lo split, 3 MMAs per K step: 22 mantissa bits) and the skinning epilogue -- the [B, 3*V] pose-offset intermediate never leaves
// the SM: it is produced in TMEM and consumed from TMEM.
//
//   off_c[v, b] = sum_k P_c[v, k] * pf[b, k]          c in {x,y,z}; K = 207 pose terms + 10 betas (SURVEY K2+K4)
//   T[v, b]     = sum_j w[v, j] * A_j[b]              3x4 per vertex and body (K6)
//   v'[b, v]    = T[v, b] . [v_template[v] + off[v, b] ; 1] (+ transl[b])      (K7)   (+ fused read-out emits)
//
// Work item = (tile of 128 vertices) x (run of 16..64 bodies), vertex-major, cut so that all CTAs get equal work.  TMEM (512 columns, lane =
// vertex): two stages of three NBI-column accumulators (x|y|z pose offsets, one body per column) + one 96-column
// accumulator for the blended transforms of 8 bodies.  Warp roles (640 threads, 1 CTA/SM):
//   warp 0      TMA producer, pose blend: per 128-byte K chunk one pose-feature stage {hi,lo} (NBI rows) and per
//               (chunk, plane) one posedirs stage {hi,lo} (128 rows) -- a 3 x 32 KB ring;
//   warp 1      pose-blend MMA issuer (+ TMEM allocation): 3 MMAs per K step (lo.hi + hi.lo + hi.hi);
//   warp 2      TMA producer, skinning: weight tile {hi,lo} when the vertex tile changes, A^T {hi,lo} per 8 bodies;
//   warp 3      skinning MMA issuer: 2 K steps x 3 MMAs (M=128, N=96, fp16 hi/lo) per 8 bodies;
//   warps 4-19  epilogue: TMEM lane quarter q = warp%4, two bodies per warp and 8-body group.  A warp pulls its
//               24 T columns and 6 offset columns into registers, releases the T accumulator at once (so the next
//               group's MMAs overlap its arithmetic and stores), applies the transform, transposes through
//               shared memory and writes three coalesced 128-byte rows per body.
// The two MMA issuers run independently; the tensor pipe interleaves their instructions, so the pose blend of
// item i+1 proceeds while item i is skinned and stored.
// Operand traffic per item (L2 -> smem): posedirs 344 KB + pose feature 0.9 KB/body + A^T 1.5 KB/body + weights
// 16 KB per vertex tile; measured ingest capability 130-140 GB/s/SM (tools/microbench/l2_ingest_bench.cu).
#pragma once
#include <cuda/std/type_traits>

#include "pose_blend_tc.cuh"
#include "readout.cuh"
#include "skin_tc.cuh"

namespace whmr {

// A-operand collector reuse (.collector::a::fill / ::lastuse) between the two MMAs that share a hi tile.  Alone in the
// tensor pipe it saves ~12 % per MMA triplet (tools/microbench/umma_rate_bench.cu: 78.7 -> 69.1 cycles); inside this
// kernel, where two issuing threads interleave their MMAs, it measured no gain (and as a run-time switch the extra
// predicated UTCHMMAs cost 2.7 us per B=256 launch), so it is a compile-time option, off by default.
#ifdef WHMR_FUSED_COLLECTOR
constexpr bool kFuCollector = true;
#else
constexpr bool kFuCollector = false;
#endif
// Skinning weights as the TMEM-resident A operand of the transform blend (tcgen05.mma [d], [a_tmem], b_desc): the four
// warps w4 == 0 of the epilogue write the tile's weight rows (64 halfs = 32 columns per vertex/lane: hi K-steps 0,1 then
// lo K-steps 0,1) into TMEM columns 480..511 with tcgen05.st when the vertex tile changes; the transform MMAs then fetch
// no A tile from shared memory (4 KB per MMA at 64 B/clk = the per-instruction floor of section 5.2 of the notes).
// Measured (B200): B=256 loop step 384.5 -> 377.3 us, 16 k bodies + H36M read-outs 15.2 -> 15.9 M bodies/s, SMPL alone
// 19.4 -> 20.8 M bodies/s.  -DWHMR_FUSED_WSMEM restores the shared-memory operand (TMA weight tile).
#ifdef WHMR_FUSED_WSMEM
constexpr bool kFuWTmem = false;
#else
constexpr bool kFuWTmem = true;
#endif
constexpr int kFuWCol = 480;                  // TMEM column of the weight operand (all plans end at 480)
// Bodies per blended-transform tile.  8 (default): one 96-column accumulator in the 64-body plan, all sixteen epilogue warps
// on every group.  4 (-DWHMR_FUSED_GB4): two 48-column accumulators in every plan and two TEAMS of eight epilogue warps, team t
// on the groups that land in accumulator t -- the transform round trip of one group overlaps the arithmetic and stores of the
// other (profiles/r02_notes.md section 4.9).
#ifdef WHMR_FUSED_GB4
constexpr int kFuGB = 4;
#else
constexpr int kFuGB = 8;
#endif
constexpr bool kFuTeams = kFuGB == 4;
constexpr int kFuTN = kFuGB * 12;             // 96 (48) accumulator columns
constexpr int kFuMaxAStages = 4;
constexpr int kFuABytes = 2 * kTcM * 128;     // 32 KB: posedirs {hi,lo} of one (K chunk, plane)
constexpr int kFuMaxPfStages = 3;
constexpr int kFuWBytes = kTcM * 128;         // 16 KB: skinning weights, fp16 hi|lo along K (64 halfs per vertex)
constexpr int kFuAtStages = kFuTeams ? 4 : 2;
constexpr int kFuAtBytes = kFuTN * 128;       // 12 KB: A^T of 8 bodies, fp16 hi|lo along K
constexpr int kFuEpiWarps = 16;
#ifdef WHMR_FUSED_STORE32
constexpr bool kFuStore64 = false;            // vertex stores of the epilogue as 4-byte (three per body and lane) ...
#else
constexpr bool kFuStore64 = true;             // ... or 8-byte accesses (one and a half): fewer LSU instructions in the busiest role
#endif
constexpr int kFuThreads = (4 + kFuEpiWarps) * 32;
// TMEM / shared-memory plan, by the template parameter MAXM = micro-items (16 bodies) per item.
// Measured (tools/microbench/umma_rate_bench.cu): a tcgen05.mma with M=128, K=16 costs max(~70, N/2) cycles -- the
// 4 KB A tile enters at 64 B/clk -- so below N=128 the cost of a pose-blend MMA does not depend on the number of
// bodies it covers.  The plan therefore packs as many bodies as TMEM allows behind every posedirs tile:
//   MAXM = 8: ONE stage of 3 x 128 pose-offset columns + one 96-column blended-transform stage (480): half the
//             pose-blend MMAs per body; the pose blend of item i+1 cannot overlap the epilogue of item i, but both
//             sides are bound by the same tensor pipe, which stays busy with the skinning MMAs meanwhile;
//   MAXM = 4: 2 x 3 x 64 pose-offset columns + one blended-transform stage (480);
//   MAXM = 3: 2 x 3 x 48 pose-offset columns + TWO blended-transform stages (480): tiny batches.
template <int MAXM> struct FuTmem {
  static constexpr int kNB = 16 * MAXM;              // max bodies per item
  static constexpr int kOffStages = MAXM <= 4 ? 2 : 1;
  static constexpr int kOffStage = 3 * kNB;          // columns per pose-offset stage
  static constexpr int kT = kOffStages * kOffStage;  // first blended-transform column
  static constexpr int kTStages = (512 - kT) / kFuTN >= 2 ? 2 : 1;
  static_assert(kT + kTStages * kFuTN <= 512, "TMEM budget");
  static_assert(!kFuTeams || kTStages == 2, "the two epilogue teams own one transform accumulator each");
  // shared memory
  static constexpr int kAStages = MAXM <= 4 ? 4 : 3;
  static constexpr int kPfPart = (kNB < 64 ? 64 : kNB) * 128;   // pose feature, one of {hi,lo}, one K chunk
  static constexpr int kPfBytes = 2 * kPfPart;
  static constexpr int kOffA = 0;
  static constexpr int kOffPf = kOffA + kAStages * kFuABytes;
  // pose-feature ring: a third stage in the room the weight tile left when it moved to TMEM (64-body and smaller plans)
  static constexpr int kPfStages = (kFuWTmem && MAXM <= 4) ? 3 : 2;
  static constexpr int kOffW = kOffPf + kPfStages * kPfBytes;
  static constexpr int kOffAt = kOffW + (kFuWTmem ? 0 : kFuWBytes);
  static constexpr int kOffStg = kOffAt + kFuAtStages * kFuAtBytes;
  static constexpr int kOffBars = kOffStg + kFuEpiWarps * 192 * 4;
  static constexpr int kSmem = kOffBars + 512 + 1024;
  static_assert(kSmem <= 232448, "fused SMPL kernel exceeds the 227 KB shared-memory limit");
};

struct FusedParams {
  const float* v_template_p;  // [3, VP]
  const float* transl;        // [nb,3] or null
  float* verts;               // [nb, V, 3]
  EmitTable emit;             // fused read-outs (readout.cuh); grp_ptr == null: off
  float* ro_out;
  int ro_B, ro_b0;
  int nb, V, VP;
  const uint4* W16;           // weights fp16 hi|lo [VP, 64] in global memory (kFuWTmem: read by the epilogue warps)
  int pieces;                 // > 0: balanced pieces per vertex tile, CTA c takes pieces c and c + grid (small batches)
  int split;                  // > 0: CTA c takes part c % split of vertex tile c / split (grid = tiles * split): one item per CTA
  int npv;                    // 16-body micro-items per vertex tile = ceil(nb / 16)
  int n_micro;                // (VP/128) * npv: the unit of work distribution
  int kch, ksteps;            // pose blend: 128-byte K chunks, 32-byte K steps (bf16: 16 elements)
  int jsteps;                 // skinning: ceil(J/16) fp16 K steps
  int store64;                // 1: `verts` is 8-byte aligned and 3 * V is even, so every body row is: the epilogue may store float2
  int merged;                 // 1: tmapP / tmapPf are the {K chunk, rows, 2 parts} maps: ONE copy brings hi AND lo (and all 64 pose-feature rows)
  unsigned backoff;           // WHMR_FUSED_BACKOFF: ns between mbarrier polls of the single-thread roles
  int dbg_mode;               // WHMR_FUSED_DBGMODE bits (timing experiments only): 1 skip vertex stores, 2 skip read-out emits, 4 skip the transposes
  long long* dbg;             // WHMR_FUSED_DEBUG: [grid][16] per-role wait/total cycles, or null
};

__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void umma_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x2(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr));
}

// packed fp32 pairs (Blackwell FFMA2 / FADD2): one instruction for the two bodies a warp handles per group
__device__ __forceinline__ uint64_t u32x2_pack(uint32_t lo, uint32_t hi) {
  uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r;
}
__device__ __forceinline__ uint64_t f32x2_pack(float lo, float hi) {
  uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void f32x2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint64_t f32x2_add(uint64_t a, uint64_t b) {
  uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

// wait that adds its duration to `acc` when instrumentation is on
#define WHMR_FU_WAIT(bar, par, acc)                       \
  do {                                                    \
    if (dbgp) { const long long _w0 = clock64(); mbar_wait(bar, par); acc += clock64() - _w0; } \
    else mbar_wait(bar, par);                             \
  } while (0)
// single-thread roles (producers, MMA issuers): polls spaced by p.backoff ns
#define WHMR_FU_WAIT_R(bar, par, acc)                     \
  do {                                                    \
    if (dbgp) { const long long _w0 = clock64(); mbar_wait_backoff(bar, par, p.backoff); acc += clock64() - _w0; } \
    else mbar_wait_backoff(bar, par, p.backoff);          \
  } while (0)

// kDbg: instrumented instantiation (WHMR_FUSED_DEBUG / WHMR_FUSED_DBGMODE); the production instantiation carries
// none of the probes' predicates or branches (the epilogue is instruction-issue bound).
// kTwo: TWO pose-blend issuing threads (warps 1 and 2).  One thread gets a tcgen05.mma into the pipe every ~64-86 cycles
// (profiles/r02_notes.md section 1), and a 48/64-body pose-blend MMA keeps the tensor core busy for 24/32: with one
// issuer the pose blend costs 126 x ~86 cycles per item whatever the item's width.  The K chunks of an item alternate
// between the two issuers (chunk kc of the n-th item of the CTA belongs to issuer (kc + n) & 1); both accumulate into the
// same TMEM columns.  The operand rings stay shared and in global order, but every stage has one "full" barrier PER
// ISSUER: a parity wait is only sound if the waiter sees every phase of its barrier, and an issuer that polled a stage
// whose previous fill belonged to the other one would pass on a stale phase.  The issuer that does not own chunk 0 waits
// for `first_issued` (the accumulate = 0 MMAs are in the pipe) before its first MMA; `off_full` takes both commits.  The
// skinning issuer (warp 3) loads its own A^T tiles, one group ahead, so no warp is added.
template <int MAXM, bool kDbg, bool kTwo = false, int kKind = 0>
__global__ void __launch_bounds__(kFuThreads, 1)
smpl_fused_tc_kernel(const __grid_constant__ CUtensorMap tmapP,    // posedirs+shapedirs bf16 [NP, 2, KP]
                     const __grid_constant__ CUtensorMap tmapPf,   // pose feature bf16 [bodies, 2, KP], box rows = nbi
                     const __grid_constant__ CUtensorMap tmapW,    // weights fp16 hi|lo [VP, 64]
                     const __grid_constant__ CUtensorMap tmapAt,   // A^T fp16 hi|lo [bodies*12, 64], box rows = 96
                     FusedParams p) {
  using TM = FuTmem<MAXM>;
  // kKind: arithmetic of the POSE blend -- 0: bf16 hi/lo (kind::f16, 64 K elements per 128-byte chunk, K = 16 per MMA),
  // 1: tf32 hi/lo = the north star's 3xTF32 (kind::tf32, 32 K elements per chunk, K = 8 per MMA: twice the chunks and MMAs
  // per item, ~2^-21 instead of ~2^-16 relative error per product).  Stage sizes, rings, barriers, TMEM plan, the transform
  // blend (fp16 hi/lo) and the epilogue are the same; the operands come from the tf32 copies (tmapP, tmapPf encoded for fp32).
  constexpr int kChunkElems = kKind == 0 ? 64 : 32;
  constexpr uint32_t kFmt = kKind == 0 ? 1u /*BF16*/ : 2u /*TF32*/;
  long long* const dbgp = kDbg ? p.dbg : nullptr;
  const int dbg_mode = kDbg ? p.dbg_mode : 0;
  extern __shared__ uint8_t smem_raw[];
  unsigned long long gt_entry = 0;
  if (dbgp && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_entry));
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_ring = smem + TM::kOffA;
  uint8_t* pf_ring = smem + TM::kOffPf;
  uint8_t* w_smem = smem + TM::kOffW;
  uint8_t* at_ring = smem + TM::kOffAt;
  float* stage_out = reinterpret_cast<float*>(smem + TM::kOffStg);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TM::kOffBars);
  constexpr int kFuAStages = TM::kAStages;
  constexpr int kFuPfBytes = TM::kPfBytes, kFuPfPart = TM::kPfPart;
  uint64_t* a_full = bars;                         // [4]
  uint64_t* a_empty = a_full + kFuMaxAStages;      // [4]
  uint64_t* pf_full = a_empty + kFuMaxAStages;     // [2]
  constexpr int kFuPfStages = TM::kPfStages;
  uint64_t* pf_empty = pf_full + kFuMaxPfStages;   // [3]
  uint64_t* w_full = pf_empty + kFuMaxPfStages;
  uint64_t* w_empty = w_full + 1;
  uint64_t* at_full = w_empty + 1;                 // [2]
  uint64_t* at_empty = at_full + kFuAtStages;      // [2]
  uint64_t* off_full = at_empty + kFuAtStages;     // [2]
  uint64_t* off_empty = off_full + 2;              // [2]
  uint64_t* t_full = off_empty + 2;                // [2]
  uint64_t* t_empty = t_full + 2;                  // [2]
  uint64_t* a_full2 = t_empty + 2;                 // [4] kTwo: stage filled for issuer 1 (a_full: issuer 0)
  uint64_t* pf_full2 = a_full2 + kFuMaxAStages;    // [3]
  uint64_t* first_issued = pf_full2 + kFuMaxPfStages;   // [2] kTwo: chunk 0 of the item on this offset stage is in the pipe
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(first_issued + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work distribution: the (vertex tile, 16-body micro-item) grid is cut into gridDim.x contiguous, equal ranges
  // (vertex-major), so every CTA gets the same number of bodies +-16.  A range is walked as items = runs of <= 4
  // micro-items (<= 64 bodies) inside one vertex tile, sized evenly (5 micro-items -> 3 + 2, not 4 + 1).
  // A CTA's work is up to two ranges of micro-items.  Small batches (p.pieces > 0): the kernel time is the CTA with the
  // most ITEMS (every item streams a whole posedirs tile, ~9 k cycles whatever its width), and equal ranges that straddle
  // a tile boundary give some CTAs three items.  So every vertex tile is cut into p.pieces balanced pieces of <= MAXM
  // micro-items, and CTA c takes piece c and piece c + gridDim.x (if it exists): never more than two items per CTA.
  int seg_b[2], seg_e[2];
  seg_b[1] = seg_e[1] = 0;
  if (p.pieces > 0) {
    const int base = p.npv / p.pieces, rem = p.npv - base * p.pieces, total = (p.n_micro / p.npv) * p.pieces;
    auto piece = [&](int q, int& b, int& e) {
      const int t = q / p.pieces, j = q - t * p.pieces;
      b = t * p.npv + j * base + min(j, rem);
      e = b + base + (j < rem ? 1 : 0);
    };
    piece(blockIdx.x, seg_b[0], seg_e[0]);
    if ((int)(blockIdx.x + gridDim.x) < total) piece(blockIdx.x + gridDim.x, seg_b[1], seg_e[1]);
  } else if (p.split > 0) {   // tile-aligned parts
    const int vt = blockIdx.x / p.split, part = blockIdx.x - vt * p.split;
    seg_b[0] = vt * p.npv + (part * p.npv) / p.split;
    seg_e[0] = vt * p.npv + ((part + 1) * p.npv) / p.split;
  } else {
    seg_b[0] = (int)(((long long)blockIdx.x * p.n_micro) / gridDim.x);
    seg_e[0] = (int)(((long long)(blockIdx.x + 1) * p.n_micro) / gridDim.x);
  }
  const int m_begin = seg_b[0];
  const bool any_work = seg_b[0] < seg_e[0];
  struct Item { int vt, body0, len, nbod, ng; };   // len: micro-items; nbod: valid bodies; ng: 8-body groups
  auto item_at = [&](int m, int m_end) {
    Item it;
    it.vt = m / p.npv;
    const int mi = m - it.vt * p.npv;
    const int L = min(m_end, (it.vt + 1) * p.npv) - m;
    const int n_it = (L + MAXM - 1) / MAXM;
    it.len = (L + n_it - 1) / n_it;
    it.body0 = mi * 16;
    it.nbod = min(it.len * 16, p.nb - it.body0);
    it.ng = (it.nbod + kFuGB - 1) / kFuGB;
    return it;
  };
#define WHMR_FU_FOR_ITEMS for (int sg = 0; sg < 2; ++sg) for (int m = seg_b[sg]; m < seg_e[sg];)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kFuAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kFuPfStages; ++s) { mbar_init(&pf_full[s], 1); mbar_init(&pf_empty[s], 1); }
    mbar_init(w_full, kFuWTmem ? 4 : 1); mbar_init(w_empty, 1);
    for (int s = 0; s < kFuAtStages; ++s) { mbar_init(&at_full[s], 1); mbar_init(&at_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&off_full[s], kTwo ? 2 : 1); mbar_init(&off_empty[s], kFuEpiWarps); }
    if (kTwo) {
      for (int s = 0; s < kFuAStages; ++s) mbar_init(&a_full2[s], 1);
      for (int s = 0; s < kFuPfStages; ++s) mbar_init(&pf_full2[s], 1);
      for (int s = 0; s < 2; ++s) mbar_init(&first_issued[s], 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], kFuTeams ? kFuEpiWarps / 2 : kFuEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (dbgp && threadIdx.x == 0) {
    unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    dbgp[blockIdx.x * 16 + 15] = (long long)(g - gt_entry);   // prologue, ns
    dbgp[(gridDim.x + blockIdx.x) * 16 + 0] = (long long)gt_entry;
  }

  if (warp == 0) {
    // ============================ TMA producer: pose-blend operands ============================
    if (elect_one()) {
      int as = 0; uint32_t aph = 0;
      int ps = 0; uint32_t pph = 0;
      long long d_pf = 0, d_a = 0;
      const long long k0 = dbgp ? clock64() : 0;
      // PDL: posedirs are constants of the handle, so the first ring fill is issued before the predecessor (the
      // chain kernel, which writes the pose feature) has finished; `pre` counts the stages already in flight
      int pre = 0;
      if (any_work) {
        const Item it0 = item_at(m_begin, seg_e[0]);
        for (int i = 0; i < kFuAStages && i < 3 * p.kch; ++i) {
          const int kc = i / 3, c = i % 3;
          uint8_t* ast = a_ring + i * kFuABytes;
          uint64_t* fb = (kTwo && (kc & 1)) ? &a_full2[i] : &a_full[i];   // item 0: chunk kc belongs to issuer kc & 1
          mbar_arrive_expect_tx(fb, kFuABytes);
          if (p.merged) {
            tma_load_3d(ast, &tmapP, fb, kc * kChunkElems, c * p.VP + it0.vt * kTcM, 0);
          } else {
            tma_load_3d(ast, &tmapP, fb, kc * kChunkElems, 0, c * p.VP + it0.vt * kTcM);
            tma_load_3d(ast + kTcM * 128, &tmapP, fb, kc * kChunkElems, 1, c * p.VP + it0.vt * kTcM);
          }
          ++pre;
        }
      }
      pdl_wait();
      int item_seq = 0;
      WHMR_FU_FOR_ITEMS {
        const Item it = item_at(m, seg_e[sg]);
        m += it.len;
        const int vt = it.vt, body0 = it.body0;
        const uint32_t pf_bytes = 2u * (uint32_t)it.len * 2048u;
        for (int kc = 0; kc < p.kch; ++kc) {
          const bool second = kTwo && ((kc + item_seq) & 1);     // the issuer this chunk is filled for
          WHMR_FU_WAIT_R(&pf_empty[ps], pph ^ 1, d_pf);
          uint8_t* pst = pf_ring + ps * kFuPfBytes;
          uint64_t* pfb = second ? &pf_full2[ps] : &pf_full[ps];
          if (p.merged) {   // one box {K chunk, kFuPfPart / 128 rows, hi + lo}: rows past the item are never multiplied (N = 16 * len)
            mbar_arrive_expect_tx(pfb, (uint32_t)kFuPfBytes);
            tma_load_3d(pst, &tmapPf, pfb, kc * kChunkElems, body0, 0);
          } else {
            mbar_arrive_expect_tx(pfb, pf_bytes);
            for (int u = 0; u < it.len; ++u) {   // 16-row boxes, stacked: same image as one (16*len)-row box
              tma_load_3d(pst + u * 2048, &tmapPf, pfb, kc * kChunkElems, 0, body0 + u * 16);
              tma_load_3d(pst + kFuPfPart + u * 2048, &tmapPf, pfb, kc * kChunkElems, 1, body0 + u * 16);
            }
          }
          if (++ps == kFuPfStages) { ps = 0; pph ^= 1; }
          for (int c = 0; c < 3; ++c) {
            if (pre > 0) {
              --pre;                       // this stage was filled by the prefetch above
            } else {
              WHMR_FU_WAIT_R(&a_empty[as], aph ^ 1, d_a);
              uint8_t* ast = a_ring + as * kFuABytes;
              uint64_t* fb = second ? &a_full2[as] : &a_full[as];
              mbar_arrive_expect_tx(fb, kFuABytes);
              if (p.merged) {
                tma_load_3d(ast, &tmapP, fb, kc * kChunkElems, c * p.VP + vt * kTcM, 0);
              } else {
                tma_load_3d(ast, &tmapP, fb, kc * kChunkElems, 0, c * p.VP + vt * kTcM);
                tma_load_3d(ast + kTcM * 128, &tmapP, fb, kc * kChunkElems, 1, c * p.VP + vt * kTcM);
              }
            }
            if (++as == kFuAStages) { as = 0; aph ^= 1; }
          }
        }
        ++item_seq;
      }
      if (dbgp) { long long* d = dbgp + blockIdx.x * 16; d[0] = d_pf; d[1] = d_a; d[2] = clock64() - k0; }
    }
  } else if (warp == 1 || (kTwo && warp == 2)) {
    // ============================ pose-blend MMA issuer(s) =====================================
    if (elect_one()) {
      const int me = warp - 1;          // kTwo: issuer 0 / 1
      uint64_t* const my_a_full = (kTwo && me) ? a_full2 : a_full;
      uint64_t* const my_pf_full = (kTwo && me) ? pf_full2 : pf_full;
      int as = 0; uint32_t aph = 0;     // one issuer: ring phase; two: bit s = parity of MY next fill of stage s
      int ps = 0; uint32_t pph = 0;
      int buf = 0; uint32_t bph = 0;
      int item_seq = 0;
      long long d_off = 0, d_pf = 0, d_a = 0;
      const long long k0 = dbgp ? clock64() : 0;
      WHMR_FU_FOR_ITEMS {
        const Item it = item_at(m, seg_e[sg]);
        m += it.len;
        const uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(it.len * 2) << 17) |
                               ((uint32_t)(kTcM >> 4) << 24);   // bf16 x bf16 -> f32, M=128, N=16*len
        WHMR_FU_WAIT_R(&off_empty[buf], bph ^ 1, d_off);
        tcgen05_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(buf * TM::kOffStage);
        bool first_seen = !kTwo || ((item_seq & 1) == me);   // I own chunk 0 (or there is only one issuer)
        for (int kc = 0; kc < p.kch; ++kc) {
          const bool mine = !kTwo || (((kc + item_seq) & 1) == me);
          if (!mine) {                  // the other issuer's chunk: only step over its stages
            if (++ps == kFuPfStages) ps = 0;
            for (int c = 0; c < 3; ++c) if (++as == kFuAStages) as = 0;
            continue;
          }
          if (kTwo) {
            WHMR_FU_WAIT_R(&my_pf_full[ps], (pph >> ps) & 1u, d_pf);
            pph ^= 1u << ps;
            if (!first_seen) {          // the accumulate = 0 MMAs of this item (chunk 0) must be in the pipe before mine
              mbar_wait_backoff(&first_issued[buf], bph, p.backoff);
              tcgen05_fence_after();
              first_seen = true;
            }
          } else {
            WHMR_FU_WAIT_R(&pf_full[ps], pph, d_pf);
          }
          const uint32_t b_hi = smem_u32(pf_ring + ps * kFuPfBytes), b_lo = b_hi + kFuPfPart;
          const int nks = min(4, p.ksteps - kc * 4);
          for (int c = 0; c < 3; ++c) {
            if (kTwo) {
              WHMR_FU_WAIT_R(&my_a_full[as], (aph >> as) & 1u, d_a);
              aph ^= 1u << as;
            } else {
              WHMR_FU_WAIT_R(&a_full[as], aph, d_a);
            }
            tcgen05_fence_after();
            const uint32_t a_hi = smem_u32(a_ring + as * kFuABytes), a_lo = a_hi + kTcM * 128;
            const uint32_t d_tmem = d_base + (uint32_t)(c * TM::kNB);
            for (int ks = 0; ks < nks; ++ks) {
              const uint64_t dA_hi = umma_desc_sw128(a_hi + ks * 32), dA_lo = umma_desc_sw128(a_lo + ks * 32);
              const uint64_t dB_hi = umma_desc_sw128(b_hi + ks * 32), dB_lo = umma_desc_sw128(b_lo + ks * 32);
              umma<kKind>(d_tmem, dA_lo, dB_hi, idesc, (kc | ks) != 0);   // small terms first
              umma<kKind, kFuCollector ? 1 : 0>(d_tmem, dA_hi, dB_lo, idesc, 1u);   // (collector: hi tile kept ...
              umma<kKind, kFuCollector ? 3 : 0>(d_tmem, dA_hi, dB_hi, idesc, 1u);   //  ... and reused, see kFuCollector)
            }
            tcgen05_commit(&a_empty[as]);
            if (++as == kFuAStages) { as = 0; if (!kTwo) aph ^= 1; }
          }
          tcgen05_commit(&pf_empty[ps]);
          if (++ps == kFuPfStages) { ps = 0; if (!kTwo) pph ^= 1; }
          if (kTwo && kc == 0) {        // chunk 0 is in the pipe: the other issuer may follow
            tcgen05_fence_before();
            mbar_arrive(&first_issued[buf]);
          }
        }
        tcgen05_commit(&off_full[buf]);
        if (++buf == TM::kOffStages) { buf = 0; bph ^= 1; }
        ++item_seq;
      }
      if (dbgp && me == 0) { long long* d = dbgp + blockIdx.x * 16; d[3] = d_off; d[4] = d_pf; d[5] = d_a; d[6] = clock64() - k0; }
    }
  } else if (!kTwo && warp == 2) {
    // ============================ TMA producer: skinning operands ==============================
    if (elect_one()) {
      int s = 0; uint32_t ph = 0, w_par = 1;
      int cur_vt = -1;
      if (any_work && !kFuWTmem) {   // the first weight tile does not depend on the predecessor either
        cur_vt = m_begin / p.npv;
        w_par ^= 1;            // first use of w_empty passes trivially
        mbar_arrive_expect_tx(w_full, kFuWBytes);
        tma_load_2d(w_smem, &tmapW, w_full, 0, cur_vt * kTcM);
      }
      pdl_wait();
      WHMR_FU_FOR_ITEMS {
        const Item it = item_at(m, seg_e[sg]);
        m += it.len;
        const int vt = it.vt;
        if (!kFuWTmem && vt != cur_vt) {
          mbar_wait_backoff(w_empty, w_par, p.backoff);   // MMAs on the previous weight tile have retired
          w_par ^= 1;
          mbar_arrive_expect_tx(w_full, kFuWBytes);
          tma_load_2d(w_smem, &tmapW, w_full, 0, vt * kTcM);
          cur_vt = vt;
        }
        const int ng = it.ng;
        for (int g = 0; g < ng; ++g) {
          mbar_wait_backoff(&at_empty[s], ph ^ 1, p.backoff);
          uint8_t* st = at_ring + s * kFuAtBytes;
          mbar_arrive_expect_tx(&at_full[s], kFuAtBytes);
          const int row0 = (it.body0 + g * kFuGB) * 12;
          tma_load_2d(st, &tmapAt, &at_full[s], 0, row0);
          if (++s == kFuAtStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ============================ skinning MMA issuer ==========================================
    if (elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kFuTN >> 3) << 17) |
                                 ((uint32_t)(kTcM >> 4) << 24);   // fp16 x fp16 -> f32, M=128, N=96
      int s = 0; uint32_t ph = 0, w_phase = 0, t_ph = 0;
      int cur_vt = -1;
      const uint32_t w_hi = smem_u32(w_smem), w_lo = w_hi + 64;   // lo half of every 128-byte row
      int ts = 0;
      long long d_t = 0, d_at = 0;
      const long long k0 = dbgp ? clock64() : 0;
      // kTwo: this thread is also the TMA producer of its A^T tiles, one 8-body group ahead (the tile of group n+1 goes
      // into the stage group n-1 used, whose MMAs retired a whole epilogue round trip ago)
      int nsg = 0, nm = seg_b[0], ng_left = 0, nbody0 = 0, n_load = 0;   // cursor of the NEXT group to load
      auto load_next = [&]() {
        while (ng_left == 0) {
          while (nsg < 2 && nm >= seg_e[nsg]) { ++nsg; if (nsg < 2) nm = seg_b[nsg]; }
          if (nsg >= 2) return;
          const Item nit = item_at(nm, seg_e[nsg]);
          nm += nit.len;
          ng_left = nit.ng; nbody0 = nit.body0;
        }
        const int st = n_load % kFuAtStages;
        if (n_load >= kFuAtStages) mbar_wait_backoff(&at_empty[st], ((uint32_t)(n_load / kFuAtStages) - 1u) & 1u, p.backoff);
        mbar_arrive_expect_tx(&at_full[st], kFuAtBytes);
        tma_load_2d(at_ring + st * kFuAtBytes, &tmapAt, &at_full[st], 0, nbody0 * 12);
        nbody0 += kFuGB; --ng_left; ++n_load;
      };
      if (kTwo) { pdl_wait(); load_next(); }
      WHMR_FU_FOR_ITEMS {
        const Item it = item_at(m, seg_e[sg]);
        m += it.len;
        const int vt = it.vt;
        if (vt != cur_vt) { mbar_wait_backoff(w_full, w_phase, p.backoff); w_phase ^= 1; cur_vt = vt; }
        const int ng = it.ng;
        for (int g = 0; g < ng; ++g) {
          if (kTwo) load_next();       // tile of the next group
          WHMR_FU_WAIT_R(&t_empty[ts], t_ph ^ 1, d_t);
          WHMR_FU_WAIT_R(&at_full[s], ph, d_at);
          const uint32_t d_tmem = tmem_base + (uint32_t)(TM::kT + ts * kFuTN);
          tcgen05_fence_after();
          const uint32_t a_hi = smem_u32(at_ring + s * kFuAtBytes), a_lo = a_hi + 64;
          for (int ks = 0; ks < p.jsteps; ++ks) {
            const uint64_t dW_hi = umma_desc_sw128(w_hi + ks * 32), dW_lo = umma_desc_sw128(w_lo + ks * 32);
            const uint64_t dA_hi = umma_desc_sw128(a_hi + ks * 32), dA_lo = umma_desc_sw128(a_lo + ks * 32);
            if (kFuWTmem) {
              const uint32_t tw_hi = tmem_base + (uint32_t)(kFuWCol + ks * 8), tw_lo = tw_hi + 16;
              umma_ta(d_tmem, tw_lo, dA_hi, idesc, ks != 0);
              umma_ta(d_tmem, tw_hi, dA_lo, idesc, 1u);
              umma_ta(d_tmem, tw_hi, dA_hi, idesc, 1u);
            } else {
              umma<0>(d_tmem, dW_lo, dA_hi, idesc, ks != 0);
              umma<0, kFuCollector ? 1 : 0>(d_tmem, dW_hi, dA_lo, idesc, 1u);
              umma<0, kFuCollector ? 3 : 0>(d_tmem, dW_hi, dA_hi, idesc, 1u);
            }
          }
          tcgen05_commit(&at_empty[s]);
          tcgen05_commit(&t_full[ts]);
          if (++s == kFuAtStages) { s = 0; ph ^= 1; }
          if (++ts == TM::kTStages) { ts = 0; t_ph ^= 1; }
        }
        const int next_vt = (m < seg_e[sg]) ? m / p.npv : ((sg == 0 && seg_b[1] < seg_e[1]) ? seg_b[1] / p.npv : -1);
        if (next_vt != vt) tcgen05_commit(w_empty);
      }
      if (dbgp) { long long* d = dbgp + blockIdx.x * 16; d[7] = d_t; d[8] = d_at; d[9] = clock64() - k0; }
    }
  } else {
    // ============================ epilogue =====================================================
    const int q = warp & 3;             // TMEM lane quarter
    const int w4 = (warp - 4) >> 2;     // which 2 bodies of every 8-body group (teams: team = w4 & 1, pair of the 4-body group = w4 >> 1)
    const int team = kFuTeams ? (w4 & 1) : 0;
    const int pair = kFuTeams ? (w4 >> 1) : w4;
    int gcount = 0;                     // teams: groups of earlier items (the skinning issuer alternates the accumulators over ALL groups)
    float* stg = stage_out + (warp - 4) * 192;   // two 96-float transpose buffers, one per body of the group
    int cur_vt = -1;
    float tx = 0.f, ty = 0.f, tz = 0.f;
    // read-out entries of this warp's 32 vertices (see skin_tc.cuh): lane l owns entries e0+l and e0+32+l
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
    int buf = 0, ts = kFuTeams ? team : 0; uint32_t bph = 0, t_ph = 0, w_par_e = 0;
    long long d_off = 0, d_t = 0, d_ld = 0, d_rel = 0;
#ifdef WHMR_FUSED_FINE_PROBES   // per-section cycles of one epilogue warp (math+stage | vertex stores | read-out emits)
    long long d_math = 0, d_st = 0, d_emit = 0;
#endif
    pdl_wait();      // outputs (and the read-out partial buffer) may still be in use by earlier kernels
    pdl_trigger();
    const long long k0 = dbgp ? clock64() : 0;
    WHMR_FU_FOR_ITEMS {
      const Item it = item_at(m, seg_e[sg]);
      m += it.len;
      const int vt = it.vt;
      if (vt != cur_vt) {
        const bool first_tile = cur_vt < 0;
        cur_vt = vt;
        const int v = vt * kTcM + q * 32 + lane;                  // < VP
        if (kFuWTmem && w4 == 0) {   // this lane-quarter's 32 weight rows -> TMEM columns kFuWCol .. kFuWCol+31
          uint32_t wr[32];
          const uint4* src = p.W16 + (size_t)v * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 t = __ldg(src + i);
            wr[i * 4 + 0] = t.x; wr[i * 4 + 1] = t.y; wr[i * 4 + 2] = t.z; wr[i * 4 + 3] = t.w;
          }
          if (!first_tile) { mbar_wait(w_empty, w_par_e); w_par_e ^= 1; }   // MMAs on the previous tile have retired
          tcgen05_fence_after();
          tmem_st_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kFuWCol, wr);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(w_full);
        }
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
      const int out_col = (vt * kTcM + q * 32) * 3 + lane;        // float index inside a body row
      const bool full_tile = (vt + 1) * kTcM <= p.V;
      const int ng = it.ng;
      WHMR_FU_WAIT(&off_full[buf], bph, d_off);
      const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
      const uint32_t off_addr = tmem_base + lane_sel + (uint32_t)(buf * TM::kOffStage);
      // teams: my groups of this item are those whose global number has my parity; the last of them releases the offsets
      const int g_last = !kFuTeams ? ng - 1 : ((((gcount + ng - 1) & 1) == team) ? ng - 1 : ng - 2);
      if (kFuTeams && g_last < 0) {        // a one-group item that belongs to the other team: nothing to read here
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&off_empty[buf]);
      }
      for (int g = 0; g < ng; ++g) {
        if (kFuTeams && (((gcount + g) & 1) != team)) continue;
        const int body_base = it.body0 + g * kFuGB + pair * 2;     // first of this warp's two bodies
        const int n_valid = min(2, p.nb - body_base);              // may be <= 0
        WHMR_FU_WAIT(&t_full[ts], t_ph, d_t);
        const uint32_t t_addr = tmem_base + lane_sel + (uint32_t)(TM::kT + ts * kFuTN + pair * 24);
        uint64_t* const t_rel = &t_empty[ts];
        if (kFuTeams) t_ph ^= 1;
        else if (++ts == TM::kTStages) { ts = 0; t_ph ^= 1; }
        tcgen05_fence_after();
        const long long l0 = dbgp ? clock64() : 0;
        uint32_t T[24], O[6];
        tmem_ld_32x32b_x16(t_addr, T);
        tmem_ld_32x32b_x8(t_addr + 16, T + 16);
        const uint32_t ocol = (uint32_t)(g * kFuGB + pair * 2);
        tmem_ld_32x32b_x2(off_addr + ocol, O);
        tmem_ld_32x32b_x2(off_addr + TM::kNB + ocol, O + 2);
        tmem_ld_32x32b_x2(off_addr + 2 * TM::kNB + ocol, O + 4);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const long long l1 = dbgp ? clock64() : 0;
        // accumulators are in registers: hand them back before the arithmetic and the stores
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(t_rel);
          if (g == g_last) mbar_arrive(&off_empty[buf]);
        }
        if (dbgp) { d_ld += l1 - l0; d_rel += clock64() - l1; }
        if (n_valid <= 0) continue;
        float* outp = p.verts + (size_t)body_base * V3 + out_col;
#ifdef WHMR_FUSED_FINE_PROBES
        const long long e0c = dbgp ? clock64() : 0;
#endif
        auto run = [&](auto guard_tag) {
          constexpr bool G = decltype(guard_tag)::value;
          // both bodies are staged before the single __syncwarp, so their shared-memory round trips overlap
          // T holds the two bodies' transforms interleaved (element k of body i in T[2k + i], see smpl_chain.cuh)
          if (!G) {   // both bodies valid: packed f32x2 arithmetic, same operation order as the scalar path
            const uint64_t t2x = f32x2_pack(tx, tx), t2y = f32x2_pack(ty, ty), t2z = f32x2_pack(tz, tz);
            const uint64_t px = f32x2_add(u32x2_pack(O[0], O[1]), t2x), py = f32x2_add(u32x2_pack(O[2], O[3]), t2y),
                           pz = f32x2_add(u32x2_pack(O[4], O[5]), t2z);
#define WHMR_T2(k) u32x2_pack(T[2 * (k)], T[2 * (k) + 1])
            const uint64_t rx = f32x2_fma(WHMR_T2(0), px, f32x2_fma(WHMR_T2(1), py, f32x2_fma(WHMR_T2(2), pz, WHMR_T2(3))));
            const uint64_t ry = f32x2_fma(WHMR_T2(4), px, f32x2_fma(WHMR_T2(5), py, f32x2_fma(WHMR_T2(6), pz, WHMR_T2(7))));
            const uint64_t rz = f32x2_fma(WHMR_T2(8), px, f32x2_fma(WHMR_T2(9), py, f32x2_fma(WHMR_T2(10), pz, WHMR_T2(11))));
#undef WHMR_T2
            float x0, x1, y0, y1, z0, z1;
            f32x2_unpack(rx, x0, x1); f32x2_unpack(ry, y0, y1); f32x2_unpack(rz, z0, z1);
            if (!(dbg_mode & 4)) {
              stg[lane * 3 + 0] = x0; stg[lane * 3 + 1] = y0; stg[lane * 3 + 2] = z0;
              stg[96 + lane * 3 + 0] = x1; stg[96 + lane * 3 + 1] = y1; stg[96 + lane * 3 + 2] = z1;
            } else if (x0 + y0 + z0 + x1 + y1 + z1 == 123.456f) stg[0] = x0;
          } else {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (G && i >= n_valid) continue;                      // warp-uniform
            const float px = __uint_as_float(O[i]) + tx, py = __uint_as_float(O[2 + i]) + ty,
                        pz = __uint_as_float(O[4 + i]) + tz;
#define WHMR_T(k) __uint_as_float(T[(k) * 2 + i])
            float rx = fmaf(WHMR_T(0), px, fmaf(WHMR_T(1), py, fmaf(WHMR_T(2), pz, WHMR_T(3))));
            float ry = fmaf(WHMR_T(4), px, fmaf(WHMR_T(5), py, fmaf(WHMR_T(6), pz, WHMR_T(7))));
            float rz = fmaf(WHMR_T(8), px, fmaf(WHMR_T(9), py, fmaf(WHMR_T(10), pz, WHMR_T(11))));
#undef WHMR_T
            if (G && has_transl) {
              const float* tr = p.transl + (size_t)(body_base + i) * 3;
              rx += tr[0]; ry += tr[1]; rz += tr[2];
            }
            float* sb = stg + i * 96;
            if (!(dbg_mode & 4)) { sb[lane * 3 + 0] = rx; sb[lane * 3 + 1] = ry; sb[lane * 3 + 2] = rz; }
            else if (rx + ry + rz == 123.456f) sb[0] = rx;
          }
          }
          __syncwarp();
#ifdef WHMR_FUSED_FINE_PROBES
          const long long e1c = dbgp ? clock64() : 0;
#endif
          if (!G && kFuStore64 && p.store64) {
            // full tile, both bodies: the 96 floats of a (warp, body) leave as 48 aligned float2 -- one 8-byte load + store per
            // lane and a second one on lanes 0..15 -- instead of three 4-byte pairs (a body row starts on an 8-byte boundary:
            // 82,680 bytes per body; 16-byte stores would be misaligned on odd bodies)
            float2 w0[2], w1[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float2* s2 = reinterpret_cast<const float2*>(stg + i * 96);
              w0[i] = (dbg_mode & 4) ? make_float2(0.f, 0.f) : s2[lane];
              w1[i] = (lane < 16 && !(dbg_mode & 4)) ? s2[32 + lane] : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              float2* o2 = reinterpret_cast<float2*>(outp - lane + (size_t)i * V3);
              if (!(dbg_mode & 1)) {
                o2[lane] = w0[i];
                if (lane < 16) o2[32 + lane] = w1[i];
              }
            }
          } else {
          float v[6];
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int r = 0; r < 3; ++r) v[i * 3 + r] = (dbg_mode & 4) ? 0.f : stg[i * 96 + r * 32 + lane];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (G && i >= n_valid) continue;
            float* ob = outp + (size_t)i * V3;
#pragma unroll
            for (int r = 0; r < 3; ++r)
              if ((!G || out_col + r * 32 < V3) && !(dbg_mode & 1)) ob[r * 32] = v[i * 3 + r];
          }
          }
#ifdef WHMR_FUSED_FINE_PROBES
          const long long e2c = dbgp ? clock64() : 0;
#endif
          if (n_e > 0 && !(dbg_mode & 2)) {   // fused read-outs: entries referencing one of this warp's 32 vertices
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (G && i >= n_valid) continue;
              const float* sb = stg + i * 96;
              const int bl = body_base + i;
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
#ifdef WHMR_FUSED_FINE_PROBES
          if (dbgp) { const long long e3c = clock64(); d_math += e1c - e0c; d_st += e2c - e1c; d_emit += e3c - e2c; }
#endif
        };
        if (n_valid == 2 && !has_transl && full_tile) run(cuda::std::false_type{}); else run(cuda::std::true_type{});
      }
      gcount += ng;
      if (++buf == TM::kOffStages) { buf = 0; bph ^= 1; }
    }
    if (dbgp && warp == 4 && lane == 0) { long long* d = dbgp + blockIdx.x * 16; d[10] = d_off; d[11] = d_t; d[12] = clock64() - k0; d[13] = d_ld; d[14] = d_rel;
#ifdef WHMR_FUSED_FINE_PROBES
      long long* d2 = dbgp + (gridDim.x + blockIdx.x) * 16; d2[2] = d_math; d2[3] = d_st; d2[4] = d_emit;
#endif
    }
  }

#undef WHMR_FU_FOR_ITEMS
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
  if (dbgp && threadIdx.x == 32) {
    unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    dbgp[(gridDim.x + blockIdx.x) * 16 + 1] = (long long)g;
  }
}

}  // namespace whmr
