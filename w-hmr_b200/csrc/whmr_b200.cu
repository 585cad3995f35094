// libwhmr_b200.so -- C ABI over the sm_100a kernels (see include/whmr_b200.h).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "backward.cuh"
#include "metrics.cuh"
#include "pose_blend_simt.cuh"
#include "pose_blend_tc.cuh"
#include "projection.cuh"
#include "readout.cuh"
#include "rotations.cuh"
#include "sampling.cuh"
#include "maf_fused_tc.cuh"
#include "skin_tc.cuh"
#include "skinning.cuh"
#include "smpl_fused_tc.cuh"
#include "smpl_chain.cuh"
#include "smpl_backward.cuh"

namespace whmr {

std::atomic<uint64_t> g_launch_count{0};

int pdl_mask() {   // default: the fused SMPL kernel (posedirs prefetch under the chain kernel) and the projections;
                   // the attribute on the read-out / chain / sampling launches measured as a loss (profiles/r01_notes.md)
  static const int mask = getenv("WHMR_PDL") ? atoi(getenv("WHMR_PDL")) : (kPdlFused | kPdlProject);
  return mask;
}

std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

// ---------------------------------------------------------------------------------------------
struct DeviceArena {   // owns the cudaMalloc'ed constants of one handle
  std::vector<void*> ptrs;
  ~DeviceArena() { for (void* p : ptrs) cudaFree(p); }
  template <typename T>
  cudaError_t upload(const std::vector<T>& h, T** out) {
    void* d = nullptr;
    const size_t bytes = std::max<size_t>(h.size() * sizeof(T), 16);
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) return e;
    ptrs.push_back(d);
    if (!h.empty()) {
      e = cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) return e;
    }
    *out = static_cast<T*>(d);
    return cudaSuccess;
  }
  cudaError_t alloc(size_t bytes, void** out) {
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) return e;
    ptrs.push_back(d);
    *out = d;
    return cudaSuccess;
  }
};

}  // namespace whmr

using namespace whmr;

// dense fp32 skinning weights [V, J] -> fp16 hi|lo rows [VP, 64] (hi in 0..31, lo in 32..63), zero padded
__global__ void split_weights_f16_kernel(const float* __restrict__ w, int V, int J, int VP, __half* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= VP * 32) return;
  const int v = i >> 5, j = i & 31;
  const float x = (v < V && j < J) ? w[(size_t)v * J + j] : 0.0f;
  const __half hi = __float2half_rn(x);
  out[(size_t)v * 64 + j] = hi;
  out[(size_t)v * 64 + 32 + j] = __float2half_rn(x - __half2float(hi));
}

struct whmr_smpl_s {
  SmplDevice d{};
  DeviceArena arena;
  int gemm_mode = WHMR_GEMM_FP32_SIMT;
  int chunk_bodies = 768;         // two-kernel path: keeps the [chunk, 3*VP] pose-offset intermediate inside the L2
  int chunk_bodies_fused = 4096;  // fused kernel: no intermediate; the chunk only bounds the read-out partial buffer
  bool chunk_from_env = false;
  TcPlan tc{};   // tensor maps etc. for the tcgen05 path
  cudaEvent_t probe_chain = nullptr, probe_blend = nullptr, probe_skin = nullptr;
  bool skin_tc = true;   // tensor-core skinning (WHMR_SKIN=simt selects the CUDA-core kernel)
  bool fused = true;     // one kernel for pose blend + skinning (WHMR_FUSED=0: GEMM kernel + skinning kernel)
  // host-buffer staging (whmr_smpl_reserve)
  int reserved_B = 0;
  float *st_betas = nullptr, *st_pose = nullptr, *st_verts = nullptr, *st_joints = nullptr;
  void* st_ws = nullptr;
  size_t st_ws_bytes = 0;
};

struct whmr_readout_s {
  DeviceArena arena;
  int R = 0, V = 0, J = 0;
  // row classes: one-hot rows (weight 1 on a single vertex: picks, markers, down-sampling), other
  // short rows (<= kShortRow non-zeros, one thread each) and long rows (one warp each)
  int n_onehot = 0, n_short = 0, n_long = 0, n_short_all = 0;
  int *row_ptr = nullptr, *col_idx = nullptr, *sub_row = nullptr, *rows_short = nullptr, *rows_long = nullptr;
  int *rows_short_all = nullptr;    // one-hot + short: used when the read-out runs stand-alone
  int *grp_prefix = nullptr, *grp_rows = nullptr;
  int *dst_ptr = nullptr, *dst_row = nullptr;   // vertex -> one-hot destination rows (CSC), [VP+1] / [n_onehot]
  int4* onehot_tab = nullptr;                   // [n_onehot] {src vertex, group prefix, group rows, row - prefix}
  // fused path (skinning epilogue emits, readout_reduce_kernel finishes)
  int* emit_grp_ptr = nullptr;                  // [VP/32 + 1]
  EmitEntry* emit_entries = nullptr;
  int* part_ptr = nullptr;                      // [R+1]
  int* rows_reduce = nullptr;                   // rows that are not vertex one-hots
  int *slot_of = nullptr, *jt_ptr = nullptr, *jt_col = nullptr;
  float* jt_val = nullptr;
  int n_partial = 0, n_reduce = 0;              // partial buffer [bodies, n_partial, 3] comes from the caller
  int n_terms = 0;                              // entries of slot_of (>= n_partial when repeated rows share slots)
  bool fusable = false;                         // partial array of one body fits in shared memory
  int dst_VP = 0;
  float* vals = nullptr;
  bool needs_joints = false;
};

extern "C" {

int whmr_abi_version(void) { return WHMR_ABI_VERSION; }
const char* whmr_last_error(void) { return last_error_ref().c_str(); }
uint64_t whmr_launch_count(void) { return g_launch_count.load(); }
void whmr_launch_count_reset(void) { g_launch_count.store(0); }
void whmr_debug_set_trap_buffer(int* host_mapped) { cudaMemcpyToSymbol(whmr::g_trap_buf, &host_mapped, sizeof(int*)); }

// =============================================================================================
// SMPL
// =============================================================================================
int whmr_smpl_create(const whmr_smpl_model_desc* m, int gemm_mode, whmr_smpl_t* out) {
  WHMR_CHECK_ARG(m && out, "whmr_smpl_create: null argument");
  WHMR_CHECK_ARG(m->v_template && m->shapedirs && m->posedirs && m->J_regressor && m->lbs_weights && m->parents,
                 "whmr_smpl_create: null model array");
  const int V = m->n_verts, J = m->n_joints, NB = m->n_betas;
  WHMR_CHECK_ARG(V > 0 && J >= 1 && J <= kMaxJoints && NB >= 0 && NB <= kMaxBetas,
                 "whmr_smpl_create: unsupported sizes V=%d J=%d n_betas=%d (J<=%d, n_betas<=%d)", V, J, NB,
                 kMaxJoints, kMaxBetas);
  WHMR_CHECK_ARG(m->parents[0] < 0, "whmr_smpl_create: parents[0] must be -1");
  for (int j = 1; j < J; ++j)
    WHMR_CHECK_ARG(m->parents[j] >= 0 && m->parents[j] < j, "whmr_smpl_create: parents[%d]=%lld not in [0,%d)", j,
                   (long long)m->parents[j], j);
  WHMR_CHECK_ARG(gemm_mode >= WHMR_GEMM_FP32_SIMT && gemm_mode <= WHMR_GEMM_TC_3XTF32,
                 "whmr_smpl_create: bad gemm_mode %d", gemm_mode);

  whmr_smpl_s* h = new (std::nothrow) whmr_smpl_s();
  if (!h) return set_error(WHMR_E_INVALID, "whmr_smpl_create: out of host memory");
  SmplDevice& d = h->d;
  d.V = V; d.J = J; d.NB = NB;
  d.VP = ceil_div(V, kVertTile) * kVertTile;
  d.NP = 3 * d.VP;
  const int nfeat = (J - 1) * 9;
  d.nfeat = nfeat;
  d.KP = std::max(16, ceil_div(nfeat + NB, 16) * 16);   // pose terms, then the NB shape coefficients, zero padded
  if (const char* e = getenv("WHMR_CHUNK_BODIES")) {
    const int c = atoi(e);
    if (c >= 8) { h->chunk_bodies = std::max(kTcBodyTile, c / kTcBodyTile * kTcBodyTile); h->chunk_from_env = true; }   // whole 256-body tiles
  }

  // kinematic tree depth
  std::vector<int> parents(J), depth(J, 0);
  int max_depth = 0;
  for (int j = 0; j < J; ++j) {
    parents[j] = (int)m->parents[j];
    depth[j] = j == 0 ? 0 : depth[parents[j]] + 1;
    max_depth = std::max(max_depth, depth[j]);
  }
  d.max_depth = max_depth;

  // pre-contracted rest-joint regressor (fp64 accumulation): J = J_template + J_shapedirs . beta
  std::vector<float> Jt((size_t)J * 3), Jd((size_t)J * 3 * std::max(NB, 1));
  for (int j = 0; j < J; ++j) {
    double t[3] = {0, 0, 0};
    std::vector<double> dd((size_t)3 * std::max(NB, 1), 0.0);
    const float* row = m->J_regressor + (size_t)j * V;
    for (int v = 0; v < V; ++v) {
      const double w = row[v];
      if (w == 0.0) continue;
      for (int c = 0; c < 3; ++c) {
        t[c] += w * m->v_template[(size_t)v * 3 + c];
        for (int k = 0; k < NB; ++k) dd[(size_t)c * NB + k] += w * m->shapedirs[((size_t)v * 3 + c) * NB + k];
      }
    }
    for (int c = 0; c < 3; ++c) {
      Jt[(size_t)j * 3 + c] = (float)t[c];
      for (int k = 0; k < NB; ++k) Jd[((size_t)j * 3 + c) * NB + k] = (float)dd[(size_t)c * NB + k];
    }
  }

  // planar padded per-vertex constants
  const int VP = d.VP;
  std::vector<float> vt((size_t)3 * VP, 0.f), sdp((size_t)3 * std::max(NB, 1) * VP, 0.f);
  for (int v = 0; v < V; ++v)
    for (int c = 0; c < 3; ++c) {
      vt[(size_t)c * VP + v] = m->v_template[(size_t)v * 3 + c];
      for (int k = 0; k < NB; ++k)
        sdp[((size_t)c * NB + k) * VP + v] = m->shapedirs[((size_t)v * 3 + c) * NB + k];
    }

  // ELL skinning weights (joint order ascending, like the dense matmul's summation order).  The
  // table width is rounded up to 4 or 8 (zero-weight padding) so the register-resident
  // specialisations of the skinning kernel apply; wider models take the generic path.
  int ell_k = 1;
  for (int v = 0; v < V; ++v) {
    int nnz = 0;
    for (int j = 0; j < J; ++j) nnz += m->lbs_weights[(size_t)v * J + j] != 0.0f;
    ell_k = std::max(ell_k, nnz);
  }
  if (ell_k <= 4) ell_k = 4; else if (ell_k <= 8) ell_k = 8;
  d.ell_k = ell_k;
  std::vector<int> eidx((size_t)ell_k * VP, 0);
  std::vector<float> ew((size_t)ell_k * VP, 0.f);
  for (int v = 0; v < V; ++v) {
    int k = 0;
    for (int j = 0; j < J; ++j) {
      const float w = m->lbs_weights[(size_t)v * J + j];
      if (w != 0.0f) { eidx[(size_t)k * VP + v] = j; ew[(size_t)k * VP + v] = w; ++k; }
    }
  }

  // posedirs, planar padded [KP, NP]
  std::vector<float> pp((size_t)d.KP * d.NP, 0.f);
  for (int k = 0; k < nfeat; ++k) {
    const float* src = m->posedirs + (size_t)k * V * 3;
    float* dst = pp.data() + (size_t)k * d.NP;
    for (int v = 0; v < V; ++v)
      for (int c = 0; c < 3; ++c) dst[(size_t)c * VP + v] = src[(size_t)v * 3 + c];
  }
  for (int k = 0; k < NB; ++k) {   // shape blend rides in the same contraction (rows nfeat..nfeat+NB-1)
    float* dst = pp.data() + (size_t)(nfeat + k) * d.NP;
    for (int v = 0; v < V; ++v)
      for (int c = 0; c < 3; ++c) dst[(size_t)c * VP + v] = m->shapedirs[((size_t)v * 3 + c) * NB + k];
  }
  // dense tf32 hi|lo skinning weights [2, VP, 32] for the tensor-core skinning kernel
  std::vector<float> wsplit((size_t)2 * VP * 32, 0.f);
  for (int v = 0; v < V; ++v)
    for (int j = 0; j < J; ++j) {
      const float w = m->lbs_weights[(size_t)v * J + j];
      const float hi = f32_to_tf32_rna(w);
      wsplit[((size_t)0 * VP + v) * 32 + j] = hi;
      wsplit[((size_t)1 * VP + v) * 32 + j] = f32_to_tf32_rna(w - hi);
    }

  cudaError_t e = cudaSuccess;
  auto up = [&](auto& vec, auto** dst) { if (e == cudaSuccess) e = h->arena.upload(vec, dst); };
  up(Jt, &d.J_template); up(Jd, &d.J_shapedirs); up(parents, &d.parents); up(depth, &d.depth);
  up(vt, &d.v_template_p); up(sdp, &d.shapedirs_p); up(eidx, &d.ell_idx); up(ew, &d.ell_w);
  up(pp, &d.posedirs_p);
  if (e != cudaSuccess) {
    delete h;
    return set_error(WHMR_E_CUDA, "whmr_smpl_create: device upload failed: %s", cudaGetErrorString(e));
  }
  // tensor-core operand (hi|lo split of posedirs, K-major) + TMA descriptors
  int rc = tc_plan_create(d, pp, h->arena, &h->tc);
  if (rc != WHMR_OK) { delete h; return rc; }
  {
    float* dw = nullptr;
    e = h->arena.upload(wsplit, &dw);
    if (e != cudaSuccess) { delete h; return set_error(WHMR_E_CUDA, "weight split upload failed: %s", cudaGetErrorString(e)); }
    h->tc.W_tf32 = dw;
    rc = tc_encode_rows32(h->tc.encode_fn, &h->tc.tmapW, dw, (size_t)VP, (size_t)VP * 128, kTcM);
    if (rc != WHMR_OK) { delete h; return rc; }
    e = cudaFuncSetAttribute(skin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkinSmem);
    if (e != cudaSuccess) { delete h; return set_error(WHMR_E_CUDA, "cudaFuncSetAttribute(skin_tc) failed: %s", cudaGetErrorString(e)); }
  }
  {   // fp16 hi|lo weights for the fused kernel, split on the device (exact round-to-nearest incl. subnormals)
    std::vector<float> wdense(m->lbs_weights, m->lbs_weights + (size_t)V * J);
    float* dwf = nullptr;
    void* dw16 = nullptr;
    e = h->arena.upload(wdense, &dwf);
    if (e == cudaSuccess) e = h->arena.alloc((size_t)VP * 64 * sizeof(__half), &dw16);
    if (e != cudaSuccess) { delete h; return set_error(WHMR_E_CUDA, "fp16 weight upload failed: %s", cudaGetErrorString(e)); }
    split_weights_f16_kernel<<<ceil_div(VP * 32, 256), 256>>>(dwf, V, J, VP, static_cast<__half*>(dw16));
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { delete h; return set_error(WHMR_E_CUDA, "split_weights_f16_kernel failed: %s", cudaGetErrorString(e)); }
    h->tc.W_f16 = dw16;
    rc = tc_encode_rows64h(h->tc.encode_fn, &h->tc.tmapW16, dw16, (size_t)VP, kTcM);
    if (rc != WHMR_OK) { delete h; return rc; }
  }
  if (const char* s = getenv("WHMR_SKIN")) h->skin_tc = strcmp(s, "simt") != 0;
  if (const char* s = getenv("WHMR_FUSED")) h->fused = atoi(s) != 0;
  e = cudaFuncSetAttribute(smpl_fused_tc_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FuTmem<4>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(smpl_fused_tc_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FuTmem<3>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(smpl_fused_tc_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FuTmem<8>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(smpl_fused_tc_kernel<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FuTmem<6>::kSmem);
  if (e != cudaSuccess) { delete h; return set_error(WHMR_E_CUDA, "cudaFuncSetAttribute(smpl_fused_tc) failed: %s", cudaGetErrorString(e)); }
  h->gemm_mode = gemm_mode;
  *out = h;
  return WHMR_OK;
}

int whmr_smpl_destroy(whmr_smpl_t h) {
  if (!h) return WHMR_OK;
  cudaFree(h->st_betas); cudaFree(h->st_pose); cudaFree(h->st_verts); cudaFree(h->st_joints); cudaFree(h->st_ws);
  delete h;
  return WHMR_OK;
}

int whmr_smpl_set_gemm_mode(whmr_smpl_t h, int gemm_mode) {
  WHMR_CHECK_ARG(h, "whmr_smpl_set_gemm_mode: null handle");
  WHMR_CHECK_ARG(gemm_mode >= WHMR_GEMM_FP32_SIMT && gemm_mode <= WHMR_GEMM_TC_3XTF32, "bad gemm_mode %d", gemm_mode);
  h->gemm_mode = gemm_mode;
  return WHMR_OK;
}

int whmr_smpl_get_info(whmr_smpl_t h, int32_t* n_verts, int32_t* n_joints, int32_t* n_betas, int32_t* ell_width,
                       int32_t* gemm_mode) {
  WHMR_CHECK_ARG(h, "whmr_smpl_get_info: null handle");
  if (n_verts) *n_verts = h->d.V;
  if (n_joints) *n_joints = h->d.J;
  if (n_betas) *n_betas = h->d.NB;
  if (ell_width) *ell_width = h->d.ell_k;
  if (gemm_mode) *gemm_mode = h->gemm_mode;
  return WHMR_OK;
}

// ---- workspace carving --------------------------------------------------------------------
static bool fused_applicable(const whmr_smpl_s* h);
static int effective_chunk(const whmr_smpl_s* h) {
  return (fused_applicable(h) && !h->chunk_from_env) ? h->chunk_bodies_fused : h->chunk_bodies;
}

static size_t carve(const whmr_smpl_s* h, int B, void* base, SmplWorkspace* ws) {
  const SmplDevice& d = h->d;
  const int chunk = std::min(B, effective_chunk(h));
  const int Bpad = ceil_div(std::max(B, 1), kTcBodyTile) * kTcBodyTile;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  const size_t oA = take((size_t)B * d.J * 12 * sizeof(float));
  const size_t oPf = take((size_t)B * d.KP * sizeof(float));
  const size_t oSplit = take((size_t)Bpad * 2 * d.KP * sizeof(float));   // sized for the tf32 variant
  const size_t oOff = take(fused_applicable(h) ? 16 : (size_t)chunk * d.NP * sizeof(float));   // unused by the fused kernel
  const size_t oAt = take((size_t)2 * Bpad * 12 * 32 * sizeof(float));
  const size_t oAt16 = take((size_t)Bpad * 12 * 64 * sizeof(__half));
  if (ws) {
    char* p = static_cast<char*>(base);
    ws->A = reinterpret_cast<float*>(p + oA);
    ws->pf = reinterpret_cast<float*>(p + oPf);
    ws->pf_split = p + oSplit;
    ws->offsets = reinterpret_cast<float*>(p + oOff);
    ws->At = reinterpret_cast<float*>(p + oAt);
    ws->At16 = p + oAt16;
    ws->At_part_stride = (size_t)Bpad * 12 * 32;
    ws->Bpad = Bpad;
    ws->chunk = chunk;
  }
  return off;
}

size_t whmr_smpl_workspace_bytes(whmr_smpl_t h, int B) {
  if (!h || B < 0) return 0;
  return carve(h, B, nullptr, nullptr) + 1024;
}

static int get_ws(whmr_smpl_t h, int B, void* workspace, size_t bytes, SmplWorkspace* ws) {
  WHMR_CHECK_ARG(h, "null SMPL handle");
  WHMR_CHECK_ARG(B >= 0, "negative batch %d", B);
  WHMR_CHECK_ARG(workspace || B == 0, "null workspace");
  const size_t need = whmr_smpl_workspace_bytes(h, B);
  if (bytes < need) return set_error(WHMR_E_WORKSPACE, "workspace too small: %zu < %zu bytes for B=%d", bytes, need, B);
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(workspace), 1024));
  carve(h, B, base, ws);
  return WHMR_OK;
}

static int chain_impl(whmr_smpl_t h, const float* betas, const float* pose, int pose_is_rotmat, const float* transl,
                      int B, float* joints, float* rel_transforms, void* workspace, size_t workspace_bytes, void* stream,
                      bool both_at_formats, const whmr_smpl_glue* glue = nullptr) {
  SmplWorkspace ws;
  int rc = get_ws(h, B, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  if (B == 0) return WHMR_OK;
  WHMR_CHECK_ARG(betas && pose, "whmr_smpl_stage_chain: null betas/pose");
  const SmplDevice& d = h->d;
  ChainParams p{};
  p.betas = betas; p.pose = pose; p.transl = transl; p.pose_is_rotmat = pose_is_rotmat;
  p.B = B; p.J = d.J; p.NB = d.NB; p.KP = d.KP; p.max_depth = d.max_depth;
  p.J_template = d.J_template; p.J_shapedirs = d.J_shapedirs; p.parents = d.parents; p.depth = d.depth;
  p.A = ws.A; p.joints = joints; p.A_user = rel_transforms;
  p.pf = h->gemm_mode == WHMR_GEMM_FP32_SIMT ? ws.pf : nullptr;
  p.pf_split = h->gemm_mode == WHMR_GEMM_TC_BF16X3 ? static_cast<__nv_bfloat16*>(ws.pf_split) : nullptr;
  p.pf_tf32 = h->gemm_mode == WHMR_GEMM_TC_3XTF32 ? static_cast<float*>(ws.pf_split) : nullptr;
  // the fused kernel reads A^T as fp16 hi|lo rows (same scratch region), the stand-alone skinning kernel as tf32 parts
  const bool f16 = fused_applicable(h);
  p.At = (h->skin_tc && (!f16 || both_at_formats)) ? ws.At : nullptr;
  p.At16 = f16 ? static_cast<__half*>(ws.At16) : nullptr;
  p.At_part_stride = ws.At_part_stride;
  if (glue) {
    WHMR_CHECK_ARG(!glue->gram_schmidt || pose_is_rotmat, "whmr_smpl_forward_regressor: gram_schmidt needs rotation-matrix input");
    p.gram_schmidt = glue->gram_schmidt; p.rotmat_out = glue->rotmat_out; p.pose_aa_out = glue->pose_aa_out;
    p.theta_out = glue->theta_out; p.cam = glue->cam; p.root_pose = glue->root_pose;
  }
  launch_pdl(kPdlChain, smpl_chain_kernel, dim3(ceil_div(B, kChainWarpsPerBlock)), dim3(kChainWarpsPerBlock * 32), 0, (cudaStream_t)stream, p);
  WHMR_LAUNCHED("smpl_chain_kernel");
  return WHMR_OK;
}

int whmr_smpl_stage_chain(whmr_smpl_t h, const float* betas, const float* pose, int pose_is_rotmat,
                          const float* transl, int B, float* joints, float* rel_transforms, void* workspace,
                          size_t workspace_bytes, void* stream) {
  // stage API: whmr_smpl_stage_skin (the stand-alone skinning kernel) may follow, so both A^T formats are written
  return chain_impl(h, betas, pose, pose_is_rotmat, transl, B, joints, rel_transforms, workspace, workspace_bytes, stream,
                    true);
}

// pose offsets for bodies [b0, b0+nb) -> ws.offsets[0..nb)
static int launch_pose_blend(whmr_smpl_t h, const SmplWorkspace& ws, int B, int b0, int nb, cudaStream_t st) {
  const SmplDevice& d = h->d;
  if (h->gemm_mode == WHMR_GEMM_FP32_SIMT) {
    dim3 grid(d.NP / kSimtBN, ceil_div(nb, kSimtBM));
    pose_blend_simt_kernel<<<grid, 256, 0, st>>>(ws.pf + (size_t)b0 * d.KP, d.posedirs_p, ws.offsets, nb, d.KP, d.NP);
    WHMR_LAUNCHED("pose_blend_simt_kernel");
    return WHMR_OK;
  }
  return tc_pose_blend_launch(h->tc, d, h->gemm_mode, ws.pf_split, B, b0, nb, ws.offsets, st);
}

// every row class of a read-out table for bodies [b0, b0+nb) in one launch
static int launch_readout_all(whmr_readout_t r, const float* verts, const float* joints, int nb, int B_total, int b0,
                              float* out, cudaStream_t st) {
  if (nb == 0 || r->R == 0) return WHMR_OK;
  ReadoutAllParams q{};
  ReadoutParams& p = q.rp;
  p.row_ptr = r->row_ptr; p.col_idx = r->col_idx; p.vals = r->vals; p.sub_row = r->sub_row;
  p.grp_prefix = r->grp_prefix; p.grp_rows = r->grp_rows;
  p.R = r->R; p.V = r->V; p.J = r->J; p.B = nb; p.B_total = B_total; p.b0 = b0;
  p.verts = verts; p.joints = joints; p.out = out;
  q.onehot_tab = r->onehot_tab; q.n_onehot = r->n_onehot;
  q.rows_long = r->rows_long; q.n_long = r->n_long;
  q.rows_short = r->rows_short; q.n_short = r->n_short;
  q.n_blocks_onehot = (int)(((long long)nb * r->n_onehot + 255) / 256);
  q.n_blocks_long = (int)(((long long)ceil_div(nb, kLongBodies) * r->n_long * 32 + 255) / 256);
  const int n_blocks_short = (int)(((long long)nb * r->n_short + 255) / 256);
  const int grid = q.n_blocks_onehot + q.n_blocks_long + n_blocks_short;
  if (grid == 0) return WHMR_OK;
  launch_pdl(kPdlReadout, readout_all_kernel, dim3(grid), dim3(256), 0, st, q);
  WHMR_LAUNCHED("readout_all_kernel");
  return WHMR_OK;
}

// fused path, second half: sum the partials the skinning epilogue emitted (+ joint-sourced terms, sub rows)
static int launch_readout_reduce(whmr_readout_t r, const float* verts, const float* joints, int nb, int B_total, int b0,
                                 const float* partial, float* out, cudaStream_t st) {
  if (nb == 0 || r->n_reduce == 0) return WHMR_OK;
  ReduceParams q{};
  ReadoutParams& p = q.rp;
  p.row_ptr = r->row_ptr; p.col_idx = r->col_idx; p.vals = r->vals; p.sub_row = r->sub_row;
  p.grp_prefix = r->grp_prefix; p.grp_rows = r->grp_rows;
  p.R = r->R; p.V = r->V; p.J = r->J; p.B = nb; p.B_total = B_total; p.b0 = b0;
  p.verts = verts; p.joints = joints; p.out = out;
  q.rows = r->rows_reduce; q.n_rows = r->n_reduce; q.part_ptr = r->part_ptr; q.partial = partial;
  q.n_partial = r->n_partial; q.n_terms = r->n_terms; q.slot_of = r->slot_of; q.jt_ptr = r->jt_ptr; q.jt_col = r->jt_col; q.jt_val = r->jt_val;
  const size_t smem = ((size_t)r->n_partial * 3 + r->n_terms) * sizeof(float);   // partials [n_partial,3] + slot list [n_terms]
  if (smem > 48 * 1024) WHMR_CUDA(ensure_dyn_smem(readout_reduce_kernel, (int)smem));   // the launching device may not be the creating one
  launch_pdl(kPdlReadout, readout_reduce_kernel, dim3(nb), dim3(128), smem, st, q);
  WHMR_LAUNCHED("readout_reduce_kernel");
  return WHMR_OK;
}

// skinning of bodies [b0, b0+nb); with `ro` the read-out entries are emitted by the same kernel
static int launch_skin(whmr_smpl_t h, const SmplWorkspace& ws, const float* betas, const float* transl, int B, int b0,
                       int nb, float* verts, whmr_readout_t ro, float* ro_out, float* ro_partial, cudaStream_t st) {
  const SmplDevice& d = h->d;
  if (h->skin_tc) {
    SkinTcParams p{};
    p.offsets = ws.offsets;
    p.v_template_p = d.v_template_p;
    p.transl = transl ? transl + (size_t)b0 * 3 : nullptr;
    p.verts = verts + (size_t)b0 * d.V * 3;
    if (ro) {   // the epilogue emits the read-out entries of each vertex (readout.cuh, path 2)
      p.emit.grp_ptr = ro->emit_grp_ptr; p.emit.entries = ro->emit_entries; p.emit.partial = ro_partial;
      p.emit.n_partial = ro->n_partial;
      p.ro_out = ro_out; p.ro_B = B; p.ro_b0 = b0;
    }
    p.nb = nb; p.V = d.V; p.VP = d.VP; p.NP = d.NP;
    p.n_groups = ceil_div(nb, kSkinGB);
    p.n_items = (d.VP / kTcM) * p.n_groups;
    p.ksteps = ceil_div(d.J, 8);
    CUtensorMap tmapAt;
    int rc = tc_encode_rows32(h->tc.encode_fn, &tmapAt, ws.At + (size_t)b0 * 12 * 32, (size_t)(ws.Bpad - b0) * 12,
                              ws.At_part_stride * sizeof(float), kSkinN);
    if (rc) return rc;
    // pose offsets: 3-D map {VP vertices, 3 planes, chunk bodies} over the planar [chunk, NP] buffer
    CUtensorMap tmapOff;
    {
      cuuint64_t dims[3] = {(cuuint64_t)d.VP, 3, (cuuint64_t)ws.chunk};
      cuuint64_t strides[2] = {(cuuint64_t)d.VP * 4, (cuuint64_t)d.NP * 4};
      cuuint32_t box[3] = {(cuuint32_t)kTcM, 3, (cuuint32_t)kSkinGB};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = reinterpret_cast<PFN_encodeTiled>(h->tc.encode_fn)(
          &tmapOff, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ws.offsets, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return set_error(WHMR_E_CUDA, "cuTensorMapEncodeTiled (offsets) failed with CUresult %d", (int)r);
    }
    const int grid = std::min(h->tc.num_sms, p.n_items);
    skin_tc_kernel<<<grid, kSkinThreads, kSkinSmem, st>>>(h->tc.tmapW, tmapAt, tmapOff, p);
    WHMR_LAUNCHED("skin_tc_kernel");
    return WHMR_OK;
  }
  SkinParams p{};
  p.offsets = ws.offsets;
  p.A = ws.A + (size_t)b0 * d.J * 12;
  p.betas = betas + (size_t)b0 * d.NB;
  p.transl = transl ? transl + (size_t)b0 * 3 : nullptr;
  p.v_template_p = d.v_template_p; p.shapedirs_p = d.shapedirs_p; p.ell_idx = d.ell_idx; p.ell_w = d.ell_w;
  p.verts = verts + (size_t)b0 * d.V * 3;
  p.B = nb; p.V = d.V; p.VP = d.VP; p.NP = d.NP; p.J = d.J; p.NB = d.NB; p.ell_k = d.ell_k;
  dim3 grid(d.VP / kVertTile, ceil_div(nb, kSkinBodies));
  const size_t smem = skin_smem_bytes(d.J);
  if (d.ell_k == 4) skin_kernel<4, 0><<<grid, kVertTile, smem, st>>>(p);
  else if (d.ell_k == 8) skin_kernel<8, 0><<<grid, kVertTile, smem, st>>>(p);
  else skin_kernel<0, 0><<<grid, kVertTile, smem, st>>>(p);
  WHMR_LAUNCHED("skin_kernel");
  return WHMR_OK;
}

// pose blend + skinning of bodies [b0, b0+nb) in one kernel (smpl_fused_tc.cuh); bf16x3 pose blend only
// (WHMR_GEMM_TC_3XTF32: the kind::tf32 instantiation, KP a multiple of the 32-element tf32 chunk; WHMR_FUSED_TF32=0 keeps
//  that mode on the two-kernel path)
static bool fused_applicable(const whmr_smpl_s* h) {
  static const bool tf32_ok = !(getenv("WHMR_FUSED_TF32") && atoi(getenv("WHMR_FUSED_TF32")) == 0);
  if (!(h->d.J <= 32 && h->fused && h->skin_tc && h->tc.ready)) return false;
  if (h->gemm_mode == WHMR_GEMM_TC_BF16X3) return h->d.KP % 16 == 0;
  if (h->gemm_mode == WHMR_GEMM_TC_3XTF32) return tf32_ok && h->d.KP % 32 == 0;
  return false;
}

static int launch_fused(whmr_smpl_t h, const SmplWorkspace& ws, const float* transl, int B, int b0, int nb, float* verts,
                        whmr_readout_t ro, float* ro_out, float* ro_partial, cudaStream_t st) {
  const SmplDevice& d = h->d;
  FusedParams p{};
  p.v_template_p = d.v_template_p;
  p.transl = transl ? transl + (size_t)b0 * 3 : nullptr;
  p.verts = verts + (size_t)b0 * d.V * 3;
  if (ro) {
    p.emit.grp_ptr = ro->emit_grp_ptr; p.emit.entries = ro->emit_entries; p.emit.partial = ro_partial;
    p.emit.n_partial = ro->n_partial;
    p.ro_out = ro_out; p.ro_B = B; p.ro_b0 = b0;
  }
  p.nb = nb; p.V = d.V; p.VP = d.VP;
  p.store64 = ((reinterpret_cast<size_t>(p.verts) & 7) == 0 && (3 * d.V) % 2 == 0) ? 1 : 0;   // any 4-byte aligned output is legal
  p.W16 = static_cast<const uint4*>(h->tc.W_f16);
  const int n_vtiles = d.VP / kTcM;
  p.npv = ceil_div(nb, 16);
  p.n_micro = n_vtiles * p.npv;
  const int kind = h->gemm_mode == WHMR_GEMM_TC_3XTF32 ? 1 : 0;   // pose-blend operands: bf16 hi|lo or tf32 hi|lo
  const int elem = kind ? 4 : 2;
  p.ksteps = d.KP * elem / 32;
  p.kch = ceil_div(p.ksteps, 4);
  p.jsteps = ceil_div(d.J, 16);
  CUtensorMap tmapPf, tmapAt;
  char* pf_base = static_cast<char*>(ws.pf_split) + (size_t)b0 * 2 * d.KP * elem;
  int rc = tc_encode(h->tc.encode_fn, &tmapPf, kind, pf_base, d.KP, ws.Bpad - b0, 16);
  if (rc) return rc;
  // merged copies (WHMR_FUSED_MERGED_TMA=1): hi and lo of an operand tile come with ONE TMA instruction, and the pose feature
  // of a chunk as one 64-row box instead of 2 x len 16-row boxes: 4 copies per K chunk instead of 6 + 2 * len.  Measured
  // neutral to -2 % in both arithmetics, with one or two issuers (profiles/r02_notes.md 4.8: the producer thread is not the
  // limit), so it is off by default.
  static const int merged_env = getenv("WHMR_FUSED_MERGED_TMA") ? atoi(getenv("WHMR_FUSED_MERGED_TMA")) : 0;
  CUtensorMap tmapPfBoth;
  bool merged = merged_env != 0 && h->tc.both_ok;
  rc = tc_encode_rows64h(h->tc.encode_fn, &tmapAt, static_cast<__half*>(ws.At16) + (size_t)b0 * 12 * 64,
                         (size_t)(ws.Bpad - b0) * 12, kFuTN);
  if (rc) return rc;
  // Item width (MAXM x 16 bodies behind one pass over a posedirs tile) -- measured, profiles/r01_notes.md section 5.
  // A pose-blend MMA costs the same ~70-100 cycles for 16 or 128 bodies (the 4 KB A tile enters the tensor core at
  // 64 B/clk: tools/microbench/umma_rate_bench.cu), which argues for wide items; but the wide plans (MAXM = 6, 8) have
  // a single pose-offset stage in TMEM, so pose blend and epilogue serialise, and the epilogue (issue-bound: ~170
  // instructions per warp and 8-body group) is the longer phase.  Measured: B=256 loop step 0.4075 ms (MAXM 8, aligned
  // halves, 108 CTAs) vs 0.3977 ms (MAXM 3); 16 k bodies 12.1 M bodies/s (MAXM 8) / 12.2 M (MAXM 6) vs 14.6 M (MAXM 4).
  // Default therefore: 48-body items with two blended-transform stages while a CTA's share is small, 64-body items
  // once posedirs re-streaming dominates; WHMR_FUSED_MAXM / WHMR_FUSED_SPLIT select the other plans.
  static const int maxm_env = getenv("WHMR_FUSED_MAXM") ? atoi(getenv("WHMR_FUSED_MAXM")) : 0;
  static const int split_env = getenv("WHMR_FUSED_SPLIT") ? atoi(getenv("WHMR_FUSED_SPLIT")) : -1;
  int maxm = p.n_micro <= 8 * h->tc.num_sms ? 3 : 4;
  if (maxm_env == 3 || maxm_env == 4 || maxm_env == 6 || maxm_env == 8) maxm = maxm_env;
  p.split = 0;
  p.pieces = 0;
  // Small batches, optional (WHMR_FUSED_PIECES=-1 automatic, k > 0 forced): at most two items per CTA (see the kernel).
  // Pieces of <= 4 micro-items (the 64-body plan), as many pieces per tile as two per CTA allow.  Measured neutral at
  // B=256 (382.6 vs 384.2 us per step): the epilogue, not the number of posedirs passes, sets a CTA's time; off by default.
  static const int pieces_env = getenv("WHMR_FUSED_PIECES") ? atoi(getenv("WHMR_FUSED_PIECES")) : 0;
  if (pieces_env != 0 && !maxm_env && split_env < 0) {
    const int k_min = ceil_div(p.npv, 4);
    const int k_one = h->tc.num_sms / n_vtiles, k_two = 2 * h->tc.num_sms / n_vtiles;
    int k = 0;
    if (k_min <= k_one) k = std::min(k_one, p.npv);
    else if (k_min <= k_two) k = std::min(k_two, p.npv);
    if (pieces_env > 0) k = pieces_env;
    if (k > 0 && ceil_div(p.npv, k) <= 4 && n_vtiles * k <= 2 * h->tc.num_sms) { p.pieces = k; maxm = 4; }
  }
  if (split_env >= 0) {
    p.split = split_env;
  } else if (!p.pieces && maxm >= 6 && p.npv <= 2 * maxm) {
    p.split = std::max(1, std::min(p.npv, h->tc.num_sms / n_vtiles));
  }
  if (p.split > 0 && (ceil_div(p.npv, p.split) > maxm || n_vtiles * p.split > h->tc.num_sms)) p.split = 0;
  const int grid = p.pieces > 0 ? std::min(h->tc.num_sms, n_vtiles * p.pieces)
                                : (p.split > 0 ? n_vtiles * p.split : std::min(h->tc.num_sms, p.n_micro));
  if (merged) {
    const int pf_rows = std::max(64, 16 * maxm);      // FuTmem<MAXM>::kPfPart / 128
    merged = tc_encode_both_parts(h->tc.encode_fn, &tmapPfBoth, kind, pf_base, d.KP, ws.Bpad - b0, pf_rows) == WHMR_OK;
  }
  p.merged = merged ? 1 : 0;
  const CUtensorMap& mapP = merged ? h->tc.tmapA_both[kind] : (kind ? h->tc.tmapA_tf32 : h->tc.tmapA_bf16);
  const CUtensorMap& mapPf = merged ? tmapPfBoth : tmapPf;
  static const bool dbg_on = getenv("WHMR_FUSED_DEBUG") != nullptr;
  static const int dbg_mode = getenv("WHMR_FUSED_DBGMODE") ? atoi(getenv("WHMR_FUSED_DBGMODE")) : 0;
  p.dbg_mode = dbg_mode;
  static const int backoff = getenv("WHMR_FUSED_BACKOFF") ? atoi(getenv("WHMR_FUSED_BACKOFF")) : 0;
  p.backoff = (unsigned)backoff;
  if (dbg_on) { cudaMalloc(&p.dbg, sizeof(long long) * 32 * grid); cudaMemsetAsync(p.dbg, 0, sizeof(long long) * 32 * grid, st); }
  const bool instrumented = dbg_on || dbg_mode != 0;
#define WHMR_FUSED_LAUNCH(M, D, T)                                                                                         \
  launch_pdl(kPdlFused, smpl_fused_tc_kernel<M, D, T>, dim3(grid), dim3(kFuThreads), FuTmem<M>::kSmem, st, mapP, \
             mapPf, h->tc.tmapW16, tmapAt, p)
  // two pose-blend issuing threads (smpl_fused_tc.cuh, kTwo): the 48- and 64-body plans only
  static const int issuers_env = getenv("WHMR_FUSED_ISSUERS") ? atoi(getenv("WHMR_FUSED_ISSUERS")) : 1;
  const bool two = issuers_env == 2 && (maxm == 3 || maxm == 4);
  if (kind == 1) {   // 3xTF32 pose blend: no instrumented build; two issuers for the 48- and 64-body plans only
    const int m = maxm;
#define WHMR_FUSED_LAUNCH_TF32(M, T)                                                                                       \
  do {                                                                                                                     \
    ensure_dyn_smem(smpl_fused_tc_kernel<M, false, T, 1>, FuTmem<M>::kSmem);                                              \
    launch_pdl(kPdlFused, smpl_fused_tc_kernel<M, false, T, 1>, dim3(grid), dim3(kFuThreads), FuTmem<M>::kSmem, st,       \
               mapP, mapPf, h->tc.tmapW16, tmapAt, p);                                                                   \
  } while (0)
    static const int issuers_tf32 = getenv("WHMR_FUSED_ISSUERS_TF32") ? atoi(getenv("WHMR_FUSED_ISSUERS_TF32")) : issuers_env;
    if (issuers_tf32 == 2 && m <= 4) { if (m == 3) WHMR_FUSED_LAUNCH_TF32(3, true); else WHMR_FUSED_LAUNCH_TF32(4, true); }
    else if (m == 3) WHMR_FUSED_LAUNCH_TF32(3, false); else if (m == 4) WHMR_FUSED_LAUNCH_TF32(4, false);
    else if (m == 6) WHMR_FUSED_LAUNCH_TF32(6, false); else WHMR_FUSED_LAUNCH_TF32(8, false);
#undef WHMR_FUSED_LAUNCH_TF32
  } else if (instrumented) {
    // the instrumented instantiations are only ever launched from here (per device)
    ensure_dyn_smem(smpl_fused_tc_kernel<3, true, false>, FuTmem<3>::kSmem);
    ensure_dyn_smem(smpl_fused_tc_kernel<4, true, false>, FuTmem<4>::kSmem);
    ensure_dyn_smem(smpl_fused_tc_kernel<6, true, false>, FuTmem<6>::kSmem);
    ensure_dyn_smem(smpl_fused_tc_kernel<8, true, false>, FuTmem<8>::kSmem);
    ensure_dyn_smem(smpl_fused_tc_kernel<3, true, true>, FuTmem<3>::kSmem);
    ensure_dyn_smem(smpl_fused_tc_kernel<4, true, true>, FuTmem<4>::kSmem);
    if (two) { if (maxm == 3) WHMR_FUSED_LAUNCH(3, true, true); else WHMR_FUSED_LAUNCH(4, true, true); }
    else if (maxm == 3) WHMR_FUSED_LAUNCH(3, true, false); else if (maxm == 4) WHMR_FUSED_LAUNCH(4, true, false);
    else if (maxm == 6) WHMR_FUSED_LAUNCH(6, true, false); else WHMR_FUSED_LAUNCH(8, true, false);
  } else if (two) {
    if (maxm == 3) { ensure_dyn_smem(smpl_fused_tc_kernel<3, false, true>, FuTmem<3>::kSmem); WHMR_FUSED_LAUNCH(3, false, true); }
    else { ensure_dyn_smem(smpl_fused_tc_kernel<4, false, true>, FuTmem<4>::kSmem); WHMR_FUSED_LAUNCH(4, false, true); }
  } else {
    if (maxm == 3) ensure_dyn_smem(smpl_fused_tc_kernel<3, false, false>, FuTmem<3>::kSmem);
    else if (maxm == 4) ensure_dyn_smem(smpl_fused_tc_kernel<4, false, false>, FuTmem<4>::kSmem);
    else if (maxm == 6) ensure_dyn_smem(smpl_fused_tc_kernel<6, false, false>, FuTmem<6>::kSmem);
    else ensure_dyn_smem(smpl_fused_tc_kernel<8, false, false>, FuTmem<8>::kSmem);
    if (maxm == 3) WHMR_FUSED_LAUNCH(3, false, false); else if (maxm == 4) WHMR_FUSED_LAUNCH(4, false, false);
    else if (maxm == 6) WHMR_FUSED_LAUNCH(6, false, false); else WHMR_FUSED_LAUNCH(8, false, false);
  }
#undef WHMR_FUSED_LAUNCH
  WHMR_LAUNCHED("smpl_fused_tc_kernel");
  if (p.dbg) {   // per-role wait/total cycles averaged over CTAs (debug only: synchronises)
    cudaStreamSynchronize(st);
    std::vector<long long> hd((size_t)32 * grid);
    cudaMemcpy(hd.data(), p.dbg, hd.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(p.dbg);
    double a[16] = {0};
    for (int c = 0; c < grid; ++c) for (int k = 0; k < 16; ++k) a[k] += (double)hd[(size_t)c * 16 + k] / grid;
    long long t_first = hd[(size_t)grid * 16], t_last_start = t_first, t_end_min = hd[(size_t)grid * 16 + 1], t_end_max = t_end_min;
    for (int c = 0; c < grid; ++c) {
      t_first = std::min(t_first, hd[(size_t)(grid + c) * 16]); t_last_start = std::max(t_last_start, hd[(size_t)(grid + c) * 16]);
      t_end_min = std::min(t_end_min, hd[(size_t)(grid + c) * 16 + 1]); t_end_max = std::max(t_end_max, hd[(size_t)(grid + c) * 16 + 1]);
    }
#ifdef WHMR_FUSED_FINE_PROBES
    {
      double m = 0, s2 = 0, e = 0;
      for (int c = 0; c < grid; ++c) { m += (double)hd[(size_t)(grid + c) * 16 + 2] / grid; s2 += (double)hd[(size_t)(grid + c) * 16 + 3] / grid; e += (double)hd[(size_t)(grid + c) * 16 + 4] / grid; }
      fprintf(stderr, "[whmr fused dbg] epilogue warp 4: math+stage %.0f | vertex stores %.0f | read-out emits %.0f cycles\n", m, s2, e);
    }
#endif
    fprintf(stderr, "[whmr fused dbg] globaltimer: CTA starts spread %lld ns, first start -> last end %lld ns, ends spread %lld ns, prologue avg %.0f ns\n",
            t_last_start - t_first, t_end_max - t_first, t_end_max - t_end_min, a[15]);
    fprintf(stderr, "[whmr fused dbg] nb=%d micro-items=%d grid=%d | pose-producer wait pf_empty %.0f a_empty %.0f of %.0f | "
            "pose-mma wait off_empty %.0f pf_full %.0f a_full %.0f of %.0f | skin-mma wait t_empty %.0f at_full %.0f of %.0f | "
            "epilogue wait off_full %.0f t_full %.0f tmem-ld %.0f release %.0f of %.0f cycles\n", nb, p.n_micro, grid, a[0], a[1], a[2], a[3], a[4],
            a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[13], a[14], a[12]);
  }
  return WHMR_OK;
}

int whmr_smpl_is_fused(whmr_smpl_t h) { return h && fused_applicable(h) ? 1 : 0; }

int whmr_smpl_stage_pose_blend(whmr_smpl_t h, int B, void* workspace, size_t workspace_bytes, void* stream) {
  SmplWorkspace ws;
  int rc = get_ws(h, B, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  if (B == 0) return WHMR_OK;
  WHMR_CHECK_ARG(!fused_applicable(h), "whmr_smpl_stage_pose_blend: the handle runs pose blend + skinning as one kernel "
                 "(no pose-offset intermediate); use whmr_smpl_forward, or WHMR_FUSED=0 / another gemm_mode for the stage API");
  WHMR_CHECK_ARG(B <= ws.chunk, "whmr_smpl_stage_pose_blend: B=%d exceeds the chunk size %d (stage calls are per chunk)", B,
                 ws.chunk);
  return launch_pose_blend(h, ws, B, 0, B, (cudaStream_t)stream);
}

int whmr_smpl_stage_skin(whmr_smpl_t h, const float* betas, int B, float* verts, void* workspace,
                         size_t workspace_bytes, void* stream) {
  SmplWorkspace ws;
  int rc = get_ws(h, B, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  if (B == 0) return WHMR_OK;
  WHMR_CHECK_ARG(betas && verts, "whmr_smpl_stage_skin: null betas/verts");
  WHMR_CHECK_ARG(!fused_applicable(h), "whmr_smpl_stage_skin: the handle runs pose blend + skinning as one kernel; use "
                 "whmr_smpl_forward, or WHMR_FUSED=0 / another gemm_mode for the stage API");
  WHMR_CHECK_ARG(B <= ws.chunk, "whmr_smpl_stage_skin: B=%d exceeds the chunk size %d", B, ws.chunk);
  return launch_skin(h, ws, betas, nullptr, B, 0, B, verts, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

size_t whmr_readout_workspace_bytes(whmr_readout_t ro, int n_bodies) {
  if (!ro || n_bodies <= 0 || !ro->fusable) return 0;
  return (size_t)n_bodies * ro->n_partial * 3 * sizeof(float) + 256;
}

int whmr_smpl_chunk_bodies(whmr_smpl_t h) { return h ? effective_chunk(h) : 0; }

int whmr_readout_finish(whmr_readout_t ro, const float* joints, int B, const void* ro_workspace, float* ro_out,
                        void* stream) {
  WHMR_CHECK_ARG(ro && B >= 0, "whmr_readout_finish: bad arguments");
  if (B == 0) return WHMR_OK;
  WHMR_CHECK_ARG(ro_workspace && ro_out, "whmr_readout_finish: null buffer");
  WHMR_CHECK_ARG(joints || !ro->needs_joints, "whmr_readout_finish: table references chain joints but joints == NULL");
  const float* partial = reinterpret_cast<const float*>(align_up(reinterpret_cast<size_t>(ro_workspace), 256));
  return launch_readout_reduce(ro, nullptr, joints, B, B, 0, partial, ro_out, (cudaStream_t)stream);
}

int whmr_readout_finish_multi(whmr_readout_t ro, int n_calls, const float* const* joints, const void* const* ro_workspaces,
                              float* const* ro_outs, int B, void* stream) {
  return whmr_readout_finish_project_multi(ro, n_calls, joints, ro_workspaces, ro_outs, B, nullptr, stream);
}

int whmr_readout_finish_project_multi(whmr_readout_t ro, int n_calls, const float* const* joints,
                                      const void* const* ro_workspaces, float* const* ro_outs, int B,
                                      const whmr_finish_projection* proj, void* stream) {
  WHMR_CHECK_ARG(ro && B >= 0 && n_calls >= 0 && n_calls <= 8, "whmr_readout_finish_multi: bad arguments (at most 8 calls)");
  if (proj) {
    WHMR_CHECK_ARG(proj->n_points >= 0 && proj->row0 >= 0 && proj->row0 + proj->n_points <= ro->R,
                   "whmr_readout_finish_project_multi: rows [%d, %d) outside the table (%d rows)", proj->row0,
                   proj->row0 + proj->n_points, ro->R);
    for (int i = 0; i < n_calls && proj->n_points > 0; ++i) {
      if (!proj->cam[i]) continue;
      WHMR_CHECK_ARG(proj->kp_weak[i], "whmr_readout_finish_project_multi: call %d has a camera but no kp_weak output", i);
      WHMR_CHECK_ARG(!proj->full[i] || (proj->kp_norm[i] && proj->bbox_height && proj->center && proj->orig_shape && proj->Tz),
                     "whmr_readout_finish_project_multi: call %d is full but lacks kp_norm / bbox inputs", i);
    }
    WHMR_CHECK_ARG(ro->n_reduce > 0 || proj->n_points == 0, "whmr_readout_finish_project_multi: table has no finishing pass");
  }
  if (B == 0 || n_calls == 0 || ro->n_reduce == 0) return WHMR_OK;
  WHMR_CHECK_ARG(joints && ro_workspaces && ro_outs, "whmr_readout_finish_multi: null pointer array");
  ReduceParams q{};
  ReadoutParams& p = q.rp;
  p.row_ptr = ro->row_ptr; p.col_idx = ro->col_idx; p.vals = ro->vals; p.sub_row = ro->sub_row;
  p.grp_prefix = ro->grp_prefix; p.grp_rows = ro->grp_rows;
  p.R = ro->R; p.V = ro->V; p.J = ro->J; p.B = B; p.B_total = B; p.b0 = 0;
  q.rows = ro->rows_reduce; q.n_rows = ro->n_reduce; q.part_ptr = ro->part_ptr;
  q.n_partial = ro->n_partial; q.n_terms = ro->n_terms; q.slot_of = ro->slot_of; q.jt_ptr = ro->jt_ptr; q.jt_col = ro->jt_col; q.jt_val = ro->jt_val;
  q.n_multi = n_calls;
  if (proj) q.pj = *proj;
  for (int i = 0; i < n_calls; ++i) {
    WHMR_CHECK_ARG(ro_workspaces[i] && ro_outs[i] && (joints[i] || !ro->needs_joints), "whmr_readout_finish_multi: null buffer");
    q.joints_m[i] = joints[i];
    q.partial_m[i] = reinterpret_cast<const float*>(align_up(reinterpret_cast<size_t>(ro_workspaces[i]), 256));
    q.out_m[i] = ro_outs[i];
  }
  const size_t smem = ((size_t)ro->n_partial * 3 + ro->n_terms) * sizeof(float);
  if (smem > 48 * 1024) WHMR_CUDA(ensure_dyn_smem(readout_reduce_kernel, (int)smem));
  launch_pdl(kPdlReadout, readout_reduce_kernel, dim3(B, n_calls), dim3(128), smem, (cudaStream_t)stream, q);
  WHMR_LAUNCHED("readout_reduce_kernel");
  return WHMR_OK;
}

int whmr_smpl_forward_readout(whmr_smpl_t h, const float* betas, const float* pose, int pose_is_rotmat,
                              const float* transl, int B, float* verts, float* joints, float* rel_transforms,
                              whmr_readout_t ro, float* ro_out, void* ro_workspace, size_t ro_workspace_bytes,
                              int defer_finish, int* finish_deferred, void* workspace, size_t workspace_bytes,
                              void* stream) {
  return whmr_smpl_forward_regressor(h, betas, pose, pose_is_rotmat, transl, B, verts, joints, rel_transforms, ro, ro_out,
                                     ro_workspace, ro_workspace_bytes, defer_finish, finish_deferred, nullptr, workspace,
                                     workspace_bytes, stream);
}

int whmr_smpl_forward_regressor(whmr_smpl_t h, const float* betas, const float* pose, int pose_is_rotmat,
                                const float* transl, int B, float* verts, float* joints, float* rel_transforms,
                                whmr_readout_t ro, float* ro_out, void* ro_workspace, size_t ro_workspace_bytes,
                                int defer_finish, int* finish_deferred, const whmr_smpl_glue* glue, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (finish_deferred) *finish_deferred = 0;
  SmplWorkspace ws;
  int rc = get_ws(h, B, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  if (B == 0) return WHMR_OK;
  WHMR_CHECK_ARG(betas && pose && verts, "whmr_smpl_forward: null betas/pose/verts");
  if (ro) {
    WHMR_CHECK_ARG(ro_out, "whmr_smpl_forward_readout: null read-out buffer");
    WHMR_CHECK_ARG(ro->V == h->d.V && ro->J == h->d.J, "whmr_smpl_forward_readout: read-out table built for V=%d J=%d", ro->V,
                   ro->J);
    WHMR_CHECK_ARG(joints || !ro->needs_joints, "whmr_smpl_forward_readout: table references chain joints but joints == NULL");
  }
  cudaStream_t st = (cudaStream_t)stream;
  rc = chain_impl(h, betas, pose, pose_is_rotmat, transl, B, joints, rel_transforms, workspace, workspace_bytes, stream,
                  false, glue);
  if (rc) return rc;
  if (h->probe_chain) WHMR_CUDA(cudaEventRecordWithFlags(h->probe_chain, st, cudaEventRecordExternal));
  // chunked so the [chunk, NP] pose-offset intermediate (and the chunk's vertices, for the read-outs)
  // stay L2-resident between the kernels
  // fused read-outs need the tensor-core skinning kernel and a caller-provided partial buffer covering one
  // chunk; otherwise the stand-alone gather kernel runs after the skinning kernel
  static const bool fuse_env = !(getenv("WHMR_FUSE_READOUT") && atoi(getenv("WHMR_FUSE_READOUT")) == 0);
  float* partial = ro_workspace ? reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(ro_workspace), 256)) : nullptr;
  const bool one_kernel = fused_applicable(h);
  const bool fused = ro && h->skin_tc && fuse_env && ro->fusable && ro->dst_VP == h->d.VP && partial &&
                     ro_workspace_bytes >= whmr_readout_workspace_bytes(ro, std::min(B, ws.chunk));
  // the finishing pass may be left to the caller (another stream) when the whole batch is one chunk
  const bool defer = fused && defer_finish && finish_deferred && B <= ws.chunk;
  if (defer) *finish_deferred = 1;
  for (int b0 = 0; b0 < B; b0 += ws.chunk) {
    const int nb = std::min(ws.chunk, B - b0);
    if (!one_kernel) {
      rc = launch_pose_blend(h, ws, B, b0, nb, st);
      if (rc) return rc;
    }
    if (h->probe_blend && b0 + nb >= B) WHMR_CUDA(cudaEventRecordWithFlags(h->probe_blend, st, cudaEventRecordExternal));
    rc = one_kernel ? launch_fused(h, ws, transl, B, b0, nb, verts, fused ? ro : nullptr, ro_out, partial, st)
                    : launch_skin(h, ws, betas, transl, B, b0, nb, verts, fused ? ro : nullptr, ro_out, partial, st);
    if (rc) return rc;
    if (h->probe_skin && b0 + nb >= B) WHMR_CUDA(cudaEventRecordWithFlags(h->probe_skin, st, cudaEventRecordExternal));
    if (ro) {
      const float* vch = verts + (size_t)b0 * h->d.V * 3;
      const float* jch = joints ? joints + (size_t)b0 * h->d.J * 3 : nullptr;
      if (defer) continue;
      rc = fused ? launch_readout_reduce(ro, vch, jch, nb, B, b0, partial, ro_out, st)
                 : launch_readout_all(ro, vch, jch, nb, B, b0, ro_out, st);
      if (rc) return rc;
    }
  }
  return WHMR_OK;
}

int whmr_smpl_forward(whmr_smpl_t h, const float* betas, const float* pose, int pose_is_rotmat, const float* transl,
                      int B, float* verts, float* joints, float* rel_transforms, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return whmr_smpl_forward_readout(h, betas, pose, pose_is_rotmat, transl, B, verts, joints, rel_transforms, nullptr,
                                   nullptr, nullptr, 0, 0, nullptr, workspace, workspace_bytes, stream);
}

int whmr_smpl_set_probe_events(whmr_smpl_t h, void* after_chain, void* after_pose_blend, void* after_skin) {
  WHMR_CHECK_ARG(h, "whmr_smpl_set_probe_events: null handle");
  h->probe_chain = (cudaEvent_t)after_chain;
  h->probe_blend = (cudaEvent_t)after_pose_blend;
  h->probe_skin = (cudaEvent_t)after_skin;
  return WHMR_OK;
}

int whmr_smpl_reserve(whmr_smpl_t h, int max_B) {
  WHMR_CHECK_ARG(h && max_B > 0, "whmr_smpl_reserve: bad arguments");
  if (max_B <= h->reserved_B) return WHMR_OK;
  cudaFree(h->st_betas); cudaFree(h->st_pose); cudaFree(h->st_verts); cudaFree(h->st_joints); cudaFree(h->st_ws);
  h->st_betas = h->st_pose = h->st_verts = h->st_joints = nullptr; h->st_ws = nullptr; h->reserved_B = 0;
  const SmplDevice& d = h->d;
  h->st_ws_bytes = whmr_smpl_workspace_bytes(h, max_B);
  WHMR_CUDA(cudaMalloc(&h->st_betas, (size_t)max_B * std::max(d.NB, 1) * sizeof(float)));
  WHMR_CUDA(cudaMalloc(&h->st_pose, (size_t)max_B * d.J * 9 * sizeof(float)));
  WHMR_CUDA(cudaMalloc(&h->st_verts, (size_t)max_B * d.V * 3 * sizeof(float)));
  WHMR_CUDA(cudaMalloc(&h->st_joints, (size_t)max_B * d.J * 3 * sizeof(float)));
  WHMR_CUDA(cudaMalloc(&h->st_ws, h->st_ws_bytes));
  h->reserved_B = max_B;
  return WHMR_OK;
}

int whmr_smpl_forward_host(whmr_smpl_t h, const float* betas, const float* pose, int pose_is_rotmat, int B,
                           float* verts, float* joints, void* stream) {
  WHMR_CHECK_ARG(h && betas && pose && verts, "whmr_smpl_forward_host: null argument");
  WHMR_CHECK_ARG(B > 0, "whmr_smpl_forward_host: B must be positive");
  if (B > h->reserved_B) {
    int rc = whmr_smpl_reserve(h, B);
    if (rc) return rc;
  }
  const SmplDevice& d = h->d;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t pose_elems = (size_t)B * d.J * (pose_is_rotmat ? 9 : 3);
  WHMR_CUDA(cudaMemcpyAsync(h->st_betas, betas, (size_t)B * d.NB * sizeof(float), cudaMemcpyHostToDevice, st));
  WHMR_CUDA(cudaMemcpyAsync(h->st_pose, pose, pose_elems * sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = whmr_smpl_forward(h, h->st_betas, h->st_pose, pose_is_rotmat, nullptr, B, h->st_verts, h->st_joints, nullptr,
                             h->st_ws, h->st_ws_bytes, stream);
  if (rc) return rc;
  WHMR_CUDA(cudaMemcpyAsync(verts, h->st_verts, (size_t)B * d.V * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (joints)
    WHMR_CUDA(cudaMemcpyAsync(joints, h->st_joints, (size_t)B * d.J * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  WHMR_CUDA(cudaStreamSynchronize(st));
  return WHMR_OK;
}

int whmr_batch_rodrigues(const float* aa, int n, float* R, void* stream) {
  WHMR_CHECK_ARG(n >= 0, "whmr_batch_rodrigues: negative n");
  if (n == 0) return WHMR_OK;
  WHMR_CHECK_ARG(aa && R, "whmr_batch_rodrigues: null pointer");
  rodrigues_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(aa, n, R);
  WHMR_LAUNCHED("rodrigues_kernel");
  return WHMR_OK;
}

#define WHMR_ROT_ENTRY(NAME, KERNEL)                                                              \
  int NAME(const float* in, int n, float* out, void* stream) {                                    \
    WHMR_CHECK_ARG(n >= 0, #NAME ": negative n");                                                 \
    if (n == 0) return WHMR_OK;                                                                   \
    WHMR_CHECK_ARG(in && out, #NAME ": null pointer");                                            \
    KERNEL<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(in, n, out);                       \
    WHMR_LAUNCHED(#KERNEL);                                                                       \
    return WHMR_OK;                                                                               \
  }
WHMR_ROT_ENTRY(whmr_rot6d_to_rotmat, rot6d_to_rotmat_kernel)
WHMR_ROT_ENTRY(whmr_unbiased_gram_schmidt, unbiased_gram_schmidt_kernel)
WHMR_ROT_ENTRY(whmr_rotmat_to_axis_angle, rotmat_to_axis_angle_kernel)
WHMR_ROT_ENTRY(whmr_batch_rodrigues_quat, batch_rodrigues_quat_kernel)
#undef WHMR_ROT_ENTRY

// =============================================================================================
// read-out
// =============================================================================================
int whmr_readout_create(int n_rows, int n_verts, int n_joints, const int32_t* row_ptr, const int32_t* col_idx,
                        const float* vals, const int32_t* sub_row, int n_groups, const int32_t* group_sizes,
                        whmr_readout_t* out) {
  WHMR_CHECK_ARG(out && row_ptr && n_rows >= 0 && n_verts > 0 && n_joints >= 0, "whmr_readout_create: bad arguments");
  WHMR_CHECK_ARG(n_groups >= 0 && (n_groups == 0 || group_sizes), "whmr_readout_create: bad groups");
  std::vector<int> gpre(n_rows, 0), grows(n_rows, n_rows);
  if (n_groups > 0) {
    int r = 0;
    for (int g = 0; g < n_groups; ++g) {
      WHMR_CHECK_ARG(group_sizes[g] >= 0 && r + group_sizes[g] <= n_rows, "whmr_readout_create: group sizes exceed n_rows");
      for (int k = 0; k < group_sizes[g]; ++k) { gpre[r + k] = r; grows[r + k] = group_sizes[g]; }
      r += group_sizes[g];
    }
    WHMR_CHECK_ARG(r == n_rows, "whmr_readout_create: group sizes sum to %d, expected %d", r, n_rows);
  }
  const int nnz = row_ptr[n_rows];
  WHMR_CHECK_ARG(row_ptr[0] == 0 && nnz >= 0, "whmr_readout_create: malformed row_ptr");
  WHMR_CHECK_ARG(nnz == 0 || (col_idx && vals), "whmr_readout_create: null col_idx/vals");
  bool needs_joints = false;
  for (int r = 0; r < n_rows; ++r) WHMR_CHECK_ARG(row_ptr[r + 1] >= row_ptr[r], "whmr_readout_create: row_ptr not monotone");
  for (int k = 0; k < nnz; ++k) {
    WHMR_CHECK_ARG(col_idx[k] >= 0 && col_idx[k] < n_verts + n_joints, "whmr_readout_create: col_idx[%d]=%d out of range", k,
                   col_idx[k]);
    needs_joints |= col_idx[k] >= n_verts;
  }
  std::vector<int> rp(row_ptr, row_ptr + n_rows + 1), ci(col_idx, col_idx + nnz), sr, rs, rl;
  std::vector<float> vv(vals, vals + nnz);
  if (sub_row) {
    sr.assign(sub_row, sub_row + n_rows);
    for (int r = 0; r < n_rows; ++r) WHMR_CHECK_ARG(sr[r] < n_rows, "whmr_readout_create: sub_row[%d] out of range", r);
  }
  // row classes
  std::vector<int> ro1, rs_all;
  std::vector<char> used_as_sub(n_rows, 0);
  if (sub_row) for (int r = 0; r < n_rows; ++r) if (sr[r] >= 0) used_as_sub[sr[r]] = 1;
  for (int r = 0; r < n_rows; ++r) {
    int len = rp[r + 1] - rp[r];
    const bool has_sub = sub_row && sr[r] >= 0;
    if (has_sub) len = std::max(len, rp[sr[r] + 1] - rp[sr[r]]);
    const bool onehot = !has_sub && !used_as_sub[r] && len == 1 && vv[rp[r]] == 1.0f;   // source: vertex or chain joint
    if (onehot) { ro1.push_back(r); rs_all.push_back(r); }
    else if (len <= kShortRow) { rs.push_back(r); rs_all.push_back(r); }
    else rl.push_back(r);
  }
  // vertex -> one-hot destination rows (CSC over the padded vertex range of the skinning kernel)
  const int VP = ceil_div(n_verts, kVertTile) * kVertTile;
  std::vector<int> dptr(VP + 1, 0), drow(ro1.size());
  for (int r : ro1) if (ci[rp[r]] < n_verts) dptr[ci[rp[r]] + 1]++;
  for (int v = 0; v < VP; ++v) dptr[v + 1] += dptr[v];
  {
    std::vector<int> fill(dptr.begin(), dptr.end() - 1);
    for (int r : ro1) if (ci[rp[r]] < n_verts) drow[fill[ci[rp[r]]]++] = r;
  }
  // ---- fused path tables: what the skinning epilogue emits per 32-vertex group, and the partial layout ----
  std::vector<int> part_ptr(n_rows + 1, 0), rows_reduce;
  std::vector<char> is_vert_onehot(n_rows, 0);
  for (int r : ro1) if (ci[rp[r]] < n_verts) is_vert_onehot[r] = 1;
  for (int r = 0; r < n_rows; ++r) {
    int nv = 0;
    if (!is_vert_onehot[r]) {
      for (int k = rp[r]; k < rp[r + 1]; ++k) nv += ci[k] < n_verts;
      rows_reduce.push_back(r);
    }
    part_ptr[r + 1] = part_ptr[r] + nv;
  }
  const int n_terms = part_ptr[n_rows];
  const int n_terms_pad = std::max((n_terms + 3) / 4 * 4, 4);
  const int n_g32 = VP / 32;
  // A row equal to an earlier regressor row (same columns, same values: e.g. kp_3d_h36m = h36m_j17[H36M_TO_J14],
  // models/whmr.py:176-180) shares that row's partial slots: the skinning epilogue emits every distinct term once.
  std::vector<int> alias_of(n_rows, -1);
  for (int r = 0; r < n_rows; ++r) {
    if (is_vert_onehot[r] || rp[r + 1] - rp[r] <= kShortRow) continue;
    for (int r0 = 0; r0 < r; ++r0) {
      if (is_vert_onehot[r0] || alias_of[r0] >= 0 || rp[r0 + 1] - rp[r0] != rp[r + 1] - rp[r]) continue;
      bool same = true;
      for (int k = 0; same && k < rp[r + 1] - rp[r]; ++k) same = ci[rp[r] + k] == ci[rp[r0] + k] && vv[rp[r] + k] == vv[rp[r0] + k];
      if (same) { alias_of[r] = r0; break; }
    }
  }
  std::vector<std::vector<EmitEntry>> grp0(n_g32), grp1(n_g32);   // kind 0 (one-hot) / kind 1 (regressor terms)
  std::vector<std::vector<int>> grp1_term(n_g32);                  // row-major term index of each kind-1 entry
  for (int r = 0; r < n_rows; ++r) {
    if (is_vert_onehot[r]) {
      const int v = ci[rp[r]];
      grp0[v / 32].push_back(EmitEntry{(v % 32), 1.0f, gpre[r], grows[r], r - gpre[r]});
    } else if (alias_of[r] < 0) {
      int t = part_ptr[r];
      for (int k = rp[r]; k < rp[r + 1]; ++k)
        if (ci[k] < n_verts) {
          grp1[ci[k] / 32].push_back(EmitEntry{(ci[k] % 32) | (1 << 8), vv[k], 0, 0, 0});
          grp1_term[ci[k] / 32].push_back(t++);
        }
    }
  }
  // emit order: per group the one-hot entries sorted by destination (neighbouring lanes -> neighbouring output
  // slots), then the regressor terms, which get consecutive slots of the partial buffer
  std::vector<int> emit_ptr(n_g32 + 1, 0), slot_of(n_terms_pad, 0);   // padded: staged with 16-byte loads
  std::vector<EmitEntry> emit_entries;
  int next_slot = 0;
  for (int g = 0; g < n_g32; ++g) {
    std::sort(grp0[g].begin(), grp0[g].end(), [](const EmitEntry& x, const EmitEntry& y) {
      return x.a != y.a ? x.a < y.a : x.d < y.d;
    });
    emit_entries.insert(emit_entries.end(), grp0[g].begin(), grp0[g].end());
    for (size_t i = 0; i < grp1[g].size(); ++i) {
      EmitEntry en = grp1[g][i];
      en.d = next_slot;
      slot_of[grp1_term[g][i]] = next_slot++;
      emit_entries.push_back(en);
    }
    emit_ptr[g + 1] = (int)emit_entries.size();
  }
  for (int r = 0; r < n_rows; ++r)      // repeated rows read the slots of the row they repeat
    if (alias_of[r] >= 0)
      for (int i = 0; i < part_ptr[r + 1] - part_ptr[r]; ++i) slot_of[part_ptr[r] + i] = slot_of[part_ptr[alias_of[r]] + i];
  const int n_partial = std::max((next_slot + 3) / 4 * 4, 4);   // slots per body; rows of the partial buffer start 16-byte aligned
  // joint-sourced terms per row (handled by the reduce kernel: the skinning kernel only sees vertices)
  std::vector<int> jt_ptr(n_rows + 1, 0), jt_col;
  std::vector<float> jt_val;
  for (int r = 0; r < n_rows; ++r) {
    for (int k = rp[r]; k < rp[r + 1]; ++k)
      if (ci[k] >= n_verts) { jt_col.push_back(ci[k] - n_verts); jt_val.push_back(vv[k]); }
    jt_ptr[r + 1] = (int)jt_col.size();
  }
  std::vector<int4> otab(ro1.size());
  for (size_t i = 0; i < ro1.size(); ++i) {
    const int r = ro1[i];
    otab[i] = make_int4(ci[rp[r]], gpre[r], grows[r], r - gpre[r]);
  }
  whmr_readout_s* h = new (std::nothrow) whmr_readout_s();
  if (!h) return set_error(WHMR_E_INVALID, "whmr_readout_create: out of host memory");
  h->R = n_rows; h->V = n_verts; h->J = n_joints;
  h->n_onehot = (int)ro1.size(); h->n_short = (int)rs.size(); h->n_long = (int)rl.size(); h->n_short_all = (int)rs_all.size();
  h->dst_VP = VP;
  h->needs_joints = needs_joints;
  cudaError_t e = cudaSuccess;
  auto up = [&](auto& vec, auto** dst) { if (e == cudaSuccess) e = h->arena.upload(vec, dst); };
  up(rp, &h->row_ptr); up(ci, &h->col_idx); up(vv, &h->vals); up(rs, &h->rows_short); up(rl, &h->rows_long);
  up(rs_all, &h->rows_short_all); up(dptr, &h->dst_ptr); up(drow, &h->dst_row); up(otab, &h->onehot_tab);
  up(gpre, &h->grp_prefix); up(grows, &h->grp_rows);
  if (sub_row) up(sr, &h->sub_row);
  up(emit_ptr, &h->emit_grp_ptr); up(emit_entries, &h->emit_entries); up(part_ptr, &h->part_ptr);
  up(rows_reduce, &h->rows_reduce); up(slot_of, &h->slot_of); up(jt_ptr, &h->jt_ptr); up(jt_col, &h->jt_col);
  up(jt_val, &h->jt_val);
  h->n_partial = n_partial; h->n_terms = n_terms_pad; h->n_reduce = (int)rows_reduce.size();
  const size_t reduce_smem = ((size_t)n_partial * 3 + n_terms_pad) * 4;
  h->fusable = reduce_smem <= 200 * 1024;   // one body's partial array + slot list must fit in shared memory
  if (h->fusable && reduce_smem > 48 * 1024) ensure_dyn_smem(readout_reduce_kernel, (int)reduce_smem);   // again at launch
  if (e != cudaSuccess) {
    delete h;
    return set_error(WHMR_E_CUDA, "whmr_readout_create: device upload failed: %s", cudaGetErrorString(e));
  }
  *out = h;
  return WHMR_OK;
}

int whmr_readout_destroy(whmr_readout_t r) { delete r; return WHMR_OK; }

int whmr_readout_apply(whmr_readout_t r, const float* verts, const float* joints, int B, float* out, void* stream) {
  WHMR_CHECK_ARG(r && B >= 0, "whmr_readout_apply: bad arguments");
  if (B == 0 || r->R == 0) return WHMR_OK;
  WHMR_CHECK_ARG(verts && out, "whmr_readout_apply: null verts/out");
  WHMR_CHECK_ARG(joints || !r->needs_joints, "whmr_readout_apply: table references chain joints but joints == NULL");
  return launch_readout_all(r, verts, joints, B, B, 0, out, (cudaStream_t)stream);
}

int whmr_gather_vertices(const float* verts, const int32_t* idx, int B, int V, int n_idx, float* out, void* stream) {
  WHMR_CHECK_ARG(B >= 0 && V > 0 && n_idx >= 0, "whmr_gather_vertices: bad sizes");
  if (B == 0 || n_idx == 0) return WHMR_OK;
  WHMR_CHECK_ARG(verts && idx && out, "whmr_gather_vertices: null pointer");
  const long long n = (long long)B * n_idx * 3;
  gather_vertices_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(verts, idx, B, V, n_idx, out);
  WHMR_LAUNCHED("gather_vertices_kernel");
  return WHMR_OK;
}

// =============================================================================================
// projection
// =============================================================================================
#define WHMR_BN_GRID(B, N) (unsigned)(((long long)(B) * (N) + 255) / 256)

int whmr_project_weak(const float* points, const float* cam, int B, int N, float focal, float img_w, float img_h,
                      float* out, void* stream) {
  WHMR_CHECK_ARG(B >= 0 && N >= 0, "whmr_project_weak: negative size");
  if (B == 0 || N == 0) return WHMR_OK;
  WHMR_CHECK_ARG(points && cam && out, "whmr_project_weak: null pointer");
  launch_pdl(kPdlProject, project_weak_kernel, dim3(WHMR_BN_GRID(B, N)), dim3(256), 0, (cudaStream_t)stream, points, cam, B, N, focal, img_w, img_h, out);
  WHMR_LAUNCHED("project_weak_kernel");
  return WHMR_OK;
}

int whmr_perspective_projection(const float* points, const float* rotation, int rot_batch, const float* translation,
                                const float* focal_dev, float focal_scalar, const float* camera_center,
                                const float* distortion, int B, int N, int retain_z, float* out, void* stream) {
  WHMR_CHECK_ARG(B >= 0 && N >= 0, "whmr_perspective_projection: negative size");
  if (B == 0 || N == 0) return WHMR_OK;
  WHMR_CHECK_ARG(points && camera_center && out, "whmr_perspective_projection: null pointer");
  WHMR_CHECK_ARG(rot_batch == 0 || rot_batch == 1 || rot_batch == B,
                 "whmr_perspective_projection: rotation batch %d must be 0, 1 or B=%d", rot_batch, B);
  WHMR_CHECK_ARG((rot_batch == 0) == (rotation == nullptr), "whmr_perspective_projection: rotation/rot_batch mismatch");
  perspective_projection_kernel<<<WHMR_BN_GRID(B, N), 256, 0, (cudaStream_t)stream>>>(
      points, rotation, rot_batch, translation, focal_dev, focal_scalar, camera_center, distortion, B, N, retain_z, out);
  WHMR_LAUNCHED("perspective_projection_kernel");
  return WHMR_OK;
}

int whmr_estimate_translation(const float* S, const float* joints_2d, int B, int N, int j0, float focal, float img_w,
                              float img_h, float* out, void* stream) {
  WHMR_CHECK_ARG(B >= 0 && N >= 0 && j0 >= 0 && j0 <= N, "whmr_estimate_translation: bad sizes");
  if (B == 0) return WHMR_OK;
  WHMR_CHECK_ARG(S && joints_2d && out, "whmr_estimate_translation: null pointer");
  launch_pdl(0, estimate_translation_kernel, dim3(ceil_div(B, 128)), dim3(128), 0, (cudaStream_t)stream, S, joints_2d, B, N,
             j0, focal, img_w, img_h, out);
  WHMR_LAUNCHED("estimate_translation_kernel");
  return WHMR_OK;
}

int whmr_project_full(const float* points, const float* cam, const float* bbox_height, const float* center,
                      const float* orig_shape, const float* Tz, int B, int N, float* kp_norm, float* kp_px,
                      float* focal_out, float* cam_t_out, void* stream) {
  WHMR_CHECK_ARG(B >= 0 && N >= 0, "whmr_project_full: negative size");
  if (B == 0 || N == 0) return WHMR_OK;
  WHMR_CHECK_ARG(points && cam && bbox_height && center && orig_shape && Tz, "whmr_project_full: null input");
  launch_pdl(kPdlProject, project_full_kernel, dim3(WHMR_BN_GRID(B, N)), dim3(256), 0, (cudaStream_t)stream, points, cam, bbox_height, center,
             orig_shape, Tz, B, N, kp_norm, kp_px, focal_out, cam_t_out, (float*)nullptr, 0.f, 0.f, 0.f);
  WHMR_LAUNCHED("project_full_kernel");
  return WHMR_OK;
}

int whmr_project_weak_full(const float* points, const float* cam, const float* bbox_height, const float* center,
                           const float* orig_shape, const float* Tz, int B, int N, float weak_focal, float weak_img_w,
                           float weak_img_h, float* kp_weak, float* kp_norm, float* kp_px, float* focal_out,
                           float* cam_t_out, void* stream) {
  WHMR_CHECK_ARG(B >= 0 && N >= 0, "whmr_project_weak_full: negative size");
  if (B == 0 || N == 0) return WHMR_OK;
  WHMR_CHECK_ARG(points && cam && bbox_height && center && orig_shape && Tz && kp_weak, "whmr_project_weak_full: null input");
  launch_pdl(kPdlProject, project_full_kernel, dim3(WHMR_BN_GRID(B, N)), dim3(256), 0, (cudaStream_t)stream, points, cam, bbox_height, center,
             orig_shape, Tz, B, N, kp_norm, kp_px, focal_out, cam_t_out, kp_weak, weak_focal, weak_img_w, weak_img_h);
  WHMR_LAUNCHED("project_full_kernel");
  return WHMR_OK;
}

int whmr_project_crop(const float* points, const float* cam, const float* center, const float* scale,
                      const float* img_focal, const float* img_center, const float* distortion, int B, int N,
                      float crop_size, float img_w, float img_h, float* full_out, float* crop_out, void* stream) {
  WHMR_CHECK_ARG(B >= 0 && N >= 0, "whmr_project_crop: negative size");
  if (B == 0 || N == 0) return WHMR_OK;
  WHMR_CHECK_ARG(points && cam && center && scale && img_focal && img_center, "whmr_project_crop: null input");
  project_crop_kernel<<<WHMR_BN_GRID(B, N), 256, 0, (cudaStream_t)stream>>>(
      points, cam, center, scale, img_focal, img_center, distortion, B, N, crop_size, img_w, img_h, full_out, crop_out);
  WHMR_LAUNCHED("project_crop_kernel");
  return WHMR_OK;
}

// =============================================================================================
// sampling
// =============================================================================================
int whmr_sample_bilinear(const float* feat, int layout, int B, int C, int H, int W, const float* points,
                         int points_shared, int N, float* out, void* stream) {
  const int pts_bstride = points_shared ? 0 : N * 2;
  WHMR_CHECK_ARG(B >= 0 && C >= 0 && N >= 0 && H > 0 && W > 0, "whmr_sample_bilinear: bad sizes");
  WHMR_CHECK_ARG(layout == WHMR_LAYOUT_NCHW || layout == WHMR_LAYOUT_NHWC, "whmr_sample_bilinear: bad layout %d", layout);
  if (B == 0 || C == 0 || N == 0) return WHMR_OK;
  WHMR_CHECK_ARG(feat && points && out, "whmr_sample_bilinear: null pointer");
  WHMR_CHECK_ARG((reinterpret_cast<size_t>(points) & 7) == 0, "whmr_sample_bilinear: points must be 8-byte aligned");
  WHMR_CHECK_ARG((long long)C * N < (1ll << 31) && B < 65536, "whmr_sample_bilinear: C*N or B too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (const int cgp = dense_channels(C, H, W, N, layout == WHMR_LAYOUT_NCHW)) {   // dense regime: whole map through shared memory
    WHMR_CUDA((layout == WHMR_LAYOUT_NCHW
                   ? launch_dense<true, false>(cgp, feat, points, pts_bstride, out, B, C, H, W, N, SampleProj{}, st)
                   : launch_dense<false, false>(cgp, feat, points, pts_bstride, out, B, C, H, W, N, SampleProj{}, st)));
    WHMR_LAUNCHED("sample_bilinear_dense_kernel");
    return WHMR_OK;
  }
  if (layout == WHMR_LAYOUT_NCHW) {
    if (const int cg = staged_channels(C, H, W, N)) {
      launch_staged<false>(feat, points, pts_bstride, out, B, C, H, W, N, cg, SampleProj{}, st);
      WHMR_LAUNCHED("sample_bilinear_nchw_staged_kernel");
      return WHMR_OK;
    }
    dim3 grid(ceil_div(C * N, 256 * kSampleItems), B);
    launch_pdl(kPdlSample, sample_bilinear_nchw_kernel<false>, grid, dim3(256), 0, st, feat, points, pts_bstride, out, C, H, W, N, SampleProj{});
    WHMR_LAUNCHED("sample_bilinear_nchw_kernel");
  } else {
    dim3 grid(ceil_div(N, 32), ceil_div(C, 64), B);
    launch_pdl(kPdlSample, sample_bilinear_nhwc_kernel<false>, grid, dim3(256), 0, st, feat, points, pts_bstride, out, C, H, W, N, SampleProj{});
    WHMR_LAUNCHED("sample_bilinear_nhwc_kernel");
  }
  return WHMR_OK;
}

int whmr_project_sample(const float* feat, int layout, int B, int C, int H, int W, const float* p, const float* cam,
                        int N, float focal, float img_w, float img_h, float* points2d_out, float* out, void* stream) {
  if (layout == WHMR_LAYOUT_NCHW) {   // one launch: the grid coordinate is computed per output element
    WHMR_CHECK_ARG(B >= 0 && C >= 0 && N >= 0 && H > 0 && W > 0, "whmr_project_sample: bad sizes");
    if (B == 0 || C == 0 || N == 0) return WHMR_OK;
    WHMR_CHECK_ARG(feat && p && cam && out, "whmr_project_sample: null pointer");
    WHMR_CHECK_ARG((long long)C * N < (1ll << 31) && B < 65536, "whmr_project_sample: C*N or B too large");
    WHMR_CHECK_ARG(!points2d_out || (reinterpret_cast<size_t>(points2d_out) & 7) == 0, "whmr_project_sample: points2d_out must be 8-byte aligned");
    SampleProj pj{cam, focal, img_w, img_h, points2d_out};
    if (const int cgp = dense_channels(C, H, W, N, 1)) {
      WHMR_CUDA((launch_dense<true, true>(cgp, feat, p, N * 3, out, B, C, H, W, N, pj, (cudaStream_t)stream)));
      WHMR_LAUNCHED("sample_bilinear_dense_kernel<project>");
      return WHMR_OK;
    }
    if (const int cg = staged_channels(C, H, W, N)) {
      launch_staged<true>(feat, p, N * 3, out, B, C, H, W, N, cg, pj, (cudaStream_t)stream);
      WHMR_LAUNCHED("sample_bilinear_nchw_staged_kernel<project>");
      return WHMR_OK;
    }
    dim3 grid(ceil_div(C * N, 256 * kSampleItems), B);
    launch_pdl(kPdlSample, sample_bilinear_nchw_kernel<true>, grid, dim3(256), 0, (cudaStream_t)stream, feat, p, N * 3, out, C, H, W, N, pj);
    WHMR_LAUNCHED("sample_bilinear_nchw_kernel<project>");
    return WHMR_OK;
  }
  WHMR_CHECK_ARG(layout == WHMR_LAYOUT_NHWC, "whmr_project_sample: bad layout %d", layout);
  WHMR_CHECK_ARG(B >= 0 && C >= 0 && N >= 0 && H > 0 && W > 0 && B < 65536, "whmr_project_sample: bad sizes");
  if (B == 0 || C == 0 || N == 0) return WHMR_OK;
  WHMR_CHECK_ARG(feat && p && cam && out, "whmr_project_sample: null pointer");
  WHMR_CHECK_ARG(!points2d_out || (reinterpret_cast<size_t>(points2d_out) & 7) == 0, "whmr_project_sample: points2d_out must be 8-byte aligned");
  SampleProj pj{cam, focal, img_w, img_h, points2d_out};
  if (const int cgp = dense_channels(C, H, W, N, 0)) {
    WHMR_CUDA((launch_dense<false, true>(cgp, feat, p, N * 3, out, B, C, H, W, N, pj, (cudaStream_t)stream)));
    WHMR_LAUNCHED("sample_bilinear_dense_kernel<project>");
    return WHMR_OK;
  }
  dim3 grid(ceil_div(N, 32), ceil_div(C, 64), B);
  launch_pdl(kPdlSample, sample_bilinear_nhwc_kernel<true>, grid, dim3(256), 0, (cudaStream_t)stream, feat, p, N * 3, out, C, H, W,
             N, pj);
  WHMR_LAUNCHED("sample_bilinear_nhwc_kernel<project>");
  return WHMR_OK;
}

// =============================================================================================
// sampling + reduce_dim MLP in one kernel (maf_fused_tc.cuh)
// =============================================================================================
struct whmr_maf_mlp_s {
  MafDims d{};
  int nx = 0, dev = 0, num_sms = 148;
  float *wx = nullptr, *w1y = nullptr, *w2y = nullptr, *bias = nullptr;
  CUtensorMap map_x, map_1, map_2;
  bool weights_set = false;
  DeviceArena arena;
};

int whmr_maf_mlp_create(int c_in, int c1, int c2, int c3, whmr_maf_mlp_t* out) {
  WHMR_CHECK_ARG(out, "whmr_maf_mlp_create: null output");
  WHMR_CHECK_ARG(c_in >= 64 && c_in % 64 == 0 && c1 >= 64 && c1 % 64 == 0 && c2 >= 64 && c2 % 64 == 0 && c3 >= 16 &&
                     c3 % 16 == 0 && c1 + c2 + c3 <= kMafMaxOut,
                 "whmr_maf_mlp_create: unsupported widths %d -> %d -> %d -> %d (need C_in, C1, C2 multiples of 64, C3 a "
                 "multiple of 16, C1+C2+C3 <= %d)", c_in, c1, c2, c3, kMafMaxOut);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return set_error(WHMR_E_CUDA, "cuTensorMapEncodeTiled unavailable: %s", cudaGetErrorString(e));
  whmr_maf_mlp_s* m = new whmr_maf_mlp_s();
  m->d = MafDims{c_in, c1, c2, c3};
  m->nx = c1 + c2 + c3;
  cudaGetDevice(&m->dev);
  cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, m->dev);
  void *a = nullptr, *b = nullptr, *c = nullptr, *bi = nullptr;
  e = m->arena.alloc((size_t)2 * m->nx * c_in * 4, &a);
  if (e == cudaSuccess) e = m->arena.alloc((size_t)2 * c2 * c1 * 4, &b);
  if (e == cudaSuccess) e = m->arena.alloc((size_t)2 * c3 * c2 * 4, &c);
  if (e == cudaSuccess) e = m->arena.alloc((size_t)kMafMaxOut * 4, &bi);
  if (e != cudaSuccess) { delete m; return set_error(WHMR_E_CUDA, "whmr_maf_mlp_create: cudaMalloc failed: %s", cudaGetErrorString(e)); }
  m->wx = (float*)a; m->w1y = (float*)b; m->w2y = (float*)c; m->bias = (float*)bi;
  int rc = maf_encode_weights(fn, &m->map_x, m->wx, c_in, m->nx);
  if (!rc) rc = maf_encode_weights(fn, &m->map_1, m->w1y, c1, c2);
  if (!rc) rc = maf_encode_weights(fn, &m->map_2, m->w2y, c2, c3);
  if (rc) { delete m; return rc; }
  *out = m;
  return WHMR_OK;
}

int whmr_maf_mlp_destroy(whmr_maf_mlp_t m) { delete m; return WHMR_OK; }

int whmr_maf_mlp_set_weights(whmr_maf_mlp_t m, const float* w0, const float* b0, const float* w1, const float* b1,
                             const float* w2, const float* b2, void* stream) {
  WHMR_CHECK_ARG(m && w0 && w1 && w2, "whmr_maf_mlp_set_weights: null handle or weight pointer");
  const int n = m->nx * m->d.c0 + m->d.c2 * m->d.c1 + m->d.c3 * m->d.c2 + m->nx;
  maf_split_weights_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w0, b0, w1, b1, w2, b2, m->d, m->wx, m->w1y,
                                                                                m->w2y, m->bias);
  WHMR_LAUNCHED("maf_split_weights_kernel");
  m->weights_set = true;
  return WHMR_OK;
}

extern "C++" {
template <bool kProject>
static int maf_launch(whmr_maf_mlp_t m, const float* feat, int layout, int B, int H, int W, const float* points,
                      int pts_bstride, int N, float* out, float* pf_out, SampleProj pj, cudaStream_t st, const char* who) {
  WHMR_CHECK_ARG(m, "%s: null handle", who);
  WHMR_CHECK_ARG(m->weights_set, "%s: whmr_maf_mlp_set_weights has not been called", who);
  WHMR_CHECK_ARG(B >= 0 && N >= 0 && H > 0 && W > 0, "%s: bad sizes", who);
  WHMR_CHECK_ARG(layout == WHMR_LAYOUT_NCHW || layout == WHMR_LAYOUT_NHWC, "%s: bad layout %d", who, layout);
  if (B == 0 || N == 0) return WHMR_OK;
  WHMR_CHECK_ARG(feat && points && out, "%s: null pointer", who);
  WHMR_CHECK_ARG((long long)B * N < (1ll << 31) - kMafRows, "%s: B*N too large", who);
  WHMR_CHECK_ARG(layout == WHMR_LAYOUT_NCHW || (reinterpret_cast<size_t>(feat) & 15) == 0, "%s: NHWC maps must be 16-byte aligned", who);
  WHMR_CHECK_ARG(kProject || (reinterpret_cast<size_t>(points) & 7) == 0, "%s: points must be 8-byte aligned", who);
  const int n_tiles = (int)(((long long)B * N + kMafRows - 1) / kMafRows);
  const int grid = std::min(n_tiles, m->num_sms);
  const size_t smem = maf_smem_bytes(m->d, layout == WHMR_LAYOUT_NHWC);
#define WHMR_MAF_LAUNCH(L)                                                                                         \
  do {                                                                                                             \
    cudaError_t e_ = ensure_dyn_smem(maf_fused_kernel<L, kProject>, (int)smem);                                    \
    if (e_ != cudaSuccess) return set_error(WHMR_E_CUDA, "%s: cudaFuncSetAttribute failed: %s", who, cudaGetErrorString(e_)); \
    launch_pdl(kPdlSample, maf_fused_kernel<L, kProject>, dim3(grid), dim3(kMafThreads), smem, st, m->map_x, m->map_1, \
               m->map_2, feat, points, pts_bstride, (const float*)m->bias, out, pf_out, B, N, H, W, m->d, n_tiles,     \
               maf_w_stages(m->d), pj);                                                                            \
  } while (0)
  if (layout == WHMR_LAYOUT_NCHW) WHMR_MAF_LAUNCH(0); else WHMR_MAF_LAUNCH(1);
#undef WHMR_MAF_LAUNCH
  WHMR_LAUNCHED("maf_fused_kernel");
  return WHMR_OK;
}
}  // extern "C++"

int whmr_sample_reduce(whmr_maf_mlp_t m, const float* feat, int layout, int B, int H, int W, const float* points,
                       int points_shared, int N, float* mesh_align_out, float* point_feat_out, void* stream) {
  return maf_launch<false>(m, feat, layout, B, H, W, points, points_shared ? 0 : N * 2, N, mesh_align_out, point_feat_out,
                           SampleProj{}, (cudaStream_t)stream, "whmr_sample_reduce");
}

int whmr_project_sample_reduce(whmr_maf_mlp_t m, const float* feat, int layout, int B, int H, int W, const float* p,
                               const float* cam, int N, float focal, float img_w, float img_h, float* points2d_out,
                               float* mesh_align_out, float* point_feat_out, void* stream) {
  WHMR_CHECK_ARG(B == 0 || N == 0 || cam, "whmr_project_sample_reduce: null camera");
  WHMR_CHECK_ARG(!points2d_out || (reinterpret_cast<size_t>(points2d_out) & 7) == 0,
                 "whmr_project_sample_reduce: points2d_out must be 8-byte aligned");
  return maf_launch<true>(m, feat, layout, B, H, W, p, N * 3, N, mesh_align_out, point_feat_out,
                          SampleProj{cam, focal, img_w, img_h, points2d_out}, (cudaStream_t)stream,
                          "whmr_project_sample_reduce");
}

// =============================================================================================
// metrics
// =============================================================================================
int whmr_joint_errors(const float* pred, const float* gt, int n, int J, float* mpjpe, float* pa_mpjpe, void* stream) {
  WHMR_CHECK_ARG(n >= 0 && J > 0 && J <= kMaxEvalJoints, "whmr_joint_errors: bad sizes n=%d J=%d", n, J);
  if (n == 0) return WHMR_OK;
  WHMR_CHECK_ARG(pred && gt, "whmr_joint_errors: null pointer");
  joint_errors_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(pred, gt, n, J, mpjpe, pa_mpjpe);
  WHMR_LAUNCHED("joint_errors_kernel");
  return WHMR_OK;
}

int whmr_vertex_errors(const float* pred, const float* gt, int n, int V, float* pve, void* stream) {
  WHMR_CHECK_ARG(n >= 0 && V > 0, "whmr_vertex_errors: bad sizes n=%d V=%d", n, V);
  if (n == 0) return WHMR_OK;
  WHMR_CHECK_ARG(pred && gt && pve, "whmr_vertex_errors: null pointer");
  vertex_errors_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(pred, gt, V, pve);
  WHMR_LAUNCHED("vertex_errors_kernel");
  return WHMR_OK;
}

// =============================================================================================
// backward
// =============================================================================================
namespace {
struct BwdCarve {
  float *A, *pf, *offsets, *g_off, *g_A, *g_pf;
  void* pf_split;
  int Bpad, chunk;
};
size_t carve_backward(const whmr_smpl_s* h, int B, void* base, BwdCarve* c) {
  const SmplDevice& d = h->d;
  const int chunk = std::min(B, h->chunk_bodies);
  const int Bpad = ceil_div(std::max(B, 1), kTcBodyTile) * kTcBodyTile;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  const size_t oA = take((size_t)B * d.J * 12 * sizeof(float));
  const size_t oPf = take((size_t)B * d.KP * sizeof(float));
  const size_t oSplit = take((size_t)Bpad * 2 * d.KP * sizeof(float));
  const size_t oOff = take((size_t)chunk * d.NP * sizeof(float));
  const size_t oGoff = take((size_t)chunk * d.NP * sizeof(float));
  const size_t oGA = take((size_t)B * d.J * 12 * sizeof(float));
  const size_t oGpf = take((size_t)B * d.KP * sizeof(float));
  if (c) {
    char* p = static_cast<char*>(base);
    c->A = reinterpret_cast<float*>(p + oA); c->pf = reinterpret_cast<float*>(p + oPf); c->pf_split = p + oSplit;
    c->offsets = reinterpret_cast<float*>(p + oOff); c->g_off = reinterpret_cast<float*>(p + oGoff);
    c->g_A = reinterpret_cast<float*>(p + oGA); c->g_pf = reinterpret_cast<float*>(p + oGpf);
    c->Bpad = Bpad; c->chunk = chunk;
  }
  return off;
}
}  // namespace

size_t whmr_smpl_backward_workspace_bytes(whmr_smpl_t h, int B) {
  if (!h || B < 0) return 0;
  return carve_backward(h, B, nullptr, nullptr) + 1024;
}

int whmr_smpl_backward(whmr_smpl_t h, const float* betas, const float* pose, int B, const float* g_verts,
                       const float* g_joints, float* g_betas, float* g_pose, void* workspace, size_t workspace_bytes,
                       void* stream) {
  WHMR_CHECK_ARG(h && B >= 0, "whmr_smpl_backward: bad arguments");
  if (B == 0) return WHMR_OK;
  WHMR_CHECK_ARG(betas && pose && g_betas && g_pose && workspace, "whmr_smpl_backward: null pointer");
  const size_t need = whmr_smpl_backward_workspace_bytes(h, B);
  if (workspace_bytes < need) return set_error(WHMR_E_WORKSPACE, "backward workspace too small: %zu < %zu bytes for B=%d", workspace_bytes, need, B);
  const SmplDevice& d = h->d;
  cudaStream_t st = (cudaStream_t)stream;
  BwdCarve c;
  carve_backward(h, B, reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(workspace), 1024)), &c);
  WHMR_CUDA(cudaMemsetAsync(c.g_A, 0, (size_t)B * d.J * 12 * sizeof(float), st));
  WHMR_CUDA(cudaMemsetAsync(c.g_pf, 0, (size_t)B * d.KP * sizeof(float), st));
  if (g_verts) {
    // forward recompute: chain (A, pose feature), then per chunk the pose offsets with the handle's GEMM mode
    ChainParams p{};
    p.betas = betas; p.pose = pose; p.pose_is_rotmat = 1;
    p.B = B; p.J = d.J; p.NB = d.NB; p.KP = d.KP; p.max_depth = d.max_depth;
    p.J_template = d.J_template; p.J_shapedirs = d.J_shapedirs; p.parents = d.parents; p.depth = d.depth;
    p.A = c.A;
    p.pf = h->gemm_mode == WHMR_GEMM_FP32_SIMT ? c.pf : nullptr;
    p.pf_split = h->gemm_mode == WHMR_GEMM_TC_BF16X3 ? static_cast<__nv_bfloat16*>(c.pf_split) : nullptr;
    p.pf_tf32 = h->gemm_mode == WHMR_GEMM_TC_3XTF32 ? static_cast<float*>(c.pf_split) : nullptr;
    launch_pdl(0, smpl_chain_kernel, dim3(ceil_div(B, kChainWarpsPerBlock)), dim3(kChainWarpsPerBlock * 32), 0, st, p);
    WHMR_LAUNCHED("smpl_chain_kernel");
    SmplWorkspace ws{};
    ws.A = c.A; ws.pf = c.pf; ws.pf_split = c.pf_split; ws.offsets = c.offsets; ws.Bpad = c.Bpad; ws.chunk = c.chunk;
    // split of the n = 3*VP reduction: enough CTAs to fill the machine, slices a multiple of the 32-wide staging step
    int ksplit = 1;
    {
      const int tiles = ceil_div(d.KP, kPbTile) * ceil_div(std::min(B, c.chunk), kPbTile);
      const int want = std::max(1, ceil_div(3 * h->tc.num_sms, tiles));
      for (int s2 = std::min(want, d.NP / kPbN); s2 >= 1; --s2) if ((d.NP / kPbN) % s2 == 0) { ksplit = s2; break; }
    }
    for (int b0 = 0; b0 < B; b0 += c.chunk) {
      const int nb = std::min(c.chunk, B - b0);
      int rc = launch_pose_blend(h, ws, B, b0, nb, st);
      if (rc) return rc;
      SkinBwdParams q{};
      q.g_verts = g_verts + (size_t)b0 * d.V * 3;
      q.offsets = c.offsets; q.A = c.A + (size_t)b0 * d.J * 12;
      q.v_template_p = d.v_template_p; q.ell_idx = d.ell_idx; q.ell_w = d.ell_w;
      q.g_offsets = c.g_off; q.g_A = c.g_A + (size_t)b0 * d.J * 12;
      q.B = nb; q.V = d.V; q.VP = d.VP; q.NP = d.NP; q.J = d.J; q.ell_k = d.ell_k;
      const size_t smem = skin_backward_smem_bytes(d.J);
      WHMR_CUDA(ensure_dyn_smem(skin_backward_kernel, (int)smem));   // per device
      skin_backward_kernel<<<dim3(d.VP / kVertTile, ceil_div(nb, kBwdBodies)), kBwdThreads, smem, st>>>(q);
      WHMR_LAUNCHED("skin_backward_kernel");
      pose_blend_backward_kernel<<<dim3(ceil_div(d.KP, kPbTile), ceil_div(nb, kPbTile), ksplit), 256, 0, st>>>(
          c.g_off, d.posedirs_p, c.g_pf + (size_t)b0 * d.KP, nb, d.KP, d.NP);
      WHMR_LAUNCHED("pose_blend_backward_kernel");
    }
  }
  ChainBwdParams cb{};
  cb.betas = betas; cb.pose = pose; cb.B = B; cb.J = d.J; cb.NB = d.NB; cb.KP = d.KP; cb.max_depth = d.max_depth;
  cb.J_template = d.J_template; cb.J_shapedirs = d.J_shapedirs; cb.parents = d.parents; cb.depth = d.depth;
  cb.g_A = c.g_A; cb.g_joints = g_joints; cb.g_pf = c.g_pf; cb.g_pose = g_pose; cb.g_betas = g_betas;
  chain_backward_kernel<<<ceil_div(B, kChainWarpsPerBlock), kChainWarpsPerBlock * 32, 0, st>>>(cb);
  WHMR_LAUNCHED("chain_backward_kernel");
  return WHMR_OK;
}

int whmr_readout_backward(whmr_readout_t r, const float* g_out, int B, float* g_verts, float* g_joints, void* stream) {
  WHMR_CHECK_ARG(r && B >= 0, "whmr_readout_backward: bad arguments");
  if (B == 0 || r->R == 0) return WHMR_OK;
  WHMR_CHECK_ARG(g_out && g_verts, "whmr_readout_backward: null pointer");
  WHMR_CHECK_ARG(g_joints || !r->needs_joints, "whmr_readout_backward: table references chain joints but g_joints == NULL");
  ReadoutBwdParams q{};
  ReadoutParams& p = q.rp;
  p.row_ptr = r->row_ptr; p.col_idx = r->col_idx; p.vals = r->vals; p.sub_row = r->sub_row;
  p.grp_prefix = r->grp_prefix; p.grp_rows = r->grp_rows;
  p.R = r->R; p.V = r->V; p.J = r->J; p.B = B; p.B_total = B; p.b0 = 0;
  p.out = const_cast<float*>(g_out);
  q.g_verts = g_verts; q.g_joints = g_joints;
  const long long n = (long long)B * r->R;
  readout_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q);
  WHMR_LAUNCHED("readout_backward_kernel");
  return WHMR_OK;
}

int whmr_project_weak_backward(const float* points, const float* cam, const float* g_out, int B, int N, float focal,
                               float img_w, float img_h, float* g_points, float* g_cam, void* stream) {
  WHMR_CHECK_ARG(B >= 0 && N >= 0, "whmr_project_weak_backward: negative size");
  if (B == 0) return WHMR_OK;
  WHMR_CHECK_ARG(points && cam && g_out && g_cam, "whmr_project_weak_backward: null pointer");
  project_weak_bwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(points, cam, g_out, N, focal, img_w, img_h, g_points, g_cam);
  WHMR_LAUNCHED("project_weak_bwd_kernel");
  return WHMR_OK;
}

int whmr_perspective_projection_backward(const float* points, const float* rotation, int rot_batch,
                                         const float* translation, const float* focal_dev, float focal_scalar,
                                         const float* g_out, int B, int N, int retain_z, float* g_points,
                                         float* g_translation, float* g_focal, float* g_center, void* stream) {
  WHMR_CHECK_ARG(B >= 0 && N >= 0, "whmr_perspective_projection_backward: negative size");
  if (B == 0) return WHMR_OK;
  WHMR_CHECK_ARG(points && g_out, "whmr_perspective_projection_backward: null pointer");
  WHMR_CHECK_ARG(rot_batch == 0 || rot_batch == 1 || rot_batch == B, "whmr_perspective_projection_backward: bad rot_batch %d", rot_batch);
  perspective_projection_bwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(points, rotation, rot_batch, translation, focal_dev,
                                                                        focal_scalar, g_out, N, retain_z, g_points,
                                                                        g_translation, g_focal, g_center);
  WHMR_LAUNCHED("perspective_projection_bwd_kernel");
  return WHMR_OK;
}

int whmr_project_full_backward(const float* points, const float* cam, const float* bbox_height, const float* center,
                               const float* orig_shape, const float* Tz, int B, int N, float weak_focal,
                               float weak_img_w, float weak_img_h, const float* g_kp_weak, const float* g_kp_norm,
                               const float* g_focal, const float* g_cam_t, float* g_points, float* g_cam, float* g_Tz,
                               void* stream) {
  WHMR_CHECK_ARG(B >= 0 && N >= 0, "whmr_project_full_backward: negative size");
  if (B == 0) return WHMR_OK;
  WHMR_CHECK_ARG(points && cam && bbox_height && center && orig_shape && Tz && g_cam && g_Tz,
                 "whmr_project_full_backward: null pointer");
  project_full_bwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(points, cam, bbox_height, center, orig_shape, Tz, N, weak_focal,
                                                              weak_img_w, weak_img_h, g_kp_weak, g_kp_norm, g_focal, g_cam_t,
                                                              g_points, g_cam, g_Tz);
  WHMR_LAUNCHED("project_full_bwd_kernel");
  return WHMR_OK;
}

int whmr_sample_bilinear_backward(const float* g_out, int layout, int B, int C, int H, int W, const float* points,
                                  int points_shared, int N, float* g_feat, void* stream) {
  WHMR_CHECK_ARG(B >= 0 && C >= 0 && N >= 0 && H > 0 && W > 0, "whmr_sample_bilinear_backward: bad sizes");
  WHMR_CHECK_ARG(layout == WHMR_LAYOUT_NCHW || layout == WHMR_LAYOUT_NHWC, "whmr_sample_bilinear_backward: bad layout %d", layout);
  if (B == 0 || C == 0 || N == 0) return WHMR_OK;
  WHMR_CHECK_ARG(g_out && points && g_feat, "whmr_sample_bilinear_backward: null pointer");
  WHMR_CHECK_ARG((reinterpret_cast<size_t>(points) & 7) == 0, "whmr_sample_bilinear_backward: points must be 8-byte aligned");
  WHMR_CHECK_ARG(B < 65536, "whmr_sample_bilinear_backward: B too large");
  const int pts_bstride = points_shared ? 0 : N * 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (layout == WHMR_LAYOUT_NCHW) {
    const long long total = (long long)C * N;
    dim3 grid((unsigned)std::min<long long>((total + 255) / 256, 1024), B);
    sample_bilinear_bwd_nchw_kernel<<<grid, 256, 0, st>>>(g_out, points, pts_bstride, g_feat, C, H, W, N);
    WHMR_LAUNCHED("sample_bilinear_bwd_nchw_kernel");
  } else {
    const long long warps = (long long)B * N;
    sample_bilinear_bwd_nhwc_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(g_out, points, pts_bstride, g_feat, B, C,
                                                                                       H, W, N);
    WHMR_LAUNCHED("sample_bilinear_bwd_nhwc_kernel");
  }
  return WHMR_OK;
}

}  // extern "C"
