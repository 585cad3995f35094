// Shared host/device helpers for libwhmr_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "../../include/whmr_b200.h"

namespace whmr {

// ---- error plumbing: nothing throws across the C ABI -------------------------------------
std::string& last_error_ref();
int set_error(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;

#define WHMR_CHECK_ARG(cond, ...)                                  \
  do {                                                             \
    if (!(cond)) return ::whmr::set_error(WHMR_E_INVALID, __VA_ARGS__); \
  } while (0)

#define WHMR_CUDA(expr)                                                                      \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::whmr::set_error(WHMR_E_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                  \
  } while (0)

// after every kernel launch: count it and surface launch-configuration errors immediately
#define WHMR_LAUNCHED(name)                                                                  \
  do {                                                                                       \
    ::whmr::g_launch_count.fetch_add(1, std::memory_order_relaxed);                          \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return ::whmr::set_error(WHMR_E_CUDA, "launch of %s failed: %s", name,                 \
                               cudaGetErrorString(_e));                                      \
  } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Every kernel of the loop is launched with cudaLaunchAttributeProgrammaticStreamSerialization, so it may become
// resident while its predecessor in the stream is still running.  Contract inside a kernel:
//   pdl_wait()    before the first access to global memory another kernel writes or reads (a no-op without the
//                 attribute); everything before it may only touch constants of the handle and on-chip state;
//   pdl_trigger() right AFTER pdl_wait(), never before: the successor can then start its own prologue, but it
//                 cannot pass ITS pdl_wait() before this grid has completed, so at most two grids overlap.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

int pdl_mask();   // WHMR_PDL: bit mask of the kernel classes launched with the attribute (whmr_b200.cu)
enum { kPdlChain = 1, kPdlFused = 2, kPdlReadout = 4, kPdlProject = 8, kPdlSample = 16 };

template <typename... KArgs, typename... Args>
static inline void launch_pdl(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_mask() & cls) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);   // errors surface through WHMR_LAUNCHED's cudaGetLastError
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel: the largest value set so far is
// remembered per (device, function), so a process that drives several GPUs (handles on cuda:1 after cuda:0,
// nn.DataParallel, multi-device tests) raises it on every device it launches on.  Only ever raises.
template <typename F>
static inline cudaError_t ensure_dyn_smem(F* func, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, int> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> g(mu);
  int& cur = done[std::make_pair(dev, reinterpret_cast<const void*>(func))];
  if (bytes > cur) {
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    cur = bytes;
  }
  return cudaSuccess;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- layout constants shared by the SMPL kernels ------------------------------------------
constexpr int kMaxJoints = 32;   // one lane per joint in the chain kernel
constexpr int kMaxBetas = 16;
constexpr int kVertTile = 128;   // vertices per skinning CTA == rows per tensor-core M tile

// SMPL model as laid out in HBM by whmr_smpl_create (all device pointers).
//  "planar padded": index c*VP + v, c in {x,y,z}, VP = V rounded up to kVertTile; pad entries are 0.
struct SmplDevice {
  int V, VP, J, NB, KP;          // KP = (J-1)*9 rounded up to 16 (pose-feature length, zero padded)
  int NP;                        // 3*VP : row pitch of the pose-offset intermediate
  int ell_k;                     // skinning ELL width (max non-zeros per vertex)
  int max_depth;                 // kinematic tree depth
  // chain kernel constants
  float* J_template;             // [J,3]   J_regressor . v_template   (pre-contracted in fp64)
  float* J_shapedirs;            // [J,3,NB] J_regressor . shapedirs
  int* parents;                  // [J]
  int* depth;                    // [J]
  // skinning kernel constants
  float* v_template_p;           // [3,VP]
  float* shapedirs_p;            // [3,NB,VP]
  int* ell_idx;                  // [ell_k,VP] joint ids (pad: 0)
  float* ell_w;                  // [ell_k,VP] weights  (pad: 0)
  // pose-blend operands
  float* posedirs_p;             // [KP,NP] fp32 planar padded (SIMT path)
  void* posedirs_split;          // tensor-core A operand: [NP, 2, KP] bf16 hi|lo, K-major (TC paths)
  int nfeat;                     // (J-1)*9 pose-feature terms; K = nfeat + NB (shape blend folded in), padded to KP
};

// scratch carved from the caller's workspace for one chunk of bodies
struct SmplWorkspace {
  float* A;          // [B,J,12]
  float* pf;         // [B,KP]       fp32 pose feature (R_j - I), zero padded
  void* pf_split;    // [Bpad,2,KP]  bf16 hi|lo split of pf  (TC paths)
  float* offsets;    // [chunk,NP]   pose offsets + shape blend, planar per body
  float* At;         // [2, Bpad*12, 32] tf32 hi|lo of the skinning transforms, transposed (TC skinning)
  size_t At_part_stride;   // floats
  void* At16;        // [Bpad*12, 64] fp16 hi|lo of the same, one 128-byte row per (body, entry) (fused kernel)
  int Bpad;
  int chunk;         // bodies per GEMM/skin chunk
};

}  // namespace whmr
