/*
 * whmr_b200.h -- C ABI of the B200-native W-HMR body-model hot path (libwhmr_b200.so).
 *
 * The reference (yw0208/W-HMR) has no FFI: the path is reached through Python attributes and
 * module-level functions.  Each entry point below names the reference
 * interface it replaces (file:line under the reference tree).  Conventions:
 *   - plain pointers and sizes only; no torch / CUDA types in signatures (`stream` is a
 *     cudaStream_t passed as void*; NULL = the legacy default stream);
 *   - unless a name ends in `_host`, every data pointer is a DEVICE pointer owned by the
 *     caller; nothing is allocated after the `*_create` / `*_reserve` calls;
 *   - all tensors are contiguous row-major float32 unless stated;
 *   - return value: 0 = ok, non-zero = WHMR_E_* ; whmr_last_error() gives the message of the
 *     calling thread's last failure.  No exception ever crosses the ABI;
 *   - kernels are launched asynchronously on `stream`; the caller synchronises.
 *   - there is NO CPU fallback: without a CUDA device every compute call returns WHMR_E_CUDA.
 */
#ifndef WHMR_B200_H_
#define WHMR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WHMR_ABI_VERSION 1

enum {
  WHMR_OK = 0,
  WHMR_E_INVALID = 1,  /* bad argument (null pointer, negative size, unsupported shape) */
  WHMR_E_CUDA = 2,     /* CUDA runtime / driver error (message has cudaGetErrorString) */
  WHMR_E_WORKSPACE = 3 /* caller workspace too small */
};

enum { WHMR_LAYOUT_NCHW = 0, WHMR_LAYOUT_NHWC = 1 };

/* pose-blend GEMM arithmetic (the one dense contraction) */
enum {
  WHMR_GEMM_FP32_SIMT = 0, /* CUDA-core FFMA, exact fp32 products (bring-up / error apportioning) */
  WHMR_GEMM_TC_BF16X3 = 1, /* tcgen05 kind::f16, 3 bf16 products of a hi/lo split (~2^-16 rel) */
  WHMR_GEMM_TC_3XTF32 = 2  /* tcgen05 kind::tf32, 3 tf32 products of a hi/lo split (~2^-21 rel) */
};

int whmr_abi_version(void);
const char* whmr_last_error(void);
/* number of kernels this library has launched from the calling process since load (bench.py's
 * `gpu_launches`); whmr_launch_count_reset() zeroes it. */
uint64_t whmr_launch_count(void);
void whmr_launch_count_reset(void);
/* Diagnostics: `host_mapped` = 16 ints of PINNED host memory (device-accessible through UVA), or NULL to clear.  A bounded
 * mbarrier wait that times out inside a kernel (a protocol bug) records {1, blockIdx.x, threadIdx.x, barrier offset in
 * shared memory, parity, line tag} there before it traps, so the host can read it although the context is lost. */
void whmr_debug_set_trap_buffer(int* host_mapped);

/* ------------------------------------------------------------------------------------------
 * SMPL body model.  Replaces pare.models.SMPL / smplx.SMPL as constructed at
 * models/whmr.py:59 and core/trainer.py:54-64 (commented twin of the wrapper: models/smpl.py:61-83;
 * in-tree statement of the math: models/smpl_webuser/lbs.py:27-79, verts.py:42-50).
 * ------------------------------------------------------------------------------------------ */
typedef struct whmr_smpl_s* whmr_smpl_t;

typedef struct {
  int32_t n_verts;          /* V, 6890 */
  int32_t n_joints;         /* J, 24 (<= 32) */
  int32_t n_betas;          /* 10 (<= 16) */
  const float* v_template;  /* HOST [V,3] */
  const float* shapedirs;   /* HOST [V,3,n_betas] */
  const float* posedirs;    /* HOST [(J-1)*9, V*3]  (smplx in-memory layout) */
  const float* J_regressor; /* HOST [J,V] dense */
  const float* lbs_weights; /* HOST [V,J] dense; any sparsity is detected and exploited */
  const int64_t* parents;   /* HOST [J], parents[0] = -1, parents[i] < i */
} whmr_smpl_model_desc;

/* Uploads and pre-arranges the model on the current device (planar padded layouts, ELL skinning
 * weights, pre-contracted rest-joint regressor, hi/lo-split tensor-core operand of posedirs). */
int whmr_smpl_create(const whmr_smpl_model_desc* desc, int gemm_mode, whmr_smpl_t* out);
int whmr_smpl_destroy(whmr_smpl_t h);
int whmr_smpl_set_gemm_mode(whmr_smpl_t h, int gemm_mode);
int whmr_smpl_get_info(whmr_smpl_t h, int32_t* n_verts, int32_t* n_joints, int32_t* n_betas,
                       int32_t* ell_width, int32_t* gemm_mode);

/* bytes of scratch whmr_smpl_forward needs for a batch of B bodies */
size_t whmr_smpl_workspace_bytes(whmr_smpl_t h, int B);

/* SMPL.forward(betas, body_pose, global_orient, pose2rot, transl) -- call sites
 * models/whmr.py:132-137,227-232,641-644; core/trainer.py:415,420,781,787-790; evaluate/eval.py:159,200.
 *   betas [B,n_betas]; pose: pose_is_rotmat ? [B,J,3,3] : [B,J*3] axis-angle (global_orient first);
 *   transl [B,3] or NULL.
 * Outputs: verts [B,V,3]; joints [B,J,3] = posed kinematic-chain joints (smplx J_transformed);
 *   rel_transforms [B,J,12] (rows of the 3x4 skinning transforms A_j) or NULL.
 * Internally: chain kernel -> pose-blend GEMM -> skinning kernel, all on `stream`. */
int whmr_smpl_forward(whmr_smpl_t h, const float* betas, const float* pose, int pose_is_rotmat,
                      const float* transl, int B, float* verts, float* joints,
                      float* rel_transforms, void* workspace, size_t workspace_bytes, void* stream);

/* The three stages individually (same workspace; used by bench.py for per-kernel CUDA-event
 * timing and by the tests).  whmr_smpl_forward == chain; pose_blend; skin. */
int whmr_smpl_stage_chain(whmr_smpl_t h, const float* betas, const float* pose, int pose_is_rotmat,
                          const float* transl, int B, float* joints, float* rel_transforms,
                          void* workspace, size_t workspace_bytes, void* stream);
int whmr_smpl_stage_pose_blend(whmr_smpl_t h, int B, void* workspace, size_t workspace_bytes,
                               void* stream);
int whmr_smpl_stage_skin(whmr_smpl_t h, const float* betas, int B, float* verts, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Profiling hook: three cudaEvent_t (passed as void*, NULL clears) that the next whmr_smpl_forward[_readout]
 * calls record on their stream right after the chain kernel, after the (last chunk's) pose-blend kernel and
 * after the (last chunk's) skinning kernel, with cudaEventRecordExternal so they become event-record nodes
 * when the call is captured into a CUDA graph.  bench.py uses them to time each kernel inside the timed region. */
int whmr_smpl_set_probe_events(whmr_smpl_t h, void* after_chain, void* after_pose_blend, void* after_skin);

/* Host-buffer variant (bench.py `e2e`): pinned or pageable HOST pointers in and out; H2D copies,
 * the three kernels and the D2H copies are enqueued on `stream` and the call returns after
 * cudaStreamSynchronize.  Device staging comes from whmr_smpl_reserve (grow-only, inside h). */
int whmr_smpl_reserve(whmr_smpl_t h, int max_B);
int whmr_smpl_forward_host(whmr_smpl_t h, const float* betas, const float* pose, int pose_is_rotmat,
                           int B, float* verts, float* joints, void* stream);

/* batch_rodrigues, smplx variant used inside SMPL.forward (angle = ||v+1e-8||,
 * R = I + sin K + (1-cos) K K).  aa [n,3] -> R [n,3,3]. */
int whmr_batch_rodrigues(const float* aa, int n, float* R, void* stream);

/* Rotation glue around the SMPL call in Regressor.forward: one thread per rotation.
 *   whmr_rot6d_to_rotmat       utils/geometry.py:243-257  x [n,6] (viewed [n,3,2]) -> R [n,3,3]
 *   whmr_unbiased_gram_schmidt utils/geometry.py:260-272  x [n,3,3] -> R [n,3,3]   (models/whmr.py:129-130)
 *   whmr_rotmat_to_axis_angle  utils/geometry.py:54-83 (kornia quaternion path, NaN -> 0)  R [n,3,3] -> aa [n,3]
 *                              (models/whmr.py:174,632) */
int whmr_rot6d_to_rotmat(const float* x, int n, float* R, void* stream);
int whmr_unbiased_gram_schmidt(const float* x, int n, float* R, void* stream);
int whmr_rotmat_to_axis_angle(const float* R, int n, float* aa, void* stream);
/* utils/geometry.py:14-51 batch_rodrigues, the QUATERNION variant (half-angle -> quaternion -> normalise -> matrix) the
 * trainer applies to ground-truth poses (core/trainer.py:244).  theta [n,3] -> R [n,3,3]. */
int whmr_batch_rodrigues_quat(const float* theta, int n, float* R, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sparse linear read-out of the posed vertices.  Replaces every "matrix x vertices" and
 * vertex-pick on the path: J_regressor_extra + joint_map (models/smpl.py:66-76), VertexJointSelector
 * (models/whmr.py:60,187,251), H36M regression + pelvis centring (models/whmr.py:176-180,240-244,647-651;
 * evaluate/eval.py:198-219), dense Dmap0/Dmap1 (models/whmr.py:182-183,246-247) and the SSM marker
 * pick (models/whmr.py:184,248).
 *   out[b,r,:] = sum_k vals[k] * src[b, col_idx[k], :]   for k in [row_ptr[r], row_ptr[r+1])
 *               - (sub_row[r] >= 0 ? same sum over row sub_row[r] : 0)
 * where src is the virtual concatenation [verts (V rows) ; joints (n_joints rows)].
 * Rows may be partitioned into consecutive groups (group_sizes, summing to n_rows); the output is
 * then group-major: group g occupies a contiguous [B, group_sizes[g], 3] block, blocks in group
 * order, so each read-out (joints, markers, sub-sampled meshes ...) is its own contiguous tensor
 * while all of them come from one launch pair.  n_groups == 0: a single [B, n_rows, 3] block.
 * ------------------------------------------------------------------------------------------ */
typedef struct whmr_readout_s* whmr_readout_t;
int whmr_readout_create(int n_rows, int n_verts, int n_joints, const int32_t* row_ptr /*HOST [n_rows+1]*/,
                        const int32_t* col_idx /*HOST [nnz]*/, const float* vals /*HOST [nnz]*/,
                        const int32_t* sub_row /*HOST [n_rows] or NULL*/, int n_groups,
                        const int32_t* group_sizes /*HOST [n_groups] or NULL*/, whmr_readout_t* out);
int whmr_readout_destroy(whmr_readout_t r);
int whmr_readout_apply(whmr_readout_t r, const float* verts /*[B,V,3]*/, const float* joints /*[B,J,3] or NULL*/,
                       int B, float* out /*[B*n_rows*3], group-major*/, void* stream);

/* SMPL forward and its read-outs in one call (what Regressor.forward does back to back,
 * models/whmr.py:132-187): same as whmr_smpl_forward followed by whmr_readout_apply(ro, verts, joints),
 * but per 768-body chunk -- the one-hot rows (vertex picks, markers, mesh down-sampling) are written
 * by the skinning kernel's epilogue and the regressor rows are summed from per-body partials it emitted.
 * ro_workspace: whmr_readout_workspace_bytes(ro, min(B, whmr_smpl_chunk_bodies(h))) bytes of scratch for the
 *   per-body partial sums the skinning epilogue emits (NULL: the read-outs run as a stand-alone gather).
 * defer_finish != 0: when the whole batch is one chunk the finishing pass (regressor rows, joint copies) is NOT
 *   enqueued; *finish_deferred is set to 1 and the caller runs whmr_readout_finish(ro, joints, B, ro_workspace,
 *   ro_out, other_stream) whenever it likes (e.g. overlapped with the feature sampling that only needs the markers,
 *   which the skinning kernel has already written). */
size_t whmr_readout_workspace_bytes(whmr_readout_t ro, int n_bodies);
int whmr_smpl_chunk_bodies(whmr_smpl_t h);
/* 1 when pose blend + skinning run as ONE kernel (smpl_fused_tc: bf16x3 mode, tensor-core skinning), else 0 */
int whmr_smpl_is_fused(whmr_smpl_t h);
int whmr_smpl_forward_readout(whmr_smpl_t h, const float* betas, const float* pose, int pose_is_rotmat,
                              const float* transl, int B, float* verts, float* joints, float* rel_transforms,
                              whmr_readout_t ro, float* ro_out /*[B*n_rows*3], group-major*/, void* ro_workspace,
                              size_t ro_workspace_bytes, int defer_finish, int* finish_deferred, void* workspace,
                              size_t workspace_bytes, void* stream);
/* Regressor.forward's rotation glue around the SMPL call, folded into the chain kernel (no extra launch, no extra
 * pass over HBM).  Replaces utils/geometry.py:260-272 unbiased_gram_schmidt (models/whmr.py:129-130, eval mode) on the
 * way in and utils/geometry.py:54-83 rotation_matrix_to_angle_axis (models/whmr.py:174) + the `theta` concatenation
 * (:190) on the way out. */
typedef struct whmr_smpl_glue {
  int32_t gram_schmidt;   /* != 0 (rotation-matrix mode only): R <- unbiased_gram_schmidt(R) before it is used */
  float* rotmat_out;      /* [B,J,9] the rotations actually used, or NULL */
  float* pose_aa_out;     /* [B,J*3] axis-angle of them ('pose'), or NULL */
  float* theta_out;       /* [B, 3+n_betas+J*3] = cat(cam, betas, pose) ('theta'), or NULL */
  const float* cam;       /* [B,3] (theta's head), or NULL (zeros) */
  const float* root_pose; /* NULL, or the root joint's rotation [B,9] / [B,3] (`global_orient`): `pose` then holds the
                             other J-1 joints only ([B,J-1,9] / [B,J-1,3], `body_pose`) -- the reference's two SMPL.forward
                             arguments go in as they are, without the concatenation smplx does (models/whmr.py:132-137) */
} whmr_smpl_glue;
/* whmr_smpl_forward_readout + the glue above (glue == NULL: identical to whmr_smpl_forward_readout; ro may be NULL). */
int whmr_smpl_forward_regressor(whmr_smpl_t h, const float* betas, const float* pose, int pose_is_rotmat,
                                const float* transl, int B, float* verts, float* joints, float* rel_transforms,
                                whmr_readout_t ro, float* ro_out, void* ro_workspace, size_t ro_workspace_bytes,
                                int defer_finish, int* finish_deferred, const whmr_smpl_glue* glue, void* workspace,
                                size_t workspace_bytes, void* stream);
int whmr_readout_finish(whmr_readout_t ro, const float* joints /*[B,J,3] or NULL*/, int B, const void* ro_workspace,
                        float* ro_out, void* stream);
/* The deferred finishing passes of up to 8 whmr_smpl_forward_readout calls (same table, same B, each with its own
 * workspace / output / chain joints) in ONE launch: host arrays of n_calls device pointers.  Inside the regressor loop
 * (models/whmr.py:550-651) nothing but the caller reads the finished rows, so the loop defers them to its end. */
int whmr_readout_finish_multi(whmr_readout_t ro, int n_calls, const float* const* joints, const void* const* ro_workspaces,
                              float* const* ro_outs, int B, void* stream);
/* The same with the joint projections of Regressor.forward / forward_init (models/whmr.py:142-173, 237) folded in: rows
 * [row0, row0 + n_points) of the table (the 49 joints) are projected right after they are finished.  Per call: cam [B,3] or
 * NULL (no projection: the global SMPL call); full != 0: weak projection + predicted-focal block (kp_weak, kp_norm [B,n,2],
 * focal_out [B], cam_t_out [B,3]), else weak projection only (kp_weak).  All device pointers. */
typedef struct whmr_finish_projection {
  int32_t row0, n_points;
  float focal, img_w, img_h;                 /* utils/geometry.py:289-307 constants */
  const float* bbox_height;                  /* [B]   (needed when any call is full) */
  const float* center;                       /* [B,2] */
  const float* orig_shape;                   /* [B,2] (h, w) */
  const float* Tz;                           /* [B] */
  const float* cam[8];
  int32_t full[8];
  float* kp_weak[8];
  float* kp_norm[8];
  float* focal_out[8];
  float* cam_t_out[8];
} whmr_finish_projection;
int whmr_readout_finish_project_multi(whmr_readout_t ro, int n_calls, const float* const* joints,
                                      const void* const* ro_workspaces, float* const* ro_outs, int B,
                                      const whmr_finish_projection* proj, void* stream);

/* ------------------------------------------------------------------------------------------
 * Projection.
 * ------------------------------------------------------------------------------------------ */
/* utils/geometry.py:289-307 projection(pred_joints, pred_camera): t = [tx, ty, 2*focal/(img_h*s + 1e-9)],
 * u = focal*(X+t)_xy/(X+t)_z, out = u / (img_w/2, img_h/2).  points [B,N,3], cam [B,3] -> out [B,N,2]. */
int whmr_project_weak(const float* points, const float* cam, int B, int N, float focal, float img_w,
                      float img_h, float* out, void* stream);

/* utils/geometry.py:310-341 perspective_projection(points, rotation, translation, focal_length,
 * camera_center, retain_z) and its sibling models/maf_extractor.py:192-235 (optional rotation /
 * translation / 5-coefficient distortion).  rotation: [rot_batch,3,3] with rot_batch in
 * {0 (NULL: identity), 1, B}; translation [B,3] or NULL; focal: per-sample [B] if focal_dev != NULL
 * else the scalar focal_scalar; distortion [B,5] or NULL; out [B,N,2] or [B,N,3] when retain_z. */
int whmr_perspective_projection(const float* points, const float* rotation, int rot_batch,
                                const float* translation, const float* focal_dev, float focal_scalar,
                                const float* camera_center, const float* distortion, int B, int N,
                                int retain_z, float* out, void* stream);

/* utils/geometry.py:386-408 estimate_translation (+ estimate_translation_np :344-383; called core/trainer.py:435):
 * confidence-weighted least-squares camera translation per sample, normal equations accumulated and solved in
 * float64 on the device.  S [B,N,3] 3-D joints, joints_2d [B,N,3] = (x, y, confidence); joints [j0, N) take part
 * (the reference uses 25: of 49).  -> out [B,3]. */
int whmr_estimate_translation(const float* S, const float* joints_2d, int B, int N, int j0, float focal, float img_w,
                              float img_h, float* out, void* stream);

/* The predicted-focal block of Regressor.forward, models/whmr.py:147-173, fused:
 * focal = s*bbox_h*Tz/2; cam_t = convert_pare_to_full_img_cam(...) (utils/geometry.py:139-157);
 * kp = perspective_projection(...); kp_norm = kp/center - 1.
 * orig_shape [B,2] = (h,w); center [B,2] = bbox centre (x,y).  Any of kp_px / focal_out / cam_t_out may be NULL. */
int whmr_project_full(const float* points, const float* cam, const float* bbox_height,
                      const float* center, const float* orig_shape, const float* Tz, int B, int N,
                      float* kp_norm, float* kp_px, float* focal_out, float* cam_t_out, void* stream);

/* whmr_project_weak + whmr_project_full on the same points in one launch (Regressor.forward evaluates
 * both back to back, models/whmr.py:142-173).  kp_weak [B,N,2] as whmr_project_weak's out. */
int whmr_project_weak_full(const float* points, const float* cam, const float* bbox_height, const float* center,
                           const float* orig_shape, const float* Tz, int B, int N, float weak_focal,
                           float weak_img_w, float weak_img_h, float* kp_weak, float* kp_norm, float* kp_px,
                           float* focal_out, float* cam_t_out, void* stream);

/* models/maf_extractor.py:145-235 MAF_Extractor.project (+get_trans, perspective_projection with the
 * optional 5-coefficient distortion): full-frame pixels and crop-normalised [-1,1] coordinates. */
int whmr_project_crop(const float* points, const float* cam, const float* center, const float* scale,
                      const float* img_focal, const float* img_center, const float* distortion /*[B,5] or NULL*/,
                      int B, int N, float crop_size, float img_w, float img_h, float* full_out /*or NULL*/,
                      float* crop_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Mesh-aligned feature sampling: models/maf_extractor.py:103-124 (the grid_sample inside
 * MAF_Extractor.sampling): bilinear, zero padding, align_corners=True, points[...,0] <-> W.
 *   feat [B,C,H,W] (NCHW) or [B,H,W,C] (NHWC); points [B,N,2], or [N,2] shared by every body when
 *   points_shared != 0 (the iteration-0 grid of models/whmr.py:338-347,596); out [B,C,N].
 *   `feat` (here and in whmr_project_sample) is a device pointer OR a pointer into page-locked host memory
 *   (cudaHostAlloc / cudaHostRegister / tensor.pin_memory(), unified addressing): a host-resident map is read in place,
 *   so only the sectors its taps touch cross PCIe -- for the sparse regime (67 points on a 128x96 map: 1/11 of the map)
 *   that is 2.4x faster than cudaMemcpyAsync of the map followed by the device gather (profiles/r02_notes.md 3.10).
 *   points / cam / out are device pointers.
 * ------------------------------------------------------------------------------------------ */
int whmr_sample_bilinear(const float* feat, int layout, int B, int C, int H, int W,
                         const float* points, int points_shared, int N, float* out, void* stream);
/* MAF_Extractor.forward (models/maf_extractor.py:126-143) = projection (weak) + sampling, fused:
 * p [B,N,3], cam [B,3]; also writes the 2-D points if points2d_out != NULL. */
int whmr_project_sample(const float* feat, int layout, int B, int C, int H, int W, const float* p,
                        const float* cam, int N, float focal, float img_w, float img_h,
                        float* points2d_out, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * MAF_Extractor.sampling / .forward INCLUDING the `reduce_dim` MLP (models/maf_extractor.py:75-101,
 * 103-143) as one kernel: grid_sample -> Conv1d(k=1) C_in->C1 -> [y;x]->C2 -> [y;x]->C3 with
 * leaky_relu / leaky_relu / relu, 3xTF32 on the tensor cores; the [B,C_in,N] point features are
 * written only when point_feat_out != NULL.  Weights are the module's conv0..2 parameters (Conv1d
 * layout [C_out, C_in_total, 1], DEVICE pointers); set_weights re-splits them (call again after an
 * optimizer step).  mesh_align_out [B, C3*N] (= y.view(B,-1) of [B,C3,N]).
 * Widths: C_in, C1, C2 multiples of 64, C3 a multiple of 16, C1+C2+C3 <= 256 (reference: 256,128,64,32).
 * ------------------------------------------------------------------------------------------ */
typedef struct whmr_maf_mlp_s* whmr_maf_mlp_t;
int whmr_maf_mlp_create(int c_in, int c1, int c2, int c3, whmr_maf_mlp_t* out);
int whmr_maf_mlp_destroy(whmr_maf_mlp_t m);
int whmr_maf_mlp_set_weights(whmr_maf_mlp_t m, const float* w0, const float* b0, const float* w1,
                             const float* b1, const float* w2, const float* b2, void* stream);
int whmr_sample_reduce(whmr_maf_mlp_t m, const float* feat, int layout, int B, int H, int W,
                       const float* points, int points_shared, int N, float* mesh_align_out,
                       float* point_feat_out /*or NULL*/, void* stream);
int whmr_project_sample_reduce(whmr_maf_mlp_t m, const float* feat, int layout, int B, int H, int W,
                               const float* p, const float* cam, int N, float focal, float img_w,
                               float img_h, float* points2d_out /*or NULL*/, float* mesh_align_out,
                               float* point_feat_out /*or NULL*/, void* stream);

/* verts[:, idx] (models/whmr.py:184 markers) -- verts [B,V,3], idx [n_idx] int32 DEVICE -> out [B,n_idx,3] */
int whmr_gather_vertices(const float* verts, const int32_t* idx, int B, int V, int n_idx, float* out,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * Evaluation metrics (BASELINE config 5): evaluate/eval.py:208-223 and utils/pose_utils.py:10-75.
 *   pred, gt [n,J,3].  mpjpe[n] = mean_j ||pred-gt||; pa_mpjpe[n] = same after the similarity
 *   (Procrustes) alignment of pred onto gt (3x3 SVD per sample, on device).
 * ------------------------------------------------------------------------------------------ */
int whmr_joint_errors(const float* pred, const float* gt, int n, int J, float* mpjpe /*or NULL*/,
                      float* pa_mpjpe /*or NULL*/, void* stream);
/* PVE, evaluate/eval.py:208-209: pve[n] = mean_v ||pred[n,v,:] - gt[n,v,:]||;  pred, gt [n,V,3] */
int whmr_vertex_errors(const float* pred, const float* gt, int n, int V, float* pve, void* stream);

/* ------------------------------------------------------------------------------------------
 * Backward entry points (what core/trainer.py:380-636 needs to train through the drop-ins).
 * Reference graph (models/whmr.py:145-173, 586-591): projections see detached joints (gradient -> pred_cam, Tz),
 * sampling points are detached (gradient -> feature maps), SMPL vertices/joints -> betas / rotation matrices.
 * ------------------------------------------------------------------------------------------ */
/* SMPL.forward in rotation-matrix mode (models/whmr.py:132-137): gradients of the vertices and of the posed chain
 * joints w.r.t. betas and the rotation matrices.  g_verts [B,V,3] and/or g_joints [B,J,3] may be NULL (= zero).
 * -> g_betas [B,n_betas], g_pose [B,J,9].  Scratch: whmr_smpl_backward_workspace_bytes(h, B). */
size_t whmr_smpl_backward_workspace_bytes(whmr_smpl_t h, int B);
int whmr_smpl_backward(whmr_smpl_t h, const float* betas, const float* pose, int B, const float* g_verts,
                       const float* g_joints, float* g_betas, float* g_pose, void* workspace, size_t workspace_bytes,
                       void* stream);
/* Transposed read-out: accumulates (atomics) the gradient of every read-out row into g_verts [B,V,3] and
 * g_joints [B,J,3] (may be NULL if no row references a chain joint); both must be initialised by the caller
 * (zeros, or the direct gradients of vertices / chain joints).  g_out: flat group-major buffer as produced by
 * whmr_readout_apply / whmr_smpl_forward_readout. */
int whmr_readout_backward(whmr_readout_t ro, const float* g_out, int B, float* g_verts, float* g_joints, void* stream);
/* utils/geometry.py:289-307.  g_out [B,N,2] -> g_points [B,N,3] (or NULL), g_cam [B,3] */
int whmr_project_weak_backward(const float* points, const float* cam, const float* g_out, int B, int N, float focal,
                               float img_w, float img_h, float* g_points, float* g_cam, void* stream);
/* utils/geometry.py:310-341 perspective_projection (no distortion): g_out [B,N,2|3] (the retained z column has zero
 * gradient) -> g_points [B,N,3] (or NULL), g_translation [B,3] (or NULL), g_focal [B] (or NULL), g_center [B,2] (or NULL).
 * rotation: NULL, one [3,3] (rot_batch 1) or [B,3,3]; it receives no gradient (the reference passes an identity). */
int whmr_perspective_projection_backward(const float* points, const float* rotation, int rot_batch,
                                         const float* translation, const float* focal_dev, float focal_scalar,
                                         const float* g_out, int B, int N, int retain_z, float* g_points,
                                         float* g_translation, float* g_focal, float* g_center, void* stream);
/* models/whmr.py:142-173 (whmr_project_weak_full / whmr_project_full).  Upstream gradients may be NULL.
 * -> g_points [B,N,3] (or NULL), g_cam [B,3], g_Tz [B] */
int whmr_project_full_backward(const float* points, const float* cam, const float* bbox_height, const float* center,
                               const float* orig_shape, const float* Tz, int B, int N, float weak_focal,
                               float weak_img_w, float weak_img_h, const float* g_kp_weak, const float* g_kp_norm,
                               const float* g_focal, const float* g_cam_t, float* g_points, float* g_cam, float* g_Tz,
                               void* stream);
/* models/maf_extractor.py:119 w.r.t. the feature maps.  g_out [B,C,N]; g_feat (layout as the forward's feat) must be
 * zero-filled by the caller; fp32 atomics (summation order over points sharing a pixel is not fixed). */
int whmr_sample_bilinear_backward(const float* g_out, int layout, int B, int C, int H, int W, const float* points,
                                  int points_shared, int N, float* g_feat, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WHMR_B200_H_ */
