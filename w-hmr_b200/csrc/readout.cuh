// Sparse linear read-out of posed vertices: out[b,r,:] = sum_k vals[k] * src[b, col[k], :]  (- sub row).
//
// One mechanism for every "regressor x vertices" / vertex-pick on the path:
// J_regressor_extra + joint_map -> 49 joints (models/smpl.py:66-76), VertexJointSelector
// (models/whmr.py:187,251), H36M 17 joints -> pelvis-centred 14 (models/whmr.py:176-180), the dense
// [1723,6890] and [431,1723] down-sampling matmuls (models/whmr.py:182-183; 71+4.5 MFLOP/body in the
// reference, a one-hot gather in fact) and the SSM marker pick (models/whmr.py:184).
// src is the virtual concatenation [verts ; chain joints] so the 24 chain joints of smplx's
// `joints` output can be routed through the same table.
//
// Two execution paths:
//  (1) stand-alone (`whmr_readout_apply`): readout_all_kernel gathers from the vertex array -- one-hot
//      rows one thread each, other rows one warp per (row, 8 bodies) with a fixed butterfly reduction.
//      2.5k random 12-byte gathers per body: measured latency/wavefront bound (profiles/r01_notes.md).
//  (2) fused with the skinning kernel (`whmr_smpl_forward_readout`): the skinning epilogue already holds
//      every vertex in registers/shared memory, so it EMITS  w * v  for every table entry that
//      references the vertex (EmitEntry lists per group of 32 vertices): one-hot rows go straight to
//      their output slot, regressor terms to a per-body partial buffer laid out in row order, which
//      readout_reduce_kernel then sums sequentially (coalesced, deterministic) and finishes
//      (sub rows, joint-sourced terms).
#pragma once
#include "common.cuh"
#include "projection.cuh"

namespace whmr {

constexpr int kShortRow = 4;

struct ReadoutParams {
  const int* row_ptr;   // [R+1]
  const int* col_idx;   // [nnz]
  const float* vals;    // [nnz]
  const int* sub_row;   // [R] or null
  const int* grp_prefix;  // [R] rows in all groups before this row's group
  const int* grp_rows;    // [R] rows in this row's group
  int R, V, J, B;       // B = bodies handled by this launch (a chunk)
  int B_total, b0;      // batch of the output buffer, first body of the chunk
  const float* verts;   // [B,V,3]  (chunk base)
  const float* joints;  // [B,J,3] or null (chunk base)
  float* out;           // group-major: group g = [B_total, R_g, 3] at float offset 3*B_total*prefix_g
};

__device__ __forceinline__ float* readout_dst(const ReadoutParams& p, int b, int r) {
  const int pre = p.grp_prefix[r], rg = p.grp_rows[r];
  return p.out + 3 * ((size_t)p.B_total * pre + (size_t)(p.b0 + b) * rg + (r - pre));
}

__device__ __forceinline__ const float* readout_src(const ReadoutParams& p, int b, int col) {
  return col < p.V ? p.verts + ((size_t)b * p.V + col) * 3
                   : p.joints + ((size_t)b * p.J + (col - p.V)) * 3;
}

__device__ __forceinline__ void readout_row_serial(const ReadoutParams& p, int b, int r, float& x,
                                                   float& y, float& z) {
  x = y = z = 0.f;
  for (int k = p.row_ptr[r]; k < p.row_ptr[r + 1]; ++k) {
    const float w = p.vals[k];
    const float* s = readout_src(p, b, p.col_idx[k]);
    x = fmaf(w, s[0], x); y = fmaf(w, s[1], y); z = fmaf(w, s[2], z);
  }
}

// ---------------------------------------------------------------------------------------------
// (1) stand-alone gather path
// ---------------------------------------------------------------------------------------------
constexpr int kLongBodies = 8;

__device__ __forceinline__ void readout_rows_warp8(const ReadoutParams& p, int r, int bbase, int nvalid, int lane,
                                                   float (&x)[kLongBodies], float (&y)[kLongBodies],
                                                   float (&z)[kLongBodies]) {
#pragma unroll
  for (int i = 0; i < kLongBodies; ++i) x[i] = y[i] = z[i] = 0.f;
  const int k1 = p.row_ptr[r + 1];
  for (int k = p.row_ptr[r] + lane; k < k1; k += 32) {
    const float w = p.vals[k];
    const int col = p.col_idx[k];
#pragma unroll
    for (int i = 0; i < kLongBodies; ++i) {
      const int b = bbase + (i < nvalid ? i : 0);
      const float* s = readout_src(p, b, col);
      x[i] = fmaf(w, s[0], x[i]); y[i] = fmaf(w, s[1], y[i]); z[i] = fmaf(w, s[2], z[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < kLongBodies; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      x[i] += __shfl_xor_sync(0xffffffffu, x[i], o);
      y[i] += __shfl_xor_sync(0xffffffffu, y[i], o);
      z[i] += __shfl_xor_sync(0xffffffffu, z[i], o);
    }
  }
}

// All row classes in ONE launch: blocks [0, n_blocks_onehot) gather the one-hot rows, the next
// n_blocks_long blocks reduce the regressor rows (8 bodies per warp), the rest take the remaining short rows.
struct ReadoutAllParams {
  ReadoutParams rp;
  const int4* onehot_tab; int n_onehot;   // {source index, group prefix, group rows, row - prefix}
  const int* rows_long; int n_long;
  const int* rows_short; int n_short;
  int n_blocks_onehot, n_blocks_long;
};

__global__ void __launch_bounds__(256) readout_all_kernel(ReadoutAllParams q) {
  pdl_wait();
  pdl_trigger();
  ReadoutParams& p = q.rp;
  if ((int)blockIdx.x < q.n_blocks_onehot) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)p.B * q.n_onehot) return;
    const int b = (int)(i / q.n_onehot);
    const int4 t = q.onehot_tab[(int)(i - (long long)b * q.n_onehot)];
    const float* s = readout_src(p, b, t.x);
    const float x = s[0], y = s[1], z = s[2];
    float* o = p.out + 3 * ((size_t)p.B_total * t.y + (size_t)(p.b0 + b) * t.z + t.w);
    o[0] = x; o[1] = y; o[2] = z;
    return;
  }
  if ((int)blockIdx.x < q.n_blocks_onehot + q.n_blocks_long) {
    const long long w = ((long long)(blockIdx.x - q.n_blocks_onehot) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int n_bblocks = (p.B + kLongBodies - 1) / kLongBodies;
    if (w >= (long long)n_bblocks * q.n_long) return;   // warp-uniform
    const int bb = (int)(w / q.n_long);
    const int r = q.rows_long[(int)(w % q.n_long)];
    const int bbase = bb * kLongBodies;
    const int nvalid = min(kLongBodies, p.B - bbase);
    float x[kLongBodies], y[kLongBodies], z[kLongBodies];
    readout_rows_warp8(p, r, bbase, nvalid, lane, x, y, z);
    const int sr = p.sub_row ? p.sub_row[r] : -1;
    if (sr >= 0) {
      float sx[kLongBodies], sy[kLongBodies], sz[kLongBodies];
      readout_rows_warp8(p, sr, bbase, nvalid, lane, sx, sy, sz);
#pragma unroll
      for (int i = 0; i < kLongBodies; ++i) { x[i] -= sx[i]; y[i] -= sy[i]; z[i] -= sz[i]; }
    }
#pragma unroll
    for (int i = 0; i < kLongBodies; ++i) {
      if (lane == i && i < nvalid) {
        float* o = readout_dst(p, bbase + i, r);
        o[0] = x[i]; o[1] = y[i]; o[2] = z[i];
      }
    }
    return;
  }
  {
    const long long i = (long long)(blockIdx.x - q.n_blocks_onehot - q.n_blocks_long) * blockDim.x + threadIdx.x;
    if (i >= (long long)p.B * q.n_short) return;
    const int b = (int)(i / q.n_short);
    const int r = q.rows_short[(int)(i % q.n_short)];
    float x, y, z;
    readout_row_serial(p, b, r, x, y, z);
    const int sr = p.sub_row ? p.sub_row[r] : -1;
    if (sr >= 0) {
      float sx, sy, sz;
      readout_row_serial(p, b, sr, sx, sy, sz);
      x -= sx; y -= sy; z -= sz;
    }
    float* o = readout_dst(p, b, r);
    o[0] = x; o[1] = y; o[2] = z;
  }
}

// ---------------------------------------------------------------------------------------------
// (2) fused path: entries emitted by the skinning epilogue + sequential reduction
// ---------------------------------------------------------------------------------------------
// One entry per table non-zero whose source is a vertex.  kind 0: one-hot row, the value goes to the
// final output (a = group prefix, c = rows in group, d = row - prefix).  kind 1: regressor term, the value
// w*v goes to partial[b][d] where d is the term's slot in EMIT order (consecutive lanes -> consecutive slots,
// so the epilogue's stores coalesce); readout_reduce_kernel finds the slots of a row through slot_of[].
struct EmitEntry {
  int lv_kind;   // local vertex (0..31) | kind << 8
  float w;
  int a, c, d;
};

struct EmitTable {              // device pointers, owned by the read-out handle
  const int* grp_ptr;           // [VP/32 + 1] entries per 32-vertex group
  const EmitEntry* entries;
  float* partial;               // [chunk, n_partial, 3]
  int n_partial;
};

// Finishing pass for the rows that are not vertex one-hots.  One CTA per body: the body's partial array
// (emit order, ~30 KB) is staged in shared memory with coalesced loads, then thread r sums row r's terms
// in the row's storage order through the slot list (deterministic), adds joint-sourced terms, subtracts
// the sub row and writes the output.
struct ReduceParams {
  ReadoutParams rp;
  const int* rows; int n_rows;          // rows handled here
  const int* part_ptr;                  // [R+1] range of each row in slot_of
  const int* slot_of;                   // [n_vertex_terms] emit slot of each vertex-sourced term, row-major
  const int* jt_ptr;                    // [R+1] range of each row's joint-sourced terms
  const int* jt_col; const float* jt_val;
  const float* partial; int n_partial;  // [chunk, n_partial, 3], n_partial % 4 == 0 (emit SLOTS per body)
  int n_terms;                          // entries of slot_of (vertex-sourced TERMS, padded to 4): >= the slots when rows
                                        // repeat (a row equal to an earlier one shares its slots: nothing emitted twice)
  // batched finish (whmr_readout_finish_multi): blockIdx.y selects one of n_multi independent (joints, partial, out)
  // triples of the same table and batch size -- the finishing passes of several SMPL calls in ONE launch
  int n_multi;
  const float* joints_m[8];
  const float* partial_m[8];
  float* out_m[8];
  // projections of the finished rows [row0, row0 + n_points) (the 49 joints) folded into the finishing pass
  // (whmr_readout_finish_project_multi): weak projection for every call with a camera, the predicted-focal block too
  // where full[call] != 0 -- the loop's four projection launches disappear (models/whmr.py:142-173, 237)
  whmr_finish_projection pj;
};

__device__ __forceinline__ void reduce_row(const ReduceParams& q, const float* joints, const float* ps, const int* slots,
                                           int b, int r, float& x, float& y, float& z) {
  const ReadoutParams& p = q.rp;
  x = y = z = 0.f;
  int k = q.part_ptr[r];
  const int k1 = q.part_ptr[r + 1];
  for (; k + 4 <= k1; k += 4) {   // four terms' loads in flight; the additions keep the row's storage order
    const float* e0 = ps + slots[k] * 3;     // slot list staged in shared memory: no dependent global load per term
    const float* e1 = ps + slots[k + 1] * 3;
    const float* e2 = ps + slots[k + 2] * 3;
    const float* e3 = ps + slots[k + 3] * 3;
    const float a0 = e0[0], a1 = e0[1], a2 = e0[2], b0 = e1[0], b1 = e1[1], b2 = e1[2];
    const float c0 = e2[0], c1 = e2[1], c2 = e2[2], d0 = e3[0], d1 = e3[1], d2 = e3[2];
    x += a0; y += a1; z += a2; x += b0; y += b1; z += b2;
    x += c0; y += c1; z += c2; x += d0; y += d1; z += d2;
  }
  for (; k < k1; ++k) {
    const float* e = ps + slots[k] * 3;
    x += e[0]; y += e[1]; z += e[2];
  }
  for (int k = q.jt_ptr[r]; k < q.jt_ptr[r + 1]; ++k) {
    const float w = q.jt_val[k];
    const float* s = joints + ((size_t)b * p.J + q.jt_col[k]) * 3;
    x = fmaf(w, s[0], x); y = fmaf(w, s[1], y); z = fmaf(w, s[2], z);
  }
}

__global__ void __launch_bounds__(128) readout_reduce_kernel(ReduceParams q) {
  extern __shared__ __align__(16) float ps[];   // [n_partial * 3]
  pdl_wait();
  pdl_trigger();
  ReadoutParams p = q.rp;
  const int b = blockIdx.x;
  const float* partial = q.partial;
  if (q.n_multi > 0) { p.joints = q.joints_m[blockIdx.y]; p.out = q.out_m[blockIdx.y]; partial = q.partial_m[blockIdx.y]; }
  int* slots = reinterpret_cast<int*>(ps + q.n_partial * 3);   // [n_terms] emit slot of every vertex-sourced term
  {   // cp.async: every 16-byte copy of a thread is in flight at once (a load -> store loop of 128 threads keeps ~8 KB in
      // flight and made this ~50 KB staging the longest phase of the kernel: 25 us per 1280-CTA launch under ncu)
    const float4* src = reinterpret_cast<const float4*>(partial + (size_t)b * q.n_partial * 3);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ps);
    const int n4 = q.n_partial * 3 / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)i * 16u), "l"(src + i) : "memory");
    const int4* ssrc = reinterpret_cast<const int4*>(q.slot_of);   // padded to n_terms entries by the host
    const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(slots);
    for (int i = threadIdx.x; i < q.n_terms / 4; i += blockDim.x)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sdst + (uint32_t)i * 16u), "l"(ssrc + i) : "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < q.n_rows; i += blockDim.x) {
    const int r = q.rows[i];
    float x, y, z;
    reduce_row(q, p.joints, ps, slots, b, r, x, y, z);
    const int sr = p.sub_row ? p.sub_row[r] : -1;
    if (sr >= 0) {
      float sx, sy, sz;
      reduce_row(q, p.joints, ps, slots, b, sr, sx, sy, sz);
      x -= sx; y -= sy; z -= sz;
    }
    float* o = readout_dst(p, b, r);
    o[0] = x; o[1] = y; o[2] = z;
  }
  if (q.n_multi > 0 && q.pj.n_points > 0 && q.pj.cam[blockIdx.y]) {
    __syncthreads();   // this body's rows written above are visible to the whole CTA (one-hot rows: by the SMPL kernel)
    const int call = blockIdx.y;
    const float* cam = q.pj.cam[call];
    for (int n = threadIdx.x; n < q.pj.n_points; n += blockDim.x) {
      const float* pt = readout_dst(p, b, q.pj.row0 + n);
      const size_t i2 = ((size_t)b * q.pj.n_points + n) * 2;
      if (q.pj.full[call])
        project_full_point(pt, b, n, cam, q.pj.bbox_height, q.pj.center, q.pj.orig_shape, q.pj.Tz, q.pj.kp_norm[call] + i2,
                           nullptr, q.pj.focal_out[call], q.pj.cam_t_out[call], q.pj.kp_weak[call] + i2, q.pj.focal,
                           q.pj.img_w, q.pj.img_h);
      else
        project_weak_point(pt, cam[b * 3 + 0], cam[b * 3 + 1], cam[b * 3 + 2], q.pj.focal, q.pj.img_w, q.pj.img_h,
                           q.pj.kp_weak[call] + i2);
    }
  }
}

// verts[:, idx]
__global__ void __launch_bounds__(256) gather_vertices_kernel(const float* __restrict__ verts,
                                                              const int* __restrict__ idx, int B, int V,
                                                              int n_idx, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over B*n_idx*3
  if (i >= (long long)B * n_idx * 3) return;
  const int c = (int)(i % 3);
  const long long t = i / 3;
  const int k = (int)(t % n_idx);
  const int b = (int)(t / n_idx);
  out[i] = verts[((size_t)b * V + idx[k]) * 3 + c];
}

}  // namespace whmr
