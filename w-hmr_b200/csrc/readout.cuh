// Sparse linear read-out of posed vertices: out[b,r,:] = sum_k vals[k] * src[b, col[k], :].
//
// One mechanism for every "regressor x vertices" / vertex-pick on the path:
// J_regressor_extra + joint_map -> 49 joints (models/smpl.py:66-76), VertexJointSelector
// (models/whmr.py:187,251), H36M 17 joints -> pelvis-centred 14 (models/whmr.py:176-180), the dense
// [1723,6890] and [431,1723] down-sampling matmuls (models/whmr.py:182-183; 71+4.5 MFLOP/body in the
// reference, a one-hot gather in fact) and the SSM marker pick (models/whmr.py:184).
// src is the virtual concatenation [verts ; chain joints] so the 24 chain joints of smplx's
// `joints` output can be routed through the same table.
//
// Rows are split by the host into "short" (<= kShortRow non-zeros: one thread per (body,row)) and
// "long" (one warp per (body,row), shuffle reduction in a fixed order => deterministic).
#pragma once
#include "common.cuh"

namespace whmr {

constexpr int kShortRow = 4;

struct ReadoutParams {
  const int* row_ptr;   // [R+1]
  const int* col_idx;   // [nnz]
  const float* vals;    // [nnz]
  const int* sub_row;   // [R] or null
  const int* grp_prefix;  // [R] rows in all groups before this row's group
  const int* grp_rows;    // [R] rows in this row's group
  const int* rows;      // row ids handled by this launch
  int n_rows_here;      // entries in `rows`
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

__global__ void __launch_bounds__(256) readout_short_kernel(ReadoutParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.B * p.n_rows_here) return;
  const int b = (int)(i / p.n_rows_here);
  const int r = p.rows[(int)(i % p.n_rows_here)];
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

__device__ __forceinline__ void readout_row_warp(const ReadoutParams& p, int b, int r, int lane,
                                                 float& x, float& y, float& z) {
  x = y = z = 0.f;
  for (int k = p.row_ptr[r] + lane; k < p.row_ptr[r + 1]; k += 32) {
    const float w = p.vals[k];
    const float* s = readout_src(p, b, p.col_idx[k]);
    x = fmaf(w, s[0], x); y = fmaf(w, s[1], y); z = fmaf(w, s[2], z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, o);
    y += __shfl_xor_sync(0xffffffffu, y, o);
    z += __shfl_xor_sync(0xffffffffu, z, o);
  }
}

__global__ void __launch_bounds__(256) readout_long_kernel(ReadoutParams p) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)p.B * p.n_rows_here) return;   // warp-uniform
  const int b = (int)(w / p.n_rows_here);
  const int r = p.rows[(int)(w % p.n_rows_here)];
  float x, y, z;
  readout_row_warp(p, b, r, lane, x, y, z);
  const int sr = p.sub_row ? p.sub_row[r] : -1;
  if (sr >= 0) {
    float sx, sy, sz;
    readout_row_warp(p, b, sr, lane, sx, sy, sz);
    x -= sx; y -= sy; z -= sz;
  }
  if (lane == 0) {
    float* o = readout_dst(p, b, r);
    o[0] = x; o[1] = y; o[2] = z;
  }
}

// One-hot rows (vertex picks, SSM markers, mesh down-sampling): compact table, one thread per (body,row).
//   tab[r] = {source vertex, rows of all groups before this row's group, rows in its group, row - prefix}
__global__ void __launch_bounds__(256)
readout_onehot_kernel(const int4* __restrict__ tab, int n_rows, const float* __restrict__ verts,
                      const float* __restrict__ joints, int V, int J, int nb, int B_total, int b0,
                      float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)nb * n_rows) return;
  const int b = (int)(i / n_rows);
  const int4 t = tab[(int)(i - (long long)b * n_rows)];
  const float* s = t.x < V ? verts + ((size_t)b * V + t.x) * 3 : joints + ((size_t)b * J + (t.x - V)) * 3;
  const float x = s[0], y = s[1], z = s[2];
  float* o = out + 3 * ((size_t)B_total * t.y + (size_t)(b0 + b) * t.z + t.w);
  o[0] = x; o[1] = y; o[2] = z;
}

// Regressor rows: one warp per (row, block of kLongBodies bodies).  The lanes fetch the row's (column,
// weight) pairs once, then every lane gathers its vertex for each body (kLongBodies independent
// loads in flight) and the warp reduces in a fixed butterfly order => deterministic.
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

__global__ void __launch_bounds__(256) readout_long8_kernel(ReadoutParams p) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int n_bblocks = (p.B + kLongBodies - 1) / kLongBodies;
  if (w >= (long long)n_bblocks * p.n_rows_here) return;   // warp-uniform
  const int bb = (int)(w / p.n_rows_here);
  const int r = p.rows[(int)(w % p.n_rows_here)];
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
  // lane i writes body i
#pragma unroll
  for (int i = 0; i < kLongBodies; ++i) {
    if (lane == i && i < nvalid) {
      float* o = readout_dst(p, bbase + i, r);
      o[0] = x[i]; o[1] = y[i]; o[2] = z[i];
    }
  }
}

// All row classes in ONE launch: blocks [0, n_blocks_onehot) gather the one-hot rows, the next
// n_blocks_long blocks reduce the regressor rows (8 bodies per warp), the rest take the remaining short rows.
struct ReadoutAllParams {
  ReadoutParams rp;            // rows / n_rows_here are set per class below
  const int4* onehot_tab; int n_onehot;
  const int* rows_long; int n_long;
  const int* rows_short; int n_short;
  int n_blocks_onehot, n_blocks_long;
};

__global__ void __launch_bounds__(256) readout_all_kernel(ReadoutAllParams q) {
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
    // regressor rows: warp = (row, 32 consecutive bodies), thread = one body.  The row's (column, weight)
    // pairs are warp-uniform loads; every thread sums its own body in the row's storage order, so the
    // result is deterministic and independent of the batch composition; no shuffles.
    const long long w = ((long long)(blockIdx.x - q.n_blocks_onehot) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int n_bgroups = (p.B + 31) >> 5;
    if (w >= (long long)n_bgroups * q.n_long) return;   // warp-uniform
    const int r = q.rows_long[(int)(w % q.n_long)];
    const int b = (int)(w / q.n_long) * 32 + lane;
    const int bc = min(b, p.B - 1);
    auto row_dot = [&](int row, float& x, float& y, float& z) {
      x = y = z = 0.f;
      const int k1 = p.row_ptr[row + 1];
      int k = p.row_ptr[row];
      for (; k + 8 <= k1; k += 8) {
        float wv[8], sx[8], sy[8], sz[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          wv[u] = p.vals[k + u];
          const float* s = readout_src(p, bc, p.col_idx[k + u]);
          sx[u] = s[0]; sy[u] = s[1]; sz[u] = s[2];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { x = fmaf(wv[u], sx[u], x); y = fmaf(wv[u], sy[u], y); z = fmaf(wv[u], sz[u], z); }
      }
      for (; k < k1; ++k) {
        const float wv = p.vals[k];
        const float* s = readout_src(p, bc, p.col_idx[k]);
        x = fmaf(wv, s[0], x); y = fmaf(wv, s[1], y); z = fmaf(wv, s[2], z);
      }
    };
    float x, y, z;
    row_dot(r, x, y, z);
    const int sr = p.sub_row ? p.sub_row[r] : -1;
    if (sr >= 0) {
      float sx, sy, sz;
      row_dot(sr, sx, sy, sz);
      x -= sx; y -= sy; z -= sz;
    }
    if (b < p.B) {
      float* o = readout_dst(p, b, r);
      o[0] = x; o[1] = y; o[2] = z;
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
