// Torch-free consumer of the C ABI (include/whmr_b200.h): what a maintainer binding libwhmr_b200.so from C / C++ /
// another FFI would write.  Plain cudart buffers, no Python, no torch.  Checks, on a synthetic SMPL-shaped model:
//   1. whmr_smpl_forward_host with the identity pose returns v_template + shapedirs . beta   (host double reference);
//   2. whmr_smpl_forward (device buffers, axis-angle) moves a single-joint-weighted vertex rigidly with the root;
//   3. whmr_project_weak and whmr_sample_bilinear against straightforward host loops;
//   4. argument errors come back as status codes with a message, never as exceptions;
//   5. whmr_maf_mlp_* + whmr_project_sample_reduce (sampling + projection + Conv1d MLP in one launch) against a host loop.
// Build (tests/test_abi_gpu.py does this):  g++ -std=c++17 -I include -I $CUDA/include abi_smoke.cpp -L w-hmr_b200
//   -lwhmr_b200 -L $CUDA/lib64 -lcudart -Wl,-rpath,... -o abi_smoke
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "whmr_b200.h"

static uint64_t g_state = 0x9E3779B97F4A7C15ull;
static float frand() {   // xorshift, uniform in [-1, 1)
  g_state ^= g_state << 13; g_state ^= g_state >> 7; g_state ^= g_state << 17;
  return (float)((g_state >> 11) * (1.0 / 9007199254740992.0)) * 2.0f - 1.0f;
}
#define CK(x) do { int rc_ = (x); if (rc_) { printf("FAIL %s -> %d: %s\n", #x, rc_, whmr_last_error()); return 1; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main() {
  const int V = 6890, J = 24, NB = 10, B = 5;
  static const int64_t parents[24] = {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};
  std::vector<float> vt((size_t)V * 3), sd((size_t)V * 3 * NB), pd((size_t)(J - 1) * 9 * V * 3), jr((size_t)J * V, 0.f),
      w((size_t)V * J, 0.f);
  for (auto& x : vt) x = 0.5f * frand();
  for (auto& x : sd) x = 0.01f * frand();
  for (auto& x : pd) x = 0.002f * frand();
  for (int j = 0; j < J; ++j) for (int k = 0; k < 8; ++k) jr[(size_t)j * V + (j * 131 + k * 17) % V] = 0.125f;
  for (int v = 0; v < V; ++v) {   // vertex 0: root only; others: two joints
    if (v == 0) { w[0] = 1.f; continue; }
    const int a = v % J, b = (v * 7 + 3) % J;
    if (a == b) w[(size_t)v * J + a] = 1.f; else { w[(size_t)v * J + a] = 0.625f; w[(size_t)v * J + b] = 0.375f; }
  }
  whmr_smpl_model_desc d{V, J, NB, vt.data(), sd.data(), pd.data(), jr.data(), w.data(), parents};
  whmr_smpl_t h = nullptr;
  if (whmr_abi_version() != WHMR_ABI_VERSION) { printf("FAIL abi version\n"); return 1; }
  CK(whmr_smpl_create(&d, WHMR_GEMM_TC_BF16X3, &h));

  // 4. errors are return codes
  if (whmr_smpl_forward(h, nullptr, nullptr, 1, nullptr, 3, nullptr, nullptr, nullptr, nullptr, 0, nullptr) == WHMR_OK ||
      whmr_last_error()[0] == 0) { printf("FAIL: null pointers accepted\n"); return 1; }

  // 1. identity pose through the host-buffer entry
  std::vector<float> betas((size_t)B * NB), rot((size_t)B * J * 9, 0.f), verts((size_t)B * V * 3), joints((size_t)B * J * 3);
  for (auto& x : betas) x = 1.5f * frand();
  for (int i = 0; i < B * J; ++i) rot[(size_t)i * 9 + 0] = rot[(size_t)i * 9 + 4] = rot[(size_t)i * 9 + 8] = 1.f;
  CK(whmr_smpl_reserve(h, B));
  CK(whmr_smpl_forward_host(h, betas.data(), rot.data(), 1, B, verts.data(), joints.data(), nullptr));
  double e1 = 0;
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < V * 3; ++i) {
      double r = vt[i];
      for (int k = 0; k < NB; ++k) r += (double)sd[(size_t)i * NB + k] * betas[(size_t)b * NB + k];
      e1 = std::fmax(e1, std::fabs(r - verts[(size_t)b * V * 3 + i]));
    }
  printf("identity pose: max |verts - v_shaped| = %.3g m\n", e1);
  if (!(e1 <= 4e-6)) { printf("FAIL identity pose\n"); return 1; }

  // 2. device buffers, axis-angle: rotate the root by theta about z; vertex 0 (root weight 1) must move rigidly:
  //    v' = Rz (v_shaped - J0) + J0 with J0 the root's rest joint (= joints[.,0] at any pose)
  float *d_betas, *d_pose, *d_verts, *d_joints; void* d_ws;
  std::vector<float> aa((size_t)B * J * 3, 0.f);
  const float theta = 0.7f;
  for (int b = 0; b < B; ++b) aa[(size_t)b * J * 3 + 2] = theta;
  const size_t ws_bytes = whmr_smpl_workspace_bytes(h, B);
  CU(cudaMalloc(&d_betas, betas.size() * 4)); CU(cudaMalloc(&d_pose, aa.size() * 4));
  CU(cudaMalloc(&d_verts, verts.size() * 4)); CU(cudaMalloc(&d_joints, joints.size() * 4)); CU(cudaMalloc(&d_ws, ws_bytes));
  CU(cudaMemcpy(d_betas, betas.data(), betas.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_pose, aa.data(), aa.size() * 4, cudaMemcpyHostToDevice));
  std::vector<float> verts0 = verts, joints0 = joints;   // identity-pose results = v_shaped, rest joints
  CK(whmr_smpl_forward(h, d_betas, d_pose, 0, nullptr, B, d_verts, d_joints, nullptr, d_ws, ws_bytes, nullptr));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(verts.data(), d_verts, verts.size() * 4, cudaMemcpyDeviceToHost));
  double e2 = 0;
  for (int b = 0; b < B; ++b) {
    const float* v0 = &verts0[(size_t)b * V * 3]; const float* j0 = &joints0[(size_t)b * J * 3];
    const double dx = v0[0] - j0[0], dy = v0[1] - j0[1], dz = v0[2] - j0[2], c = std::cos(theta), s = std::sin(theta);
    const double rx = c * dx - s * dy + j0[0], ry = s * dx + c * dy + j0[1], rz = dz + j0[2];
    const float* v = &verts[(size_t)b * V * 3];
    // the pose-corrective offsets of vertex 0 do not depend on the root rotation, so compare with identity pose + rigid move
    e2 = std::fmax(e2, std::fmax(std::fabs(rx - v[0]), std::fmax(std::fabs(ry - v[1]), std::fabs(rz - v[2]))));
  }
  printf("root rotation: rigid-vertex error = %.3g m\n", e2);
  if (!(e2 <= 4e-6)) { printf("FAIL rigid vertex\n"); return 1; }

  // 3. projection + sampling against host loops
  const int N = 49, C = 8, H = 9, W = 7;
  std::vector<float> pts((size_t)B * N * 3), cam((size_t)B * 3), kp((size_t)B * N * 2), feat((size_t)B * C * H * W),
      pf((size_t)B * C * N);
  for (auto& x : pts) x = 0.4f * frand();
  for (int b = 0; b < B; ++b) { cam[b * 3] = 0.9f + 0.2f * frand(); cam[b * 3 + 1] = 0.1f * frand(); cam[b * 3 + 2] = 0.1f * frand(); }
  for (auto& x : feat) x = frand();
  float *d_pts, *d_cam, *d_kp, *d_feat, *d_pf;
  CU(cudaMalloc(&d_pts, pts.size() * 4)); CU(cudaMalloc(&d_cam, cam.size() * 4)); CU(cudaMalloc(&d_kp, kp.size() * 4));
  CU(cudaMalloc(&d_feat, feat.size() * 4)); CU(cudaMalloc(&d_pf, pf.size() * 4));
  CU(cudaMemcpy(d_pts, pts.data(), pts.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_cam, cam.data(), cam.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_feat, feat.data(), feat.size() * 4, cudaMemcpyHostToDevice));
  CK(whmr_project_weak(d_pts, d_cam, B, N, 1000.f, 256.f, 256.f, d_kp, nullptr));
  CK(whmr_sample_bilinear(d_feat, WHMR_LAYOUT_NCHW, B, C, H, W, d_kp, 0, N, d_pf, nullptr));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(kp.data(), d_kp, kp.size() * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(pf.data(), d_pf, pf.size() * 4, cudaMemcpyDeviceToHost));
  double e3 = 0, e4 = 0;
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < N; ++n) {
      const double tz = 2.0 * 1000.0 / (256.0 * cam[b * 3] + 1e-9);
      const double x = pts[((size_t)b * N + n) * 3] + cam[b * 3 + 1], y = pts[((size_t)b * N + n) * 3 + 1] + cam[b * 3 + 2],
                   z = pts[((size_t)b * N + n) * 3 + 2] + tz;
      const double u = 1000.0 * x / z / 128.0, v = 1000.0 * y / z / 128.0;
      e3 = std::fmax(e3, std::fmax(std::fabs(u - kp[((size_t)b * N + n) * 2]), std::fabs(v - kp[((size_t)b * N + n) * 2 + 1])) * 128.0);
      const double ix = (kp[((size_t)b * N + n) * 2] + 1.0) * 0.5 * (W - 1), iy = (kp[((size_t)b * N + n) * 2 + 1] + 1.0) * 0.5 * (H - 1);
      const int x0 = (int)std::floor(ix), y0 = (int)std::floor(iy);
      for (int c = 0; c < C; ++c) {
        double acc = 0;
        for (int dy = 0; dy < 2; ++dy)
          for (int dx = 0; dx < 2; ++dx) {
            const int xx = x0 + dx, yy = y0 + dy;
            if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
            const double wgt = (dx ? ix - x0 : 1.0 - (ix - x0)) * (dy ? iy - y0 : 1.0 - (iy - y0));
            acc += wgt * feat[(((size_t)b * C + c) * H + yy) * W + xx];
          }
        e4 = std::fmax(e4, std::fabs(acc - pf[((size_t)b * C + c) * N + n]));
      }
    }
  printf("projection: %.3g px   sampling: %.3g\n", e3, e4);
  if (!(e3 <= 1e-3) || !(e4 <= 1e-4)) { printf("FAIL projection/sampling\n"); return 1; }
  // 5. MAF_Extractor.forward as one call: weak projection + sampling + the reduce_dim MLP (3xTF32 tensor-core kernel),
  //    weights in Conv1d layout, against a host double-precision loop; the [B,256,N] output is optional
  {
    const int C0 = 256, C1 = 128, C2 = 64, C3 = 32, Hm = 10, Wm = 6;
    std::vector<float> fm((size_t)B * C0 * Hm * Wm), w0((size_t)C1 * C0), b0(C1), w1((size_t)C2 * (C1 + C0)), b1(C2),
        w2((size_t)C3 * (C2 + C0)), b2(C3), maf((size_t)B * C3 * N), pfm((size_t)B * C0 * N);
    for (auto& x : fm) x = frand();
    for (auto& x : w0) x = frand() * 0.12f;
    for (auto& x : w1) x = frand() * 0.10f;
    for (auto& x : w2) x = frand() * 0.11f;
    for (auto& x : b0) x = 0.3f * frand();
    for (auto& x : b1) x = 0.3f * frand();
    for (auto& x : b2) x = 0.3f * frand();
    float *d_fm, *d_w0, *d_b0, *d_w1, *d_b1, *d_w2, *d_b2, *d_maf, *d_pfm;
    CU(cudaMalloc(&d_fm, fm.size() * 4)); CU(cudaMalloc(&d_w0, w0.size() * 4)); CU(cudaMalloc(&d_b0, b0.size() * 4));
    CU(cudaMalloc(&d_w1, w1.size() * 4)); CU(cudaMalloc(&d_b1, b1.size() * 4)); CU(cudaMalloc(&d_w2, w2.size() * 4));
    CU(cudaMalloc(&d_b2, b2.size() * 4)); CU(cudaMalloc(&d_maf, maf.size() * 4)); CU(cudaMalloc(&d_pfm, pfm.size() * 4));
    CU(cudaMemcpy(d_fm, fm.data(), fm.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_w0, w0.data(), w0.size() * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_b0, b0.data(), b0.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_w1, w1.data(), w1.size() * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_b1, b1.data(), b1.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_w2, w2.data(), w2.size() * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_b2, b2.data(), b2.size() * 4, cudaMemcpyHostToDevice));
    whmr_maf_mlp_t mlp = nullptr;
    if (whmr_maf_mlp_create(256, 100, 64, 32, &mlp) == WHMR_OK) { printf("FAIL: unsupported MLP widths accepted\n"); return 1; }
    CK(whmr_maf_mlp_create(C0, C1, C2, C3, &mlp));
    if (whmr_sample_reduce(mlp, d_fm, WHMR_LAYOUT_NCHW, B, Hm, Wm, d_kp, 0, N, d_maf, nullptr, nullptr) == WHMR_OK) {
      printf("FAIL: sample_reduce before set_weights accepted\n"); return 1;
    }
    CK(whmr_maf_mlp_set_weights(mlp, d_w0, d_b0, d_w1, d_b1, d_w2, d_b2, nullptr));
    CK(whmr_project_sample_reduce(mlp, d_fm, WHMR_LAYOUT_NCHW, B, Hm, Wm, d_pts, d_cam, N, 1000.f, 256.f, 256.f, nullptr, d_maf,
                                  d_pfm, nullptr));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(maf.data(), d_maf, maf.size() * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(pfm.data(), d_pfm, pfm.size() * 4, cudaMemcpyDeviceToHost));
    double e5 = 0, e6 = 0, ymax = 0;
    std::vector<double> x(C0), y0(C1), y1(C2);
    for (int b = 0; b < B; ++b)
      for (int n = 0; n < N; ++n) {
        const double gx = kp[((size_t)b * N + n) * 2], gy = kp[((size_t)b * N + n) * 2 + 1];   // weak projection checked in 3.
        const double ix = (gx + 1.0) * 0.5 * (Wm - 1), iy = (gy + 1.0) * 0.5 * (Hm - 1);
        const int x0 = (int)std::floor(ix), yy0 = (int)std::floor(iy);
        for (int c = 0; c < C0; ++c) {
          double acc = 0;
          for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
              const int xx = x0 + dx, yy = yy0 + dy;
              if (xx < 0 || xx >= Wm || yy < 0 || yy >= Hm) continue;
              acc += (dx ? ix - x0 : 1.0 - (ix - x0)) * (dy ? iy - yy0 : 1.0 - (iy - yy0)) * fm[(((size_t)b * C0 + c) * Hm + yy) * Wm + xx];
            }
          x[c] = acc;
          e5 = std::fmax(e5, std::fabs(acc - pfm[((size_t)b * C0 + c) * N + n]));
        }
        for (int o = 0; o < C1; ++o) {      // models/maf_extractor.py:75-101
          double a = b0[o];
          for (int c = 0; c < C0; ++c) a += (double)w0[(size_t)o * C0 + c] * x[c];
          y0[o] = a > 0 ? a : 0.01 * a;
        }
        for (int o = 0; o < C2; ++o) {
          double a = b1[o];
          for (int c = 0; c < C1; ++c) a += (double)w1[(size_t)o * (C1 + C0) + c] * y0[c];
          for (int c = 0; c < C0; ++c) a += (double)w1[(size_t)o * (C1 + C0) + C1 + c] * x[c];
          y1[o] = a > 0 ? a : 0.01 * a;
        }
        for (int o = 0; o < C3; ++o) {
          double a = b2[o];
          for (int c = 0; c < C2; ++c) a += (double)w2[(size_t)o * (C2 + C0) + c] * y1[c];
          for (int c = 0; c < C0; ++c) a += (double)w2[(size_t)o * (C2 + C0) + C2 + c] * x[c];
          a = a > 0 ? a : 0.0;
          ymax = std::fmax(ymax, a);
          e6 = std::fmax(e6, std::fabs(a - maf[(size_t)b * C3 * N + (size_t)o * N + n]));
        }
      }
    printf("fused sampling + MLP: point features %.3g, mesh_align_feat %.3g (max %.3g)\n", e5, e6, ymax);
    if (!(e5 <= 1e-4) || !(e6 <= 1e-4 * ymax) || !(ymax > 0.1)) { printf("FAIL fused sampling + MLP\n"); return 1; }
    CK(whmr_maf_mlp_destroy(mlp));
  }
  CK(whmr_smpl_destroy(h));
  printf("launches: %llu\nABI SMOKE OK\n", (unsigned long long)whmr_launch_count());
  return 0;
}
