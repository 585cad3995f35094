// Store-throughput microbenchmark: how fast can N warps per SM write rows of `row_bytes` with a large
// stride between consecutive rows (the access pattern of the tcgen05 epilogues)?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int VEC>   // floats per lane per store: 1, 2, 4
__global__ void store_kernel(float* out, size_t row_stride_floats, int rows_per_warp, int iters) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  float* base = out + (size_t)warp * 32 * VEC + lane * VEC;   // each warp owns a 32*VEC-float column block
  float v = (float)lane;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 8
    for (int r = 0; r < rows_per_warp; ++r) {
      float* p = base + (size_t)r * row_stride_floats;
      if (VEC == 1) *p = v;
      else if (VEC == 2) *reinterpret_cast<float2*>(p) = make_float2(v, v);
      else *reinterpret_cast<float4*>(p) = make_float4(v, v, v, v);
    }
    v += 1.0f;
  }
}

template <int VEC>
void run(const char* name, float* buf, size_t stride, int warps_per_cta, int rows, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  store_kernel<VEC><<<148, warps_per_cta * 32>>>(buf, stride, rows, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  store_kernel<VEC><<<148, warps_per_cta * 32>>>(buf, stride, rows, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double bytes = 148.0 * warps_per_cta * rows * iters * 32 * VEC * 4;
  printf("%-28s warps/SM %2d  stride %8zu B  %7.1f GB/s  (%.3f ms, err=%s)\n", name, warps_per_cta, stride * 4,
         bytes / ms * 1e-6, ms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* buf;
  const size_t stride = 20736;   // floats: the pose-offset row pitch (82,944 B)
  cudaMalloc(&buf, stride * 4 * 768 + (1 << 20));
  for (int w : {4, 8, 16, 32}) {
    run<1>("STG.32  strided rows", buf, stride, w, 256, 8);
    run<2>("STG.64  strided rows", buf, stride, w, 256, 8);
    run<4>("STG.128 strided rows", buf, stride, w, 256, 8);
  }
  // same bytes, small stride (rows packed): is it the stride or the store width?
  for (int w : {4, 16}) {
    run<1>("STG.32  packed rows", buf, 148 * 32 * 32 * 1, w, 256, 8);
    run<4>("STG.128 packed rows", buf, 148 * 32 * 32 * 4, w, 64, 8);
  }
  return 0;
}
