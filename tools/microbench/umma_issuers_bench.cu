// tcgen05.mma microbenchmark, part 4 (round 2): P issuing threads: is the ~70-86-cycle cost of an M=128, K=16 MMA with N <= 128 a throughput
// floor or the latency of the accumulate dependency?  ONE issuing thread, consecutive MMAs rotate over R independent
// accumulators (R = 1: every MMA accumulates into the accumulator of its predecessor, the pattern of a plain K loop).
//   SS: A and B from shared memory (distinct A tiles);  TS: A from TMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_dep_bench umma_dep_bench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int R, bool TS>
__global__ void __launch_bounds__(256, 1) dep_kernel(int N, int iters, long long* out, int P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar4[4];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  uint64_t& bar = bar4[warp & 3];
  if (threadIdx.x == 0) {
    for (int q = 0; q < 4; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar4[q])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  if ((threadIdx.x & 31) == 0 && warp < P) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_base = smem_u32(smem);                 // 8 A tiles of 16 KB (128 rows x 128 B), K steps at +32 B
    const uint32_t b_base = a_base + 8 * 16384;             // 2 B tiles of 32 KB (256 rows x 128 B)
    const uint32_t a_t = tmem + 448;                        // TS: A operand columns (8 per K step), 448..511
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int ks = i & 3;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint32_t d = tmem + (uint32_t)((warp * R + r) * N);
        const uint64_t b = desc_sw128(b_base + (r & 1) * 32768 + ks * 32);
        if (TS) umma_ts(d, a_t + ((i * R + r) & 7) * 8, b, idesc, 1u);
        else umma_ss(d, desc_sw128(a_base + (((i >> 2) * R + r + 2 * warp) & 7) * 16384 + ks * 32), b, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    if (warp == 0) out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
  }
}


template <int R, bool TS>
double run(int N, int P, int sms, size_t smem, long long* out, int iters) {
  if (P * R * N > 448) return -1.0;
  cudaFuncSetAttribute(dep_kernel<R, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dep_kernel<R, TS><<<sms, 256, smem>>>(N, iters, out, P);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
  double s = 0;
  for (int i = 0; i < sms; ++i) s += (double)out[i];
  return s / sms / ((double)R * iters * P);
}
int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = 8 * 16384 + 2 * 32768 + 1024;
  long long* out;
  cudaMallocManaged(&out, sms * sizeof(long long));
  const int iters = 4000;
  printf("cycles per MMA (M=128, K=16, kind::f16), P issuing threads (one per warp, own accumulator each); ideal = N/2\n");
  for (int N : {16, 32, 48, 64, 96, 112, 128}) {
    printf("N=%3d (ideal %5.1f): SS P=1 %.1f P=2 %.1f P=3 %.1f P=4 %.1f | TS P=1 %.1f P=2 %.1f P=3 %.1f P=4 %.1f\n", N, N / 2.0,
           run<1, false>(N, 1, sms, smem, out, iters), run<1, false>(N, 2, sms, smem, out, iters), run<1, false>(N, 3, sms, smem, out, iters),
           run<1, false>(N, 4, sms, smem, out, iters),
           run<1, true>(N, 1, sms, smem, out, iters), run<1, true>(N, 2, sms, smem, out, iters), run<1, true>(N, 3, sms, smem, out, iters),
           run<1, true>(N, 4, sms, smem, out, iters));
  }
  return 0;
}
