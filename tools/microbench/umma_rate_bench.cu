// tcgen05.mma issue-rate microbenchmark: cycles per MMA (M=128, K=16, kind::f16) as a function of N, of whether
// consecutive MMAs read different A tiles, and of the A-operand collector hints (.collector::a::fill / lastuse).
// One CTA per SM, one issuing thread, operands are whatever shared memory holds (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate_bench umma_rate_bench.cu
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
template <int C>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
#define M_ASM(COLL)                                                                                       \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                           \
               "tcgen05.mma.cta_group::1.kind::f16" COLL " [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), \
               "r"(idesc), "r"(acc) : "memory")
  if (C == 1) M_ASM(".collector::a::fill");
  else if (C == 2) M_ASM(".collector::a::use");
  else if (C == 3) M_ASM(".collector::a::lastuse");
  else M_ASM("");
#undef M_ASM
}

// mode 0: every MMA a different A tile (ring of 8 x 4 KB), different B tile
// mode 1: triplets (A0,B0) (A1,B1) (A1,B0) without hints      [the lo.hi / hi.lo / hi.hi pattern]
// mode 2: the same triplets with fill / lastuse on the A1 pair
// mode 3: same A tile for every MMA, no hints
// mode 4: same A tile for every MMA, fill then use
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_base = smem_u32(smem);                 // 8 A tiles of 16 KB (128 rows x 128 B), K steps at +32 B
    const uint32_t b_base = a_base + 8 * 16384;             // 2 B tiles of 32 KB (256 rows x 128 B)
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int ks = i & 3;
      const uint64_t a0 = desc_sw128(a_base + ((i >> 2) & 7) * 16384 + ks * 32);
      const uint64_t a1 = desc_sw128(a_base + (((i >> 2) + 4) & 7) * 16384 + ks * 32);
      const uint64_t b0 = desc_sw128(b_base + ks * 32), b1 = desc_sw128(b_base + 32768 + ks * 32);
      if (mode == 0) {
        umma<0>(tmem, a0, b0, idesc, 1u);
        umma<0>(tmem, a1, b1, idesc, 1u);
        umma<0>(tmem, desc_sw128(a_base + (((i >> 2) + 2) & 7) * 16384 + ks * 32), b0, idesc, 1u);
      } else if (mode == 1) {
        umma<0>(tmem, a0, b0, idesc, 1u);
        umma<0>(tmem, a1, b1, idesc, 1u);
        umma<0>(tmem, a1, b0, idesc, 1u);
      } else if (mode == 2) {
        umma<0>(tmem, a0, b0, idesc, 1u);
        umma<1>(tmem, a1, b1, idesc, 1u);
        umma<3>(tmem, a1, b0, idesc, 1u);
      } else if (mode == 3) {
        const uint64_t a = desc_sw128(a_base);
        umma<0>(tmem, a, b0, idesc, 1u);
        umma<0>(tmem, a, b1, idesc, 1u);
        umma<0>(tmem, a, b0, idesc, 1u);
      } else {
        const uint64_t a = desc_sw128(a_base);
        umma<1>(tmem, a, b0, idesc, 1u);
        umma<2>(tmem, a, b1, idesc, 1u);
        umma<2>(tmem, a, b0, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = 8 * 16384 + 2 * 32768 + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long* out;
  cudaMallocManaged(&out, sms * sizeof(long long));
  const int iters = 4000;
  const char* names[5] = {"distinct A", "lo/hi/hi, no hint", "lo/hi/hi, fill+lastuse", "same A, no hint", "same A, fill+use"};
  printf("cycles per MMA (M=128, K=16, kind::f16 bf16), %d SMs, %d MMAs per CTA; ideal = N/2\n", sms, 3 * iters);
  for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
    printf("N=%3d (ideal %5.1f):", N, N / 2.0);
    for (int mode = 0; mode < 5; ++mode) {
      rate_kernel<<<sms, 128, smem>>>(N, mode, iters, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" error %s\n", cudaGetErrorString(e)); return 1; }
      double s = 0;
      for (int i = 0; i < sms; ++i) s += (double)out[i];
      printf("  %s %.1f", names[mode], s / sms / (3.0 * iters));
    }
    printf("\n");
  }
  return 0;
}
