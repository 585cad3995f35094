// tcgen05.mma issue-rate microbenchmark, part 2 (round 2): cycles per MMA (M=128, K=16, kind::f16) as a function of N
//   mode 0: A from shared memory (SS), distinct A tiles                     [reference point, = umma_rate_bench mode 0]
//   mode 1: A from TMEM (TS: tcgen05.mma [d], [a_tmem], b_desc), distinct A columns per MMA
//   mode 2: TS, same A columns for every MMA
//   mode 3: SS, two issuing threads (warps 0 and 1) interleaving MMAs into disjoint accumulators (the fused kernel's
//           pose-blend + transform-blend pattern); cycles per MMA of the pair
// plus the TMEM read rate of the epilogue: cycles for 4 / 16 warps to pull C columns per lane with tcgen05.ld.32x32b.x32.
// One CTA per SM; operands are whatever shared memory / TMEM holds (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_ts_bench umma_ts_bench.cu
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
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t par) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
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
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t a_base = smem_u32(smem);                 // 8 A tiles of 16 KB (128 rows x 128 B), K steps at +32 B
  const uint32_t b_base = a_base + 8 * 16384;             // 2 B tiles of 32 KB (256 rows x 128 B)
  if (mode == 3) {
    if ((threadIdx.x & 31) == 0 && warp < 2) {
      const uint32_t d = tmem + warp * 256;
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int ks = i & 3;
        for (int r = 0; r < 3; ++r) {
          const uint64_t a = desc_sw128(a_base + (((i >> 2) + 2 * r + 4 * warp) & 7) * 16384 + ks * 32);
          umma_ss(d, a, desc_sw128(b_base + (r & 1) * 32768 + ks * 32), idesc, 1u);
        }
      }
      commit(&bar[warp]);
      wait_bar(&bar[warp], 0);
      if (warp == 0) out[blockIdx.x] = (clock64() - t0) / 2;   // two issuers: cycles per MMA of the pair
    }
  } else if (threadIdx.x == 0) {
    const uint32_t a_t = tmem + 256;    // A operand columns in TMEM (K=16 halfs = 8 columns per K step)
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int ks = i & 3;
      const uint64_t b0 = desc_sw128(b_base + ks * 32), b1 = desc_sw128(b_base + 32768 + ks * 32);
      if (mode == 0) {
        umma_ss(tmem, desc_sw128(a_base + ((i >> 2) & 7) * 16384 + ks * 32), b0, idesc, 1u);
        umma_ss(tmem, desc_sw128(a_base + (((i >> 2) + 4) & 7) * 16384 + ks * 32), b1, idesc, 1u);
        umma_ss(tmem, desc_sw128(a_base + (((i >> 2) + 2) & 7) * 16384 + ks * 32), b0, idesc, 1u);
      } else if (mode == 1) {
        umma_ts(tmem, a_t + ((3 * i) & 15) * 8, b0, idesc, 1u);
        umma_ts(tmem, a_t + ((3 * i + 1) & 15) * 8, b1, idesc, 1u);
        umma_ts(tmem, a_t + ((3 * i + 2) & 15) * 8, b0, idesc, 1u);
      } else {
        umma_ts(tmem, a_t, b0, idesc, 1u);
        umma_ts(tmem, a_t, b1, idesc, 1u);
        umma_ts(tmem, a_t, b0, idesc, 1u);
      }
    }
    commit(&bar[0]);
    wait_bar(&bar[0], 0);
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
  }
}

// TMEM -> register read rate: `nwarps` warps (4 per lane quarter max), each pulls `cols` columns per iteration
__global__ void __launch_bounds__(512, 1) ldtm_kernel(int cols, int iters, long long* out, float* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    for (int c = 0; c < cols; c += 32) {
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(base + (uint32_t)((c + i * 32) & 255)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int k = 0; k < 32; ++k) acc += __uint_as_float(v[k]);
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512u));
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = 8 * 16384 + 2 * 32768 + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long* out;
  float* sink;
  cudaMallocManaged(&out, sms * sizeof(long long));
  cudaMalloc(&sink, 4);
  const int iters = 4000;
  const char* names[4] = {"SS distinct A", "TS distinct A", "TS same A", "SS two issuers"};
  printf("cycles per MMA (M=128, K=16, kind::f16 bf16), %d SMs, %d MMAs per CTA; ideal = N/2\n", sms, 3 * iters);
  for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
    printf("N=%3d (ideal %5.1f):", N, N / 2.0);
    for (int mode = 0; mode < 4; ++mode) {
      rate_kernel<<<sms, 128, smem>>>(N, mode, iters, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" error %s\n", cudaGetErrorString(e)); return 1; }
      double s = 0;
      for (int i = 0; i < sms; ++i) s += (double)out[i];
      printf("  %s %.1f", names[mode], s / sms / (3.0 * iters));
    }
    printf("\n");
  }
  printf("tcgen05.ld.32x32b.x32 + wait: cycles per 32-column load per warp\n");
  for (int nw : {4, 8, 16}) {
    for (int cols : {32, 96}) {
      ldtm_kernel<<<sms, nw * 32>>>(cols, 2000, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" error %s\n", cudaGetErrorString(e)); return 1; }
      double s = 0;
      for (int i = 0; i < sms; ++i) s += (double)out[i];
      const double per = s / sms / (2000.0 * (cols / 32));
      printf("  warps=%2d cols=%3d: %.1f cycles per x32 load per warp  => %.1f B/clk/SM\n", nw, cols, per, nw * 32 * 32 * 4 / per);
    }
  }
  return 0;
}
