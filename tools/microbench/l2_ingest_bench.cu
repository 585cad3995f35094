// L2 -> shared-memory ingest microbenchmark: how many bytes per second can ONE CTA per SM pull out of an
// L2-resident buffer with bulk async copies (cp.async.bulk, the TMA engine), as a function of the bytes in flight?
// This is the rate that bounds the operand streams of the tcgen05 kernels (posedirs / pose-feature / A tiles).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int STAGE_BYTES>
__global__ void __launch_bounds__(128, 1) ingest_kernel(const uint8_t* __restrict__ src, size_t src_bytes, int stages,
                                                        int copies_per_cta, int split, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * STAGE_BYTES);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t n_chunks = src_bytes / STAGE_BYTES;
    size_t chunk = ((size_t)blockIdx.x * 7919) % n_chunks;
    auto issue = [&](int s) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(STAGE_BYTES) : "memory");
      const int piece = STAGE_BYTES / split;
      for (int k = 0; k < split; ++k)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + (size_t)s * STAGE_BYTES + k * piece)), "l"(src + chunk * STAGE_BYTES + (size_t)k * piece),
                       "r"(piece), "r"(smem_u32(&bars[s])) : "memory");
      chunk += gridDim.x; if (chunk >= n_chunks) chunk -= n_chunks;
    };
    int issued = 0;
    for (; issued < stages && issued < copies_per_cta; ++issued) issue(issued);
    uint32_t phase = 0; int s = 0;
    for (int done = 0; done < copies_per_cta; ++done) {
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bars[s])), "r"(phase) : "memory");
      if (issued < copies_per_cta) { issue(s); ++issued; }
      if (++s == stages) { s = 0; phase ^= 1; }
    }
    if (sink && smem[0] == 123 && smem[STAGE_BYTES] == 77) *sink = 1;
  }
}

template <int STAGE_BYTES>
void run(const uint8_t* buf, size_t bytes, int stages, int split, int grid) {
  const int copies = (int)(((size_t)6 << 30) / STAGE_BYTES / grid);   // ~6 GB total
  const size_t smem = (size_t)stages * STAGE_BYTES + 64 * 8;
  cudaFuncSetAttribute(ingest_kernel<STAGE_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  ingest_kernel<STAGE_BYTES><<<grid, 128, smem>>>(buf, bytes, stages, copies / 8, split, nullptr);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  ingest_kernel<STAGE_BYTES><<<grid, 128, smem>>>(buf, bytes, stages, copies, split, nullptr);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double total = (double)copies * STAGE_BYTES * grid;
  printf("src %5zu MB  stage %3d KB x %d stages (%3d KB in flight, %d pieces/stage)  grid %3d : %7.1f GB/s  = %5.1f GB/s/SM  (%s)\n",
         bytes >> 20, STAGE_BYTES >> 10, stages, stages * STAGE_BYTES >> 10, split, grid, total / ms * 1e-6,
         total / ms * 1e-6 / grid, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  uint8_t* buf;
  const size_t big = (size_t)1 << 30;
  cudaMalloc(&buf, big);
  cudaMemset(buf, 1, big);
  for (size_t bytes : {(size_t)16 << 20, (size_t)64 << 20, big}) {   // L2-resident, L2-resident, DRAM
    for (int stages : {1, 2, 4, 6}) run<32768>(buf, bytes, stages, 1, 148);
    run<32768>(buf, bytes, 6, 8, 148);
    for (int stages : {2, 4, 8, 12}) run<16384>(buf, bytes, stages, 1, 148);
    run<98304>(buf, bytes, 2, 6, 148);
  }
  run<32768>(buf, (size_t)16 << 20, 6, 1, 74);
  run<32768>(buf, (size_t)16 << 20, 6, 1, 37);
  run<32768>(buf, (size_t)16 << 20, 6, 1, 8);
  return 0;
}
