// Micro-benchmark: L2 -> shared memory streaming with cp.async.bulk, the access pattern of the conv kernels'
// weight ring: every CTA (one per SM) pulls the SAME `set_kb` KB working set over and over in `chunk_kb` pieces
// with `depth` copies in flight.  Reports aggregate GB/s and bytes per SM clock.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
}
__global__ void __launch_bounds__(32, 1) stream(const uint8_t* src, int set_bytes, int chunk, int depth, int iters, int distinct, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    const uint8_t* base = src + (distinct ? (size_t)blockIdx.x * set_bytes : 0);
    const int per_set = set_bytes / chunk;
    long long t0 = clock64();
    for (int it = 0; it < iters + depth; ++it) {
      const int s = it % depth;
      if (it >= depth) wait(smem_u32(&bars[s]), ((it / depth) - 1) & 1);
      if (it < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)s * chunk)),
                     "l"(base + (size_t)(it % per_set) * chunk), "r"(chunk), "r"(smem_u32(&bars[s])) : "memory");
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  uint8_t* w; long long* d;
  const size_t total = (size_t)148 * 2048 * 1024;
  CK(cudaMalloc(&w, total)); CK(cudaMemset(w, 1, total));
  CK(cudaMalloc(&d, 148 * sizeof(long long)));
  CK(cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  printf("%8s %8s %6s %9s | %10s %12s %10s\n", "set_KB", "chunk_KB", "depth", "distinct", "ms", "GB/s", "B/clk/SM");
  for (int distinct : {0, 1})
    for (int set_kb : {360, 1440})
      for (int chunk_kb : {8, 16, 32})
        for (int depth : {2, 4, 6}) {
          if (chunk_kb * depth > 192) continue;
          const int iters = 4096;
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          stream<<<148, 32, 200 * 1024>>>(w, set_kb * 1024, chunk_kb * 1024, depth, 64, distinct, d);  // warm L2
          cudaEventRecord(e0);
          stream<<<148, 32, 200 * 1024>>>(w, set_kb * 1024, chunk_kb * 1024, depth, iters, distinct, d);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          long long h[148]; CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
          double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i];
          const double bytes = 148.0 * iters * chunk_kb * 1024;
          printf("%8d %8d %6d %9d | %10.3f %12.0f %10.1f\n", set_kb, chunk_kb, depth, distinct, ms, bytes / ms / 1e6, (double)iters * chunk_kb * 1024 / (cyc / 148));
        }
  return 0;
}
