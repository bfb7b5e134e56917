// Micro-benchmark: tcgen05.mma issue/execute rate on sm_100a for different smem layouts, N and
// accumulator interleavings.  Operand data is whatever is in shared memory (zeros); only timing matters.
//   usage: mma_probe            (prints a table)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
}

struct Cfg {
  int N;          // MMA N
  int swz;        // 0 = no swizzle (core matrices, LBO = rows*16), 1 = 128B swizzle K-major
  int nacc;       // accumulators cycled through by consecutive MMAs
  int a_rows;     // rows per K chunk in the no-swizzle A buffer (LBO = a_rows * 16)
  int iters;      // MMAs issued
  int a_step;     // 16-byte units added to the A start address per MMA group (no-swizzle), cycles through 8 positions
  int b_vary;     // B start address changes every nacc MMAs (like streaming weights)
  int poll;       // 0 none, 1 = 8 warps x 32 lanes spin on an mbarrier meanwhile, 2 = 8 warps x 1 lane
};

__global__ void __launch_bounds__(320, 1) probe(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 320) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tslot;
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | (8u << 24);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 128 * 1024;
    uint32_t alo, ahi, blo, bhi;
    if (c.swz) {
      alo = (a_base >> 4) | (1u << 16);
      blo = (b_base >> 4) | (1u << 16);
      ahi = bhi = (1024u >> 4) | (1u << 14) | (2u << 29);
    } else {
      alo = (a_base >> 4) | ((uint32_t)c.a_rows << 16);
      blo = (b_base >> 4) | ((uint32_t)c.N << 16);
      ahi = bhi = (128u >> 4) | (1u << 14);
    }
    long long t0 = clock64();
    if (elect_one()) {
      const int sh = c.nacc == 1 ? 0 : (c.nacc == 2 ? 1 : 2);
      const uint32_t jstep_a = c.swz ? 1024u : 128u, gstep_a = c.swz ? 2u : (uint32_t)c.a_step;
      const uint32_t gstep_b = c.swz ? 2u : (c.b_vary ? (uint32_t)(2 * c.N) : 0u);
#pragma unroll 8
      for (int i = 0; i < c.iters; ++i) {
        const uint32_t j = (uint32_t)i & (uint32_t)(c.nacc - 1);
        const uint32_t g = ((uint32_t)i >> sh) & (c.swz ? 3u : 7u);
        mma(tm + j * (uint32_t)c.N, alo + g * gstep_a + j * jstep_a, ahi, blo + g * gstep_b, bhi, idesc, i >= c.nacc);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) {
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar2)) : "memory");
    }
  } else if (warp >= 2 && c.poll) {
    if (c.poll == 1 || (threadIdx.x & 31) == 0) {
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar2)), "r"(0u) : "memory");
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 148 * 2 * sizeof(long long)));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  printf("%5s %4s %5s %6s %5s | %10s %10s  (cycles per MMA; floor = N/2)\n", "N", "swz", "nacc", "b_vary", "poll", "issue", "complete");
  for (int N : {32, 64, 128, 256})
    for (int nacc : {1, 2, 4}) {
      if (nacc * N > 512) continue;
      for (int b_vary : {0, 1})
        for (int poll : {0, 1, 2}) {
          Cfg c{N, 0, nacc, 563, 2048, 3, b_vary, poll};
          probe<<<148, 320, 200 * 1024>>>(c, d);
          CK(cudaDeviceSynchronize());
          long long h[296];
          CK(cudaMemcpy(h, d, 148 * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
          double iss = 0, comp = 0;
          for (int i = 0; i < 148; ++i) { iss += h[2 * i]; comp += h[2 * i + 1]; }
          printf("%5d %4d %5d %6d %5d | %10.1f %10.1f\n", N, 0, nacc, b_vary, poll, iss / 148 / c.iters, comp / 148 / c.iters);
        }
    }
  return 0;
}
