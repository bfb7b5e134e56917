// Micro-benchmark: how fast can ONE CTA per SM with ~200 KB of shared memory (the tensor-core kernels' situation: little L1 left)
// stream a large fp32 array from HBM, by mechanism?
//   mode 0: ld.global.nc 128-bit, U loads in flight per thread        mode 1: 256-bit loads (LDG.E.256)
//   mode 2: cp.async.cg 16 B into a shared-memory ring                 mode 3: cp.async.bulk (TMA 1-D) of CHUNK bytes, D copies in flight
// Every thread/CTA consumes what it loaded (sum) so that nothing is optimised away.  usage: load_probe [smem_kb]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int U>
__global__ void __launch_bounds__(256, 1) k_ld128(const float4* __restrict__ x, size_t n4, float* out) {
  extern __shared__ uint8_t sm[];
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i + (U - 1) * stride < n4; i += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(x + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) out[0] = acc + sm[0];
}
template <int U>
__global__ void __launch_bounds__(256, 1) k_ld256(const float* __restrict__ x, size_t n8, float* out) {
  extern __shared__ uint8_t sm[];
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i + (U - 1) * stride < n8; i += U * stride) {
    float v[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u)
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[u][0]), "=f"(v[u][1]), "=f"(v[u][2]), "=f"(v[u][3]), "=f"(v[u][4]), "=f"(v[u][5]), "=f"(v[u][6]), "=f"(v[u][7]) : "l"(x + (i + u * stride) * 8));
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[u][j];
  }
  if (acc == 123.456f) out[0] = acc + sm[0];
}
// cp.async 16 B: each thread keeps D groups of 16 B in flight into its own ring slots
template <int D>
__global__ void __launch_bounds__(256, 1) k_cpasync(const float4* __restrict__ x, size_t n4, float* out) {
  extern __shared__ __align__(16) uint8_t sm[];
  float4* ring = reinterpret_cast<float4*>(sm);  // [D][256]
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * 256;
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t iters = n4 / stride;
  size_t issued = 0;
  for (; issued < D && issued < iters; ++issued) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(ring + (issued % D) * 256 + threadIdx.x)), "l"(x + i + issued * stride));
    asm volatile("cp.async.commit_group;");
  }
  for (size_t it = 0; it < iters; ++it) {
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1));
    const float4 v = ring[(it % D) * 256 + threadIdx.x];
    acc += v.x + v.y + v.z + v.w;
    if (issued < iters) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(ring + (issued % D) * 256 + threadIdx.x)), "l"(x + i + issued * stride));
      ++issued;
    }
    asm volatile("cp.async.commit_group;");
  }
  if (acc == 123.456f) out[0] = acc;
}
// TMA 1-D bulk copies: D slots of CHUNK bytes; thread 0 issues, all 256 threads consume a landed slot
template <int CHUNK, int D>
__global__ void __launch_bounds__(256, 1) k_bulk(const uint8_t* __restrict__ x, size_t bytes, float* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t full[D], empty[D];
  if (threadIdx.x == 0) {
    for (int s = 0; s < D; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const size_t nchunks = bytes / CHUNK, per = nchunks / gridDim.x;
  const size_t c0 = (size_t)blockIdx.x * per;
  float acc = 0.f;
  auto issue = [&](size_t c) {
    const int s = (int)(c % D);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"((uint32_t)CHUNK) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + (size_t)s * CHUNK)),
                 "l"(x + (c0 + c) * CHUNK), "r"((uint32_t)CHUNK), "r"(smem_u32(&full[s])) : "memory");
  };
  if (threadIdx.x == 0)
    for (size_t c = 0; c < D && c < per; ++c) issue(c);
  for (size_t c = 0; c < per; ++c) {
    const int s = (int)(c % D);
    const uint32_t par = (uint32_t)((c / D) & 1);
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&full[s])), "r"(par) : "memory");
    const float4* src = reinterpret_cast<const float4*>(sm + (size_t)s * CHUNK);
    for (int e = threadIdx.x; e < CHUNK / 16; e += 256) { const float4 v = src[e]; acc += v.x + v.y + v.z + v.w; }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    if (threadIdx.x == 0 && c + D < per) {
      ok = 0;
      while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&empty[s])), "r"(par) : "memory");
      issue(c + D);
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <typename F>
void timeit(const char* name, size_t bytes, F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 5; ++i) launch();
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%-44s %8.1f GB/s\n", name, bytes * 5.0 / (ms * 1e-3) / 1e9);
}
int main(int argc, char** argv) {
  const int smem_kb = argc > 1 ? atoi(argv[1]) : 200;
  const size_t bytes = (size_t)2 << 30;
  uint8_t* x; float* out;
  CK(cudaMalloc(&x, bytes)); CK(cudaMemset(x, 0, bytes)); CK(cudaMalloc(&out, 4));
  const int grid = 148; const size_t smem = (size_t)smem_kb * 1024;
  printf("one CTA of 256 threads per SM, %d KB dynamic shared memory, 2 GiB streamed\n", smem_kb);
#define L128(U) CK(cudaFuncSetAttribute(k_ld128<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
  timeit("ld.global.nc 128-bit, U=" #U, bytes, [&] { k_ld128<U><<<grid, 256, smem>>>((const float4*)x, bytes / 16, out); });
  L128(4) L128(8) L128(16)
#define L256(U) CK(cudaFuncSetAttribute(k_ld256<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
  timeit("ld.global.nc 256-bit, U=" #U, bytes, [&] { k_ld256<U><<<grid, 256, smem>>>((const float*)x, bytes / 32, out); });
  L256(2) L256(4) L256(8)
#define LCP(D) CK(cudaFuncSetAttribute(k_cpasync<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
  timeit("cp.async.cg 16 B, depth " #D, bytes, [&] { k_cpasync<D><<<grid, 256, smem>>>((const float4*)x, bytes / 16, out); });
  LCP(8) LCP(16) LCP(32)
#define LB(C, D) CK(cudaFuncSetAttribute(k_bulk<C, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
  timeit("cp.async.bulk " #C " B x " #D " in flight", bytes, [&] { k_bulk<C, D><<<grid, 256, smem>>>(x, bytes, out); });
  LB(8192, 4) LB(8192, 8) LB(8192, 16) LB(16384, 8) LB(32768, 4) LB(2048, 16) LB(512, 32)
  return 0;
}
