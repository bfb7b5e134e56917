// Micro-benchmark 2: tcgen05.mma rate with the per-stage machinery of the conv kernels added step by step:
//   mode 0: bare issue loop               mode 1: + tcgen05.commit to an mbarrier every G MMAs
//   mode 2: + cp.async.bulk producer warp streaming `stage_bytes` per G MMAs through a full/empty ring
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
__device__ __forceinline__ void mma_pred(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc, uint32_t lead) {
  asm volatile("{\n.reg .pred p, q;\n.reg .b64 da, db;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\nsetp.ne.b32 q, %7, 0;\n"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(lead) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
}
struct Cfg { int N, nacc, kk, mode, groups, stages, swz, pad; };  // group = kk * nacc MMAs sharing one weight stage of kk*16 x N
template <int KK, int NACC>
__global__ void __launch_bounds__(320, 1) probe(Cfg c, const uint8_t* wsrc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[20];
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 320) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 20; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
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
  const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[8]), fin = smem_u32(&bars[16]);
  const uint32_t stage_bytes = (uint32_t)(c.kk * 16 * c.N * 2);
  const uint32_t wst = smem_u32(smem) + 96 * 1024;
  if (warp == 2 && c.mode == 99) {
    if (lane == 0)
      for (int it = 0; it < c.groups; ++it) {
        const int s = it % c.stages;
        wait(empty + 8 * s, ((it / c.stages) & 1) ^ 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * s), "r"(stage_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(wst + s * stage_bytes),
                     "l"(wsrc + (size_t)(it % 64) * stage_bytes), "r"(stage_bytes), "r"(full + 8 * s) : "memory");
      }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | (8u << 24);
    const uint32_t rows_pad = 128 * NACC + c.pad;
    const uint32_t alo0 = c.swz ? ((smem_u32(smem) >> 4) | (1u << 16)) : ((smem_u32(smem) >> 4) | (rows_pad << 16));
    const uint32_t blo0 = c.swz ? ((wst >> 4) | (1u << 16)) : ((wst >> 4) | ((uint32_t)c.N << 16));
    const uint32_t hi = c.swz ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : ((128u >> 4) | (1u << 14));
    const uint32_t a_kk = c.swz ? 2u : 2 * rows_pad, b_kk = c.swz ? 2u : 2 * (uint32_t)c.N;
    const uint32_t a_j = c.swz ? 1024u : 128u, a_it = c.swz ? 8u : 3u;
    long long t0 = clock64();
    if (c.mode == 0) {
      for (int it = 0; it < c.groups; ++it) {
        if (elect_one()) {
          uint32_t alo = alo0 + (uint32_t)(it & 7) * a_it;
          uint32_t blo = blo0;
          _Pragma("unroll") for (int kk = 0; kk < KK; ++kk) {
            _Pragma("unroll") for (int j = 0; j < NACC; ++j) mma(tm + (uint32_t)(j * c.N), alo + (uint32_t)j * a_j, hi, blo, hi, idesc, (it | kk) != 0);
            alo += a_kk;
            blo += b_kk;
          }
        }
        __syncwarp();
      }
    } else if (c.mode == 1) {
      if (elect_one()) {
        for (int it = 0; it < c.groups; ++it) {
          uint32_t alo = alo0 + (uint32_t)(it & 7) * a_it;
          uint32_t blo = blo0;
          _Pragma("unroll") for (int kk = 0; kk < KK; ++kk) {
            _Pragma("unroll") for (int j = 0; j < NACC; ++j) mma(tm + (uint32_t)(j * c.N), alo + (uint32_t)j * a_j, hi, blo, hi, idesc, (it | kk) != 0);
            alo += a_kk;
            blo += b_kk;
          }
        }
      }
      __syncwarp();
    } else {
      const uint32_t lead = lane == 0;
      for (int it = 0; it < c.groups; ++it) {
        uint32_t alo = alo0 + (uint32_t)(it & 7) * a_it;
        uint32_t blo = blo0;
        _Pragma("unroll") for (int kk = 0; kk < KK; ++kk) {
          _Pragma("unroll") for (int j = 0; j < NACC; ++j) mma_pred(tm + (uint32_t)(j * c.N), alo + (uint32_t)j * a_j, hi, blo, hi, idesc, (it | kk) != 0, lead);
          alo += a_kk;
          blo += b_kk;
        }
      }
    }
    if (elect_one()) commit(fin);
    __syncwarp();
    long long t1 = clock64();
    wait(fin, 0);
    long long t2 = clock64();
    if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
template <int KK, int NACC>
void run(int N, long long* d, const uint8_t* w) {
  CK(cudaFuncSetAttribute(probe<KK, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int swz = 0; swz < 2; ++swz)
    for (int mode = 0; mode < 3; ++mode) {
      const int grid = 148;
      Cfg c{N, NACC, KK, mode, 264, 4, swz, swz ? 0 : 50};
      probe<KK, NACC><<<grid, 320, 200 * 1024>>>(c, w, d);
      CK(cudaDeviceSynchronize());
      long long h[296];
      CK(cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
      double iss = 0, comp = 0;
      for (int i = 0; i < grid; ++i) { iss += h[2 * i]; comp += h[2 * i + 1]; }
      const double nm = (double)c.groups * KK * NACC;
      printf("%5d %5d %3d %5d  swz=%d | %9.1f %9.1f   %9.1f\n", N, NACC, KK, mode, swz, iss / grid / nm, comp / grid / nm, comp / grid / c.groups);
    }
}
int main() {
  long long* d; uint8_t* w;
  CK(cudaMalloc(&d, 148 * 2 * sizeof(long long)));
  CK(cudaMalloc(&w, 64 * 32768)); CK(cudaMemset(w, 0, 64 * 32768));
  printf("%5s %5s %3s %5s        | %9s %9s (cycles per MMA)   per group\n", "N", "nacc", "kk", "mode", "issue", "complete");
  run<2, 4>(32, d, w); run<4, 4>(64, d, w); run<4, 2>(128, d, w); run<4, 4>(128, d, w); run<4, 1>(256, d, w); run<4, 2>(256, d, w);
  return 0;
}
