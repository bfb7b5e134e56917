// Micro-benchmark 5: tcgen05.mma.cta_group::2 (a CTA pair: M = 256 split over the two SMs, each CTA supplies its own 128 rows of A
// and HALF of B) against cta_group::1, for the M = time, N = C <= 128 shapes of the ResBlock kernels, plus a numeric check of the
// operand / accumulator placement (each CTA's accumulator = its own 128 rows x all N columns).
//   A: [k8][rows][8] bf16, K-major no-swizzle, same layout and shared-memory offset in both CTAs
//   B: CTA r holds columns r*N/2 .. (r+1)*N/2-1 as [k8][N/2][8]
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma2(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\n"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma1(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
}
// commit to the barrier at the same shared-memory offset in both CTAs of the pair
__device__ __forceinline__ void commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void commit1(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool wait(uint32_t bar, uint32_t par) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    if (clock64() - t0 > 1000000000LL) return false;
  }
  return true;
}

// PAIR = 1: cluster of two CTAs, leader issues cta_group::2 MMAs.  PAIR = 0: plain cta_group::1 (every CTA for itself).
template <int KK, int NACC, int PAIR>
__global__ void __launch_bounds__(128, 1) probe(int N, int groups, const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, long long* out, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tslot;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = PAIR ? (int)cluster.block_rank() : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = 128 * NACC, K = 16 * KK;
  const int nb = PAIR ? N / 2 : N;  // B columns held by this CTA
  __nv_bfloat16* sa = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sb = reinterpret_cast<__nv_bfloat16*>(smem + 96 * 1024);
  // A rows of this CTA: global rows [cta * rows, +rows); B columns [rank * nb, +nb)
  const int cta = blockIdx.x;
  for (int e = threadIdx.x; e < rows * K; e += 128) {
    const int r = e / K, kx = e % K;
    sa[((kx >> 3) * rows + r) * 8 + (kx & 7)] = A ? A[(size_t)((cta % 2) * rows + r) * K + kx] : __float2bfloat16(0.f);
  }
  for (int e = threadIdx.x; e < nb * K; e += 128) {
    const int n = e / K, kx = e % K;
    sb[((kx >> 3) * nb + n) * 8 + (kx & 7)] = B ? B[(size_t)(rank * nb + n) * K + kx] : __float2bfloat16(0.f);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  if (PAIR) cluster.sync(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tslot, fin = smem_u32(&bars[0]);
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 1) {
    t0 = clock64();
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
      const uint32_t hi = (128u >> 4) | (1u << 14);
      const uint32_t alo0 = (smem_u32(sa) >> 4) | ((uint32_t)rows << 16);
      const uint32_t blo0 = (smem_u32(sb) >> 4) | ((uint32_t)nb << 16);
      for (int it = 0; it < groups; ++it) {
        uint32_t alo = alo0, blo = blo0;
        _Pragma("unroll") for (int kk = 0; kk < KK; ++kk) {
          _Pragma("unroll") for (int j = 0; j < NACC; ++j) {
            if (PAIR) mma2(tm + (uint32_t)(j * N), alo + (uint32_t)j * 128u, hi, blo, hi, idesc, (it | kk) != 0);
            else mma1(tm + (uint32_t)(j * N), alo + (uint32_t)j * 128u, hi, blo, hi, idesc, (it | kk) != 0);
          }
          alo += 2 * (uint32_t)rows;
          blo += 2 * (uint32_t)nb;
        }
      }
      if (PAIR) commit2(fin); else commit1(fin);
    }
    __syncwarp();
    t1 = clock64();
  }
  const bool ok = wait(fin, 0);
  t2 = clock64();
  if (!ok && threadIdx.x == 0) *status = 1;
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 1 && lane == 0 && out) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  if (ok && D) {
    for (int j = 0; j < NACC; ++j)
      for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * N + c0)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) D[(size_t)((cta % 2) * rows + j * 128 + warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
      }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  if (PAIR) cluster.sync(); else __syncthreads();
  if (warp == 0) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
  }
}

template <int KK, int NACC, int PAIR>
void launch(int grid, int N, int groups, const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, long long* out, int* st) {
  CK(cudaFuncSetAttribute(probe<KK, NACC, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 160 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, probe<KK, NACC, PAIR>, N, groups, A, B, D, out, st));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d pair=%d: CUDA error %s\n", N, PAIR, cudaGetErrorString(e)); exit(2); }
}

template <int KK, int NACC>
void run(int N) {
  const int rows = 128 * NACC, K = 16 * KK;
  // ---- numeric check: one pair, D = A x B^T for the 2 * rows rows of the pair ----
  std::vector<__nv_bfloat16> A(2 * rows * K), B(N * K);
  std::vector<float> Af(A.size()), Bf(B.size());
  for (size_t i = 0; i < A.size(); ++i) { Af[i] = (float)((int)((i * 7 + (i / K) * 3) % 17) - 8) * 0.125f; A[i] = __float2bfloat16(Af[i]); }
  for (size_t i = 0; i < B.size(); ++i) { Bf[i] = (float)((int)((i * 5 + (i / K)) % 13) - 6) * 0.25f; B[i] = __float2bfloat16(Bf[i]); }
  __nv_bfloat16 *dA, *dB; float* dD; int* dS; long long* dT;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, (size_t)2 * rows * N * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMalloc(&dT, 148 * 2 * sizeof(long long)));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  for (int pair = 1; pair >= 0; --pair) {
    CK(cudaMemset(dD, 0xff, (size_t)2 * rows * N * 4)); CK(cudaMemset(dS, 0, 4));
    if (pair) launch<KK, NACC, 1>(2, N, 1, dA, dB, dD, nullptr, dS); else launch<KK, NACC, 0>(2, N, 1, dA, dB, dD, nullptr, dS);
    std::vector<float> D((size_t)2 * rows * N); int st = 0;
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    int bad = 0; double worst = 0;
    for (int r = 0; r < 2 * rows; ++r)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int kx = 0; kx < K; ++kx) ref += Af[(size_t)r * K + kx] * Bf[(size_t)n * K + kx];
        const double err = fabs((double)D[(size_t)r * N + n] - ref);
        if (!(err <= 1e-3)) ++bad;
        if (err > worst || err != err) worst = err;
      }
    printf("check N=%3d nacc=%d cta_group::%d: timeout=%d mismatches=%d of %d worst=%g\n", N, NACC, pair ? 2 : 1, st, bad, 2 * rows * N, worst);
  }
  // ---- rate: all 148 SMs ----
  for (int pair = 0; pair < 2; ++pair) {
    const int groups = 264;
    CK(cudaMemset(dS, 0, 4));
    if (pair) launch<KK, NACC, 1>(148, N, groups, nullptr, nullptr, nullptr, dT, dS); else launch<KK, NACC, 0>(148, N, groups, nullptr, nullptr, nullptr, dT, dS);
    long long h[296];
    CK(cudaMemcpy(h, dT, sizeof(h), cudaMemcpyDeviceToHost));
    double comp = 0; int n = 0;
    for (int i = 0; i < 148; ++i) if (!pair || i % 2 == 0) { comp += h[2 * i + 1]; ++n; }
    const double nm = (double)groups * KK * NACC;
    printf("%5d %5d %3d  cta_group::%d | %9.1f cycles per MMA instruction = %6.1f per 128 rows x N x 16\n", N, NACC, KK, pair ? 2 : 1,
           comp / n / nm, comp / n / nm / (pair ? 2 : 1));
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS); cudaFree(dT);
}
int main() {
  run<2, 4>(32);
  run<4, 4>(64);
  run<4, 2>(64);
  run<4, 2>(128);
  run<4, 1>(256);
  return 0;
}
