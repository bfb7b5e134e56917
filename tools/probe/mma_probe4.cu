// Micro-benchmark 4: does the weight-stationary form of tcgen05.mma (tcgen05.mma.ws, B held in a collector
// buffer across MMAs) lift the shared-memory operand bound of the M = time, N = C <= 64 mapping?
//   variant 0: plain tcgen05.mma                         (reads A 4 KB + B N*32 B from shared memory per MMA)
//   variant 1: tcgen05.mma.ws, no collector reuse
//   variant 2: tcgen05.mma.ws, B filled by the first tile's MMA of a (tap, k-step) and reused by the other tiles
//   variant 3: plain tcgen05.mma with collector::a::fill on every MMA (control)
// plus a numeric check of the .ws result layout in tensor memory (lane = row, column = n expected).
//   usage: mma_probe4 [N ...]     (default 64 128; pass 32 separately: it may not be a legal .ws shape)
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
#define MMA_ASM(NAME, OPC)                                                                                                   \
  __device__ __forceinline__ void NAME(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,  \
                                       uint32_t acc) {                                                                       \
    asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\n" \
                 OPC " [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)     \
                 : "memory");                                                                                                \
  }
MMA_ASM(mma_plain, "tcgen05.mma.cta_group::1.kind::f16")
MMA_ASM(mma_afill, "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill")
MMA_ASM(mma_ws, "tcgen05.mma.ws.cta_group::1.kind::f16")
MMA_ASM(mma_ws_fill, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill")
MMA_ASM(mma_ws_use, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::use")
MMA_ASM(mma_ws_last, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse")
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool wait(uint32_t bar, uint32_t par) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    if (clock64() - t0 > 2000000000LL) return false;
  }
  return true;
}
template <int V>
__device__ __forceinline__ void issue(int j, int nacc, uint32_t d, uint32_t alo, uint32_t hi, uint32_t blo, uint32_t idesc, uint32_t acc) {
  if (V == 0) mma_plain(d, alo, hi, blo, hi, idesc, acc);
  if (V == 1) mma_ws(d, alo, hi, blo, hi, idesc, acc);
  if (V == 2) {
    if (j == 0) mma_ws_fill(d, alo, hi, blo, hi, idesc, acc);
    else if (j == nacc - 1) mma_ws_last(d, alo, hi, blo, hi, idesc, acc);
    else mma_ws_use(d, alo, hi, blo, hi, idesc, acc);
  }
  if (V == 3) mma_afill(d, alo, hi, blo, hi, idesc, acc);
}
struct Cfg { int N, groups, pad; };
template <int KK, int NACC, int V>
__global__ void __launch_bounds__(128, 1) probe(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
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
  const uint32_t fin = smem_u32(&bars[0]);
  const uint32_t wst = smem_u32(smem) + 96 * 1024;
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | (8u << 24);
    const uint32_t rows_pad = 128 * NACC + c.pad;
    const uint32_t alo0 = (smem_u32(smem) >> 4) | (rows_pad << 16);
    const uint32_t blo0 = (wst >> 4) | ((uint32_t)c.N << 16);
    const uint32_t hi = (128u >> 4) | (1u << 14);
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < c.groups; ++it) {
        uint32_t alo = alo0 + (uint32_t)(it & 7) * 3u;
        uint32_t blo = blo0;
        _Pragma("unroll") for (int kk = 0; kk < KK; ++kk) {
          _Pragma("unroll") for (int j = 0; j < NACC; ++j)
            issue<V>(j, NACC, tm + (uint32_t)(j * c.N), alo + (uint32_t)j * 128u, hi, blo, idesc, (it | kk) != 0);
          alo += 2 * rows_pad;
          blo += 2 * (uint32_t)c.N;
        }
      }
      commit(fin);
    }
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

// ---- numeric check: D[128 x N] = A[128 x 32] * B[32 x N] (two K = 16 steps), two row tiles sharing B through the collector ----
template <int V>
__global__ void __launch_bounds__(128, 1) check(int N, const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A: 256 rows x 32 k -> [k/8][row][k%8];  B: N rows x 32 k -> [k/8][n][k%8]
  __nv_bfloat16* sa = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sb = reinterpret_cast<__nv_bfloat16*>(smem + 96 * 1024);
  for (int e = threadIdx.x; e < 256 * 32; e += 128) { const int r = e / 32, kx = e % 32; sa[((kx >> 3) * 256 + r) * 8 + (kx & 7)] = A[e]; }
  for (int e = threadIdx.x; e < N * 32; e += 128) { const int n = e / 32, kx = e % 32; sb[((kx >> 3) * N + n) * 8 + (kx & 7)] = B[e]; }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
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
  const uint32_t tm = tslot, fin = smem_u32(&bars[0]);
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
      const uint32_t hi = (128u >> 4) | (1u << 14);
      uint32_t alo = (smem_u32(sa) >> 4) | (256u << 16), blo = (smem_u32(sb) >> 4) | ((uint32_t)N << 16);
      for (int kk = 0; kk < 2; ++kk) {
        for (int j = 0; j < 2; ++j) issue<V>(j, 2, tm + (uint32_t)(j * N), alo + (uint32_t)j * 128u, hi, blo, idesc, kk != 0);
        alo += 2 * 256u;
        blo += 2 * (uint32_t)N;
      }
      commit(fin);
    }
    __syncwarp();
  }
  const bool ok = wait(fin, 0);
  if (!ok && threadIdx.x == 0) *status = 1;
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (ok) {
    for (int j = 0; j < 2; ++j)
      for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * N + c0)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) D[(size_t)(j * 128 + warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
      }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

template <int V>
void run_check(int N) {
  std::vector<__nv_bfloat16> A(256 * 32), B(N * 32);
  std::vector<float> Af(256 * 32), Bf(N * 32);
  for (int i = 0; i < 256 * 32; ++i) { Af[i] = (float)((i * 7 + (i / 32) * 3) % 17 - 8) * 0.125f; A[i] = __float2bfloat16(Af[i]); }
  for (int i = 0; i < N * 32; ++i) { Bf[i] = (float)((i * 5 + (i / 32)) % 13 - 6) * 0.25f; B[i] = __float2bfloat16(Bf[i]); }
  __nv_bfloat16 *dA, *dB; float* dD; int* dS;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 256 * N * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 256 * N * 4)); CK(cudaMemset(dS, 0, 4));
  CK(cudaFuncSetAttribute(check<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  check<V><<<1, 128, 160 * 1024>>>(N, dA, dB, dD, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("check N=%d variant %d: CUDA error %s\n", N, V, cudaGetErrorString(e)); exit(2); }
  std::vector<float> D(256 * N); int st = 0;
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  double worst = 0; int bad = 0;
  for (int r = 0; r < 256; ++r)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int kx = 0; kx < 32; ++kx) ref += Af[r * 32 + kx] * Bf[n * 32 + kx];
      const double err = fabs((double)D[r * N + n] - ref);
      if (!(err <= 1e-3)) ++bad;
      if (err > worst || err != err) worst = err;
    }
  printf("check N=%3d variant %d: timeout=%d  mismatches=%d of %d  worst=%g\n", N, V, st, bad, 256 * N, worst);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
}

template <int KK, int NACC, int V>
void run(int N, long long* d) {
  CK(cudaFuncSetAttribute(probe<KK, NACC, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  const int grid = 148;
  Cfg c{N, 264, 50};
  probe<KK, NACC, V><<<grid, 128, 160 * 1024>>>(c, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d variant %d: CUDA error %s\n", N, V, cudaGetErrorString(e)); exit(2); }
  long long h[296];
  CK(cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
  double iss = 0, comp = 0;
  for (int i = 0; i < grid; ++i) { iss += h[2 * i]; comp += h[2 * i + 1]; }
  const double nm = (double)c.groups * KK * NACC;
  printf("%5d %5d %3d  variant %d | %9.1f %9.1f\n", N, NACC, KK, V, iss / grid / nm, comp / grid / nm);
}
template <int KK, int NACC>
void run_all(int N, long long* d) {
  run_check<0>(N); run_check<1>(N); run_check<2>(N);
  run<KK, NACC, 0>(N, d); run<KK, NACC, 1>(N, d); run<KK, NACC, 2>(N, d); run<KK, NACC, 3>(N, d);
}
int main(int argc, char** argv) {
  long long* d;
  CK(cudaMalloc(&d, 148 * 2 * sizeof(long long)));
  printf("%5s %5s %3s            | %9s %9s (cycles per MMA)\n", "N", "nacc", "kk", "issue", "complete");
  std::vector<int> ns;
  for (int i = 1; i < argc; ++i) ns.push_back(atoi(argv[i]));
  if (ns.empty()) ns = {64, 128};
  for (int N : ns) {
    if (N == 32) { run_all<2, 4>(32, d); run_all<2, 8>(32, d); }
    else if (N == 64) { run_all<4, 4>(64, d); run_all<4, 2>(64, d); }
    else if (N == 128) run_all<4, 2>(128, d);
    else if (N == 256) run_all<4, 1>(256, d);
  }
  return 0;
}
