// Issue / throughput probe: packed fp32x2 (FFMA2 / FADD2) against scalar FFMA on sm_100a.
// Each warp runs ITER iterations of 8 independent accumulators; reports cycles per warp-instruction at 1, 2, 4, 8 warps/SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

__global__ void k_ffma(float* out, float a, float b) {
  float acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(t1 - t0) * 0.0f;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out + 65536)[0] = t1 - t0;
}

__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long acc[8], av, bv;
  float2 t;
  t.x = a; t.y = a; av = *reinterpret_cast<unsigned long long*>(&t);
  t.x = b; t.y = b; bv = *reinterpret_cast<unsigned long long*>(&t);
  for (int i = 0; i < 8; ++i) { t.x = threadIdx.x + i; t.y = threadIdx.x - i; acc[i] = *reinterpret_cast<unsigned long long*>(&t); }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) { t = *reinterpret_cast<float2*>(&acc[i]); s += t.x + t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out + 65536)[0] = t1 - t0;
}

int main() {
  float* out;
  cudaMalloc(&out, sizeof(float) * (65536 + 16));
  for (int warps = 4; warps <= 32; warps *= 2) {  // warps per CTA, one CTA per SM
    long long c1 = 0, c2 = 0;
    k_ffma<<<148, warps * 32>>>(out, 1.0001f, 0.5f);
    cudaMemcpy(&c1, out + 65536, 8, cudaMemcpyDeviceToHost);
    k_ffma2<<<148, warps * 32>>>(out, 1.0001f, 0.5f);
    cudaMemcpy(&c2, out + 65536, 8, cudaMemcpyDeviceToHost);
    // flops per SM per cycle: warps * 32 lanes * ITER * 16 fma * 2 / cycles
    printf("warps/SM %2d: FFMA %lld cyc (%.1f flop/clk/SM)   FFMA2 %lld cyc (%.1f flop/clk/SM)\n", warps, c1,
           warps * 32.0 * ITER * 16 * 2 / c1, c2, warps * 32.0 * ITER * 16 * 2 / c2);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
