// Weight gradient of a stride-1 tap-list convolution on the tcgen05 tensor cores.
//
//   G[j][ca][cb] = sum_{b,t} U[b, t + off_j, ca] * V[b, t, cb]        (grad.cuh, WgradArgs)
//
// is a GEMM whose reduction dimension is TIME: M = ca, N = cb, K = t.  Both operands are staged in
// the activation-tile layout of conv_tc.cu, [c/8][row][c%8] bf16 -- which, read with time as K, is the
// canonical no-swizzle MN-MAJOR UMMA layout: a core matrix is 8 consecutive rows (K) x 8 channels
// (16 B, MN), 128 contiguous bytes; K groups are 128 B apart (LBO), channel groups one tile plane
// apart (SBO).  A tap is a K shift of the A operand = its descriptor start address + off_j * 16 B, so
// one staged U tile (with halo) and one V tile feed every tap: the mirror image of the forward kernel,
// where a tap was an M shift.  Accumulators (one [128 x Cb] fp32 block per tap) stay in tensor memory
// over all row chunks of the CTA's split; partial sums go to the scratch that grad.cu's reduce kernel
// adds in a fixed order (bit-reproducible).  bf16 operands, fp32 accumulation.
#include "grad.cuh"
#include "tc_ptx.cuh"

#include <cstdlib>

namespace nvse {

namespace {

using namespace tc;

constexpr int kWgThreads = 256;
constexpr int kWgStageUnroll = 4;

struct WgTcKernelArgs {
  WgradArgs a;
  int min_off, span;
  int ru_pad, rv_pad;     // staged rows of the U / V tile, padded to an odd count (conflict-free staging stores)
  int tg, ngroups;        // taps per CTA (tg * Cb <= 256 TMEM columns: two CTAs share an SM's 512), tap groups
  int chunks_per_b;       // row chunks per utterance
  long long nchunks;      // B * chunks_per_b
  int cps;                // chunks per split (blockIdx.y)
  float* partial;         // [nsplit][ntaps][Ca][Cb]
};

// fp32 channels-last rows -> bf16 [c/8][row][c%8] plane; lanes run over the 8-channel groups of a row
__device__ __forceinline__ void wg_stage(uint8_t* tile, int rows_pad, int nrows, int ngr, const float* __restrict__ src_b,
                                         int t_first, int T, int C, int c0, float slope, int tid) {
  const int items = ngr * nrows;
  for (int e0 = tid; e0 < items; e0 += kWgThreads * kWgStageUnroll) {
    float4 f0[kWgStageUnroll], f1[kWgStageUnroll];
    int dst[kWgStageUnroll];
#pragma unroll
    for (int u = 0; u < kWgStageUnroll; ++u) {
      const int e = e0 + u * kWgThreads;
      const int grp = e % ngr, r = e / ngr;
      const int t = t_first + r;
      dst[u] = e < items ? (grp * rows_pad + r) * 16 : -1;
      f0[u] = f1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < items && t >= 0 && t < T) {
        const float4* src = reinterpret_cast<const float4*>(src_b + (int64_t)t * C + c0 + grp * 8);
        f0[u] = __ldg(src);
        f1[u] = __ldg(src + 1);
      }
    }
#pragma unroll
    for (int u = 0; u < kWgStageUnroll; ++u) {
      if (dst[u] < 0) continue;
      uint4 v;
      v.x = pack_bf16(lrelu(f0[u].x, slope), lrelu(f0[u].y, slope));
      v.y = pack_bf16(lrelu(f0[u].z, slope), lrelu(f0[u].w, slope));
      v.z = pack_bf16(lrelu(f1[u].x, slope), lrelu(f1[u].y, slope));
      v.w = pack_bf16(lrelu(f1[u].z, slope), lrelu(f1[u].w, slope));
      *reinterpret_cast<uint4*>(tile + dst[u]) = v;
    }
  }
}

// R = V rows (K) per chunk.  grid: (ca tiles * tap groups, splits)
template <int R>
__global__ void __launch_bounds__(kWgThreads, 2) wgrad_tc_kernel(const __grid_constant__ WgTcKernelArgs k) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const WgradArgs& a = k.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ca = a.Ca, Cb = a.Cb;
  const int ca_tile = (int)blockIdx.x / k.ngroups, grp_i = (int)blockIdx.x % k.ngroups;
  const int tap_lo = grp_i * k.tg;
  const int ntl = a.ntaps - tap_lo < k.tg ? a.ntaps - tap_lo : k.tg;
  const int ca0 = ca_tile * 128;
  const int nca = Ca - ca0 < 128 ? Ca - ca0 : 128;
  const int ngr_a = nca >> 3, ngr_b = Cb >> 3;
  const uint32_t u_bytes = (16u * (uint32_t)k.ru_pad * 16u + 127u) & ~127u;
  const uint32_t v_bytes = ((uint32_t)ngr_b * (uint32_t)k.rv_pad * 16u + 127u) & ~127u;
  uint8_t* u_tile = smem_raw;
  uint8_t* v_tile = smem_raw + u_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_tile + v_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);
  const uint32_t bar_mma = smem_u32(bars);
  uint32_t ncols = 32;
  while ((int)ncols < k.tg * Cb) ncols <<= 1;

  if (tid == 0) {
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), ncols);
  // channel groups above Ca (M is always 128): zero once, never written again
  for (int e = ngr_a * k.ru_pad + tid; e < 16 * k.ru_pad; e += kWgThreads)
    *reinterpret_cast<uint4*>(u_tile + (size_t)e * 16) = make_uint4(0u, 0u, 0u, 0u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long c_lo = (long long)blockIdx.y * k.cps;
  const long long c_hi = c_lo + k.cps < k.nchunks ? c_lo + k.cps : k.nchunks;
  // bits 4: fp32 accumulate; 7 / 10: bf16 A / B; 15 / 16: A / B MN-major; N >> 3 at 17, M >> 4 at 24
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(Cb >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t parity = 0, first = 1;
  bool ok = true;
  for (long long c = c_lo; c < c_hi && ok; ++c) {
    const long long b = c / k.chunks_per_b;
    const int t0 = (int)(c - b * k.chunks_per_b) * R;
    wg_stage(u_tile, k.ru_pad, R + k.span, ngr_a, a.U + b * a.u_bstride, t0 + k.min_off, a.Tu, Ca, ca0, a.u_slope, tid);
    wg_stage(v_tile, k.rv_pad, R, ngr_b, a.V + b * a.v_bstride, t0, a.Tv, Cb, 0, a.v_slope, tid);
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        const uint32_t a_lo0 = umma_desc_lo(smem_u32(u_tile), 128u);  // LBO: K groups (8 rows) are 128 B apart
        const uint32_t b_lo0 = umma_desc_lo(smem_u32(v_tile), 128u);
        const uint32_t a_hi = umma_desc_hi((uint32_t)k.ru_pad * 16u);  // SBO: 8-channel groups are one plane apart
        const uint32_t b_hi = umma_desc_hi((uint32_t)k.rv_pad * 16u);
        for (int tl = 0; tl < ntl; ++tl) {
          const uint32_t a_lo = a_lo0 + (uint32_t)(a.off[tap_lo + tl] - k.min_off);  // a tap = a K shift, in rows of 16 B
          const uint32_t d_tmem = tmem_base + (uint32_t)(tl * Cb);
#pragma unroll
          for (int kk = 0; kk < R / 16; ++kk)
            tc_mma_bf16_lohi(d_tmem, a_lo + (uint32_t)(kk * 16), a_hi, b_lo0 + (uint32_t)(kk * 16), b_hi, idesc,
                             (first && kk == 0) ? 0u : 1u);
        }
        tc_commit(bar_mma);  // arrives when every MMA above has read its operands and written TMEM
      }
      __syncwarp();
    }
    first = 0;
    ok = mbar_wait(bar_mma, parity);  // the tiles may be overwritten / the accumulators read
    parity ^= 1u;
  }
  tc_fence_after();

  // epilogue: TMEM lane = ca row, columns = (tap, cb)
  if (ok && c_lo < c_hi) {
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    for (int tl = 0; tl < ntl; ++tl) {
      float* pt = k.partial + (((int64_t)blockIdx.y * a.ntaps + tap_lo + tl) * Ca + ca0 + row) * Cb;
      for (int c0 = half * 16; c0 < Cb; c0 += 32) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tl * Cb + c0), v);
        tmem_ld_wait();
        if (row < nca) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(pt + c0 + 4 * q) = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                        __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

struct WgTcPlan {
  bool ok = false;
  int R = 0, min_off = 0, span = 0, ru_pad = 0, rv_pad = 0, tg = 0, ngroups = 0, chunks_per_b = 0, cps = 0, nsplit = 0;
  long long nchunks = 0;
  size_t smem = 0;
};

WgTcPlan wg_tc_plan(int Ca, int Cb, int ntaps, const int* off, int u_stride, int64_t B, int Tv) {
  WgTcPlan p;
  if (u_stride > 1 || Ca % 8 != 0 || Ca < 16 || !(Cb == 32 || Cb == 64 || Cb == 128 || Cb == 256) || B < 1 || Tv < 1) return p;
  int lo = 0, hi = 0;
  if (off) {
    lo = hi = off[0];
    for (int j = 1; j < ntaps; ++j) { lo = std::min(lo, off[j]); hi = std::max(hi, off[j]); }
  }
  p.min_off = lo; p.span = hi - lo;
  static const int env_r = [] { const char* e = std::getenv("NVSE_WG_R"); return e ? std::atoi(e) : 0; }();
  static const int env_cols = [] { const char* e = std::getenv("NVSE_WG_TMEM"); return e ? std::atoi(e) : 0; }();
  // measured (tools/wg_ab.sh, training step of HiFi-GAN V1): 128-row chunks + <= 256 TMEM columns per CTA -- two CTAs per
  // SM, each staging while the other multiplies -- beat 256 rows / 512 columns (one CTA per SM) by 10 % of the
  // wgrad time although every operand tile is then staged for twice as many tap groups
  p.R = 128;
  if (env_r == 128 || env_r == 256) p.R = (Cb > 128) ? 128 : env_r;
  p.ru_pad = (p.R + p.span) | 1;
  p.rv_pad = p.R | 1;
  const int cols = (env_cols == 256 || env_cols == 512) ? env_cols : 256;
  p.tg = std::max(1, std::min(ntaps, cols / Cb));
  p.ngroups = (ntaps + p.tg - 1) / p.tg;
  p.chunks_per_b = (Tv + p.R - 1) / p.R;
  p.nchunks = (long long)B * p.chunks_per_b;
  const long long base = (long long)((Ca + 127) / 128) * p.ngroups;
  long long want = std::max<long long>(1, std::min<long long>((296 + base - 1) / base, p.nchunks));
  p.cps = (int)((p.nchunks + want - 1) / want);
  p.nsplit = (int)((p.nchunks + p.cps - 1) / p.cps);
  const size_t u_bytes = ((size_t)16 * p.ru_pad * 16 + 127) & ~(size_t)127, v_bytes = ((size_t)(Cb / 8) * p.rv_pad * 16 + 127) & ~(size_t)127;
  p.smem = u_bytes + v_bytes + 64;
  p.ok = p.smem <= 200 * 1024 && p.nsplit <= 65535;
  return p;
}

}  // namespace

bool wgrad_tc_supported(int Ca, int Cb, int ntaps, const int* off, int u_stride, int64_t B, int Tv) {
  return wg_tc_plan(Ca, Cb, ntaps, off, u_stride, B, Tv).ok;
}

size_t wgrad_tc_scratch_elems(int Ca, int Cb, int ntaps, int64_t B, int Tv) {
  const WgTcPlan p = wg_tc_plan(Ca, Cb, ntaps, nullptr, 1, B, Tv);  // the split count does not depend on the offsets
  return p.ok ? (size_t)p.nsplit * ntaps * Ca * Cb : 0;
}

int launch_wgrad_tc(const WgradArgs& a, int64_t B, float* scratch, cudaStream_t st) {
  const WgTcPlan p = wg_tc_plan(a.Ca, a.Cb, a.ntaps, a.off, a.u_stride, B, a.Tv);
  NVSE_REQUIRE(p.ok, NVSE_ERR_UNSUPPORTED, "tensor-core wgrad: unsupported shape Ca=%d Cb=%d", a.Ca, a.Cb);
  NVSE_REQUIRE(a.U && a.V && a.dst && scratch, NVSE_ERR_INVALID, "tensor-core wgrad: null argument");
  WgTcKernelArgs k{};
  k.a = a;
  k.min_off = p.min_off; k.span = p.span; k.ru_pad = p.ru_pad; k.rv_pad = p.rv_pad; k.tg = p.tg; k.ngroups = p.ngroups;
  k.chunks_per_b = p.chunks_per_b; k.nchunks = p.nchunks; k.cps = p.cps; k.partial = scratch;
  {
    const double rows = (double)B * a.Tv;
    ProfScope prof("wgrad_tc", a.Ca, a.Cb, 2.0 * rows * a.Ca * a.Cb * a.ntaps, rows * 4.0 * (a.Ca + a.Cb) * p.ngroups, st);
    dim3 grid((unsigned)(((a.Ca + 127) / 128) * p.ngroups), (unsigned)p.nsplit);
    if (p.R == 256) {
      NVSE_CUDA_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
      wgrad_tc_kernel<256><<<grid, kWgThreads, p.smem, st>>>(k);
    } else {
      NVSE_CUDA_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
      wgrad_tc_kernel<128><<<grid, kWgThreads, p.smem, st>>>(k);
    }
    NVSE_LAUNCH_CHECK("wgrad_tc_kernel");
  }
  return launch_wgrad_reduce(scratch, p.nsplit, a.ntaps, a.Ca, a.Cb, a.dst, a.scale, st);
}

NVSE_TC_ABORT_IMPL(wgrad)

}  // namespace nvse
