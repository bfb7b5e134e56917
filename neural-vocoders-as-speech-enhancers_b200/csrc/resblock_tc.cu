// Fused ResBlock1 chain on the 5th-gen tensor cores (sm_100a): Models/hifigan.py:43-50,
//
//   for m in pairs:  xt = c1_m(lrelu(x));  xt = c2_m(lrelu(xt));  x = xt + x
//
// as ONE kernel per (utterance, time tile).  A CTA pushes R = 128 * ntile rows through the whole
// chain ("overlapped tiling": every conv is evaluated on all R rows, rows closer to the tile edge
// than the receptive field go stale and only the central R - 2 * halo rows are written), so the
// residual stream and every intermediate stay on chip:
//
//   * X, the fp32 residual stream, lives in TENSOR MEMORY (ntile * C columns).  c2's MMAs are
//     issued with accumulate = 1 straight onto X, i.e. the tensor core performs the "+ x".  The
//     c2 biases are never added to X: X_true = X + cb_m with cb_m = b2_0 + .. + b2_m per channel,
//     which the epilogues add on the fly.
//   * ACC, the accumulator of c1, is a second TMEM region of the same size.
//   * OP, the bf16 operand buffer in shared memory, canonical no-swizzle K-major UMMA layout
//     [ci/8][row][ci%8] with P zero rows on either side.  A conv tap is a row shift of the A
//     descriptor's start address.  The epilogues overwrite OP in place:
//        load:  OP = bf16(lrelu(x))          epi1:  OP = bf16(lrelu(ACC + b1))
//        epi2:  OP = bf16(lrelu(X + cb_m))   (rows outside [0, T) forced to 0 = "same" padding)
//   * weights stream through an mbarrier ring of (tap, 64-channel K chunk) stages, one
//     cp.async.bulk per stage, each stage feeding the MMAs of all ntile tiles.
//
// Warp roles (320 threads): warps 0-7 load/epilogue (TMEM lane quarter = warp & 3, the two warps of
// a quarter split the 16-column chunks), warp 8 weight producer, warp 9 TMEM allocator + MMA issuer.
#include "resblock_tc.cuh"

#include <cuda_fp16.h>

#include <cstdlib>

#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace nvse {

namespace {

using namespace tc;

constexpr int kWorkWarps = 8;
constexpr int kThreads = (kWorkWarps + 2) * 32;
constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
constexpr int kNumBars = 2 * kMaxStages + 4;
// N = 64: weight slices shared by the row tiles of a CTA go through the tensor core's B collector (tc_ptx.cuh)
#ifndef NVSE_WS_B
#define NVSE_WS_B 1
#endif
constexpr bool kWsB = NVSE_WS_B != 0;
// ... but not in the pipelined kernel: same-box A/B (profiles/r02_ws_ab.txt) k = 7: 1.87 -> 2.01 ms, k = 11: 2.75 -> 2.90 ms WITH it
#ifndef NVSE_WS_PIPE
#define NVSE_WS_PIPE 0
#endif
constexpr bool kWsPipe = NVSE_WS_PIPE != 0;

struct RbKernelArgs {
  ResblockTcArgs a;
  int ntile;     // 128-row tiles per CTA
  int halo;      // rows of context the chain consumes on each side
  int V;         // 128 * ntile - 2 * halo: output rows one CTA produces
  int P;         // rows on each side of the operand buffer beyond the tile (largest conv padding)
  int P0;        // padding of the FIRST conv: that many rows beyond the tile on either side are loaded from
                 // global memory (real data), so the first conv costs no halo
  int rows_pad;  // operand buffer rows per 8-channel chunk (P + R + P)
  int stages, kc;
  int sm_count;
  int ntx, nsets;    // persistent pipelined kernel: tile sets per utterance and in all
  long long* trace;  // debug: clock64 stamps of one CTA's phase boundaries (NVSE_RB_TRACE), else null
};

__device__ __forceinline__ bool mbar_wait_warp(uint32_t bar, uint32_t parity) {
  // every lane polls (try_wait suspends in hardware); a timeout in any lane is seen by all
  const bool ok = mbar_wait(bar, parity);
  return __all_sync(0xffffffffu, ok);
}

// 16 activated values of one row -> two 16-byte core-matrix rows of the operand buffer, as bf16 or (F16) as IEEE
// half saturated to the finite range (three more mantissa bits for the c1 -> c2 intermediate of the last stage)
template <bool F16>
__device__ __forceinline__ void store_operand(uint8_t* op, int rows_pad, int brow, int c0, const float (&f)[16], float slope, bool keep) {
  uint32_t hi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a0 = keep ? lrelu(f[2 * i], slope) : 0.f, a1 = keep ? lrelu(f[2 * i + 1], slope) : 0.f;
    if (F16) {
      hi[i] = pack_f16(a0, a1);
    } else {
      hi[i] = pack_bf16(a0, a1);
    }
  }
  const size_t o0 = ((size_t)(c0 >> 3) * rows_pad + brow) * 16, o1 = o0 + (size_t)rows_pad * 16;
  *reinterpret_cast<uint4*>(op + o0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(op + o1) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
}

// C and the tile count are compile-time: the MMA issue loop must unroll completely -- a tcgen05.mma issued
// from a loop with run-time trip counts costs 60-200 cycles of issue (tools/probe/mma_probe3.cu), more
// than the MMA itself takes on the tensor core.
template <int C, int NT, bool H16>
__global__ void __launch_bounds__(kThreads, (2 * NT * C <= 256) ? 2 : 1) resblock_tc_kernel(const __grid_constant__ RbKernelArgs k) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const ResblockTcArgs& a = k.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int n = NT, R = n * kTileM, nchunk = C >> 3, KC = C < 64 ? C : 64;
  constexpr int TPS = C == 32 ? 4 : (C == 64 ? 2 : 1);  // taps per weight stage (nkc == 1 whenever TPS > 1)
  const int npairs = a.npairs;
  const int64_t b = blockIdx.y;
  const int64_t bstride = a.bstride;
  const int ws4 = a.t32 ? 32 : 1;  // float4 stride between the 4-channel groups of one row
  const int t_in0 = (int)blockIdx.x * k.V - k.halo;  // time index of tile row 0
  const int Tb = valid_rows(a.lens, b, a.T);            // this utterance's valid rows (ragged batches)
  const uint32_t op_bytes = (((uint32_t)nchunk * k.rows_pad * 16u) + 127u) & ~127u;
  constexpr uint32_t tap_bytes = (uint32_t)KC * C * 2u, stage_bytes = TPS * tap_bytes;
  uint8_t* op = smem_raw;
  uint8_t* wst = smem_raw + op_bytes;
  float* bsm = reinterpret_cast<float*>(wst + (size_t)k.stages * stage_bytes);  // b1[m][C] then cb[m][C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + 2 * kRbMaxPairs * C);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kMaxStages);
  const uint32_t bar_op = smem_u32(bars + 2 * kMaxStages), bar_acc = bar_op + 8, bar_h = bar_op + 16, bar_x = bar_op + 24;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * n * C)) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < k.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_op, kWorkWarps);
    mbar_init(bar_acc, 1);
    mbar_init(bar_h, kWorkWarps);
    mbar_init(bar_x, 1);
    fence_barrier_init();
  }
  if (warp == kWorkWarps + 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  // zero rows on both sides of every chunk of the operand buffer (never written again)
  {
    const int zr = 2 * k.P;
    for (int e = tid; e < nchunk * zr; e += kThreads) {
      const int chunk = e / zr, i = e - chunk * zr;
      const int row = i < k.P ? i : R + i;  // [0, P) and [P + R, 2P + R)
      *reinterpret_cast<uint4*>(op + ((size_t)chunk * k.rows_pad + row) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int e = tid; e < npairs * C; e += kThreads) {
      const int m = e / C, c = e - m * C;
      bsm[e] = __ldg(a.pair[m].b1 + c);
      float cb = 0.f;
      for (int i = 0; i <= m; ++i) cb += __ldg(a.pair[i].b2 + c);
      bsm[kRbMaxPairs * C + e] = cb;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_x = tmem_base, tmem_acc = tmem_base + (uint32_t)(n * C);
  constexpr int nkc = C / KC;
  const bool tracing = k.trace != nullptr && blockIdx.x == 3 && blockIdx.y == gridDim.y / 2 && lane == 0;  // a CTA of a middle wave
  int tr_i = 0;
#define RB_STAMP(role) do { if (tracing) k.trace[(role) * 64 + tr_i++] = clock64(); } while (0)

  if (warp == kWorkWarps) {
    // ===== weight producer: one bulk copy per (conv, tap, K chunk) stage =====
    if (lane == 0) {
      const uint32_t nstage = (uint32_t)k.stages;
      uint32_t s = 0, ph = 1;
      for (int m = 0; m < npairs; ++m)
        for (int half = 0; half < 2; ++half) {
          const __nv_bfloat16* wimg = half ? a.pair[m].w2 : a.pair[m].w1;
          for (int st = 0; st < a.k * nkc; st += TPS) {  // a stage = TPS consecutive (tap, K chunk) slices of the image
            if (!mbar_wait(bar_empty + 8 * s, ph)) goto done;
            const uint32_t cp_bytes = (uint32_t)min(TPS, a.k * nkc - st) * tap_bytes;
            mbar_arrive_expect_tx(bar_full + 8 * s, cp_bytes);
            bulk_copy_g2s(smem_u32(wst + (size_t)s * stage_bytes), wimg + (size_t)st * (tap_bytes / 2), cp_bytes, bar_full + 8 * s);
            if (++s == nstage) { s = 0; ph ^= 1u; }
          }
        }
    }
  } else if (warp == kWorkWarps + 1) {
    // ===== MMA issuer: ONE elected thread runs the whole loop.  Everything on its path between two
    // MMAs is a handful of instructions (no divisions, no votes): the tensor-core queue is only a few
    // MMAs deep, so issue-side latency beyond ~100 cycles per stage shows up as tensor idle time. =====
    if (elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      constexpr uint32_t idesc_f16 = idesc & ~((1u << 7) | (1u << 10));  // A and B formats 0 = IEEE half
      // descriptor low words advance in 16-byte units: one operand-buffer row = 1, one weight row = 1
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(op), (uint32_t)k.rows_pad * 16u);
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(wst), (uint32_t)C * 16u);
      const uint32_t hi = umma_desc_hi(128u);
      constexpr uint32_t stage_units = stage_bytes >> 4;
      constexpr int kkn = KC >> 4;
      const uint32_t nstage = (uint32_t)k.stages;
      uint32_t s = 0, ph = 0;  // ring slot and its phase parity
      long long tr_wait = 0;
      for (int m = 0; m < npairs; ++m)
        for (int half = 0; half < 2; ++half) {
          RB_STAMP(0);
          if (!mbar_wait(half ? bar_h : bar_op, m & 1)) goto mma_exit;
          tc_fence_after();
          RB_STAMP(0);
          const int d = half ? 1 : a.pair[m].dil;
          const uint32_t d_tmem = half ? tmem_x : tmem_acc;
          uint32_t acc = half ? 1u : 0u;  // c1 starts a fresh accumulator, c2 accumulates onto X
          uint32_t a_tap = a_lo0 + (uint32_t)(k.P - (a.k - 1) / 2 * d);
          for (int tap = 0; tap < a.k; tap += TPS) {
#pragma unroll
            for (int kc = 0; kc < (TPS > 1 ? 1 : nkc); ++kc) {
              const long long tw0 = tracing ? clock64() : 0;
              if (!mbar_wait(bar_full + 8 * s, ph)) goto mma_exit;
              if (tracing) tr_wait += clock64() - tw0;
              tc_fence_after();
              uint32_t b_lo = b_lo0 + s * stage_units;
#pragma unroll
              for (int tt = 0; tt < TPS; ++tt) {
                if (tap + tt < a.k) {
                  uint32_t a_lo = a_tap + (uint32_t)(kc * (KC >> 3)) * (uint32_t)k.rows_pad;
#pragma unroll
                  for (int kk = 0; kk < kkn; ++kk) {
#pragma unroll
                    for (int j = 0; j < n; ++j)
                      tc_mma_group<kWsB && C == 64, n>(j, d_tmem + (uint32_t)(j * C), a_lo + (uint32_t)(j * kTileM), hi, b_lo, hi,
                                                       (H16 && half) ? idesc_f16 : idesc, acc);
                    acc = 1u;
                    a_lo += 2u * (uint32_t)k.rows_pad;
                    b_lo += 2u * (uint32_t)C;
                  }
                  if (TPS > 1) a_tap += (uint32_t)d;
                }
              }
              tc_commit(bar_empty + 8 * s);
              if (++s == nstage) { s = 0; ph ^= 1u; }
            }
            if (TPS == 1) a_tap += (uint32_t)d;
          }
          tc_commit(half ? bar_x : bar_acc);
        }
      if (tracing) k.trace[63] = tr_wait;
    mma_exit:;
    }
    __syncwarp();
  } else {
    // ===== load / epilogue warps =====
    const int q = warp & 3, h = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const float slope = a.slope;
#define RB_WSTAMP() do { if (warp == 0) RB_STAMP(1); } while (0)
    RB_WSTAMP();
    // Work items of this warp: (tile j, 16-column chunk c0).  Global loads are issued U items ahead of
    // their use: the load and the final phase are pure memory latency otherwise.
    constexpr int CPW = C / 32;          // chunks per warp and tile
    constexpr int IT = n * CPW;          // items per warp
    constexpr int UMAX = (2 * NT * C <= 256) ? 2 : 4;  // two resident CTAs: half the registers each
    constexpr int U = IT < UMAX ? IT : UMAX;           // items in flight (U * 64 bytes per thread)
    // Rows just outside the tile that the FIRST conv reads: real data (zero outside the sequence) instead of
    // stale rows -- the first conv is then valid on the whole tile and the chain's halo excludes its padding.
    {
      const int n2 = 2 * k.P0;
      for (int e = (warp * 32 + lane); e < n2 * nchunk; e += kWorkWarps * 32) {
        const int chunk = e / n2, i = e - chunk * n2;
        const int orow = i < k.P0 ? k.P - k.P0 + i : k.P + R + (i - k.P0);  // operand-buffer row
        const int t = t_in0 + orow - k.P;
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (t >= 0 && t < Tb) {
          const float4* src = reinterpret_cast<const float4*>(a.x + b * bstride + (a.t32 ? t32_off(t, chunk * 8, C) : (int64_t)t * C + chunk * 8));
          const float4 f0 = __ldg(src), f1 = __ldg(src + ws4);
          pk.x = pack_bf16(lrelu(f0.x, slope), lrelu(f0.y, slope));
          pk.y = pack_bf16(lrelu(f0.z, slope), lrelu(f0.w, slope));
          pk.z = pack_bf16(lrelu(f1.x, slope), lrelu(f1.y, slope));
          pk.w = pack_bf16(lrelu(f1.z, slope), lrelu(f1.w, slope));
        }
        *reinterpret_cast<uint4*>(op + ((size_t)chunk * k.rows_pad + orow) * 16) = pk;
      }
    }
    // load: x -> X (TMEM, fp32) and OP = bf16(lrelu(x))
#pragma unroll
    for (int i0 = 0; i0 < IT; i0 += U) {
      float4 v[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u, c0 = ((i % CPW) * 2 + h) * 16;
        const int r = (i / CPW) * kTileM + q * 32 + lane, t = t_in0 + r;
        const float4* src = reinterpret_cast<const float4*>(a.x + b * bstride + (a.t32 ? t32_off(t, c0, C) : (int64_t)t * C + c0));
        const bool inb = t >= 0 && t < Tb;
#pragma unroll
        for (int w = 0; w < 4; ++w) v[u][w] = inb ? __ldg(src + w * ws4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u, c0 = ((i % CPW) * 2 + h) * 16, jt = i / CPW;
        const int r = jt * kTileM + q * 32 + lane, t = t_in0 + r;
        float f[16];
        uint32_t bits[16];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          f[4 * w] = v[u][w].x; f[4 * w + 1] = v[u][w].y; f[4 * w + 2] = v[u][w].z; f[4 * w + 3] = v[u][w].w;
        }
#pragma unroll
        for (int w = 0; w < 16; ++w) bits[w] = __float_as_uint(f[w]);
        tmem_st_32x16(tmem_x + lane_sel + (uint32_t)(jt * C + c0), bits);
        store_operand<false>(op, k.rows_pad, k.P + r, c0, f, slope, t >= 0 && t < Tb);
      }
    }
    tmem_st_wait();
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_op);
    RB_WSTAMP();
    // While the tensor core works on the first conv: pull the x tile of the CTA that will be scheduled one
    // wave from now into L2 (blocks are dispatched in linear order), so that its load phase hits L2.
    if (a.t32) {
      const unsigned resident = (unsigned)k.sm_count * ((2 * NT * C <= 256) ? 2u : 1u);
      const unsigned lin = blockIdx.y * gridDim.x + blockIdx.x + resident;
      const unsigned nb = lin / gridDim.x, nx = lin - nb * gridDim.x;
      if (nb < gridDim.y && (lane & 7) == 0) {
        const int nt0 = (int)nx * k.V - k.halo;
#pragma unroll
        for (int i = 0; i < IT; ++i) {
          const int c0 = ((i % CPW) * 2 + h) * 16, t = nt0 + (i / CPW) * kTileM + q * 32 + lane;
          if (t >= 0 && t < Tb) {
            const float* src = a.x + (int64_t)nb * bstride + t32_off(t, c0, C);
#pragma unroll
            for (int w = 0; w < 4; ++w) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + w * 128));
          }
        }
      }
    }

    for (int m = 0; m < npairs; ++m) {
      // epi1: OP = bf16(lrelu(ACC + b1_m)), zero outside the sequence
      if (!mbar_wait_warp(bar_acc, m & 1)) break;
      tc_fence_after();
      RB_WSTAMP();
      const float* b1 = bsm + m * C;
      {
        // TMEM loads run one item ahead of the conversion (tcgen05.wait::ld covers every load in flight)
        uint32_t v[2][16];
        tmem_ld_32x16(tmem_acc + lane_sel + (uint32_t)(h * 16), v[0]);
#pragma unroll
        for (int i = 0; i < IT; ++i) {
          const int c0 = ((i % CPW) * 2 + h) * 16, jt = i / CPW;
          const int r = jt * kTileM + q * 32 + lane, t = t_in0 + r;
          tmem_ld_wait();
          if (i + 1 < IT)
            tmem_ld_32x16(tmem_acc + lane_sel + (uint32_t)(((i + 1) / CPW) * C + (((i + 1) % CPW) * 2 + h) * 16), v[(i + 1) & 1]);
          float f[16];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 bq = *reinterpret_cast<const float4*>(b1 + c0 + 4 * u);
            f[4 * u] = __uint_as_float(v[i & 1][4 * u]) + bq.x;
            f[4 * u + 1] = __uint_as_float(v[i & 1][4 * u + 1]) + bq.y;
            f[4 * u + 2] = __uint_as_float(v[i & 1][4 * u + 2]) + bq.z;
            f[4 * u + 3] = __uint_as_float(v[i & 1][4 * u + 3]) + bq.w;
          }
          store_operand<H16>(op, k.rows_pad, k.P + r, c0, f, slope, t >= 0 && t < Tb);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_h);
      RB_WSTAMP();

      if (!mbar_wait_warp(bar_x, m & 1)) break;
      tc_fence_after();
      RB_WSTAMP();
      const float* cb = bsm + (kRbMaxPairs + m) * C;
      if (m + 1 < npairs) {
        // epi2: OP = bf16(lrelu(X + cb_m)): the input of the next pair
        {
          uint32_t v[2][16];
          tmem_ld_32x16(tmem_x + lane_sel + (uint32_t)(h * 16), v[0]);
#pragma unroll
          for (int i = 0; i < IT; ++i) {
            const int c0 = ((i % CPW) * 2 + h) * 16, jt = i / CPW;
            const int r = jt * kTileM + q * 32 + lane, t = t_in0 + r;
            tmem_ld_wait();
            if (i + 1 < IT)
              tmem_ld_32x16(tmem_x + lane_sel + (uint32_t)(((i + 1) / CPW) * C + (((i + 1) % CPW) * 2 + h) * 16), v[(i + 1) & 1]);
            float f[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 bq = *reinterpret_cast<const float4*>(cb + c0 + 4 * u);
              f[4 * u] = __uint_as_float(v[i & 1][4 * u]) + bq.x;
              f[4 * u + 1] = __uint_as_float(v[i & 1][4 * u + 1]) + bq.y;
              f[4 * u + 2] = __uint_as_float(v[i & 1][4 * u + 2]) + bq.z;
              f[4 * u + 3] = __uint_as_float(v[i & 1][4 * u + 3]) + bq.w;
            }
            store_operand<false>(op, k.rows_pad, k.P + r, c0, f, slope, t >= 0 && t < Tb);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_op);
        RB_WSTAMP();
      } else {
        // final: y = [y +] out_scale * (X + cb_last) for the central V rows.  The product is rounded before the sum (no FMA): the
        // MRF total is then (r0/3 + r1/3) + r2/3 with every term rounded, bit-identical to the small-job mode, where the three
        // ResBlocks write their own buffers and add3 sums them -- a batch's results must not depend on which mode its size selects
#pragma unroll
        for (int i0 = 0; i0 < IT; i0 += U) {
          float4 yq[U][4];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = i0 + u, c0 = ((i % CPW) * 2 + h) * 16;
            const int r = (i / CPW) * kTileM + q * 32 + lane, t = t_in0 + r;
            const bool valid = r >= k.halo && r < R - k.halo && t < Tb;
            const float4* src = reinterpret_cast<const float4*>(a.y + b * bstride + (a.t32 ? t32_off(t, c0, C) : (int64_t)t * C + c0));
#pragma unroll
            for (int w = 0; w < 4; ++w) yq[u][w] = (valid && a.accumulate) ? src[w * ws4] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = i0 + u, c0 = ((i % CPW) * 2 + h) * 16, jt = i / CPW;
            const int r = jt * kTileM + q * 32 + lane, t = t_in0 + r;
            const bool valid = r >= k.halo && r < R - k.halo && t < Tb;
            uint32_t v[16];
            tmem_ld_32x16(tmem_x + lane_sel + (uint32_t)(jt * C + c0), v);
            tmem_ld_wait();
            if (valid) {
              float4* dst = reinterpret_cast<float4*>(a.y + b * bstride + (a.t32 ? t32_off(t, c0, C) : (int64_t)t * C + c0));
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const float4 bq = *reinterpret_cast<const float4*>(cb + c0 + 4 * w);
                float4 o;
                o.x = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * w]) + bq.x, a.out_scale), yq[u][w].x);
                o.y = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * w + 1]) + bq.y, a.out_scale), yq[u][w].y);
                o.z = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * w + 2]) + bq.z, a.out_scale), yq[u][w].z);
                o.w = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * w + 3]) + bq.w, a.out_scale), yq[u][w].w);
                dst[w * ws4] = o;
              }
            }
          }
        }
      }
    }
  }
done:
  if (warp == 0) RB_STAMP(1);
  if (warp == kWorkWarps + 1) RB_STAMP(0);
  tc_fence_before();
  __syncthreads();
  if (warp == kWorkWarps + 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------
// Software-pipelined, PERSISTENT variant of the chain kernel (C <= 64, NT even): the NT tiles of a tile set are
// worked on as two halves A and B.  The MMA thread issues conv l on A, then on B, then conv l+1 on A, ...; the
// epilogue of (l, A) runs while the tensor core works on (l, B), the epilogue of (l, B) while it works on (l+1, A):
// except for the few rows of B that conv l+1 on A reads across the A/B boundary, the tensor core never waits
// for an epilogue.  What that takes:
//   * two operand buffers: the input S_l of conv l lives in OP[l & 1], its epilogue writes S_(l+1) into
//     OP[(l+1) & 1] -- half A may not overwrite rows conv l still reads for half B;
//   * one mbarrier per TILE ("S_l of tile j is written"): conv l on a half waits for its own tiles and the
//     neighbouring tile of the other half; one "accumulator ready" mbarrier per half;
//   * the weights of every conv stream through the ring twice (once per half): L2 traffic x2, which at
//     C <= 64 is ~20 B/clk/SM.
// Persistent: a CTA walks tile sets first, first + gridDim.x, ... (consecutive CTAs take consecutive tile sets of an
// utterance).  With X and ACC filling tensor memory there is one CTA per SM at C = 64, so nothing else would hide the
// load / final phases (16 % of a k = 11 block, 22 % at k = 7: profiles/r02_pipe_trace.txt).  The epilogue warps
// therefore bring in tile j of the NEXT tile set right after writing out tile j of this one -- half A under the MMAs
// of the last conv on half B, half B under the next set's first MMAs -- with the global loads of the next tile issued
// BEFORE the final phase of the current one, so that their (DRAM) latency passes under it.  Every resource such an
// early load touches (the X columns and OP[0] rows of that tile) was last used before the tile's final accumulator
// completed.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPipeBars = 2 * kMaxStages + 8 + 2;

template <int C, int NT, bool H16>
__global__ void __launch_bounds__(kThreads, (2 * NT * C <= 256) ? 2 : 1) resblock_pipe_kernel(const __grid_constant__ RbKernelArgs k) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  static_assert(NT % 2 == 0 && NT <= 8, "the tiles are split into two halves");
  const ResblockTcArgs& a = k.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int n = NT, HN = NT / 2, R = n * kTileM, nchunk = C >> 3, KC = C < 64 ? C : 64;
  constexpr int nkc = C / KC, kkn = KC >> 4;
  constexpr int TPS = C == 32 ? 4 : 2;  // taps per weight stage (nkc == 1 at C <= 64)
  static_assert(nkc == 1, "the pipelined kernel is for C <= 64");
  const int npairs = a.npairs, L = 2 * npairs;
  const int64_t bstride = a.bstride;
  const int ws4 = a.t32 ? 32 : 1;
  const int ntx = k.ntx;      // tile sets per utterance
  const int nsets = k.nsets;  // tile sets in all (ntx * B)
  const uint32_t op_bytes = (((uint32_t)nchunk * k.rows_pad * 16u) + 127u) & ~127u;
  constexpr uint32_t tap_bytes = (uint32_t)KC * C * 2u, stage_bytes = TPS * tap_bytes;
  uint8_t* op0 = smem_raw;  // OP[0], OP[1]
  uint8_t* wst = smem_raw + 2 * op_bytes;
  float* bsm = reinterpret_cast<float*>(wst + (size_t)k.stages * stage_bytes);  // b1[m][C] then cb[m][C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + 2 * kRbMaxPairs * C);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kPipeBars);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kMaxStages);
  const uint32_t bar_tile = smem_u32(bars + 2 * kMaxStages);  // [NT]  S_l of tile j written   (8 warps)
  const uint32_t bar_accf = bar_tile + 8 * 8;                 // [2]   accumulator of a half ready (commit)
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * n * C)) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < k.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int j = 0; j < n; ++j) mbar_init(bar_tile + 8 * j, kWorkWarps);
    mbar_init(bar_accf, 1);
    mbar_init(bar_accf + 8, 1);
    fence_barrier_init();
  }
  if (warp == kWorkWarps + 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  {
    const int zr = 2 * k.P;
    for (int e = tid; e < 2 * nchunk * zr; e += kThreads) {
      const int bufi = e / (nchunk * zr), e2 = e - bufi * nchunk * zr;
      const int chunk = e2 / zr, i = e2 - chunk * zr;
      const int row = i < k.P ? i : R + i;
      *reinterpret_cast<uint4*>(op0 + (size_t)bufi * op_bytes + ((size_t)chunk * k.rows_pad + row) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int e = tid; e < npairs * C; e += kThreads) {
      const int m = e / C, c = e - m * C;
      bsm[e] = __ldg(a.pair[m].b1 + c);
      float cb = 0.f;
      for (int i = 0; i <= m; ++i) cb += __ldg(a.pair[i].b2 + c);
      bsm[kRbMaxPairs * C + e] = cb;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_x = tmem_base, tmem_acc = tmem_base + (uint32_t)(n * C);
  const bool tracing = k.trace != nullptr && blockIdx.x == 3 && lane == 0;
  int tr_i = 0;
#define RP_STAMP(role) do { if (tracing && tr_i < 60) k.trace[(role) * 64 + tr_i++] = clock64(); } while (0)
  const int first = (int)blockIdx.x, step = (int)gridDim.x;
  const int my_sets = first < nsets ? (nsets - first + step - 1) / step : 0;

  if (warp == kWorkWarps) {
    // ===== weight producer: every conv's stages twice (half A, half B), for every tile set of this CTA =====
    if (lane == 0) {
      const uint32_t nstage = (uint32_t)k.stages;
      uint32_t s = 0, ph = 1;
      for (int it = 0; it < my_sets; ++it)
        for (int l = 0; l < L; ++l) {
          const __nv_bfloat16* wimg = (l & 1) ? a.pair[l >> 1].w2 : a.pair[l >> 1].w1;
          for (int hf = 0; hf < 2; ++hf)
            for (int st = 0; st < a.k; st += TPS) {
              if (!mbar_wait(bar_empty + 8 * s, ph)) goto done;
              const uint32_t cp_bytes = (uint32_t)min(TPS, a.k - st) * tap_bytes;
              mbar_arrive_expect_tx(bar_full + 8 * s, cp_bytes);
              bulk_copy_g2s(smem_u32(wst + (size_t)s * stage_bytes), wimg + (size_t)st * (tap_bytes / 2), cp_bytes, bar_full + 8 * s);
              if (++s == nstage) { s = 0; ph ^= 1u; }
            }
        }
    }
  } else if (warp == kWorkWarps + 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      constexpr uint32_t idesc_f16 = idesc & ~((1u << 7) | (1u << 10));
      const uint32_t a_lo_even = umma_desc_lo(smem_u32(op0), (uint32_t)k.rows_pad * 16u);
      const uint32_t a_lo_odd = umma_desc_lo(smem_u32(op0 + op_bytes), (uint32_t)k.rows_pad * 16u);
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(wst), (uint32_t)C * 16u);
      const uint32_t hi = umma_desc_hi(128u);
      constexpr uint32_t stage_units = stage_bytes >> 4;
      const uint32_t nstage = (uint32_t)k.stages;
      uint32_t s = 0, ph = 0;
      uint32_t tphase = 0;  // phase of the tile / accumulator barriers: one per conv, L per tile set
      for (int it = 0; it < my_sets; ++it)
        for (int l = 0; l < L; ++l, ++tphase) {
          const int c2 = l & 1;
          const int d = c2 ? 1 : a.pair[l >> 1].dil;
          const uint32_t d_tmem = c2 ? tmem_x : tmem_acc;
          const uint32_t id = (H16 && c2) ? idesc_f16 : idesc;
          const uint32_t a_buf = (c2 ? a_lo_odd : a_lo_even) + (uint32_t)(k.P - (a.k - 1) / 2 * d);
          for (int hf = 0; hf < 2; ++hf) {
            // S_l of this half's tiles and of the adjacent tile of the other half.  Half A reads its right-hand
            // neighbour (tile HN, the first tile of half B, whose epilogue only starts when conv l-1 on B is complete)
            // through the taps with a positive row offset alone: that wait is deferred to the first such tap, so the
            // taps with offsets <= 0 run under the neighbour's epilogue instead of after it.
            const int j_lo = hf ? HN - 1 : 0, j_hi = hf ? n - 1 : HN - 1;
            for (int j = j_lo; j <= j_hi; ++j)
              if (!mbar_wait(bar_tile + 8 * j, tphase & 1u)) goto mma_exit;
            bool right_ready = hf != 0;
            tc_fence_after();
            RP_STAMP(0);
            uint32_t acc = c2 ? 1u : 0u;
            uint32_t a_tap = a_buf + (uint32_t)(hf * HN * kTileM);
            for (int tap = 0; tap < a.k; tap += TPS) {
              if (!right_ready && tap + TPS - 1 > (a.k - 1) / 2) {
                if (!mbar_wait(bar_tile + 8 * HN, tphase & 1u)) goto mma_exit;
                tc_fence_after();
                right_ready = true;
              }
              if (!mbar_wait(bar_full + 8 * s, ph)) goto mma_exit;
              tc_fence_after();
              uint32_t b_lo = b_lo0 + s * stage_units;
#pragma unroll
              for (int tt = 0; tt < TPS; ++tt) {
                if (tap + tt < a.k) {
                  uint32_t a_lo = a_tap;
#pragma unroll
                  for (int kk = 0; kk < kkn; ++kk) {
#pragma unroll
                    for (int j = 0; j < HN; ++j)
                      tc_mma_group<kWsPipe && C == 64, HN>(j, d_tmem + (uint32_t)((hf * HN + j) * C), a_lo + (uint32_t)(j * kTileM), hi, b_lo, hi, id, acc);
                    acc = 1u;
                    a_lo += 2u * (uint32_t)k.rows_pad;
                    b_lo += 2u * (uint32_t)C;
                  }
                  a_tap += (uint32_t)d;
                }
              }
              tc_commit(bar_empty + 8 * s);
              if (++s == nstage) { s = 0; ph ^= 1u; }
            }
            if (!right_ready && !mbar_wait(bar_tile + 8 * HN, tphase & 1u)) goto mma_exit;  // (k = 1: no tap looks right)
            tc_commit(bar_accf + 8 * hf);
            RP_STAMP(0);
          }
        }
    mma_exit:;
    }
    __syncwarp();
  } else {
    // ===== load / epilogue warps =====
    const int q = warp & 3, h = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const float slope = a.slope;
    constexpr int CPW = C / 32;  // 16-column chunks per warp and tile
    if (warp == 0) RP_STAMP(1);
    // Tile jt of tile set `set`, in two steps so that the latency of the global loads can pass under other work:
    //   load_issue:  this thread's row of x -> registers
    //   load_commit: registers -> X (TMEM, fp32) and S_0 = bf16(lrelu(x)) -> OP[0], one "tile written" arrival; with the
    //                first / last tile also the P0 rows just outside the tile set that the FIRST conv reads (real data,
    //                zero outside the sequence, instead of stale rows: the first conv is then valid on the whole tile
    //                set and the chain's halo excludes its padding).
    auto load_issue = [&](int set, int jt, float4 (&v)[CPW][4]) {
      const int64_t b = set / ntx;
      const int t = (set - (int)b * ntx) * k.V - k.halo + jt * kTileM + q * 32 + lane;
      const bool inb = t >= 0 && t < valid_rows(a.lens, b, a.T);
#pragma unroll
      for (int ci = 0; ci < CPW; ++ci) {
        const int c0 = (ci * 2 + h) * 16;
        const float4* src = reinterpret_cast<const float4*>(a.x + b * bstride + (a.t32 ? t32_off(t, c0, C) : (int64_t)t * C + c0));
#pragma unroll
        for (int w = 0; w < 4; ++w) v[ci][w] = inb ? __ldg(src + w * ws4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto load_commit = [&](int set, int jt, const float4 (&v)[CPW][4]) {
      const int64_t b = set / ntx;
      const int t_in0 = (set - (int)b * ntx) * k.V - k.halo;
      const int Tb = valid_rows(a.lens, b, a.T);
      if (jt == 0 || jt == n - 1) {
        const float* xb = a.x + b * bstride;
        for (int e = (warp * 32 + lane); e < k.P0 * nchunk; e += kWorkWarps * 32) {
          const int chunk = e / k.P0, i = e - chunk * k.P0;
          const int orow = jt == 0 ? k.P - k.P0 + i : k.P + R + i;  // operand-buffer row
          const int t = t_in0 + orow - k.P;
          uint4 pk = make_uint4(0u, 0u, 0u, 0u);
          if (t >= 0 && t < Tb) {
            const float4* src = reinterpret_cast<const float4*>(xb + (a.t32 ? t32_off(t, chunk * 8, C) : (int64_t)t * C + chunk * 8));
            const float4 f0 = __ldg(src), f1 = __ldg(src + ws4);
            pk.x = pack_bf16(lrelu(f0.x, slope), lrelu(f0.y, slope));
            pk.y = pack_bf16(lrelu(f0.z, slope), lrelu(f0.w, slope));
            pk.z = pack_bf16(lrelu(f1.x, slope), lrelu(f1.y, slope));
            pk.w = pack_bf16(lrelu(f1.z, slope), lrelu(f1.w, slope));
          }
          *reinterpret_cast<uint4*>(op0 + ((size_t)chunk * k.rows_pad + orow) * 16) = pk;
        }
      }
      const int r = jt * kTileM + q * 32 + lane, t = t_in0 + r;
      const bool inb = t >= 0 && t < Tb;
#pragma unroll
      for (int ci = 0; ci < CPW; ++ci) {
        const int c0 = (ci * 2 + h) * 16;
        float f[16];
        uint32_t bits[16];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          f[4 * w] = v[ci][w].x; f[4 * w + 1] = v[ci][w].y; f[4 * w + 2] = v[ci][w].z; f[4 * w + 3] = v[ci][w].w;
        }
#pragma unroll
        for (int w = 0; w < 16; ++w) bits[w] = __float_as_uint(f[w]);
        tmem_st_32x16(tmem_x + lane_sel + (uint32_t)(jt * C + c0), bits);
        store_operand<false>(op0, k.rows_pad, k.P + r, c0, f, slope, inb);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tile + 8 * jt);
    };
    // L2 prefetch of a half of a tile set (T32: one 128-byte line is 8 rows of a 4-channel group)
    auto prefetch_half = [&](int set, int hf) {
      if (!a.t32 || set >= nsets || (lane & 7) != 0) return;
      const int64_t b = set / ntx;
      const int t_in0 = (set - (int)b * ntx) * k.V - k.halo;
#pragma unroll
      for (int jj = 0; jj < HN; ++jj) {
        const int t = t_in0 + (hf * HN + jj) * kTileM + q * 32 + lane;
        if (t >= 0 && t < a.T) {
#pragma unroll
          for (int ci = 0; ci < CPW; ++ci) {
            const float* src = a.x + b * bstride + t32_off(t, (ci * 2 + h) * 16, C);
#pragma unroll
            for (int w = 0; w < 4; ++w) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + w * 128));
          }
        }
      }
    };

    // first tile set: two tiles of loads in flight
    if (my_sets > 0) {
#pragma unroll
      for (int jt0 = 0; jt0 < n; jt0 += 2) {
        float4 v0[CPW][4], v1[CPW][4];
        load_issue(first, jt0, v0);
        load_issue(first, jt0 + 1, v1);
        load_commit(first, jt0, v0);
        load_commit(first, jt0 + 1, v1);
      }
    }
    if (warp == 0) RP_STAMP(1);
    bool alive = true;
    uint32_t tphase = 0;
    for (int it = 0; it < my_sets && alive; ++it) {
      const int set = first + it * step, next = set + step;
      const bool has_next = next < nsets;
      const int64_t b = set / ntx;
      const int t_in0 = (set - (int)b * ntx) * k.V - k.halo;
      const int Tb = valid_rows(a.lens, b, a.T);
      for (int l = 0; l < L && alive; ++l, ++tphase) {
        const int m = l >> 1, c2 = l & 1;
        const bool last = (l == L - 1);
        const float* bias = c2 ? bsm + (kRbMaxPairs + m) * C : bsm + m * C;  // cb_m for X, b1_m for ACC
        const uint32_t src_tmem = c2 ? tmem_x : tmem_acc;
        uint8_t* dst_op = op0 + (size_t)((l + 1) & 1) * op_bytes;
        for (int hf = 0; hf < 2 && alive; ++hf) {
          if (l == 1) prefetch_half(next, hf);  // the next tile set's rows into L2 well before they are loaded
          alive = mbar_wait_warp(bar_accf + 8 * hf, tphase & 1u);
          if (!alive) break;
          tc_fence_after();
          if (warp == 0) RP_STAMP(1);
#pragma unroll
          for (int jj = 0; jj < HN; ++jj) {
            const int jt = hf * HN + jj;
            const int r = jt * kTileM + q * 32 + lane, t = t_in0 + r;
            if (!last) {
              // S_(l+1) = bf16 / half (lrelu(acc + bias)), zero outside the sequence
              uint32_t v[2][16];
              tmem_ld_32x16(src_tmem + lane_sel + (uint32_t)(jt * C + h * 16), v[0]);
#pragma unroll
              for (int ci = 0; ci < CPW; ++ci) {
                const int c0 = (ci * 2 + h) * 16;
                tmem_ld_wait();
                if (ci + 1 < CPW) tmem_ld_32x16(src_tmem + lane_sel + (uint32_t)(jt * C + ((ci + 1) * 2 + h) * 16), v[(ci + 1) & 1]);
                float f[16];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float4 bq = *reinterpret_cast<const float4*>(bias + c0 + 4 * u);
                  f[4 * u] = __uint_as_float(v[ci & 1][4 * u]) + bq.x;
                  f[4 * u + 1] = __uint_as_float(v[ci & 1][4 * u + 1]) + bq.y;
                  f[4 * u + 2] = __uint_as_float(v[ci & 1][4 * u + 2]) + bq.z;
                  f[4 * u + 3] = __uint_as_float(v[ci & 1][4 * u + 3]) + bq.w;
                }
                if (c2) store_operand<false>(dst_op, k.rows_pad, k.P + r, c0, f, slope, t >= 0 && t < Tb);
                else store_operand<H16>(dst_op, k.rows_pad, k.P + r, c0, f, slope, t >= 0 && t < Tb);
              }
              fence_proxy_async_smem();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_tile + 8 * jt);
            } else {
              // the same tile of this CTA's next tile set: loads go out now, are used after this tile is written out
              float4 nx[CPW][4];
              if (has_next) load_issue(next, jt, nx);
              // final: y = [y +] out_scale * (X + cb_last) for the central V rows
              const bool valid = r >= k.halo && r < R - k.halo && t < Tb;
              float4 yq[CPW][4];
#pragma unroll
              for (int ci = 0; ci < CPW; ++ci) {
                const int c0 = (ci * 2 + h) * 16;
                const float4* src = reinterpret_cast<const float4*>(a.y + b * bstride + (a.t32 ? t32_off(t, c0, C) : (int64_t)t * C + c0));
#pragma unroll
                for (int w = 0; w < 4; ++w) yq[ci][w] = (valid && a.accumulate) ? src[w * ws4] : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int ci = 0; ci < CPW; ++ci) {
                const int c0 = (ci * 2 + h) * 16;
                uint32_t v[16];
                tmem_ld_32x16(tmem_x + lane_sel + (uint32_t)(jt * C + c0), v);
                tmem_ld_wait();
                if (valid) {
                  float4* dst = reinterpret_cast<float4*>(a.y + b * bstride + (a.t32 ? t32_off(t, c0, C) : (int64_t)t * C + c0));
#pragma unroll
                  for (int w = 0; w < 4; ++w) {
                    const float4 bq = *reinterpret_cast<const float4*>(bias + c0 + 4 * w);
                    float4 o;
                    o.x = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * w]) + bq.x, a.out_scale), yq[ci][w].x);
                    o.y = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * w + 1]) + bq.y, a.out_scale), yq[ci][w].y);
                    o.z = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * w + 2]) + bq.z, a.out_scale), yq[ci][w].z);
                    o.w = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * w + 3]) + bq.w, a.out_scale), yq[ci][w].w);
                    dst[w * ws4] = o;
                  }
                }
              }
              // this tile's X columns / OP[0] rows are free now
              if (has_next) load_commit(next, jt, nx);
            }
          }
          if (warp == 0) RP_STAMP(1);
        }
      }
    }
  }
done:
  if (warp == 0) RP_STAMP(1);
  tc_fence_before();
  __syncthreads();
  if (warp == kWorkWarps + 1) tmem_dealloc(tmem_base, tmem_cols);
}

constexpr size_t kSmemBudget = 224 * 1024;

struct RbPlan {
  int ntile, halo, V, P, P0, rows_pad, stages, kc;
  bool pipe;  // software-pipelined kernel (two operand buffers)
  size_t smem;
};

// ntile: as many 128-row tiles as tensor memory (X + ACC = 2 * ntile * C columns <= 512) and shared memory allow
bool make_plan(int C, int k, const int* dil, int npairs, RbPlan* p) {
  if (!(C == 32 || C == 64 || C == 128 || C == 256)) return false;
  if (k < 1 || !(k & 1) || k > kMaxTaps || npairs < 1 || npairs > kRbMaxPairs) return false;
  int halo = 0, P = (k - 1) / 2;
  for (int m = 0; m < npairs; ++m) {
    if (dil[m] < 1) return false;
    halo += (k - 1) / 2 * dil[m] + (k - 1) / 2;
    P = std::max(P, (k - 1) / 2 * dil[m]);
  }
  const int P0 = (k - 1) / 2 * dil[0];
  halo -= P0;  // the first conv reads its context from global memory (kernel: "rows just outside the tile")
  static const int forced = [] { const char* e = std::getenv("NVSE_RB_NTILE"); return e ? std::atoi(e) : 0; }();
  int ntile = std::min(C <= 64 ? 4 : 8, 256 / C);
  // C = 128, whole k = 3 ResBlock: one tile per CTA lets two CTAs share an SM (TMEM 256 columns each), and
  // the overlap of one CTA's load / final phases with the other's MMAs outweighs the larger halo share
  if (C == 128 && npairs > 1 && k <= 3) ntile = 1;
  static const int pipe_kmin = [] { const char* e = std::getenv("NVSE_RB_PIPE_KMIN"); return e ? std::atoi(e) : 5; }();
  if (C == 64 && k <= 3 && pipe_kmin > 3) ntile = 2;  // same trade at C = 64: 1.37 -> 1.13 ms for the k = 3 ResBlock of stage 3
  if (forced > 0) ntile = std::min(forced, 256 / C);
  const int kc = tc_kchunk(C);
  const size_t stage_bytes = (size_t)(C == 32 ? 4 : (C == 64 ? 2 : 1)) * kc * C * 2;  // TPS taps per stage, as in the kernels
  const int min_tile = C == 32 ? 4 : (C == 64 ? 2 : 1);  // instantiated kernels: see RB_LAUNCH
  for (; ntile >= min_tile; ntile >>= 1) {
    const int R = kTileM * ntile;
    const int rows_pad = R + 2 * P;
    static const bool pipe_env = [] { const char* e = std::getenv("NVSE_RB_PIPE"); return !(e && e[0] == '0'); }();
    const bool pipe = pipe_env && C <= 64 && k >= pipe_kmin && ntile == 4;  // measured: -3..-9 % from k = 5 up, a loss at k = 3
    const size_t opb = (pipe ? 2 : 1) * (((size_t)(C / 8) * rows_pad * 16 + 127) & ~(size_t)127);
    const size_t tail = sizeof(float) * 2 * kRbMaxPairs * C + sizeof(uint64_t) * (pipe ? kPipeBars : kNumBars) + 16;
    if (R - 2 * halo < 32) return false;  // not enough useful rows per tile: per-layer kernels do better
    if (opb + 2 * stage_bytes + tail > kSmemBudget) continue;
    static const int max_stages = [] { const char* e = std::getenv("NVSE_RB_STAGES"); return e ? std::atoi(e) : kMaxStages; }();
    int stages = (int)std::min<size_t>((kSmemBudget - opb - tail) / stage_bytes, (size_t)std::max(2, std::min(max_stages, kMaxStages)));
    // A deep ring buys nothing at C >= 128 (a stage is >= 512 tensor-core cycles) while every 16-32 KB of
    // shared memory it takes is L1 that the load / final phases' global accesses need for misses in flight
    // (measured: C = 256 pairs 0.47 -> 0.41 ms (k = 3) with 3 stages, C = 128 pairs 0.89 -> 0.78 ms with 4)
    if (C >= 128) stages = std::min(stages, C == 256 ? 3 : 4);
    if (C <= 64) stages = std::min(stages, std::min(max_stages, pipe ? 3 : 4));  // stages hold 2-4 taps: a few are enough
    // keep two CTAs per SM resident when tensor memory allows it (2 * ntile * C <= 256 columns)
    if (2 * ntile * C <= 256)
      while (stages > 2 && opb + stages * stage_bytes + tail > 110 * 1024) --stages;
    p->pipe = pipe;
    p->ntile = ntile; p->halo = halo; p->V = R - 2 * halo; p->P = P; p->P0 = P0; p->rows_pad = rows_pad;
    p->stages = stages; p->kc = kc;
    p->smem = opb + stages * stage_bytes + tail;
    return true;
  }
  return false;
}

}  // namespace

static long long* trace_buffer() {
  static long long* buf = [] {
    long long* ptr = nullptr;
    if (std::getenv("NVSE_RB_TRACE")) {
      cudaMalloc(&ptr, 128 * sizeof(long long));
      cudaMemset(ptr, 0, 128 * sizeof(long long));
    }
    return ptr;
  }();
  return buf;
}

bool rb_supported(int C, int k, const int* dil, int npairs) {
  RbPlan p;
  return make_plan(C, k, dil, npairs, &p);
}

// Per-CTA cycle model fitted to the phase traces of tools/rb_trace.py (profiles/): load + final phases,
// 2 * npairs MMA phases at the tensor-core/shared-memory floor, 2 * npairs - 1 epilogues.
double rb_cost_per_row(int C, int k, const int* dil, int npairs) {
  RbPlan p;
  if (!make_plan(C, k, dil, npairs, &p)) return -1.0;
  const double mma = std::max(C / 2.0, 32.0 + C / 4.0) * 1.12;
  const double conv = (double)p.ntile * k * (C / 16) * mma;
  const double epi = 600.0 + 7.5 * p.ntile * C;
  const double io = 5000.0 + 110.0 * C * p.ntile / (C == 128 ? 2.0 : (C == 256 ? 1.0 : 4.0));
  return (io + 2.0 * npairs * conv + (2.0 * npairs - 1.0) * epi) / p.V;
}

int launch_resblock_tc(const ResblockTcArgs& a, int64_t B, cudaStream_t st) {
  NVSE_REQUIRE(B <= 65535, NVSE_ERR_INVALID, "fused resblock: batch %lld exceeds 65535 per launch", (long long)B);
  if (B == 0 || a.T <= 0) return NVSE_OK;
  int dil[kRbMaxPairs] = {1, 1, 1};
  for (int m = 0; m < a.npairs && m < kRbMaxPairs; ++m) dil[m] = a.pair[m].dil;
  RbPlan p;
  NVSE_REQUIRE(make_plan(a.C, a.k, dil, a.npairs, &p), NVSE_ERR_UNSUPPORTED, "fused resblock: C=%d k=%d unsupported", a.C, a.k);
  RbKernelArgs k;
  k.a = a;
  if (k.a.bstride == 0) k.a.bstride = (a.t32 ? t32_rows(a.T) : (int64_t)a.T) * a.C;
  k.trace = nullptr;
  k.trace = trace_buffer();
  const int sm_count = device_sm_count();
  k.sm_count = sm_count;
  k.ntile = p.ntile; k.halo = p.halo; k.V = p.V; k.P = p.P; k.P0 = p.P0; k.rows_pad = p.rows_pad; k.stages = p.stages; k.kc = p.kc;
  dim3 grid((unsigned)((a.T + p.V - 1) / p.V), (unsigned)B);
  // persistent pipelined kernel: one CTA per resident slot, walking the tile sets with stride gridDim.x
  k.ntx = (int)grid.x;
  const int64_t nsets = (int64_t)grid.x * B;
  NVSE_REQUIRE(nsets < (int64_t)1 << 30, NVSE_ERR_INVALID, "fused resblock: too many tiles");
  k.nsets = (int)nsets;
  static const int pipe_cap = [] { const char* e = std::getenv("NVSE_RB_PIPE_CTAS"); return e ? std::atoi(e) : 0; }();
  const int slots = pipe_cap > 0 ? pipe_cap : sm_count * ((2 * p.ntile * a.C <= 256) ? 2 : 1);
  dim3 pgrid((unsigned)std::min<int64_t>(nsets, slots));
  const double rows = (double)B * a.T;
  ProfScope prof("resblock_tc", a.C, a.C, 2.0 * rows * a.C * a.C * a.k * 2.0 * a.npairs,
                 rows * a.C * 4.0 * (a.accumulate ? 3.0 : 2.0), st);
#define RB_LAUNCH(CC, NN, SP)                                                                                               \
  if (a.C == CC && p.ntile == NN && (a.h_fp16 != 0) == SP) {                                                               \
    NVSE_CUDA_CHECK(cudaFuncSetAttribute(resblock_tc_kernel<CC, NN, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)); \
    resblock_tc_kernel<CC, NN, SP><<<grid, kThreads, p.smem, st>>>(k);                                                      \
  } else
#define RB_LAUNCH_PIPE(CC, NN, SP)                                                                                          \
  if (p.pipe && a.C == CC && p.ntile == NN && (a.h_fp16 != 0) == SP) {                                                      \
    NVSE_CUDA_CHECK(cudaFuncSetAttribute(resblock_pipe_kernel<CC, NN, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)); \
    resblock_pipe_kernel<CC, NN, SP><<<pgrid, kThreads, p.smem, st>>>(k);                                                   \
  } else
  RB_LAUNCH_PIPE(32, 4, false) RB_LAUNCH_PIPE(32, 4, true) RB_LAUNCH_PIPE(64, 4, false) RB_LAUNCH_PIPE(64, 4, true)
  RB_LAUNCH(32, 4, false) RB_LAUNCH(32, 4, true) RB_LAUNCH(32, 8, false) RB_LAUNCH(64, 4, false) RB_LAUNCH(64, 4, true)
  RB_LAUNCH(64, 2, false) RB_LAUNCH(128, 2, false) RB_LAUNCH(128, 1, false) RB_LAUNCH(256, 1, false)
  return fail(NVSE_ERR_UNSUPPORTED, "fused resblock: no kernel for C=%d with %d tiles", a.C, p.ntile);
#undef RB_LAUNCH
#undef RB_LAUNCH_PIPE
  NVSE_LAUNCH_CHECK("resblock_tc_kernel");
  return NVSE_OK;
}

long long* rb_trace_buffer() { return trace_buffer(); }

NVSE_TC_ABORT_IMPL(rb)

}  // namespace nvse

// debug (tools/rb_trace.py): copy the 2 x 64 phase stamps of the traced CTA to the host
extern "C" int nvse_debug_rb_trace(long long* out128) {
  if (!nvse::rb_trace_buffer() || !out128) return -1;
  return cudaMemcpy(out128, nvse::rb_trace_buffer(), 128 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}

