// Persistent tcgen05 / TMEM ConvTranspose1d for the HiFi-GAN upsamplers (reference Models/hifigan.py:93-96,111-112:
// kernel = 2 * stride, padding = stride / 2), T32 activations in and out.
//
// With k = 2u and padding u/2 every output phase p (output row u*t + p) has exactly two taps,
//     p <  u/2 :  x[t] * W[p + u/2]  +  x[t-1] * W[p + u/2 + u]
//     p >= u/2 :  x[t] * W[p + u/2]  +  x[t+1] * W[p + u/2 - u]
// so ALL u phases of a slice of Cs output channels are computed side by side as the N = u * Cs columns of one
// accumulator D[128 input rows, N]:
//     D[:, 0:N]     += A(row shift  0) x W0      (every phase)
//     D[:, 0:N/2]   += A(row shift -1) x Wprev   (phases <  u/2)
//     D[:, N/2:N]   += A(row shift +1) x Wnext   (phases >= u/2)
// A = one staged tile of lrelu(x) in IEEE half (canonical no-swizzle K-major layout, a tap is a row shift of the
// descriptor), W* = pre-packed half images streamed (or, for the small layers, resident) in (slice, 32-channel K chunk)
// stages of 64 * N bytes.  Columns are ordered [half][4-channel group][phase in half][4] so that a thread -- a TMEM lane
// is an INPUT row -- reads, per channel group, its u consecutive OUTPUT rows x 4 channels: u * 16 contiguous bytes of the
// T32 output.  A warp's 32 input rows are transposed through a 4 KB shared-memory scratch so that every global store
// instruction writes one contiguous 512-byte run (the per-phase launch stored 16 bytes per lane into 32 different lines:
// profiles/r02_ups_store_ab.txt -- 54 % of the 256 -> 128 layer's time was those stores).
//
// Persistent CTAs (one per SM, two where shared and tensor memory allow) walk 128-row tiles; the activation tile is
// double-buffered where it fits, the accumulator always is: staging of tile i+1, the MMAs of tile i and the write-out
// of the previous accumulator overlap.  Warps 0-7 stage and write out, warp 8 streams weights, warp 9 issues MMAs.
#include "ups_tc.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "tc_ptx.cuh"

namespace nvse {

namespace {

using namespace tc;

constexpr int kWorkWarps = 8;
constexpr int kThreads = (kWorkWarps + 2) * 32;
constexpr int kTileM = 128;
constexpr int kRowsPad = 131;       // staged rows (128 + one either side), odd pitch: conflict-free staging stores
constexpr int kMaxSlots = 8;        // weight ring slots / resident stages
constexpr int kNumBars = 2 * kMaxSlots + 8;
constexpr int kScratchPerWarp = 4096;
constexpr size_t kSmemBudget = 224 * 1024;

struct UpsKernelArgs {
  UpsTcArgs a;
  int N, Cs, nslice, nkc;  // accumulator columns (stride * Cs), channels per slice, slices, 32-channel K chunks
  int nbufA, nslots, resident;
  int ntx, ntiles;         // tiles per utterance and in all
};

// every lane polls (try_wait suspends in hardware); a timeout in any lane is seen by all
__device__ __forceinline__ bool mbar_wait_all(uint32_t bar, uint32_t parity) {
  const bool ok = mbar_wait(bar, parity);
  return __all_sync(0xffffffffu, ok);
}

// MINB = resident CTAs per SM the register budget is set for; with one CTA per SM staging keeps twice the loads in flight
// XCL: channels-last input (the first upsampler reads conv_pre's output) instead of T32
template <int U, int MINB, bool XCL>
__global__ void __launch_bounds__(kThreads, MINB) ups_tc_kernel(const __grid_constant__ UpsKernelArgs k) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const UpsTcArgs& a = k.a;
  constexpr int kStageUnroll = MINB == 1 ? 8 : 4;  // 32-byte loads in flight per thread while staging
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = k.N, Cin = a.Cin, Cout = a.Cout, nchunk = Cin >> 3;
  const uint32_t a_bytes = (((uint32_t)nchunk * kRowsPad * 16u) + 127u) & ~127u;
  const uint32_t stage_bytes = 64u * (uint32_t)N;
  uint8_t* abuf = smem_raw;
  uint8_t* wst = abuf + (size_t)k.nbufA * a_bytes;
  uint8_t* scratch = wst + (size_t)k.nslots * stage_bytes;
  float* bias_s = reinterpret_cast<float*>(scratch + kWorkWarps * kScratchPerWarp);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + Cout);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  const uint32_t bar_wfull = smem_u32(bars), bar_wempty = bar_wfull + 8 * kMaxSlots;
  const uint32_t bar_afull = bar_wempty + 8 * kMaxSlots, bar_afree = bar_afull + 16;
  const uint32_t bar_accfull = bar_afree + 16, bar_accfree = bar_accfull + 16;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * N)) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < kMaxSlots; ++s) {
      mbar_init(bar_wfull + 8 * s, 1);
      mbar_init(bar_wempty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_afull + 8 * i, kWorkWarps);
      mbar_init(bar_afree + 8 * i, 1);
      mbar_init(bar_accfull + 8 * i, 1);
      mbar_init(bar_accfree + 8 * i, kWorkWarps);
    }
    fence_barrier_init();
  }
  if (warp == kWorkWarps + 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  for (int c = tid; c < Cout; c += kThreads) bias_s[c] = a.bias ? __ldg(a.bias + c) : 0.0f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int first = (int)blockIdx.x, step = (int)gridDim.x;
  const int my_tiles = first < k.ntiles ? (k.ntiles - first + step - 1) / step : 0;
  const int nst = k.nslice * k.nkc * 2;  // weight stages per tile

  if (warp == kWorkWarps) {
    // ===== weight producer =====
    if (lane == 0 && my_tiles > 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wimg);
      if (k.resident) {
        for (int st = 0; st < nst; ++st) {
          mbar_arrive_expect_tx(bar_wfull + 8 * st, stage_bytes);
          bulk_copy_g2s(smem_u32(wst + (size_t)st * stage_bytes), wsrc + (size_t)st * stage_bytes, stage_bytes, bar_wfull + 8 * st);
        }
      } else {
        uint32_t s = 0, ph = 1;
        for (int it = 0; it < my_tiles; ++it)
          for (int st = 0; st < nst; ++st) {
            if (!mbar_wait(bar_wempty + 8 * s, ph)) goto done;
            mbar_arrive_expect_tx(bar_wfull + 8 * s, stage_bytes);
            bulk_copy_g2s(smem_u32(wst + (size_t)s * stage_bytes), wsrc + (size_t)st * stage_bytes, stage_bytes, bar_wfull + 8 * s);
            if (++s == (uint32_t)k.nslots) { s = 0; ph ^= 1u; }
          }
      }
    }
  } else if (warp == kWorkWarps + 1) {
    // ===== MMA issuer: one elected thread =====
    if (elect_one()) {
      // IEEE-half operands (format bits 7 / 10 clear), fp32 accumulate, M = 128; N columns or N / 2
      const uint32_t idesc_n = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t idesc_h = (1u << 4) | ((uint32_t)(N >> 4) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t hi = umma_desc_hi(128u);
      const uint32_t b_full0 = umma_desc_lo(smem_u32(wst), (uint32_t)N * 16u);         // K-adjacent core matrices N rows apart
      const uint32_t b_half0 = umma_desc_lo(smem_u32(wst), (uint32_t)(N >> 1) * 16u);
      const uint32_t stage_units = stage_bytes >> 4, a_units = a_bytes >> 4;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(abuf), (uint32_t)kRowsPad * 16u);
      const uint32_t nh = (uint32_t)(N >> 1);
      uint32_t s = 0, ph = 0, acc_idx = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int buf = k.nbufA == 2 ? (it & 1) : 0;
        if (!mbar_wait(bar_afull + 8 * buf, (uint32_t)(it / k.nbufA) & 1u)) goto mma_exit;
        tc_fence_after();
        const uint32_t a_tile = a_lo0 + (uint32_t)buf * a_units;
        int st = 0;
        for (int sl = 0; sl < k.nslice; ++sl, ++acc_idx) {
          const uint32_t ab = acc_idx & 1u;
          if (acc_idx >= 2) {  // the write-out of the accumulator used two slices ago must be complete
            if (!mbar_wait(bar_accfree + 8 * ab, ((acc_idx >> 1) - 1u) & 1u)) goto mma_exit;
            tc_fence_after();
          }
          const uint32_t d_tmem = tmem_base + ab * (uint32_t)N;
          for (int kc = 0; kc < k.nkc; ++kc) {
            const uint32_t a_kc = a_tile + (uint32_t)(kc * 4) * (uint32_t)kRowsPad;
            // stage 1 of the chunk: W0, every phase, centre rows
            {
              const uint32_t slot = k.resident ? (uint32_t)st : s;
              if (!mbar_wait(bar_wfull + 8 * slot, k.resident ? 0u : ph)) goto mma_exit;
              tc_fence_after();
              const uint32_t b_lo = b_full0 + slot * stage_units;
#pragma unroll
              for (int kk = 0; kk < 2; ++kk)
                tc_mma_bf16_lohi(d_tmem, a_kc + 1u + (uint32_t)(kk * 2) * (uint32_t)kRowsPad, hi, b_lo + (uint32_t)(kk * 2) * (uint32_t)N, hi,
                                 idesc_n, (kc | kk) != 0 ? 1u : 0u);
              if (!k.resident) {
                tc_commit(bar_wempty + 8 * s);
                if (++s == (uint32_t)k.nslots) { s = 0; ph ^= 1u; }
              }
              ++st;
            }
            // stage 2: Wprev (row shift -1 -> columns [0, N/2)) and Wnext (row shift +1 -> columns [N/2, N))
            {
              const uint32_t slot = k.resident ? (uint32_t)st : s;
              if (!mbar_wait(bar_wfull + 8 * slot, k.resident ? 0u : ph)) goto mma_exit;
              tc_fence_after();
              const uint32_t b_lo = b_half0 + slot * stage_units;
#pragma unroll
              for (int kk = 0; kk < 2; ++kk)
                tc_mma_bf16_lohi(d_tmem, a_kc + 0u + (uint32_t)(kk * 2) * (uint32_t)kRowsPad, hi, b_lo + (uint32_t)(kk * 2) * nh, hi, idesc_h, 1u);
#pragma unroll
              for (int kk = 0; kk < 2; ++kk)
                tc_mma_bf16_lohi(d_tmem + nh, a_kc + 2u + (uint32_t)(kk * 2) * (uint32_t)kRowsPad, hi, b_lo + 4u * nh + (uint32_t)(kk * 2) * nh, hi,
                                 idesc_h, 1u);
              if (!k.resident) {
                tc_commit(bar_wempty + 8 * s);
                if (++s == (uint32_t)k.nslots) { s = 0; ph ^= 1u; }
              }
              ++st;
            }
          }
          tc_commit(bar_accfull + 8 * ab);
        }
        tc_commit(bar_afree + 8 * buf);  // the tile buffer may be restaged once these MMAs have read it
      }
    mma_exit:;
    }
    __syncwarp();
  } else {
    // ===== worker warps: stage activation tiles, write accumulators out =====
    const int q = warp & 3, hh = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const int wtid = warp * 32 + lane;
    const float slope = a.in_slope;
    uint8_t* sc = scratch + warp * kScratchPerWarp;

    auto stage_tile = [&](int tile, int buf) {
      const int64_t b = tile / k.ntx;
      const int t0 = (tile - (int)b * k.ntx) * kTileM;
      const float* xb = a.x + b * a.x_bstride;
      const int Tin_b = valid_rows(a.in_lens, b, a.Tin);
      uint8_t* dstb = abuf + (size_t)buf * a_bytes;
      const int rows = kTileM + 2, items = rows * nchunk;
      for (int e0 = wtid; e0 < items; e0 += kWorkWarps * 32 * kStageUnroll) {
        float4 f0[kStageUnroll], f1[kStageUnroll];
        int dst[kStageUnroll];
        if constexpr (XCL) {
          // channels-last input: consecutive lanes take consecutive 8-channel chunks of a row (contiguous); the odd row pitch
          // of the tile keeps the shared-memory stores conflict-free.  nchunk is a power of two.
          const int cshift = 31 - __clz(nchunk);
#pragma unroll
          for (int u = 0; u < kStageUnroll; ++u) {
            const int e = e0 + u * kWorkWarps * 32;
            const int chunk = e & (nchunk - 1), r = e >> cshift;
            const int t = t0 - 1 + r;
            dst[u] = e < items ? (chunk * kRowsPad + r) * 16 : -1;
            f0[u] = f1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < items && t >= 0 && t < Tin_b) {
              const float4* src = reinterpret_cast<const float4*>(xb + (int64_t)t * Cin + chunk * 8);
              f0[u] = __ldg(src);
              f1[u] = __ldg(src + 1);
            }
          }
        } else {
#pragma unroll
          for (int u = 0; u < kStageUnroll; ++u) {
            const int e = e0 + u * kWorkWarps * 32;
            const int chunk = e / rows, r = e - chunk * rows;  // consecutive lanes: consecutive rows of a chunk (T32: contiguous)
            const int t = t0 - 1 + r;
            dst[u] = e < items ? (chunk * kRowsPad + r) * 16 : -1;
            f0[u] = f1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < items && t >= 0 && t < Tin_b) {
              const float4* src = reinterpret_cast<const float4*>(xb + t32_off(t, chunk * 8, Cin));
              f0[u] = __ldg(src);
              f1[u] = __ldg(src + 32);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kStageUnroll; ++u) {
          if (dst[u] < 0) continue;
          uint4 v;
          v.x = pack_f16(lrelu(f0[u].x, slope), lrelu(f0[u].y, slope));
          v.y = pack_f16(lrelu(f0[u].z, slope), lrelu(f0[u].w, slope));
          v.z = pack_f16(lrelu(f1[u].x, slope), lrelu(f1[u].y, slope));
          v.w = pack_f16(lrelu(f1[u].z, slope), lrelu(f1[u].w, slope));
          *reinterpret_cast<uint4*>(dstb + dst[u]) = v;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_afull + 8 * buf);
    };

    uint32_t acc_idx = 0;
    // write the accumulators of tile `tile` out (all slices); false on a timed-out wait
    auto write_tile = [&](int tile) -> bool {
      const int64_t b = tile / k.ntx;
      const int t0 = (tile - (int)b * k.ntx) * kTileM;
      float* yb = a.y + b * a.y_bstride;
      const int cg4 = Cout >> 2;  // 4-channel groups per output row
      for (int sl = 0; sl < k.nslice; ++sl, ++acc_idx) {
        const uint32_t ab = acc_idx & 1u;
        if (!mbar_wait_all(bar_accfull + 8 * ab, (acc_idx >> 1) & 1u)) return false;
        tc_fence_after();
        const uint32_t t_acc = tmem_base + lane_sel + ab * (uint32_t)N;
        const int nitem = U == 8 ? (k.Cs >> 2) : (k.Cs >> 4);  // U = 8: one 4-channel group, U = 2: four groups per item
        for (int itx = hh; itx < nitem; itx += 2) {
          uint32_t v0[16], v1[16];
          tmem_ld_32x16(t_acc + (uint32_t)(itx * 16), v0);
          tmem_ld_32x16(t_acc + (uint32_t)(N >> 1) + (uint32_t)(itx * 16), v1);
          tmem_ld_wait();
          if (U == 8) {
            const float4 bq = *reinterpret_cast<const float4*>(bias_s + sl * k.Cs + itx * 4);
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              const uint32_t* s4 = (p < 4 ? v0 : v1) + (p & 3) * 4;
              const float4 o = make_float4(__uint_as_float(s4[0]) + bq.x, __uint_as_float(s4[1]) + bq.y, __uint_as_float(s4[2]) + bq.z,
                                           __uint_as_float(s4[3]) + bq.w);
              *reinterpret_cast<float4*>(sc + (lane * 8 + (p ^ (lane & 7))) * 16) = o;
            }
            __syncwarp();
            const int64_t G = (int64_t)sl * (k.Cs >> 2) + itx;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int tl = 4 * i + (lane >> 3), pp = lane & 7;
              const float4 o = *reinterpret_cast<const float4*>(sc + (tl * 8 + (pp ^ (tl & 7))) * 16);
              const int trow = t0 + q * 32 + tl;
              if (trow < a.Tin) {
                const int64_t ob = (int64_t)(t0 + q * 32 + 4 * i) >> 2;
                *reinterpret_cast<float4*>(yb + (ob * cg4 + G) * 128 + lane * 4) = o;
              }
            }
          } else {
            const float* bp = bias_s + sl * k.Cs + itx * 16;
#pragma unroll
            for (int gl = 0; gl < 4; ++gl) {
              const float4 bq = *reinterpret_cast<const float4*>(bp + gl * 4);
#pragma unroll
              for (int p = 0; p < 2; ++p) {
                const uint32_t* s4 = (p ? v1 : v0) + gl * 4;
                const float4 o = make_float4(__uint_as_float(s4[0]) + bq.x, __uint_as_float(s4[1]) + bq.y, __uint_as_float(s4[2]) + bq.z,
                                             __uint_as_float(s4[3]) + bq.w);
                *reinterpret_cast<float4*>(sc + (gl * 64 + lane * 2 + p) * 16) = o;
              }
            }
            __syncwarp();
            const int64_t G0 = ((int64_t)sl * k.Cs + itx * 16) >> 2;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int gl = i >> 1, blk = i & 1;
              const float4 o = *reinterpret_cast<const float4*>(sc + (32 * i + lane) * 16);
              const int trow = t0 + q * 32 + 16 * blk + (lane >> 1);
              if (trow < a.Tin) {
                const int64_t ob = (int64_t)(t0 + q * 32 + 16 * blk) >> 4;
                *reinterpret_cast<float4*>(yb + (ob * cg4 + G0 + gl) * 128 + lane * 4) = o;
              }
            }
          }
          __syncwarp();  // scratch is reused by the next item
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_accfree + 8 * ab);
      }
      return true;
    };

    if (my_tiles > 0) stage_tile(first, 0);
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = first + it * step, next = tile + step;
      const bool has_next = it + 1 < my_tiles;
      if (k.nbufA == 2) {
        // the other buffer is free once the MMAs of tile it - 1 have read it
        if (has_next) {
          if (it >= 1 && !mbar_wait_all(bar_afree + 8 * ((it + 1) & 1), (uint32_t)((it - 1) >> 1) & 1u)) break;
          stage_tile(next, (it + 1) & 1);
        }
        if (!write_tile(tile)) break;
      } else {
        if (!write_tile(tile)) break;
        if (has_next) {
          if (!mbar_wait_all(bar_afree, (uint32_t)it & 1u)) break;
          stage_tile(next, 0);
        }
      }
    }
  }
done:
  tc_fence_before();
  __syncthreads();
  if (warp == kWorkWarps + 1) tmem_dealloc(tmem_base, tmem_cols);
}

// fp32 [k][Cin][Cout] -> half image, stage order [slice][32-channel K chunk]{ W0 [k8][n][8] | Wprev [k8][n'][8] | Wnext [k8][n'][8] }
__global__ void __launch_bounds__(256) pack_weight_ups_kernel(const float* __restrict__ w, __half* __restrict__ img, int Cin,
                                                               int Cout, int u, int Cs) {
  const int N = u * Cs, nh = N >> 1, pad = u >> 1, nkc = Cin >> 5;
  const int64_t total = (int64_t)Cin * Cout * 2 * u;
  const int64_t blk = (int64_t)64 * N;  // halfs per (slice, K chunk)
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bi = e / blk;
    const int idx = (int)(e - bi * blk);
    const int sl = (int)(bi / nkc), kc = (int)(bi - (int64_t)sl * nkc);
    int piece, k8, n, kr = idx & 7;
    if (idx < 32 * N) {
      piece = 0; k8 = idx / (8 * N); n = (idx >> 3) % N;
    } else {
      const int i2 = idx - 32 * N;
      piece = 1 + i2 / (16 * N);
      const int i3 = i2 % (16 * N);
      k8 = i3 / (4 * N); n = (i3 >> 3) % nh;
    }
    const int half = piece == 0 ? n / nh : piece - 1, rem = piece == 0 ? n % nh : n;
    const int g = rem / (2 * u), pl = (rem % (2 * u)) >> 2, c4 = rem & 3;
    const int p = half * (u >> 1) + pl;
    const int co = sl * Cs + 4 * g + c4, ci = kc * 32 + k8 * 8 + kr;
    const int j = piece == 0 ? p + pad : (piece == 1 ? p + pad + u : p + pad - u);
    img[e] = __float2half_rn(w[((int64_t)j * Cin + ci) * Cout + co]);
  }
}

int slice_channels(int Cout, int u) { return std::min(Cout, 256 / u); }

}  // namespace

bool ups_tc_supported(int Cin, int Cout, int k, int stride, int padding) {
  if (!(stride == 2 || stride == 8) || k != 2 * stride || padding != stride / 2) return false;
  if (Cin < 32 || Cin % 32 || Cin > 512 || Cout < 16 || Cout % 16 || Cout > 512) return false;
  const int Cs = slice_channels(Cout, stride);
  if (Cout % Cs || Cs % 16) return false;
  const int N = stride * Cs;
  if (N < 32 || N > 256 || (N & (N - 1))) return false;
  // one tile buffer + two weight slots + scratch must fit
  const size_t a_bytes = ((size_t)(Cin / 8) * kRowsPad * 16 + 127) & ~(size_t)127;
  return a_bytes + 2 * (size_t)64 * N + kWorkWarps * kScratchPerWarp + Cout * 4 + kNumBars * 8 + 16 <= kSmemBudget;
}

int launch_pack_weight_ups(const float* w_kio, void* img, int Cin, int Cout, int stride, cudaStream_t st) {
  NVSE_REQUIRE(ups_tc_supported(Cin, Cout, 2 * stride, stride, stride / 2), NVSE_ERR_UNSUPPORTED, "ups_tc: Cin=%d Cout=%d stride=%d unsupported",
               Cin, Cout, stride);
  const int64_t n = (int64_t)Cin * Cout * 2 * stride;
  pack_weight_ups_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, st>>>(w_kio, reinterpret_cast<__half*>(img), Cin, Cout,
                                                                                             stride, slice_channels(Cout, stride));
  NVSE_LAUNCH_CHECK("pack_weight_ups_kernel");
  return NVSE_OK;
}

int launch_ups_tc(const UpsTcArgs& a, int64_t B, cudaStream_t st) {
  NVSE_REQUIRE(ups_tc_supported(a.Cin, a.Cout, 2 * a.stride, a.stride, a.stride / 2), NVSE_ERR_UNSUPPORTED,
               "ups_tc: Cin=%d Cout=%d stride=%d unsupported", a.Cin, a.Cout, a.stride);
  NVSE_REQUIRE(!a.x_cl || (a.stride == 8 && !((a.Cin / 8) & (a.Cin / 8 - 1))), NVSE_ERR_UNSUPPORTED,
               "ups_tc: channels-last input needs stride 8 and a power-of-two channel count");
  if (B == 0 || a.Tin <= 0) return NVSE_OK;
  UpsKernelArgs k;
  k.a = a;
  k.Cs = slice_channels(a.Cout, a.stride);
  k.N = a.stride * k.Cs;
  k.nslice = a.Cout / k.Cs;
  k.nkc = a.Cin / 32;
  const size_t a_bytes = ((size_t)(a.Cin / 8) * kRowsPad * 16 + 127) & ~(size_t)127;
  const size_t stage_bytes = (size_t)64 * k.N;
  const size_t tail = kWorkWarps * kScratchPerWarp + (size_t)a.Cout * 4 + kNumBars * 8 + 16;
  const int nst = k.nslice * k.nkc * 2;
  k.resident = nst <= kMaxSlots && a_bytes + nst * stage_bytes + tail <= kSmemBudget;
  const int want_slots = k.resident ? nst : 3;
  k.nbufA = a_bytes * 2 + want_slots * stage_bytes + tail <= kSmemBudget ? 2 : 1;
  k.nslots = k.resident ? nst : (int)std::min<size_t>(kMaxSlots, (kSmemBudget - k.nbufA * a_bytes - tail) / stage_bytes);
  NVSE_REQUIRE(k.nslots >= 2 || k.resident, NVSE_ERR_UNSUPPORTED, "ups_tc: tile does not fit shared memory");
  if (!k.resident) k.nslots = std::min(k.nslots, 4);
  const size_t smem = k.nbufA * a_bytes + k.nslots * stage_bytes + tail;
  k.ntx = (a.Tin + kTileM - 1) / kTileM;
  const int64_t ntiles = (int64_t)k.ntx * B;
  NVSE_REQUIRE(ntiles < (int64_t)1 << 30, NVSE_ERR_INVALID, "ups_tc: too many tiles");
  k.ntiles = (int)ntiles;
  const int sm_count = device_sm_count();
  // two CTAs per SM where both fit (shared memory incl. the per-CTA reservation, 2 * N tensor-memory columns each)
  const int per_sm = (a.stride == 2 && 2 * (smem + 1024) <= 227 * 1024 && 4 * k.N <= 512) ? 2 : 1;
  dim3 grid((unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count * per_sm));
  const double rows = (double)B * a.Tin;
  ProfScope prof("ups_tc", a.Cin, a.Cout, 2.0 * rows * a.Cin * a.Cout * 2.0 * a.stride,
                 rows * (a.Cin * 4.0 + (double)a.stride * a.Cout * 4.0), st);
#define UPS_LAUNCH(UU, MB, XC)                                                                                                \
  if (a.stride == UU && per_sm == MB && (a.x_cl != 0) == XC) {                                                                \
    NVSE_CUDA_CHECK(cudaFuncSetAttribute(ups_tc_kernel<UU, MB, XC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)); \
    ups_tc_kernel<UU, MB, XC><<<grid, kThreads, smem, st>>>(k);                                                               \
  } else
  UPS_LAUNCH(8, 1, false) UPS_LAUNCH(8, 1, true) UPS_LAUNCH(2, 1, false) UPS_LAUNCH(2, 2, false)
  return fail(NVSE_ERR_UNSUPPORTED, "ups_tc: no kernel for stride %d%s", a.stride, a.x_cl ? " with channels-last input" : "");
#undef UPS_LAUNCH
  NVSE_LAUNCH_CHECK("ups_tc_kernel");
  return NVSE_OK;
}

NVSE_TC_ABORT_IMPL(ups)

}  // namespace nvse
