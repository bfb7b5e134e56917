// tcgen05 / TMEM implicit-GEMM convolution over channels-last activations (sm_100a).
//
//   D[128 time rows, Cout] (fp32, TMEM)  +=  A[128 rows, 16 ch] (bf16, smem)  x  B[Cout, 16 ch] (bf16, smem)
//
// summed over the taps of a tap list and over Cin in steps of 16.  The activation tile is
// staged ONCE per CTA with its halo (rows t0+min_off .. t0+127+max_off), with the input
// leaky_relu and the fp32->bf16 conversion fused into the staging pass, in the canonical
// no-swizzle K-major UMMA layout  [ci/8][row][ci%8]  (8x16-byte core matrices, row pitch
// 16 B).  In that layout a tap is nothing but a start-address offset of (row shift)*16 B in
// the A descriptor, so dilation costs nothing and no im2col copy is ever made.  Weights are
// pre-packed (at load time, weight-norm already folded) as the exact shared-memory image of
// each (tap, 64-channel K chunk) stage and streamed through an mbarrier ring with
// cp.async.bulk (TMA 1-D).  One elected thread issues tcgen05.mma; tcgen05.commit releases
// weight stages and finally signals the epilogue warps, which pull the accumulator out of
// TMEM with tcgen05.ld and apply bias / residual / MRF scale+accumulate (fp32 output) or
// bias + leaky_relu + bf16 pack (the intermediate of a ResBlock pair).
//
// Warp roles (320 threads): warps 0-7 epilogue (two per TMEM lane quarter, interleaved over the
// 16-column chunks), warp 8 weight producer, warp 9 TMEM allocator + MMA issuer.  All ten warps
// stage activations first; the epilogue warps also prefetch their residual / accumulate rows into
// L2 before staging so that the loads of the epilogue do not pay HBM latency.
#include <cuda_fp16.h>

#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <cstdlib>

namespace nvse {

namespace {

using namespace tc;

constexpr int kEpiWarps = 8;                     // two per TMEM lane quarter, splitting the column chunks
constexpr int kThreads = (kEpiWarps + 2) * 32;   // + weight producer + MMA issuer
constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
constexpr int kMaxPhases = 8;

struct KernelArgs {
  ConvTcArgs a;
  int nphase;                      // output phases sharing one staged activation tile (ConvTranspose1d); 1 for Conv1d
  int ph_per_cta;                  // phases one CTA computes (blockIdx.z selects the group): nphase, or fewer for small grids
  int gsz;                         // phases per accumulator group: the phases of a group are adjacent output rows
                                   // (out_mul * t + p, p + 1, ..), accumulate side by side in TMEM and are stored
                                   // together, so that every store instruction pair fills whole 32-byte sectors
  ConvTaps ptaps[kMaxPhases];      // tap list of each phase
  int pout_add[kMaxPhases];        // output row = out_mul * t + pout_add[phase]
  int ntile;       // 128-row M tiles per CTA: every weight stage feeds ntile MMAs (weight reuse from smem)
  int min_off;     // smallest tap offset over all phases
  int rows;        // staged rows = 128 * ntile + (max_off - min_off)
  int rows_pad;    // rows rounded up to an odd count (conflict-free staging stores)
  int stages;      // weight ring depth
  int kc;          // channels per weight stage (min(Cin, 64))
  int dbg;         // experiments: 1 = skip the epilogue's global stores, 2 = skip the activation staging loads
};

constexpr int kStageUnroll = 4;  // independent 32-byte loads in flight per thread while staging

// KK (16-channel MMA steps per weight stage), the tile count and the split flag are compile-time so that the
// MMA issue loop unrolls completely (an MMA issued from a loop with run-time trip counts costs more issue
// cycles than the tensor core needs to execute it: tools/probe/mma_probe3.cu).  MINB = resident CTAs per SM
// the register budget is set for: 3 for the small memory-bound tiles, 1 where shared memory allows one CTA.
template <int KK, int NT, bool SPLIT, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) conv_tc_kernel(const __grid_constant__ KernelArgs k) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const ConvTcArgs& a = k.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.y;
  constexpr int ntile = NT;
  const int t0 = blockIdx.x * kTileM * ntile;
  const int Cin = a.Cin, Cout = a.Cout;
  const int nchunk = Cin >> 3;
  const uint32_t act_bytes = (((uint32_t)nchunk * k.rows_pad * 16u) + 127u) & ~127u;  // one bf16 plane
  const uint32_t stage_bytes = (uint32_t)k.kc * Cout * 2u;
  constexpr bool split = SPLIT;  // activations as hi + lo bf16 planes (two MMAs per K step)
  uint8_t* act = smem_raw;
  uint8_t* wst = smem_raw + (split ? 2u : 1u) * act_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + (size_t)k.stages * stage_bytes);
  // barriers: full[kMaxStages], empty[kMaxStages], accum_full[2], tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  // bias in shared memory: the epilogue adds it to every 16-column chunk, and a global load there sits on the critical
  // path of each chunk (ncu, round 2: 21 % of all stall samples of the 256 -> 128 upsampler were that one FADD)
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);
  for (int c = tid; c < Cout; c += kThreads) bias_s[c] = a.bias ? __ldg(a.bias + c) : 0.0f;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kMaxStages);
  const uint32_t bar_accum = smem_u32(bars + 2 * kMaxStages), bar_tfree = smem_u32(bars + 2 * kMaxStages + 2);
  const int ph_lo = (int)blockIdx.z * k.ph_per_cta;                                   // this CTA's phases
  const int ph_n = k.nphase - ph_lo < k.ph_per_cta ? k.nphase - ph_lo : k.ph_per_cta;
  const int gsz = k.gsz;                                // phases per accumulator group
  const int ngrp = (ph_n + gsz - 1) / gsz;              // groups this CTA computes
  const int nbuf = k.ph_per_cta > gsz ? 2 : 1;          // TMEM accumulator double-buffering across groups
  const int gcols = gsz * ntile * Cout;                 // TMEM columns of one group: [phase in group][tile][Cout]
  const uint32_t tmem_cols = (uint32_t)(gcols * nbuf) < 32u ? 32u : (uint32_t)(gcols * nbuf);  // power of two by construction

  if (tid == 0) {
    for (int s = 0; s < k.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_accum, 1);
    mbar_init(bar_accum + 8, 1);
    mbar_init(bar_tfree, kEpiWarps);
    mbar_init(bar_tfree + 8, kEpiWarps);
    fence_barrier_init();
  }
  if (warp == kEpiWarps + 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);

  // ---- epilogue warps: pull the rows they will read in the epilogue (residual, accumulate target)
  //      into L2 now; staging + MMAs give the prefetch time to land ------------------------------
  if (warp < kEpiWarps && !a.out_bf16 && !a.y_t32 && (a.residual || a.accumulate)) {
    const int q = warp & 3, half = warp >> 2;
    for (int ph = 0; ph < ph_n; ++ph)
      for (int j = 0; j < ntile; ++j) {
        const int t = t0 + j * kTileM + q * 32 + lane;
        const int64_t orow = (int64_t)a.out_mul * t + k.pout_add[ph_lo + ph];
        if (t < a.Trows && orow < a.Tout) {
          const int64_t base = b * a.y_bstride + orow * (a.y_ld ? a.y_ld : Cout);
          for (int c = half * 32; c < Cout; c += 64) {  // one 128-byte line per 32 fp32 channels
            if (a.residual) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.residual + base + c));
            if (a.accumulate) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const float*>(a.y) + base + c));
          }
        }
      }
  }

  // ---- stage the activation tile (all warps): lrelu + bf16 + K-major core-matrix layout ------
  {
    const int Tin_b = valid_rows(a.in_lens, b, a.Tin);  // this utterance's valid rows (ragged batches)
    const int items = k.rows * nchunk;
    const int cshift = 31 - __clz(nchunk);
    if (a.in_bf16) {
      const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(a.x) + b * a.x_bstride;
      for (int e0 = tid; e0 < items; e0 += kThreads * kStageUnroll) {
        uint4 v[kStageUnroll];
        int dst[kStageUnroll];
#pragma unroll
        for (int u = 0; u < kStageUnroll; ++u) {
          const int e = e0 + u * kThreads;
          const int chunk = e & (nchunk - 1), r = e >> cshift;
          const int t = t0 + k.min_off + r;
          dst[u] = e < items ? (chunk * k.rows_pad + r) * 16 : -1;
          v[u] = make_uint4(0u, 0u, 0u, 0u);
          if (e < items && t >= 0 && t < Tin_b) v[u] = __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)t * Cin + chunk * 8));
        }
#pragma unroll
        for (int u = 0; u < kStageUnroll; ++u)
          if (dst[u] >= 0) *reinterpret_cast<uint4*>(act + dst[u]) = v[u];
      }
    } else {
      const float* xb = reinterpret_cast<const float*>(a.x) + b * a.x_bstride;
      const float s = a.in_slope;
      for (int e0 = tid; e0 < items; e0 += kThreads * kStageUnroll) {
        float4 f0[kStageUnroll], f1[kStageUnroll];
        int dst[kStageUnroll];
#pragma unroll
        for (int u = 0; u < kStageUnroll; ++u) {
          const int e = e0 + u * kThreads;
          // channels-last: consecutive lanes take consecutive 8-channel chunks of a row; T32: consecutive rows of a chunk
          const int chunk = a.x_t32 ? e / k.rows : e & (nchunk - 1), r = a.x_t32 ? e - chunk * k.rows : e >> cshift;
          const int t = t0 + k.min_off + r;
          dst[u] = e < items ? (chunk * k.rows_pad + r) * 16 : -1;
          f0[u] = f1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e < items && t >= 0 && t < Tin_b && !(k.dbg & 2)) {
            const float4* src = reinterpret_cast<const float4*>(xb + (a.x_t32 ? t32_off(t, chunk * 8, Cin) : (int64_t)t * Cin + chunk * 8));
            f0[u] = __ldg(src);
            f1[u] = __ldg(src + (a.x_t32 ? 32 : 1));
          }
        }
#pragma unroll
        for (int u = 0; u < kStageUnroll; ++u) {
          if (dst[u] < 0) continue;
          const float f[8] = {lrelu(f0[u].x, s), lrelu(f0[u].y, s), lrelu(f0[u].z, s), lrelu(f0[u].w, s),
                              lrelu(f1[u].x, s), lrelu(f1[u].y, s), lrelu(f1[u].z, s), lrelu(f1[u].w, s)};
          uint4 v;
          if (a.ops_f16) {
            v.x = pack_f16(f[0], f[1]);
            v.y = pack_f16(f[2], f[3]);
            v.z = pack_f16(f[4], f[5]);
            v.w = pack_f16(f[6], f[7]);
          } else {
            v.x = pack_bf16(f[0], f[1]);
            v.y = pack_bf16(f[2], f[3]);
            v.z = pack_bf16(f[4], f[5]);
            v.w = pack_bf16(f[6], f[7]);
          }
          *reinterpret_cast<uint4*>(act + dst[u]) = v;
          if (split) {  // residual of the first rounding, itself rounded to bf16: ~16 mantissa bits in total
            uint4 lo;
            lo.x = pack_bf16(f[0] - bf16_lo(v.x), f[1] - bf16_hi(v.x));
            lo.y = pack_bf16(f[2] - bf16_lo(v.y), f[3] - bf16_hi(v.y));
            lo.z = pack_bf16(f[4] - bf16_lo(v.z), f[5] - bf16_hi(v.z));
            lo.w = pack_bf16(f[6] - bf16_lo(v.w), f[7] - bf16_hi(v.w));
            *reinterpret_cast<uint4*>(act + act_bytes + dst[u]) = lo;
          }
        }
      }
    }
  }
  fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nkc = Cin / k.kc;

  if (warp == kEpiWarps) {
    // ===== weight producer: one bulk copy per (phase, tap, K chunk) stage =====
    if (lane == 0) {
      int it = 0;
      for (int ph = 0; ph < ph_n; ++ph) {
        const ConvTaps& tp = k.ptaps[ph_lo + ph];
        for (int tap = 0; tap < tp.ntaps; ++tap) {
          const __nv_bfloat16* wsrc = a.wimg + (size_t)tp.widx[tap] * nkc * (stage_bytes / 2);
          for (int kc = 0; kc < nkc; ++kc, ++it) {
            const int s = it % k.stages;
            const uint32_t par = ((it / k.stages) & 1) ^ 1;
            if (!mbar_wait(bar_empty + 8 * s, par)) goto done;
            mbar_arrive_expect_tx(bar_full + 8 * s, stage_bytes);
            bulk_copy_g2s(smem_u32(wst + (size_t)s * stage_bytes), wsrc + (size_t)kc * (stage_bytes / 2), stage_bytes,
                          bar_full + 8 * s);
          }
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===== MMA issuer: one elected thread, a few instructions between consecutive MMAs =====
    if (elect_one()) {
      // operand formats: bits 7 / 10 set = bf16, clear = IEEE half
      const uint32_t idesc = (1u << 4) | (a.ops_f16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(Cout >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(act), (uint32_t)k.rows_pad * 16u);  // advances in rows (16-byte units)
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(wst), (uint32_t)Cout * 16u);
      const uint32_t hi = umma_desc_hi(128u);
      const uint32_t stage_units = stage_bytes >> 4, nstage = (uint32_t)k.stages, lo_plane = act_bytes >> 4;
      const uint32_t kc_rows = (uint32_t)(KK * 2) * (uint32_t)k.rows_pad;
      uint32_t s = 0, par = 0;
      for (int ph = 0; ph < ph_n; ++ph) {
        const ConvTaps& tp = k.ptaps[ph_lo + ph];
        const int grp = ph / gsz, pq = ph - grp * gsz, buf = grp & 1;
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * gcols + pq * ntile * Cout);
        if (pq == 0 && grp >= 2) {  // the epilogue must have drained this accumulator (group grp - 2)
          if (!mbar_wait(bar_tfree + 8 * buf, ((grp >> 1) - 1) & 1)) goto mma_exit;
          tc_fence_after();
        }
        uint32_t acc = 0u;
        for (int tap = 0; tap < tp.ntaps; ++tap) {
          uint32_t a_kc = a_lo0 + (uint32_t)(tp.off[tap] - k.min_off);
          for (int kc = 0; kc < nkc; ++kc, a_kc += kc_rows) {
            if (!mbar_wait(bar_full + 8 * s, par)) goto mma_exit;
            tc_fence_after();
            uint32_t a_lo = a_kc, b_lo = b_lo0 + s * stage_units;
#pragma unroll
            for (int kk = 0; kk < KK; ++kk) {
#pragma unroll
              for (int j = 0; j < NT; ++j) {  // the same weight stage feeds every M tile of this CTA
                tc_mma_bf16_lohi(d_tmem + (uint32_t)(j * Cout), a_lo + (uint32_t)(j * kTileM), hi, b_lo, hi, idesc, acc);
                if (SPLIT) tc_mma_bf16_lohi(d_tmem + (uint32_t)(j * Cout), a_lo + lo_plane + (uint32_t)(j * kTileM), hi, b_lo, hi, idesc, 1u);
              }
              acc = 1u;
              a_lo += 2u * (uint32_t)k.rows_pad;
              b_lo += 2u * (uint32_t)Cout;
            }
            tc_commit(bar_empty + 8 * s);  // stage reusable once these MMAs have read it
            if (++s == nstage) { s = 0; par ^= 1u; }
          }
        }
        if (pq == gsz - 1 || ph == ph_n - 1) tc_commit(bar_accum + 8 * buf);  // accumulators of this group complete -> epilogue
      }
    mma_exit:;
    }
    __syncwarp();
  } else {
    // ===== epilogue warps: TMEM -> registers -> global =====
    const int quarter = warp & 3, chalf = warp >> 2;  // TMEM lane quarter; which 16-column chunks (even / odd)
    const int row = quarter * 32 + lane;
    for (int grp = 0; grp < ngrp; ++grp) {
      const int buf = grp & 1;
      if (!mbar_wait(bar_accum + 8 * buf, (grp >> 1) & 1)) break;
      tc_fence_after();
      const int gq = ph_n - grp * gsz < gsz ? ph_n - grp * gsz : gsz;  // phases in this group
     for (int j = 0; j < ntile; ++j) {
      const int t = t0 + j * kTileM + row;
      for (int c0 = chalf * 16; c0 < Cout; c0 += 16 * (kEpiWarps / 4)) {
       for (int pq = 0; pq < gq; ++pq) {  // adjacent output rows back to back: whole sectors
        const int ph = grp * gsz + pq;
        const int64_t orow = (int64_t)a.out_mul * t + k.pout_add[ph_lo + ph];
        const bool valid = t < a.Trows && orow < a.Tout && !(k.dbg & 1);
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * gcols + (pq * ntile + j) * Cout);
        uint32_t v[16];
        tmem_ld_32x16(t_addr + (uint32_t)c0, v);
        if (a.out_bf16) {
          float bq[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(bq + 4 * q) = *(reinterpret_cast<const float4*>(bias_s + c0) + q);
          tmem_ld_wait();
          if (valid) {
            __nv_bfloat16* yr = reinterpret_cast<__nv_bfloat16*>(a.y) + b * a.y_bstride + orow * Cout + c0;
            uint4 o[2];
            uint32_t* op = reinterpret_cast<uint32_t*>(o);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              op[j] = pack_bf16(lrelu(__uint_as_float(v[2 * j]) + bq[2 * j], a.out_slope),
                                lrelu(__uint_as_float(v[2 * j + 1]) + bq[2 * j + 1], a.out_slope));
            *reinterpret_cast<uint4*>(yr) = o[0];
            *reinterpret_cast<uint4*>(yr + 8) = o[1];
          }
        } else {
          // issue every global load of this chunk before waiting on TMEM (latency overlap)
          float4 bq[4], rq[4], yq[4];
          const int64_t yoff = b * a.y_bstride + (a.y_t32 ? t32_off(orow, c0, Cout) : orow * (a.y_ld ? a.y_ld : Cout) + c0);
          const int qs = a.y_t32 ? 128 : 4;  // floats between consecutive 4-channel groups of a row
          float* yr = reinterpret_cast<float*>(a.y) + yoff;
          const float* rr = a.residual ? a.residual + yoff : nullptr;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            bq[q] = *(reinterpret_cast<const float4*>(bias_s + c0) + q);
            rq[q] = (valid && rr) ? *reinterpret_cast<const float4*>(rr + qs * q) : make_float4(0.f, 0.f, 0.f, 0.f);
            yq[q] = (valid && a.accumulate) ? *reinterpret_cast<const float4*>(yr + qs * q) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 o = make_float4(__uint_as_float(v[q * 4 + 0]) + bq[q].x, __uint_as_float(v[q * 4 + 1]) + bq[q].y,
                                     __uint_as_float(v[q * 4 + 2]) + bq[q].z, __uint_as_float(v[q * 4 + 3]) + bq[q].w);
              if (a.out_slope != 1.0f) {
                o.x = lrelu(o.x, a.out_slope); o.y = lrelu(o.y, a.out_slope);
                o.z = lrelu(o.z, a.out_slope); o.w = lrelu(o.w, a.out_slope);
              }
              if (a.mask) {  // dgrad: the derivative of the leaky_relu in front of the layer
                const float4 mq = *reinterpret_cast<const float4*>(a.mask + yoff + qs * q);
                if (!(mq.x > 0.0f)) o.x *= a.mask_slope;
                if (!(mq.y > 0.0f)) o.y *= a.mask_slope;
                if (!(mq.z > 0.0f)) o.z *= a.mask_slope;
                if (!(mq.w > 0.0f)) o.w *= a.mask_slope;
              }
              o.x = __fadd_rn(__fmul_rn(o.x + rq[q].x, a.out_scale), yq[q].x);
              o.y = __fadd_rn(__fmul_rn(o.y + rq[q].y, a.out_scale), yq[q].y);
              o.z = __fadd_rn(__fmul_rn(o.z + rq[q].z, a.out_scale), yq[q].z);
              o.w = __fadd_rn(__fmul_rn(o.w + rq[q].w, a.out_scale), yq[q].w);
              *reinterpret_cast<float4*>(yr + qs * q) = o;
            }
          }
        }
       }
      }
     }
      if (nbuf > 1) {  // hand the accumulator back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tfree + 8 * buf);
      }
    }
  }
done:
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) tmem_dealloc(tmem_base, tmem_cols);
}

// fp32 [j][ci][co] -> bf16 [j][ci/KC][(ci%KC)/8][co][ci%8]; the source may be wider (cout_src, first column co0), or narrower
// (columns at or beyond cout_src are zero), and have fewer input channels (cin_src <= Cin: the rest of the image is zero)
__global__ void __launch_bounds__(256) pack_weight_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ img,
                                                              int Cin, int Cout, int ktaps, int kc, int as_fp16, int cin_src,
                                                              int cout_src, int co0) {
  const int64_t n = (int64_t)Cin * Cout * ktaps;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int co = e % Cout, ci = (e / Cout) % Cin, j = e / ((int64_t)Cout * Cin);
    const int64_t dst = ((((int64_t)j * (Cin / kc) + ci / kc) * (kc / 8) + (ci % kc) / 8) * Cout + co) * 8 + (ci % 8);
    const float v = (ci < cin_src && co0 + co < cout_src) ? w[((int64_t)j * cin_src + ci) * cout_src + co0 + co] : 0.0f;
    if (as_fp16) reinterpret_cast<__half*>(img)[dst] = __float2half_rn(v);  // same 16-bit slots, IEEE half
    else img[dst] = __float2bfloat16_rn(v);
  }
}

}  // namespace

int launch_pack_weight_tc_slice(const float* w_kio, int cin_src, int cout_src, int co0, __nv_bfloat16* img, int Cin, int Cout,
                                int ktaps, cudaStream_t st, bool as_fp16) {
  NVSE_REQUIRE(tc_supported(Cin, Cout), NVSE_ERR_UNSUPPORTED, "tensor-core conv: Cin=%d / Cout=%d unsupported", Cin, Cout);
  NVSE_REQUIRE(cin_src <= Cin && co0 >= 0 && co0 < cout_src, NVSE_ERR_INVALID, "tensor-core conv: bad weight slice");
  const int64_t n = (int64_t)Cin * Cout * ktaps;
  pack_weight_tc_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, st>>>(w_kio, img, Cin, Cout, ktaps, tc_kchunk(Cin),
                                                                                            as_fp16 ? 1 : 0, cin_src, cout_src, co0);
  NVSE_LAUNCH_CHECK("pack_weight_tc_kernel");
  return NVSE_OK;
}

int launch_pack_weight_tc(const float* w_kio, __nv_bfloat16* img, int Cin, int Cout, int ktaps, cudaStream_t st, bool as_fp16) {
  return launch_pack_weight_tc_slice(w_kio, Cin, Cout, 0, img, Cin, Cout, ktaps, st, as_fp16);
}

static constexpr size_t kSmemBudget = 224 * 1024;  // of the 227 KB a CTA may opt in to
static constexpr size_t kSmemTail = sizeof(uint64_t) * (2 * kMaxStages + 4) + 16 + 256 * sizeof(float);  // barriers, TMEM slot, bias

static size_t tc_smem_bytes_rows(int Cin, int Cout, int rows, bool split, int stages) {
  const int rows_pad = rows | 1;
  const size_t plane = ((size_t)(Cin / 8) * rows_pad * 16 + 127) & ~(size_t)127;
  return (split ? 2 : 1) * plane + (size_t)stages * tc_kchunk(Cin) * Cout * 2 + kSmemTail;
}
size_t tc_smem_bytes(int Cin, int Cout, int tap_span, bool split, int stages) {
  const int rows = kTileM + tap_span;
  const int rows_pad = rows | 1;
  const size_t plane = ((size_t)(Cin / 8) * rows_pad * 16 + 127) & ~(size_t)127;
  return (split ? 2 : 1) * plane + (size_t)stages * tc_kchunk(Cin) * Cout * 2 + kSmemTail;
}
bool tc_split_fits(int Cin, int Cout, int tap_span) { return tc_smem_bytes(Cin, Cout, tap_span, true, 2) <= kSmemBudget; }

// `a.taps` / `a.out_add` describe phase 0; `extra` adds phases 1.. (ConvTranspose1d) that share the staged tile.
int launch_conv_tc_phases(const ConvTcArgs& a, const ConvTaps* phase_taps, const int* phase_out_add, int nphase, int64_t B,
                          cudaStream_t st) {
  NVSE_REQUIRE(tc_supported(a.Cin, a.Cout), NVSE_ERR_UNSUPPORTED, "tensor-core conv: Cin=%d / Cout=%d unsupported", a.Cin, a.Cout);
  NVSE_REQUIRE(nphase >= 1 && nphase <= kMaxPhases, NVSE_ERR_INVALID, "tensor-core conv: %d phases per launch", nphase);
  NVSE_REQUIRE(B <= 65535, NVSE_ERR_INVALID, "tensor-core conv: batch %lld exceeds 65535 per launch", (long long)B);
  NVSE_REQUIRE(!(a.out_bf16 && (a.residual || a.accumulate)), NVSE_ERR_INVALID, "tensor-core conv: bf16 output takes no residual");
  NVSE_REQUIRE(!(a.out_bf16 && (a.mask || !a.bias)), NVSE_ERR_INVALID, "tensor-core conv: bf16 output needs a bias and takes no mask");
  NVSE_REQUIRE(!(a.split_act && a.in_bf16), NVSE_ERR_INVALID, "tensor-core conv: split activations need fp32 input");
  NVSE_REQUIRE(!(a.x_t32 && a.in_bf16) && !(a.y_t32 && a.out_bf16), NVSE_ERR_INVALID, "tensor-core conv: the T32 layout is fp32 only");
  NVSE_REQUIRE(!(a.ops_f16 && (a.in_bf16 || a.split_act)), NVSE_ERR_INVALID, "tensor-core conv: half operands need fp32 input and no split");
  NVSE_REQUIRE(!(a.y_ld && (a.y_t32 || a.out_bf16 || a.y_ld < a.Cout)), NVSE_ERR_INVALID, "tensor-core conv: a row pitch needs fp32 channels-last output");
  if (B == 0 || a.Trows <= 0) return NVSE_OK;
  KernelArgs k;
  k.a = a;
  k.nphase = nphase;
  k.ph_per_cta = nphase;
  k.gsz = 1;
  int mn = phase_taps[0].off[0], mx = mn, total_taps = 0;
  for (int p = 0; p < nphase; ++p) {
    NVSE_REQUIRE(phase_taps[p].ntaps >= 1 && phase_taps[p].ntaps <= kMaxTaps, NVSE_ERR_INVALID, "tensor-core conv: bad tap count");
    k.ptaps[p] = phase_taps[p];
    k.pout_add[p] = phase_out_add[p];
    total_taps += phase_taps[p].ntaps;
    for (int i = 0; i < phase_taps[p].ntaps; ++i) {
      mn = std::min(mn, phase_taps[p].off[i]);
      mx = std::max(mx, phase_taps[p].off[i]);
    }
  }
  k.min_off = mn;
  static const int dbg_env = [] { const char* e = std::getenv("NVSE_TC_DBG"); return e ? std::atoi(e) : 0; }();
  k.dbg = dbg_env;
  // M tiles per CTA: weight bytes streamed per FLOP fall as 1/ntile (L2 -> smem weight traffic is what
  // bounds the 128-row tile at C >= 128), and per-CTA fixed costs are amortised for the small-C layers.
  // Phases per accumulator group (see KernelArgs::gsz): consecutive output rows stored together.  Two where the
  // phases come in consecutive-row order (ConvTranspose1d polyphase launches) and two accumulators of 2 * Cout
  // columns still fit TMEM next to each other (double buffering) -- or the launch only HAS two phases.
  static const int gsz_env = [] { const char* e = std::getenv("NVSE_TC_GSZ"); return e ? std::atoi(e) : 0; }();
  int gsz = 1;
  if (nphase >= 2 && !a.out_bf16) {
    bool consecutive = true;
    for (int p = 1; p < nphase; ++p) consecutive = consecutive && phase_out_add[p] == phase_out_add[p - 1] + 1;
    const int want_g = gsz_env > 0 ? gsz_env : 2;
    if (consecutive && nphase % want_g == 0 && a.Cout * want_g * (nphase > want_g ? 2 : 1) <= 512) gsz = want_g;
  }
  k.gsz = gsz;
  const int nbuf = nphase > gsz ? 2 : 1;
  int ntile = 1;
  {
    static const int forced = [] { const char* e = std::getenv("NVSE_TC_NTILE"); return e ? std::atoi(e) : 0; }();
    // ConvTranspose1d launches (several phases, double-buffered TMEM) are memory-bound: small tiles keep
    // several CTAs resident per SM (TMEM columns = Cout * 2 * ntile), which hides the staging latency
    const int want = forced > 0 ? forced : a.ntile_hint > 0 ? a.ntile_hint : (nphase > 1 ? (a.Cout >= 64 ? 1 : 2) : (a.Cout >= 128 ? 2 : 4));
    const int64_t tiles_needed = ((int64_t)a.Trows + kTileM - 1) / kTileM;
    while (ntile * 2 <= want && a.Cout * gsz * nbuf * ntile * 2 <= 512 && ntile * 2 <= tiles_needed &&
           tc_smem_bytes_rows(a.Cin, a.Cout, kTileM * ntile * 2 + (mx - mn), a.split_act != 0, 2) <= kSmemBudget)
      ntile *= 2;
  }
  k.ntile = ntile;
  k.rows = kTileM * ntile + (mx - mn);
  k.rows_pad = k.rows | 1;  // odd row pitch (in 16-byte units): the 8 lanes of a store phase hit 8 distinct bank groups
  k.kc = tc_kchunk(a.Cin);
  const size_t act_bytes = (a.split_act ? 2 : 1) * (((size_t)(a.Cin / 8) * k.rows_pad * 16 + 127) & ~(size_t)127);
  const size_t stage_bytes = (size_t)k.kc * a.Cout * 2;
  const size_t tail = kSmemTail, budget = kSmemBudget;
  NVSE_REQUIRE(act_bytes + 2 * stage_bytes + tail <= budget, NVSE_ERR_UNSUPPORTED,
               "tensor-core conv: tile needs %zu B of shared memory", act_bytes + 2 * stage_bytes + tail);
  const int n_iters = total_taps * (a.Cin / k.kc);
  static const int max_stages = [] { const char* e = std::getenv("NVSE_TC_STAGES"); const int v = e ? std::atoi(e) : 0; return v >= 2 && v <= kMaxStages ? v : 4; }();
  int stages = (int)std::min<size_t>((budget - act_bytes - tail) / stage_bytes, (size_t)max_stages);
  stages = std::max(2, std::min(stages, std::max(2, n_iters)));
  // more resident CTAs per SM beat a deeper weight ring: one CTA's staging / epilogue overlaps another's MMAs
  // (only where that is attainable: a tile that is > 56 KB by itself keeps its 4-stage ring)
  if (act_bytes + 2 * stage_bytes + tail <= 56 * 1024)
    while (stages > 2 && act_bytes + stages * stage_bytes + tail > 56 * 1024) --stages;
  k.stages = stages;
  const size_t smem = act_bytes + stages * stage_bytes + tail;
  dim3 grid((unsigned)((a.Trows + kTileM * ntile - 1) / (kTileM * ntile)), (unsigned)B);
  // Small grids (the batch-1 case): one phase per CTA, the phases of a tile on different SMs -- each CTA then
  // streams 1 / nphase of the weights, which is what a few-CTA ConvTranspose1d launch is bound by.
  const int sm_count = device_sm_count();
  if (nphase > 1 && (int64_t)grid.x * grid.y * nphase <= 2 * sm_count) {
    k.ph_per_cta = 1;
    k.gsz = 1;
    grid.z = (unsigned)nphase;
  }
  const double rows = (double)B * a.Trows;
  ProfScope prof("conv_tc", a.Cin, a.Cout, 2.0 * rows * a.Cin * a.Cout * total_taps,
                 rows * (a.Cin * (a.in_bf16 ? 2.0 : 4.0) +
                         nphase * (a.Cout * (a.out_bf16 ? 2.0 : 4.0) * (a.accumulate ? 2.0 : 1.0) + (a.residual ? 4.0 * a.Cout : 0.0))),
                 st);
  const bool sp = a.split_act != 0;
  const bool small = smem <= 72 * 1024;  // three CTAs of this size fit an SM
#define TC_LAUNCH(KKV, NTV, SPV, MB)                                                                                      \
  if (k.kc == 16 * KKV && ntile == NTV && sp == SPV && small == (MB == 3)) {                                              \
    NVSE_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<KKV, NTV, SPV, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget)); \
    conv_tc_kernel<KKV, NTV, SPV, MB><<<grid, kThreads, smem, st>>>(k);                                                   \
  } else
#define TC_LAUNCH2(KKV, NTV, SPV) TC_LAUNCH(KKV, NTV, SPV, 1) TC_LAUNCH(KKV, NTV, SPV, 3)
  TC_LAUNCH2(2, 1, false) TC_LAUNCH2(2, 2, false) TC_LAUNCH2(2, 4, false) TC_LAUNCH2(4, 1, false) TC_LAUNCH2(4, 2, false) TC_LAUNCH2(4, 4, false)
  TC_LAUNCH2(2, 1, true) TC_LAUNCH2(2, 2, true) TC_LAUNCH2(2, 4, true) TC_LAUNCH2(4, 1, true) TC_LAUNCH2(4, 2, true) TC_LAUNCH2(4, 4, true)
  return fail(NVSE_ERR_UNSUPPORTED, "tensor-core conv: no kernel for kc=%d ntile=%d", k.kc, ntile);
#undef TC_LAUNCH2
#undef TC_LAUNCH
  NVSE_LAUNCH_CHECK("conv_tc_kernel");
  return NVSE_OK;
}

int launch_conv_tc(const ConvTcArgs& a, int64_t B, cudaStream_t st) {
  return launch_conv_tc_phases(a, &a.taps, &a.out_add, 1, B, st);
}

NVSE_TC_ABORT_IMPL(tc)

}  // namespace nvse
