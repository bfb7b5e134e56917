// Persistent, software-pipelined ResBlock1 PAIR on the 5th-gen tensor cores (sm_100a), for the channel
// counts whose whole-ResBlock chain does not fit tensor memory (C = 128): Models/hifigan.py:45-49,
//
//   y = [y +] out_scale * ( c2(lrelu(c1(lrelu(x)) + b1)) + b2 + x )
//
// One CTA per SM loops over (utterance, time tile) items.  Unlike resblock_tc.cu, where the phases of a
// tile run back to back, FOUR roles work on DIFFERENT items at the same time, decoupled by mbarriers:
//
//   loader warps (4) : x tile of item i+1 (T32 layout, coalesced) -> lrelu -> bf16 -> operand buffer OP[(i+1)&1]
//   producer thread  : cp.async.bulk weight stages of c1 / c2 through a ring
//   MMA thread       : c1(i): OP[i&1] -> ACC1      c2(i): OP[i&1] (now H) -> ACC2
//   epilogue warps(8): epi1(i): ACC1 + b1 -> lrelu -> bf16 -> OP[i&1] in place
//                      final(i): ACC2 + b2 + x (re-read, L2-resident) -> y         (overlaps c1(i+1))
//
// so that per item the tensor core only idles during epi1 (the c1 -> c2 dependency of one tile).
// TMEM: ACC1 and ACC2, NT * C fp32 columns each.  Shared memory: two operand buffers, the weight ring.
#include "resblock_tc.cuh"

#include <cstdlib>

#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace nvse {

namespace {

using namespace tc;

#ifndef NVSE_PAIR_LOADWARPS
#define NVSE_PAIR_LOADWARPS 4
#endif
#ifndef NVSE_PAIR_FINAL_U
#define NVSE_PAIR_FINAL_U 2
#endif
constexpr int kEpiWarps = 8, kLoadWarps = NVSE_PAIR_LOADWARPS;
constexpr int kThreads = (kEpiWarps + 2 + kLoadWarps) * 32;
constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
constexpr int kNumBars = 2 * kMaxStages + 10;

struct PairKernelArgs {
  ResblockTcArgs a;  // npairs == 1, t32 == 1
  int halo, V, P, rows_pad, stages;
  int tiles_per_seq, n_items;
  long long* trace;  // debug (NVSE_RB_TRACE): clock64 stamps of CTA 3, 3 roles x 40 slots
};

__device__ __forceinline__ bool wait_warp(uint32_t bar, uint32_t parity) {
  const bool ok = mbar_wait(bar, parity);
  return __all_sync(0xffffffffu, ok);
}

template <int C, int NT, bool ACCUM>
__global__ void __launch_bounds__(kThreads, 1) pair_tc_kernel(const __grid_constant__ PairKernelArgs k) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const ResblockTcArgs& a = k.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int R = NT * kTileM, nchunk = C >> 3, KC = 64, nkc = C / KC, kkn = KC >> 4;
  constexpr uint32_t stage_bytes = (uint32_t)KC * C * 2u;
  const uint32_t op_bytes = (((uint32_t)nchunk * k.rows_pad * 16u) + 127u) & ~127u;
  uint8_t* op0 = smem_raw;  // OP[0], OP[1]
  uint8_t* wst = smem_raw + 2 * op_bytes;
  float* bsm = reinterpret_cast<float*>(wst + (size_t)k.stages * stage_bytes);  // b1[C], b2[C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + 2 * C);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kMaxStages);
  const uint32_t bar0 = smem_u32(bars + 2 * kMaxStages);
  // per operand buffer b: op_ready[b] (loaders -> MMA), op_free[b] (c2 done -> loaders), h_ready[b] (epi1 -> MMA)
  const uint32_t bar_op_ready = bar0, bar_op_free = bar0 + 16, bar_h_ready = bar0 + 32;
  const uint32_t bar_acc1 = bar0 + 48, bar_acc2 = bar0 + 56;  // MMA commits -> epilogue (one per item)
  // acc2_free: final(i) has drained ACC2 -> c2(i+1) may overwrite it
  const uint32_t bar_acc2_free = bar0 + 64;
  constexpr uint32_t tmem_cols = 2 * NT * C <= 32 ? 32 : (2 * NT * C <= 64 ? 64 : (2 * NT * C <= 128 ? 128 : (2 * NT * C <= 256 ? 256 : 512)));
  const int64_t bstride = a.bstride;

  if (tid == 0) {
    for (int s = 0; s < k.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_op_ready + 8 * b, kLoadWarps);
      mbar_init(bar_op_free + 8 * b, 1);
      mbar_init(bar_h_ready + 8 * b, kEpiWarps);
    }
    mbar_init(bar_acc1, 1);
    mbar_init(bar_acc2, 1);
    mbar_init(bar_acc2_free, kEpiWarps);
    fence_barrier_init();
  }
  if (warp == kEpiWarps + 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  {
    // zero rows on both sides of every chunk of both operand buffers (never written again)
    const int zr = 2 * k.P;
    for (int e = tid; e < 2 * nchunk * zr; e += kThreads) {
      const int bufi = e / (nchunk * zr), e2 = e - bufi * nchunk * zr;
      const int chunk = e2 / zr, i = e2 - chunk * zr;
      const int row = i < k.P ? i : R + i;
      *reinterpret_cast<uint4*>(op0 + (size_t)bufi * op_bytes + ((size_t)chunk * k.rows_pad + row) * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int e = tid; e < C; e += kThreads) {
      bsm[e] = __ldg(a.pair[0].b1 + e);
      bsm[C + e] = __ldg(a.pair[0].b2 + e);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc1 = tmem_base, tmem_acc2 = tmem_base + (uint32_t)(NT * C);
  const int my_items = ((int)k.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const float slope = a.slope;
  const bool tracing = k.trace != nullptr && blockIdx.x == 3 && lane == 0;
  int tr_i = 0;
#define PT_STAMP(role) do { if (tracing && tr_i < 40) k.trace[(role) * 40 + tr_i++] = clock64(); } while (0)

  if (warp == kEpiWarps) {
    // ===== weight producer =====
    if (lane == 0) {
      const uint32_t nstage = (uint32_t)k.stages;
      uint32_t s = 0, ph = 1;
      for (int i = 0; i < my_items; ++i)
        for (int half = 0; half < 2; ++half) {
          const __nv_bfloat16* wimg = half ? a.pair[0].w2 : a.pair[0].w1;
          for (int st = 0; st < a.k * nkc; ++st) {
            if (!mbar_wait(bar_empty + 8 * s, ph)) goto done;
            mbar_arrive_expect_tx(bar_full + 8 * s, stage_bytes);
            bulk_copy_g2s(smem_u32(wst + (size_t)s * stage_bytes), wimg + (size_t)st * (stage_bytes / 2), stage_bytes,
                          bar_full + 8 * s);
            if (++s == nstage) { s = 0; ph ^= 1u; }
          }
        }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===== MMA issuer (one elected thread; fully unrolled issue) =====
    if (elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t a_lo_buf[2] = {umma_desc_lo(smem_u32(op0), (uint32_t)k.rows_pad * 16u),
                                    umma_desc_lo(smem_u32(op0 + op_bytes), (uint32_t)k.rows_pad * 16u)};
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(wst), (uint32_t)C * 16u);
      const uint32_t hi = umma_desc_hi(128u);
      constexpr uint32_t stage_units = stage_bytes >> 4;
      const uint32_t nstage = (uint32_t)k.stages;
      uint32_t s = 0, ph = 0;
      for (int i = 0; i < my_items; ++i) {
        const int b = i & 1;
        const uint32_t bpar = (uint32_t)(i >> 1) & 1u;  // phase of the per-buffer barriers
        for (int half = 0; half < 2; ++half) {
          if (half == 0) {
            if (!mbar_wait(bar_op_ready + 8 * b, bpar)) goto mma_exit;
          } else {
            if (!mbar_wait(bar_h_ready + 8 * b, bpar)) goto mma_exit;
            if (i > 0 && !mbar_wait(bar_acc2_free, (uint32_t)(i - 1) & 1u)) goto mma_exit;
          }
          tc_fence_after();
          PT_STAMP(0);
          const int d = half ? 1 : a.pair[0].dil;
          const uint32_t d_tmem = half ? tmem_acc2 : tmem_acc1;
          uint32_t acc = 0u;
          uint32_t a_tap = a_lo_buf[b] + (uint32_t)(k.P - (a.k - 1) / 2 * d);
          for (int tap = 0; tap < a.k; ++tap, a_tap += (uint32_t)d) {
#pragma unroll
            for (int kc = 0; kc < nkc; ++kc) {
              if (!mbar_wait(bar_full + 8 * s, ph)) goto mma_exit;
              tc_fence_after();
              uint32_t a_lo = a_tap + (uint32_t)(kc * (KC >> 3)) * (uint32_t)k.rows_pad;
              uint32_t b_lo = b_lo0 + s * stage_units;
#pragma unroll
              for (int kk = 0; kk < kkn; ++kk) {
#pragma unroll
                for (int j = 0; j < NT; ++j)
                  tc_mma_bf16_lohi(d_tmem + (uint32_t)(j * C), a_lo + (uint32_t)(j * kTileM), hi, b_lo, hi, idesc, acc);
                acc = 1u;
                a_lo += 2u * (uint32_t)k.rows_pad;
                b_lo += 2u * (uint32_t)C;
              }
              tc_commit(bar_empty + 8 * s);
              if (++s == nstage) { s = 0; ph ^= 1u; }
            }
          }
          tc_commit(half ? bar_acc2 : bar_acc1);
          PT_STAMP(0);
          if (half) tc_commit(bar_op_free + 8 * b);  // OP[b] may be refilled once c2 has read it
        }
      }
    mma_exit:;
    }
    __syncwarp();
  } else if (warp >= kEpiWarps + 2) {
    // ===== loader warps: x tile -> lrelu -> bf16 -> OP[i & 1] =====
    const int lw = warp - (kEpiWarps + 2);
    constexpr int UB = 8;  // (row, 8-channel chunk) units in flight per thread (16 LDG.128)
    const int rows_all = R + 2 * k.P;  // the tile and the P rows of context on either side that c1 reads
    const int units = rows_all * nchunk;
    for (int i = 0; i < my_items; ++i) {
      const int b = i & 1;
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int ub = item / k.tiles_per_seq, tx = item - ub * k.tiles_per_seq;
      const int t_in0 = tx * k.V - k.halo;
      const float* xb = a.x + (int64_t)ub * bstride;
      const int Tb = valid_rows(a.lens, ub, a.T);  // this utterance's valid rows (ragged batches)
      uint8_t* op = op0 + (size_t)b * op_bytes;
      if (lw == 0) PT_STAMP(2);
      if (i >= 2 && !wait_warp(bar_op_free + 8 * b, (uint32_t)((i >> 1) - 1) & 1u)) break;  // c2(i-2) has read OP[b]
      if (lw == 0) PT_STAMP(2);
#pragma unroll 1
      for (int e0 = lw * 32 + lane; e0 < units; e0 += UB * kLoadWarps * 32) {
        float4 f0[UB], f1[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int e = e0 + u * (kLoadWarps * 32);
          const int chunk = e / rows_all, orow = e - chunk * rows_all, t = t_in0 + orow - k.P;
          f0[u] = f1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e < units && t >= 0 && t < Tb) {
            const float4* src = reinterpret_cast<const float4*>(xb + t32_off(t, chunk * 8, C));
            f0[u] = __ldg(src);
            f1[u] = __ldg(src + 32);
          }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int e = e0 + u * (kLoadWarps * 32);
          if (e >= units) continue;
          const int chunk = e / rows_all, orow = e - chunk * rows_all;
          uint4 v;
          v.x = pack_bf16(lrelu(f0[u].x, slope), lrelu(f0[u].y, slope));
          v.y = pack_bf16(lrelu(f0[u].z, slope), lrelu(f0[u].w, slope));
          v.z = pack_bf16(lrelu(f1[u].x, slope), lrelu(f1[u].y, slope));
          v.w = pack_bf16(lrelu(f1[u].z, slope), lrelu(f1[u].w, slope));
          *reinterpret_cast<uint4*>(op + ((size_t)chunk * k.rows_pad + orow) * 16) = v;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_op_ready + 8 * b);
      if (lw == 0) PT_STAMP(2);
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3, h = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    // items of the final phase in flight: without the accumulate read the registers of yq go to a deeper batch of x
    constexpr int CPW = C / 32, IT = NT * CPW, UF = ACCUM ? NVSE_PAIR_FINAL_U : 2 * NVSE_PAIR_FINAL_U, U = IT < UF ? IT : UF;
    const float* b1 = bsm;
    const float* b2 = bsm + C;
    for (int i = 0; i < my_items; ++i) {
      const int b = i & 1;
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int ub = item / k.tiles_per_seq, tx = item - ub * k.tiles_per_seq;
      const int t_in0 = tx * k.V - k.halo;
      const int Tb = valid_rows(a.lens, ub, a.T);
      uint8_t* op = op0 + (size_t)b * op_bytes;
      // epi1: OP[b] = bf16(lrelu(ACC1 + b1)), zero outside the sequence
      if (!wait_warp(bar_acc1, (uint32_t)i & 1u)) break;
      tc_fence_after();
      if (warp == 0) PT_STAMP(1);
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int c0 = ((it % CPW) * 2 + h) * 16, jt = it / CPW;
        const int r = jt * kTileM + q * 32 + lane, t = t_in0 + r;
        uint32_t v[16];
        tmem_ld_32x16(tmem_acc1 + lane_sel + (uint32_t)(jt * C + c0), v);
        tmem_ld_wait();
        uint32_t hi[8];
        const bool inb = t >= 0 && t < Tb;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float a0 = lrelu(__uint_as_float(v[2 * w]) + b1[c0 + 2 * w], slope);
          const float a1 = lrelu(__uint_as_float(v[2 * w + 1]) + b1[c0 + 2 * w + 1], slope);
          hi[w] = inb ? pack_bf16(a0, a1) : 0u;
        }
        const size_t o0 = ((size_t)(c0 >> 3) * k.rows_pad + k.P + r) * 16;
        *reinterpret_cast<uint4*>(op + o0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(op + o0 + (size_t)k.rows_pad * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_h_ready + 8 * b);
      if (warp == 0) PT_STAMP(1);

      // final: y = [y +] out_scale * (ACC2 + b2 + x) for the central V rows; x / y loads are issued before the wait
      bool alive = true;
#pragma unroll
      for (int i0 = 0; i0 < IT; i0 += U) {
        float4 xq[U][4], yq[ACCUM ? U : 1][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int it = i0 + u, c0 = ((it % CPW) * 2 + h) * 16;
          const int r = (it / CPW) * kTileM + q * 32 + lane, t = t_in0 + r;
          const bool valid = r >= k.halo && r < R - k.halo && t < Tb;
          const int64_t off = (int64_t)ub * bstride + t32_off(t, c0, C);
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            xq[u][w] = valid ? __ldg(reinterpret_cast<const float4*>(a.x + off) + w * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (ACCUM) yq[u][w] = valid ? reinterpret_cast<const float4*>(a.y + off)[w * 32] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (i0 == 0) {
          alive = wait_warp(bar_acc2, (uint32_t)i & 1u);
          tc_fence_after();
          if (warp == 0) PT_STAMP(1);
        }
        if (!alive) break;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int it = i0 + u, c0 = ((it % CPW) * 2 + h) * 16, jt = it / CPW;
          const int r = jt * kTileM + q * 32 + lane, t = t_in0 + r;
          const bool valid = r >= k.halo && r < R - k.halo && t < Tb;
          uint32_t v[16];
          tmem_ld_32x16(tmem_acc2 + lane_sel + (uint32_t)(jt * C + c0), v);
          tmem_ld_wait();
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(a.y + (int64_t)ub * bstride + t32_off(t, c0, C));
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const float4 bq = *reinterpret_cast<const float4*>(b2 + c0 + 4 * w);
              float4 o;
              o.x = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * w]) + bq.x + xq[u][w].x, a.out_scale), (ACCUM ? yq[ACCUM ? u : 0][w].x : 0.f));
              o.y = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * w + 1]) + bq.y + xq[u][w].y, a.out_scale), (ACCUM ? yq[ACCUM ? u : 0][w].y : 0.f));
              o.z = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * w + 2]) + bq.z + xq[u][w].z, a.out_scale), (ACCUM ? yq[ACCUM ? u : 0][w].z : 0.f));
              o.w = __fadd_rn(__fmul_rn(__uint_as_float(v[4 * w + 3]) + bq.w + xq[u][w].w, a.out_scale), (ACCUM ? yq[ACCUM ? u : 0][w].w : 0.f));
              dst[w * 32] = o;
            }
          }
        }
      }
      if (!alive) break;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc2_free);
      if (warp == 0) PT_STAMP(1);
    }
  }
done:
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) tmem_dealloc(tmem_base, tmem_cols);
}

constexpr size_t kSmemBudget = 224 * 1024;

struct PairPlan {
  int halo, V, P, rows_pad, stages;
  size_t smem;
};

bool make_pair_plan(int C, int k, int dil, PairPlan* p) {
  // C = 128: two 128-row tiles per item (ACC1 + ACC2 = 512 TMEM columns).  C = 256, one tile per item in the same 512 columns,
  // is instantiated but OFF (NVSE_PAIR_C256=1 to try it): two operand buffers of 70-80 KB leave room for a two-stage ring of
  // 32 KB weight stages only (and for six of HiFi-GAN V1's nine C = 256 pairs only: context P = (k-1)/2 * d <= 12 rows), and with
  // 1 024 MMA cycles of weights buffered the ring starves -- 916 TFLOP/s against 956 for resblock_tc_kernel<256, 1>, whose phases
  // run back to back but whose ring is 3-4 stages deep (same-box A/B, 32 x 862 frames)
  static const bool c256 = [] { const char* e = std::getenv("NVSE_PAIR_C256"); return e && e[0] == '1'; }();
  if ((C != 128 && !(C == 256 && c256)) || k < 1 || !(k & 1) || k > kMaxTaps || dil < 1) return false;
  const int NT = C == 128 ? 2 : 1, R = kTileM * NT;
  // c1 reads its context (P rows on either side of the tile) from global memory: only c2 costs halo
  const int halo = (k - 1) / 2, P = (k - 1) / 2 * dil;
  if (R - 2 * halo < 64) return false;
  const int rows_pad = R + 2 * P;
  const size_t opb = ((size_t)(C / 8) * rows_pad * 16 + 127) & ~(size_t)127;
  const size_t stage_bytes = (size_t)64 * C * 2;
  const size_t tail = sizeof(float) * 2 * C + sizeof(uint64_t) * kNumBars + 16;
  if (2 * opb + 2 * stage_bytes + tail > kSmemBudget) return false;
  static const int max_stages = [] { const char* e = std::getenv("NVSE_PAIR_STAGES"); return e ? std::atoi(e) : 3; }();
  // shared memory left unused is L1 for the loaders' and the final phase's global accesses (see resblock_tc.cu)
  p->stages = (int)std::min<size_t>((kSmemBudget - 2 * opb - tail) / stage_bytes, (size_t)std::max(2, std::min(max_stages, kMaxStages)));
  p->halo = halo; p->V = R - 2 * halo; p->P = P; p->rows_pad = rows_pad;
  p->smem = 2 * opb + p->stages * stage_bytes + tail;
  return true;
}

}  // namespace

bool pair_supported(int C, int k, int dil) {
  PairPlan p;
  return make_pair_plan(C, k, dil, &p);
}

int launch_pair_tc(const ResblockTcArgs& a, int64_t B, cudaStream_t st) {
  NVSE_REQUIRE(a.npairs == 1 && a.t32 && !a.h_fp16, NVSE_ERR_INVALID, "pipelined pair kernel: one pair, T32 layout, no split");
  if (B == 0 || a.T <= 0) return NVSE_OK;
  PairPlan p;
  NVSE_REQUIRE(make_pair_plan(a.C, a.k, a.pair[0].dil, &p), NVSE_ERR_UNSUPPORTED, "pipelined pair kernel: C=%d k=%d d=%d unsupported", a.C,
               a.k, a.pair[0].dil);
  PairKernelArgs k;
  k.a = a;
  if (k.a.bstride == 0) k.a.bstride = t32_rows(a.T) * a.C;
  k.halo = p.halo; k.V = p.V; k.P = p.P; k.rows_pad = p.rows_pad; k.stages = p.stages;
  k.trace = rb_trace_buffer();
  k.tiles_per_seq = (a.T + p.V - 1) / p.V;
  const int64_t n_items = (int64_t)k.tiles_per_seq * B;
  NVSE_REQUIRE(n_items < (int64_t)1 << 30, NVSE_ERR_INVALID, "pipelined pair kernel: too many tiles");
  k.n_items = (int)n_items;
  const int sm_count = device_sm_count();
  const unsigned grid = (unsigned)std::min<int64_t>(n_items, sm_count);
  const double rows = (double)B * a.T;
  ProfScope prof("pair_tc", a.C, a.C, 2.0 * rows * a.C * a.C * a.k * 2.0, rows * a.C * 4.0 * (a.accumulate ? 3.0 : 2.0), st);
#define PAIR_LAUNCH(CC, NN, AC)                                                                                                  \
  if (a.C == CC && (a.accumulate != 0) == AC) {                                                                                  \
    NVSE_CUDA_CHECK(cudaFuncSetAttribute(pair_tc_kernel<CC, NN, AC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)); \
    pair_tc_kernel<CC, NN, AC><<<grid, kThreads, p.smem, st>>>(k);                                                               \
  }
  PAIR_LAUNCH(128, 2, true) PAIR_LAUNCH(128, 2, false) PAIR_LAUNCH(256, 1, true) PAIR_LAUNCH(256, 1, false)
#undef PAIR_LAUNCH
  NVSE_LAUNCH_CHECK("pair_tc_kernel");
  return NVSE_OK;
}

NVSE_TC_ABORT_IMPL(pair)

}  // namespace nvse
