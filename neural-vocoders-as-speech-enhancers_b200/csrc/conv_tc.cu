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
// Warp roles (192 threads): warps 0-3 epilogue (one TMEM lane quarter each), warp 4 weight
// producer, warp 5 TMEM allocator + MMA issuer.  All six warps stage activations first.
#include "conv_tc.cuh"

#include <cstdlib>

namespace nvse {

namespace {

constexpr int kThreads = 192;
constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
constexpr long long kTimeoutCycles = 400000000LL;  // ~0.2 s: no legitimate wait is within 1000x of this

__device__ unsigned int g_tc_abort = 0;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error flag, never as a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long start = clock64();
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xffu) == 0) {
      if (clock64() - start > kTimeoutCycles || *(volatile unsigned int*)&g_tc_abort) {
        atomicExch(&g_tc_abort, 1u);
        return false;
      }
    }
  }
  return true;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 B (rows 16 B apart).
//   lbo = byte distance between the two core matrices along K of one K=16 MMA
//   sbo = byte distance between consecutive 8-row groups along M / N
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float lrelu(float v, float s) { return v >= 0.0f ? v : v * s; }
// the two bf16 halves of a packed word, widened back to fp32 (exact)
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

struct KernelArgs {
  ConvTcArgs a;
  int min_off;     // smallest tap offset
  int rows;        // staged rows = 128 + (max_off - min_off)
  int rows_pad;    // rows rounded up to 8m+1 (conflict-free staging stores)
  int stages;      // weight ring depth
  int kc;          // channels per weight stage (min(Cin, 64))
};

__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const KernelArgs k) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  if (*(volatile unsigned int*)&g_tc_abort) return;  // a previous launch tripped a wait timeout
  const ConvTcArgs& a = k.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.y;
  const int t0 = blockIdx.x * kTileM;
  const int Cin = a.Cin, Cout = a.Cout;
  const int nchunk = Cin >> 3;
  const uint32_t act_bytes = (((uint32_t)nchunk * k.rows_pad * 16u) + 127u) & ~127u;  // one bf16 plane
  const uint32_t stage_bytes = (uint32_t)k.kc * Cout * 2u;
  const int split = a.split_act;  // activations as hi + lo bf16 planes (two MMAs per K step)
  uint8_t* act = smem_raw;
  uint8_t* wst = smem_raw + (split ? 2u : 1u) * act_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + (size_t)k.stages * stage_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kMaxStages), bar_accum = smem_u32(bars + 2 * kMaxStages);
  const uint32_t tmem_cols = Cout < 32 ? 32u : (uint32_t)Cout;  // power of two >= 32 by construction

  if (tid == 0) {
    for (int s = 0; s < k.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_accum, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(smem_u32(tmem_slot), tmem_cols);

  // ---- stage the activation tile (all warps): lrelu + bf16 + K-major core-matrix layout ------
  {
    const int items = k.rows * nchunk;
    const int cshift = 31 - __clz(nchunk);
    if (a.in_bf16) {
      const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(a.x) + b * a.x_bstride;
      for (int e = tid; e < items; e += kThreads) {
        const int chunk = e & (nchunk - 1), r = e >> cshift;
        const int t = t0 + k.min_off + r;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (t >= 0 && t < a.Tin) v = __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)t * Cin + chunk * 8));
        *reinterpret_cast<uint4*>(act + ((size_t)chunk * k.rows_pad + r) * 16) = v;
      }
    } else {
      const float* xb = reinterpret_cast<const float*>(a.x) + b * a.x_bstride;
      const float s = a.in_slope;
      for (int e = tid; e < items; e += kThreads) {
        const int chunk = e & (nchunk - 1), r = e >> cshift;
        const int t = t0 + k.min_off + r;
        uint4 v = make_uint4(0u, 0u, 0u, 0u), lo = make_uint4(0u, 0u, 0u, 0u);
        if (t >= 0 && t < a.Tin) {
          const float4* src = reinterpret_cast<const float4*>(xb + (int64_t)t * Cin + chunk * 8);
          const float4 f0 = __ldg(src), f1 = __ldg(src + 1);
          const float f[8] = {lrelu(f0.x, s), lrelu(f0.y, s), lrelu(f0.z, s), lrelu(f0.w, s),
                              lrelu(f1.x, s), lrelu(f1.y, s), lrelu(f1.z, s), lrelu(f1.w, s)};
          v.x = pack_bf16(f[0], f[1]);
          v.y = pack_bf16(f[2], f[3]);
          v.z = pack_bf16(f[4], f[5]);
          v.w = pack_bf16(f[6], f[7]);
          if (split) {  // residual of the first rounding, itself rounded to bf16: ~16 mantissa bits in total
            lo.x = pack_bf16(f[0] - bf16_lo(v.x), f[1] - bf16_hi(v.x));
            lo.y = pack_bf16(f[2] - bf16_lo(v.y), f[3] - bf16_hi(v.y));
            lo.z = pack_bf16(f[4] - bf16_lo(v.z), f[5] - bf16_hi(v.z));
            lo.w = pack_bf16(f[6] - bf16_lo(v.w), f[7] - bf16_hi(v.w));
          }
        }
        *reinterpret_cast<uint4*>(act + ((size_t)chunk * k.rows_pad + r) * 16) = v;
        if (split) *reinterpret_cast<uint4*>(act + act_bytes + ((size_t)chunk * k.rows_pad + r) * 16) = lo;
      }
    }
  }
  fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nkc = Cin / k.kc;
  const int n_iters = a.taps.ntaps * nkc;

  if (warp == 4) {
    // ===== weight producer: one bulk copy per (tap, K chunk) stage =====
    if (lane == 0) {
      int it = 0;
      for (int tap = 0; tap < a.taps.ntaps; ++tap) {
        const __nv_bfloat16* wsrc = a.wimg + (size_t)a.taps.widx[tap] * nkc * (stage_bytes / 2);
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
  } else if (warp == 5) {
    // ===== MMA issuer: a single thread drives the tensor core =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cout >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t act_base = smem_u32(act);
      const uint32_t a_lbo = (uint32_t)k.rows_pad * 16u, b_lbo = (uint32_t)Cout * 16u, sbo = 128u;
      int it = 0;
      for (int tap = 0; tap < a.taps.ntaps; ++tap) {
        const uint32_t row_shift = (uint32_t)(a.taps.off[tap] - k.min_off);
        for (int kc = 0; kc < nkc; ++kc, ++it) {
          const int s = it % k.stages;
          const uint32_t par = (it / k.stages) & 1;
          if (!mbar_wait(bar_full + 8 * s, par)) goto done;
          tc_fence_after();
          const uint32_t w_base = smem_u32(wst + (size_t)s * stage_bytes);
          for (int kk = 0; kk < k.kc / 16; ++kk) {
            const uint32_t a_addr = act_base + ((uint32_t)(kc * (k.kc / 8) + 2 * kk) * k.rows_pad + row_shift) * 16u;
            const uint32_t b_addr = w_base + (uint32_t)(2 * kk) * b_lbo;
            const uint64_t bd = umma_desc(b_addr, b_lbo, sbo);
            tc_mma_bf16(tmem_base, umma_desc(a_addr, a_lbo, sbo), bd, idesc, (it | kk) != 0 ? 1u : 0u);
            if (split) tc_mma_bf16(tmem_base, umma_desc(a_addr + act_bytes, a_lbo, sbo), bd, idesc, 1u);
          }
          tc_commit(bar_empty + 8 * s);  // stage reusable once these MMAs have read it
        }
      }
      tc_commit(bar_accum);  // accumulator complete -> epilogue
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> global =====
    const bool ok = mbar_wait(bar_accum, 0);
    tc_fence_after();
    if (ok) {
      const int row = warp * 32 + lane;
      const int t = t0 + row;
      const int64_t orow = (int64_t)a.out_mul * t + a.out_add;
      const bool valid = t < a.Trows && orow < a.Tout;
      for (int c0 = 0; c0 < Cout; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (!valid) continue;
        if (a.out_bf16) {
          __nv_bfloat16* yr = reinterpret_cast<__nv_bfloat16*>(a.y) + b * a.y_bstride + orow * Cout + c0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = q * 8 + j * 2;
              const float f0 = lrelu(__uint_as_float(v[c]) + __ldg(a.bias + c0 + c), a.out_slope);
              const float f1 = lrelu(__uint_as_float(v[c + 1]) + __ldg(a.bias + c0 + c + 1), a.out_slope);
              op[j] = pack_bf16(f0, f1);
            }
            *reinterpret_cast<uint4*>(yr + q * 8) = o;
          }
        } else {
          float* yr = reinterpret_cast<float*>(a.y) + b * a.y_bstride + orow * Cout + c0;
          const float* rr = a.residual ? a.residual + b * a.y_bstride + orow * Cout + c0 : nullptr;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bq = __ldg(reinterpret_cast<const float4*>(a.bias + c0 + q * 4));
            float4 o = make_float4(__uint_as_float(v[q * 4 + 0]) + bq.x, __uint_as_float(v[q * 4 + 1]) + bq.y,
                                   __uint_as_float(v[q * 4 + 2]) + bq.z, __uint_as_float(v[q * 4 + 3]) + bq.w);
            if (a.out_slope != 1.0f) {
              o.x = lrelu(o.x, a.out_slope); o.y = lrelu(o.y, a.out_slope);
              o.z = lrelu(o.z, a.out_slope); o.w = lrelu(o.w, a.out_slope);
            }
            if (rr) {
              const float4 r4 = *reinterpret_cast<const float4*>(rr + q * 4);
              o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
            }
            o.x *= a.out_scale; o.y *= a.out_scale; o.z *= a.out_scale; o.w *= a.out_scale;
            if (a.accumulate) {
              const float4 y4 = *reinterpret_cast<const float4*>(yr + q * 4);
              o.x += y4.x; o.y += y4.y; o.z += y4.z; o.w += y4.w;
            }
            *reinterpret_cast<float4*>(yr + q * 4) = o;
          }
        }
      }
    }
  }
done:
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, tmem_cols);
}

// fp32 [j][ci][co] -> bf16 [j][ci/KC][(ci%KC)/8][co][ci%8]
__global__ void __launch_bounds__(256) pack_weight_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ img,
                                                              int Cin, int Cout, int ktaps, int kc) {
  const int64_t n = (int64_t)Cin * Cout * ktaps;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int co = e % Cout, ci = (e / Cout) % Cin, j = e / ((int64_t)Cout * Cin);
    const int64_t dst = ((((int64_t)j * (Cin / kc) + ci / kc) * (kc / 8) + (ci % kc) / 8) * Cout + co) * 8 + (ci % 8);
    img[dst] = __float2bfloat16_rn(w[e]);
  }
}

}  // namespace

int launch_pack_weight_tc(const float* w_kio, __nv_bfloat16* img, int Cin, int Cout, int ktaps, cudaStream_t st) {
  NVSE_REQUIRE(tc_supported(Cin, Cout), NVSE_ERR_UNSUPPORTED, "tensor-core conv: Cin=%d / Cout=%d unsupported", Cin, Cout);
  const int64_t n = (int64_t)Cin * Cout * ktaps;
  pack_weight_tc_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, st>>>(w_kio, img, Cin, Cout, ktaps,
                                                                                            tc_kchunk(Cin));
  NVSE_LAUNCH_CHECK("pack_weight_tc_kernel");
  return NVSE_OK;
}

static constexpr size_t kSmemBudget = 200 * 1024;
static constexpr size_t kSmemTail = sizeof(uint64_t) * (2 * kMaxStages + 1) + 16;

size_t tc_smem_bytes(int Cin, int Cout, int tap_span, bool split, int stages) {
  const int rows = kTileM + tap_span;
  const int rows_pad = ((rows + 6) / 8) * 8 + 1;
  const size_t plane = ((size_t)(Cin / 8) * rows_pad * 16 + 127) & ~(size_t)127;
  return (split ? 2 : 1) * plane + (size_t)stages * tc_kchunk(Cin) * Cout * 2 + kSmemTail;
}
bool tc_split_fits(int Cin, int Cout, int tap_span) { return tc_smem_bytes(Cin, Cout, tap_span, true, 2) <= kSmemBudget; }

int launch_conv_tc(const ConvTcArgs& a, int64_t B, cudaStream_t st) {
  NVSE_REQUIRE(tc_supported(a.Cin, a.Cout), NVSE_ERR_UNSUPPORTED, "tensor-core conv: Cin=%d / Cout=%d unsupported", a.Cin, a.Cout);
  NVSE_REQUIRE(a.taps.ntaps >= 1 && a.taps.ntaps <= kMaxTaps, NVSE_ERR_INVALID, "tensor-core conv: bad tap count");
  NVSE_REQUIRE(B <= 65535, NVSE_ERR_INVALID, "tensor-core conv: batch %lld exceeds 65535 per launch", (long long)B);
  NVSE_REQUIRE(!(a.out_bf16 && (a.residual || a.accumulate)), NVSE_ERR_INVALID, "tensor-core conv: bf16 output takes no residual");
  if (B == 0 || a.Trows <= 0) return NVSE_OK;
  KernelArgs k;
  k.a = a;
  int mn = a.taps.off[0], mx = a.taps.off[0];
  for (int i = 1; i < a.taps.ntaps; ++i) {
    mn = std::min(mn, a.taps.off[i]);
    mx = std::max(mx, a.taps.off[i]);
  }
  k.min_off = mn;
  k.rows = kTileM + (mx - mn);
  k.rows_pad = ((k.rows + 6) / 8) * 8 + 1;  // smallest 8m+1 >= rows
  k.kc = tc_kchunk(a.Cin);
  NVSE_REQUIRE(!(a.split_act && a.in_bf16), NVSE_ERR_INVALID, "tensor-core conv: split activations need fp32 input");
  const size_t act_bytes = (a.split_act ? 2 : 1) * (((size_t)(a.Cin / 8) * k.rows_pad * 16 + 127) & ~(size_t)127);
  const size_t stage_bytes = (size_t)k.kc * a.Cout * 2;
  const size_t tail = kSmemTail, budget = kSmemBudget;
  NVSE_REQUIRE(act_bytes + 2 * stage_bytes + tail <= budget, NVSE_ERR_UNSUPPORTED,
               "tensor-core conv: tile needs %zu B of shared memory", act_bytes + 2 * stage_bytes + tail);
  const int n_iters = a.taps.ntaps * (a.Cin / k.kc);
  int stages = (int)std::min<size_t>((budget - act_bytes - tail) / stage_bytes, (size_t)4);
  stages = std::max(2, std::min(stages, std::max(2, n_iters)));
  // prefer two CTAs per SM (one CTA's staging / epilogue overlaps the other's MMAs)
  while (stages > 2 && act_bytes + stages * stage_bytes + tail > 110 * 1024 &&
         act_bytes + 2 * stage_bytes + tail <= 110 * 1024) --stages;
  k.stages = stages;
  const size_t smem = act_bytes + stages * stage_bytes + tail;
  NVSE_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(budget + 1024)));
  dim3 grid((unsigned)((a.Trows + kTileM - 1) / kTileM), (unsigned)B);
  const double rows = (double)B * a.Trows;
  ProfScope prof("conv_tc", a.Cin, a.Cout, 2.0 * rows * a.Cin * a.Cout * a.taps.ntaps,
                 rows * (a.Cin * (a.in_bf16 ? 2.0 : 4.0) + a.Cout * (a.out_bf16 ? 2.0 : 4.0) * (a.accumulate ? 2.0 : 1.0) +
                         (a.residual ? 4.0 * a.Cout : 0.0)),
                 st);
  conv_tc_kernel<<<grid, kThreads, smem, st>>>(k);
  NVSE_LAUNCH_CHECK("conv_tc_kernel");
  return NVSE_OK;
}

int tc_abort_status(bool reset, unsigned int* flag) {
  unsigned int v = 0;
  NVSE_CUDA_CHECK(cudaMemcpyFromSymbol(&v, g_tc_abort, sizeof(v)));
  if (reset && v) {
    const unsigned int z = 0;
    NVSE_CUDA_CHECK(cudaMemcpyToSymbol(g_tc_abort, &z, sizeof(z)));
  }
  *flag = v;
  return NVSE_OK;
}

}  // namespace nvse
