// Inline-PTX wrappers for the sm_100a tensor-core kernels: mbarrier, TMA bulk copy, tcgen05
// (TMEM alloc / mma / commit / ld) and the UMMA shared-memory descriptor.
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"

namespace nvse {
namespace tc {

constexpr long long kTimeoutCycles = 400000000LL;  // ~0.2 s: no legitimate wait is within 1000x of this
// set when a bounded wait times out; one copy per translation unit that includes this header
// (conv_tc.cu, resblock_tc.cu); nvse_tc_abort_status reports the OR of the copies
static __device__ unsigned int g_tc_abort = 0;
// device pointer of ONE pinned, mapped host word shared by every translation unit and device (tc_abort.cu): a timeout
// also raises it, so that the host sees the failure with a plain memory read at the next library call -- no
// synchronising cudaMemcpyFromSymbol on the product path.  Null until tc_abort_bind() ran for this device.
static __device__ unsigned int* g_tc_abort_host = nullptr;

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
        if (g_tc_abort_host) {
          *(volatile unsigned int*)g_tc_abort_host = 1u;
          __threadfence_system();
        }
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
// two floats -> packed IEEE half, saturated to the finite range
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float lrelu(float v, float s) { return v >= 0.0f ? v : v * s; }
// the two bf16 halves of a packed word, widened back to fp32 (exact)
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }


__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// One lane of the (converged) warp; the rest of the warp runs the same uniform code so that the
// descriptors live in uniform registers and an MMA issue is a couple of instructions.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// descriptor halves: lo = (addr >> 4) | (lbo >> 4) << 16  (both 14-bit fields), hi = (sbo >> 4) | version bit 46
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo) {
  return ((saddr & 0x3FFFFu) >> 4) | (((lbo >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t umma_desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ void tc_mma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Weight-stationary form (tcgen05.mma.ws): the B tile is latched in collector buffer b0 by the first MMA of a group
// (FILL), re-used by the following ones (USE) and released by the last (LAST), so that a (tap, k-step) weight slice
// shared by the row tiles of a CTA is read from shared memory once.  Legal for N >= 64 only (N = 32 traps); measured
// (tools/probe/mma_probe4.cu, profiles/r02_mma_probe4.txt): N = 64 48.1 -> 42.0 cycles per MMA over 4 tiles, 45.1 over 2;
// a loss at N = 128 (64 -> 69).  Same accumulator layout as the plain form (lane = row).
enum WsMode { kWsNone = 0, kWsFill = 1, kWsUse = 2, kWsLast = 3 };
template <int MODE>
__device__ __forceinline__ void tc_mma_ws_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
#define NVSE_WS_ASM(OPC)                                                                                                  \
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\n" \
               OPC " [%0], da, db, %5, p;\n}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc),  \
               "r"(accumulate) : "memory")
  if (MODE == kWsFill) NVSE_WS_ASM("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill");
  else if (MODE == kWsUse) NVSE_WS_ASM("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::use");
  else if (MODE == kWsLast) NVSE_WS_ASM("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse");
  else NVSE_WS_ASM("tcgen05.mma.cta_group::1.kind::f16");
#undef NVSE_WS_ASM
}
// MMA number j of a group of NGRP that share one B tile (j is a constant after unrolling)
template <bool WS, int NGRP>
__device__ __forceinline__ void tc_mma_group(int j, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  if (!WS || NGRP < 2) tc_mma_ws_lohi<kWsNone>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
  else if (j == 0) tc_mma_ws_lohi<kWsFill>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
  else if (j == NGRP - 1) tc_mma_ws_lohi<kWsLast>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
  else tc_mma_ws_lohi<kWsUse>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
}

// registers -> TMEM: thread i writes 16 consecutive fp32 columns of lane base+i
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace tc

// Host side of the per-translation-unit flag: `prefix`_abort_status (read, optionally reset; synchronous -- tests and
// bench.py), `prefix`_abort_bind (point this TU's copy at the mapped host word), `prefix`_abort_clear (stream-ordered reset).
#define NVSE_TC_ABORT_IMPL(prefix)                                                                    \
  int prefix##_abort_status(bool reset, unsigned int* flag) {                                         \
    unsigned int v = 0;                                                                               \
    NVSE_CUDA_CHECK(cudaMemcpyFromSymbol(&v, tc::g_tc_abort, sizeof(v)));                             \
    if (reset && v) {                                                                                 \
      const unsigned int z = 0;                                                                       \
      NVSE_CUDA_CHECK(cudaMemcpyToSymbol(tc::g_tc_abort, &z, sizeof(z)));                             \
    }                                                                                                 \
    *flag = v;                                                                                        \
    return NVSE_OK;                                                                                   \
  }                                                                                                   \
  int prefix##_abort_bind(unsigned int* host_word_dev) {                                              \
    NVSE_CUDA_CHECK(cudaMemcpyToSymbol(tc::g_tc_abort_host, &host_word_dev, sizeof(host_word_dev)));  \
    return NVSE_OK;                                                                                   \
  }                                                                                                   \
  int prefix##_abort_clear(cudaStream_t st) {                                                         \
    static const unsigned int z = 0;                                                                  \
    NVSE_CUDA_CHECK(cudaMemcpyToSymbolAsync(tc::g_tc_abort, &z, sizeof(z), 0, cudaMemcpyHostToDevice, st)); \
    return NVSE_OK;                                                                                   \
  }

}  // namespace nvse
