// Shared host/device helpers for the nvse_b200 library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/nvse_b200.h"

namespace nvse {

// ---- error plumbing ------------------------------------------------------------------
std::string& last_error_slot();
int fail(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define NVSE_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::nvse::fail(NVSE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                          __FILE__, __LINE__);                                             \
  } while (0)

// Call right after a <<<>>> launch: counts it and surfaces launch-configuration errors.
#define NVSE_LAUNCH_CHECK(name)                                                            \
  do {                                                                                     \
    ::nvse::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess)                                                                 \
      return ::nvse::fail(NVSE_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)

#define NVSE_REQUIRE(cond, code, ...)                     \
  do {                                                    \
    if (!(cond)) return ::nvse::fail(code, __VA_ARGS__);  \
  } while (0)

// SM count of the CURRENT device (cached per device index: one process may drive several GPUs)
int device_sm_count();

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Optional per-launch timing (nvse_profile_begin / nvse_profile_end): when enabled every launch
// site brackets its kernel with a pair of CUDA events on the launching stream and tags it with
// its algorithmic FLOPs and bytes; bench.py turns the totals into the roofline numbers.
struct ProfScope {
  int slot = -1;
  cudaStream_t st;
  ProfScope(const char* kernel, int c_in, int c_out, double flops, double bytes, cudaStream_t stream);
  ~ProfScope();
};

// "T32" activation layout of the tensor-core path: fp32 [T, C] stored as [T/32][C/4][32][4], i.e.
// 32-row blocks in which the 4-channel groups of consecutive rows are contiguous.  The tensor-core
// kernels move activations with one thread per row (a TMEM lane is a row): in this layout a warp's
// 16-byte accesses to 32 consecutive rows form ONE 512-byte segment, where channels-last scatters
// them over 32 cache lines.  Rows are padded to a multiple of 32 per utterance.
__host__ __device__ inline int64_t t32_off(int64_t t, int c, int C) {
  return ((t >> 5) * (C >> 2) + (c >> 2)) * 128 + (t & 31) * 4 + (c & 3);
}
__host__ __device__ inline int64_t t32_rows(int64_t T) { return (T + 31) / 32 * 32; }

// Per-utterance valid row counts of a padded (ragged) batch: utterance b has min(T, lens[b] * mul + add) valid rows at this
// stage (lens = mel frames per utterance on the device; every stage's length is an affine function of it), T when lens is
// null.  Rows at or beyond it are treated exactly like rows beyond T: they read as zero (the convolutions' zero padding).
struct RowLens {
  const int* lens;
  int mul, add;
};
__device__ __forceinline__ int valid_rows(const RowLens& l, int64_t b, int T) {
  if (!l.lens) return T;
  const int v = l.lens[b] * l.mul + l.add;
  return v < T ? (v > 0 ? v : 0) : T;
}

constexpr int kMaxTaps = 16;

// One "tap-list" convolution over channels-last activations.  A plain dilated Conv1d is a
// tap list with offsets j*d - pad; one phase of a ConvTranspose1d is a short tap list with
// an output row map  row = out_mul * t + out_add  (polyphase decomposition).
struct ConvTaps {
  int ntaps;
  int off[kMaxTaps];   // input row offset of tap i relative to the output row index t
  int widx[kMaxTaps];  // which [Cin x Cout] slice of the packed weight tensor tap i uses
};

}  // namespace nvse
