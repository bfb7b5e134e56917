// Layer-level C ABI entry points of the tensor-core path (test convenience: they pack the
// weights on the fly, which the generator handle does once at load time).
#include <cstdlib>

#include "conv_f32.cuh"
#include "conv_tc.cuh"
#include "generator.cuh"
#include "resblock_tc.cuh"
#include "ups_tc.cuh"
#include "grad.cuh"

using namespace nvse;

namespace {
struct Scratch {
  void* p = nullptr;
  cudaStream_t st;
  explicit Scratch(cudaStream_t s) : st(s) {}
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes, st); }
  ~Scratch() {
    if (p) cudaFreeAsync(p, st);
  }
};
}  // namespace

extern "C" int nvse_conv1d_bf16(const float* x, const float* w, const float* bias, const float* residual, float* y,
                                int64_t B, int64_t T, int Cin, int Cout, int k, int dilation, float in_slope,
                                float out_scale, int accumulate, void* stream) {
  NVSE_REQUIRE(x && w && y && bias, NVSE_ERR_INVALID, "nvse_conv1d_bf16: null argument");
  NVSE_REQUIRE(B >= 0 && T >= 0 && T <= 0x7fffffff && dilation >= 1, NVSE_ERR_INVALID, "nvse_conv1d_bf16: bad shape");
  NVSE_REQUIRE(k >= 1 && (k & 1) && k <= kMaxTaps, NVSE_ERR_UNSUPPORTED, "nvse_conv1d_bf16: k=%d (odd k <= %d only)", k, kMaxTaps);
  NVSE_REQUIRE(tc_supported(Cin, Cout), NVSE_ERR_UNSUPPORTED, "nvse_conv1d_bf16: Cin=%d / Cout=%d unsupported", Cin, Cout);
  cudaStream_t st = as_stream(stream);
  Scratch wk(st), img(st);
  NVSE_CUDA_CHECK(wk.alloc(sizeof(float) * (size_t)Cin * Cout * k));
  NVSE_CUDA_CHECK(img.alloc(sizeof(__nv_bfloat16) * (size_t)Cin * Cout * k));
  if (int rc = launch_repack_weight(w, (float*)wk.p, Cin, Cout, k, false, st)) return rc;
  if (int rc = launch_pack_weight_tc((const float*)wk.p, (__nv_bfloat16*)img.p, Cin, Cout, k, st)) return rc;
  ConvTcArgs a{};
  a.x = x; a.x_bstride = T * Cin; a.Tin = (int)T; a.Cin = Cin; a.Cout = Cout;
  a.wimg = (const __nv_bfloat16*)img.p; a.bias = bias; a.residual = residual;
  a.y = y; a.y_bstride = T * Cout; a.Tout = (int)T;
  conv1d_taps(k, dilation, &a.taps);
  a.out_mul = 1; a.out_add = 0; a.Trows = (int)T;
  a.in_slope = in_slope; a.out_slope = 1.0f; a.out_scale = out_scale; a.accumulate = accumulate;
  return launch_conv_tc(a, B, st);
}

extern "C" int nvse_conv_transpose1d_bf16(const float* x, const float* w, const float* bias, float* y, int64_t B,
                                          int64_t T, int Cin, int Cout, int k, int stride, int padding, float in_slope,
                                          void* stream) {
  NVSE_REQUIRE(x && w && y && bias, NVSE_ERR_INVALID, "nvse_conv_transpose1d_bf16: null argument");
  NVSE_REQUIRE(B >= 0 && T >= 1 && stride >= 1 && padding >= 0 && k >= 1, NVSE_ERR_INVALID, "nvse_conv_transpose1d_bf16: bad shape");
  NVSE_REQUIRE(tc_supported(Cin, Cout), NVSE_ERR_UNSUPPORTED, "nvse_conv_transpose1d_bf16: Cin=%d / Cout=%d unsupported", Cin, Cout);
  const int64_t Tout = (T - 1) * stride - 2 * padding + k;
  NVSE_REQUIRE(Tout > 0 && Tout <= 0x7fffffff, NVSE_ERR_INVALID, "nvse_conv_transpose1d_bf16: bad output length");
  cudaStream_t st = as_stream(stream);
  Scratch wk(st), img(st);
  NVSE_CUDA_CHECK(wk.alloc(sizeof(float) * (size_t)Cin * Cout * k));
  NVSE_CUDA_CHECK(img.alloc(sizeof(__nv_bfloat16) * (size_t)Cin * Cout * k));
  const char* f16env = std::getenv("NVSE_TC_F16");  // tests: IEEE-half operands, as the generator runs its upsamplers
  const bool f16 = f16env && f16env[0] != '0';
  if (int rc = launch_repack_weight(w, (float*)wk.p, Cin, Cout, k, true, st)) return rc;
  if (int rc = launch_pack_weight_tc((const float*)wk.p, (__nv_bfloat16*)img.p, Cin, Cout, k, st, f16)) return rc;
  const int nph = (int)std::min<int64_t>(stride, Tout);
  for (int r0 = 0; r0 < nph; r0 += kTcMaxPhases) {
    const int n = std::min(kTcMaxPhases, nph - r0);
    ConvTaps taps[kTcMaxPhases];
    int out_add[kTcMaxPhases];
    for (int p = 0; p < n; ++p) {
      NVSE_REQUIRE(conv_transpose_phase_taps(k, stride, padding, r0 + p, &taps[p]) > 0, NVSE_ERR_UNSUPPORTED,
                   "nvse_conv_transpose1d_bf16: k < stride is not supported");
      out_add[p] = r0 + p;
    }
    ConvTcArgs a{};
    a.x = x; a.x_bstride = T * Cin; a.Tin = (int)T; a.Cin = Cin; a.Cout = Cout;
    a.wimg = (const __nv_bfloat16*)img.p; a.bias = bias;
    a.y = y; a.y_bstride = Tout * Cout; a.Tout = (int)Tout;
    a.out_mul = stride; a.Trows = (int)((Tout - r0 + stride - 1) / stride);
    a.in_slope = in_slope; a.out_slope = 1.0f; a.out_scale = 1.0f;
    a.ops_f16 = f16;
    if (int rc = launch_conv_tc_phases(a, taps, out_add, n, B, st)) return rc;
  }
  return NVSE_OK;
}

extern "C" int nvse_resblock1_bf16(const float* x, const float* const* w1, const float* const* b1, const float* const* w2,
                                   const float* const* b2, const int* dilations, int npairs, float* y, int64_t B, int64_t T,
                                   int C, int k, float out_scale, int accumulate, void* stream) {
  NVSE_REQUIRE(x && w1 && b1 && w2 && b2 && dilations && y, NVSE_ERR_INVALID, "nvse_resblock1_bf16: null argument");
  NVSE_REQUIRE(B >= 0 && T >= 0 && T <= 0x7fffffff, NVSE_ERR_INVALID, "nvse_resblock1_bf16: bad shape");
  NVSE_REQUIRE(npairs >= 1 && npairs <= kRbMaxPairs, NVSE_ERR_UNSUPPORTED, "nvse_resblock1_bf16: %d pairs (1..%d)", npairs, kRbMaxPairs);
  NVSE_REQUIRE(rb_supported(C, k, dilations, npairs), NVSE_ERR_UNSUPPORTED, "nvse_resblock1_bf16: C=%d k=%d unsupported", C, k);
  cudaStream_t st = as_stream(stream);
  const size_t wn = (size_t)C * C * k;
  Scratch wk(st), img(st);
  NVSE_CUDA_CHECK(wk.alloc(sizeof(float) * wn));
  NVSE_CUDA_CHECK(img.alloc(sizeof(__nv_bfloat16) * wn * 2 * npairs));
  ResblockTcArgs a{};
  a.x = x; a.y = y; a.T = (int)T; a.C = C; a.k = k; a.npairs = npairs;
  a.slope = 0.1f; a.out_scale = out_scale; a.accumulate = accumulate;
  // NVSE_RB_T32 (timing, tools/rb_bench.py): treat x / y as T32 buffers as they are (needs T % 32 == 0).
  // NVSE_RB_LAYER_T32 (tests): convert x (and y when accumulating) to T32 scratch copies, run the T32
  // kernels the generator runs -- including the pipelined pair kernel where it applies -- and convert back.
  const bool bench_t32 = std::getenv("NVSE_RB_T32") != nullptr && (T % 32 == 0);
  const bool layer_t32 = !bench_t32 && std::getenv("NVSE_RB_LAYER_T32") != nullptr;
  const char* pp = std::getenv("NVSE_PAIRPIPE");
  const bool pairpipe = !(pp && pp[0] == '0');
  a.t32 = bench_t32 || layer_t32;
  const char* h16 = std::getenv("NVSE_RB_H16");  // tests / experiments: IEEE-half c1 -> c2 intermediate (what the generator uses at C <= 32)
  a.h_fp16 = h16 && h16[0] != '0' && C <= 64;
  Scratch xt(st), yt(st);
  if (layer_t32) {
    const size_t bytes = sizeof(float) * (size_t)B * t32_rows(T) * C;
    NVSE_CUDA_CHECK(xt.alloc(bytes));
    NVSE_CUDA_CHECK(yt.alloc(bytes));
    if (int rc = launch_relayout_t32(x, (float*)xt.p, B, T, C, true, st)) return rc;
    if (accumulate)
      if (int rc = launch_relayout_t32(y, (float*)yt.p, B, T, C, true, st)) return rc;
    a.x = (const float*)xt.p;
    a.y = (float*)yt.p;
  }
  for (int m = 0; m < npairs; ++m) {
    NVSE_REQUIRE(w1[m] && w2[m] && b1[m] && b2[m], NVSE_ERR_INVALID, "nvse_resblock1_bf16: null tensor in pair %d", m);
    __nv_bfloat16* i1 = (__nv_bfloat16*)img.p + (size_t)(2 * m) * wn;
    __nv_bfloat16* i2 = i1 + wn;
    if (int rc = launch_repack_weight(w1[m], (float*)wk.p, C, C, k, false, st)) return rc;
    if (int rc = launch_pack_weight_tc((const float*)wk.p, i1, C, C, k, st)) return rc;
    if (int rc = launch_repack_weight(w2[m], (float*)wk.p, C, C, k, false, st)) return rc;
    if (int rc = launch_pack_weight_tc((const float*)wk.p, i2, C, C, k, st, a.h_fp16 != 0)) return rc;
    a.pair[m] = RbPair{i1, i2, b1[m], b2[m], dilations[m]};
  }
  int rc;
  if (a.t32 && pairpipe && npairs == 1 && !a.h_fp16 && pair_supported(C, k, dilations[0])) rc = launch_pair_tc(a, B, st);
  else rc = launch_resblock_tc(a, B, st);
  if (rc) return rc;
  if (layer_t32) return launch_relayout_t32((const float*)yt.p, y, B, T, C, false, st);
  return NVSE_OK;
}

extern "C" int nvse_tc_abort_status(int reset, int* flag) {
  NVSE_REQUIRE(flag, NVSE_ERR_INVALID, "nvse_tc_abort_status: null argument");
  unsigned int v = 0, v2 = 0;
  if (int rc = tc_abort_status(reset != 0, &v)) return rc;
  if (int rc = rb_abort_status(reset != 0, &v2)) return rc;
  unsigned int v3 = 0, v4 = 0;
  if (int rc = pair_abort_status(reset != 0, &v3)) return rc;
  if (int rc = wgrad_abort_status(reset != 0, &v4)) return rc;
  unsigned int v5 = 0;
  if (int rc = ups_abort_status(reset != 0, &v5)) return rc;
  *flag = (int)(v | v2 | v3 | v4 | v5);
  if (reset) tc_abort_host_clear();
  return NVSE_OK;
}
