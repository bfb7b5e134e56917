// Generator handle: HiFiGAN.forward (reference Models/hifigan.py:108-124) and
// iSTFTNet.forward (Models/istftnet.py:299-318) as a fixed launch sequence over
// channels-last activations.  The handle owns the packed weights; the caller owns the
// workspace (four ping-pong activation buffers).
#include "generator.cuh"
#include "conv_tc.cuh"
#include "resblock_tc.cuh"
#include "ups_tc.cuh"

#include <cstdlib>
#include <cstring>

namespace nvse {

static int build_layers(nvse_generator* g) {
  const nvse_generator_config& c = g->cfg;
  auto add = [&](const std::string& name, bool transposed, int cin, int cout, int k, int dil, int stride, int pad) {
    Layer L;
    L.name = name;
    L.transposed = transposed;
    L.Cin = cin; L.Cout = cout; L.k = k; L.dilation = dil; L.stride = stride; L.padding = pad;
    L.grad_off = g->layers.empty() ? 0 : g->layers.back().grad_off + (int64_t)g->layers.back().Cin * g->layers.back().Cout * g->layers.back().k + g->layers.back().Cout;
    g->index[name] = (int)g->layers.size();
    g->layers.push_back(L);
  };
  const int c0 = c.initial_channel;
  add("conv_pre", false, c.in_channels, c0, 7, 1, 1, 3);  // hifigan.py:89
  for (int i = 0; i < c.num_upsamples; ++i) {             // hifigan.py:93-96
    const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i];
    add("ups." + std::to_string(i), true, c0 >> i, c0 >> (i + 1), k, 1, u, (k - u) / 2);
  }
  int ch = c0;
  for (int i = 0; i < c.num_upsamples; ++i) {  // hifigan.py:98-102
    ch = c0 >> (i + 1);
    for (int j = 0; j < c.num_kernels; ++j) {
      const int k = c.resblock_kernel_sizes[j];
      const std::string p = "resblocks." + std::to_string(i * c.num_kernels + j);
      for (int m = 0; m < c.num_dilations[j]; ++m) {
        const int d = c.resblock_dilations[j][m];
        if (c.resblock_type == 1) {
          add(p + ".convs1." + std::to_string(m), false, ch, ch, k, d, 1, (k * d - d) / 2);
          add(p + ".convs2." + std::to_string(m), false, ch, ch, k, 1, 1, (k - 1) / 2);
        } else {
          add(p + ".convs." + std::to_string(m), false, ch, ch, k, d, 1, (k * d - d) / 2);
        }
      }
    }
  }
  const int out_ch = c.kind == NVSE_GEN_HIFIGAN ? 1 : c.istft_n_fft + 2;  // hifigan.py:104, istftnet.py:293
  add("conv_post", false, ch, out_ch, 7, 1, 1, 3);
  return NVSE_OK;
}

// floats per batch item of the largest activation the forward ever holds
static int64_t max_activation_elems(const nvse_generator* g, int64_t F) {
  const nvse_generator_config& c = g->cfg;
  int64_t m = std::max<int64_t>(F * c.in_channels, F * c.initial_channel);
  int64_t T = F;
  for (int i = 0; i < c.num_upsamples; ++i) {
    const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i];
    T = (T - 1) * u - 2 * ((k - u) / 2) + k;
    m = std::max<int64_t>(m, t32_rows(T) * (c.initial_channel >> (i + 1)));  // T32 pads rows to a multiple of 32
  }
  if (c.kind == NVSE_GEN_ISTFTNET) m = std::max<int64_t>(m, (T + 1) * (c.istft_n_fft + 2));
  return m;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// One Conv1d layer ("same" padding) with the fused prologue/epilogue, on the tensor-core path when
// `tc` is set and the layer has a bf16 image, otherwise on the fp32 CUDA-core path.
//   y = [accumulate ? y : 0] + out_scale * (conv(lrelu(x, in_slope)) + bias [+ residual])          (fp32 out)
//   y = bf16(lrelu(conv(lrelu(x, in_slope)) + bias, out_slope))                                      (bf16 out)
struct ConvIO {
  const void* x; bool x_bf16; const float* residual; void* y; bool y_bf16;
  float in_slope, out_slope, out_scale; int accumulate;
};
// fp32 out with out_slope != 1:  y = lrelu(conv + bias, out_slope)  (the fp32 c1 -> c2 intermediate of split layers)

static int run_conv(const Layer& L, bool tc, const ConvIO& io, int64_t B, int64_t T, cudaStream_t st) {
  ConvTaps taps;
  conv1d_taps(L.k, L.dilation, &taps);
  if (tc && L.w_bf16) {
    ConvTcArgs a{};
    a.x = io.x; a.x_bstride = T * L.Cin; a.Tin = (int)T; a.Cin = L.Cin; a.Cout = L.Cout; a.in_bf16 = io.x_bf16;
    a.wimg = reinterpret_cast<const __nv_bfloat16*>(L.w_bf16); a.bias = L.bias; a.residual = io.residual;
    a.y = io.y; a.y_bstride = T * L.Cout; a.Tout = (int)T; a.out_bf16 = io.y_bf16;
    a.taps = taps; a.out_mul = 1; a.out_add = 0; a.Trows = (int)T;
    a.in_slope = io.in_slope; a.out_slope = io.out_slope; a.out_scale = io.out_scale; a.accumulate = io.accumulate;
    a.split_act = L.tc_split && !io.x_bf16;
    return launch_conv_tc(a, B, st);
  }
  NVSE_REQUIRE(!io.x_bf16 && !io.y_bf16 && io.out_slope == 1.0f, NVSE_ERR_STATE, "layer %s: bf16 activations on the fp32 path", L.name.c_str());
  ConvF32Args a{};
  a.x = reinterpret_cast<const float*>(io.x); a.x_bstride = T * L.Cin; a.Tin = (int)T; a.Cin = L.Cin;
  a.w = L.w; a.bias = L.bias; a.residual = io.residual;
  a.y = reinterpret_cast<float*>(io.y); a.y_bstride = T * L.Cout; a.Tout = (int)T; a.Cout = L.Cout;
  a.taps = taps; a.out_mul = 1; a.out_add = 0; a.Trows = (int)T;
  a.in_slope = io.in_slope; a.out_scale = io.out_scale; a.accumulate = io.accumulate;
  return launch_conv_f32(a, B, st);
}

// ConvTranspose1d as `stride` polyphase tap-list convolutions.  On the tensor-core
// path up to 8 phases share one launch: the activation tile is staged once and the phases ping-pong
// between two TMEM accumulators.
int run_conv_transpose(const Layer& L, bool tc, const float* x, int64_t B, int64_t Tin, float* y, float in_slope,
                       cudaStream_t st, bool x_t32, bool y_t32, RowLens in_lens) {
  const int64_t Tout = (Tin - 1) * L.stride - 2 * L.padding + L.k;
  const int nph = (int)std::min<int64_t>(L.stride, Tout);
  static const bool ups_env = [] { const char* e = std::getenv("NVSE_UPS_TC"); return !(e && e[0] == '0'); }();
  static const bool ups_xcl_env = [] { const char* e = std::getenv("NVSE_UPS_XCL"); return !(e && e[0] == '0'); }();
  if (tc && ups_env && L.w_ups && y_t32 && (x_t32 || (ups_xcl_env && L.stride == 8 && !((L.Cin / 8) & (L.Cin / 8 - 1))))) {  // all phases in one persistent launch (ups_tc.cu)
    UpsTcArgs a{};
    a.x = x; a.x_cl = x_t32 ? 0 : 1; a.x_bstride = (x_t32 ? t32_rows(Tin) : Tin) * L.Cin; a.Tin = (int)Tin; a.Cin = L.Cin; a.Cout = L.Cout; a.stride = L.stride;
    a.wimg = L.w_ups; a.bias = L.bias; a.y = y; a.y_bstride = t32_rows(Tout) * L.Cout; a.in_slope = in_slope;
    a.in_lens = in_lens;
    return launch_ups_tc(a, B, st);
  }
  if (tc && L.w_bf16) {
    for (int r0 = 0; r0 < nph; r0 += kTcMaxPhases) {
      const int n = std::min(kTcMaxPhases, nph - r0);
      ConvTaps taps[kTcMaxPhases];
      int out_add[kTcMaxPhases];
      for (int p = 0; p < n; ++p) {
        NVSE_REQUIRE(conv_transpose_phase_taps(L.k, L.stride, L.padding, r0 + p, &taps[p]) > 0, NVSE_ERR_UNSUPPORTED,
                     "ConvTranspose1d %s: unsupported k/stride", L.name.c_str());
        out_add[p] = r0 + p;
      }
      ConvTcArgs a{};
      a.x = x; a.x_bstride = (x_t32 ? t32_rows(Tin) : Tin) * L.Cin; a.Tin = (int)Tin; a.Cin = L.Cin; a.Cout = L.Cout;
      a.wimg = reinterpret_cast<const __nv_bfloat16*>(L.tc_f16 ? L.w_f16 : L.w_bf16); a.bias = L.bias;
      a.ops_f16 = L.tc_f16;
      a.y = y; a.y_bstride = (y_t32 ? t32_rows(Tout) : Tout) * L.Cout; a.Tout = (int)Tout;
      a.x_t32 = x_t32; a.y_t32 = y_t32;
      a.out_mul = L.stride; a.Trows = (int)((Tout - r0 + L.stride - 1) / L.stride);
      a.in_slope = in_slope; a.out_slope = 1.0f; a.out_scale = 1.0f;
      a.split_act = L.tc_split && !L.tc_f16;
      a.in_lens = in_lens;
      if (int rc = launch_conv_tc_phases(a, taps, out_add, n, B, st)) return rc;
    }
    return NVSE_OK;
  }
  NVSE_REQUIRE(!x_t32 && !y_t32 && !in_lens.lens, NVSE_ERR_STATE, "ConvTranspose1d %s: the T32 layout / ragged batches need the tensor-core path", L.name.c_str());
  for (int r = 0; r < nph; ++r) {
    ConvF32Args a{};
    a.x = x; a.x_bstride = Tin * L.Cin; a.Tin = (int)Tin; a.Cin = L.Cin;
    a.w = L.w; a.bias = L.bias;
    a.y = y; a.y_bstride = Tout * L.Cout; a.Tout = (int)Tout; a.Cout = L.Cout;
    NVSE_REQUIRE(conv_transpose_phase_taps(L.k, L.stride, L.padding, r, &a.taps) > 0, NVSE_ERR_UNSUPPORTED,
                 "ConvTranspose1d %s: unsupported k/stride", L.name.c_str());
    a.out_mul = L.stride; a.out_add = r; a.Trows = (int)((Tout - r + L.stride - 1) / L.stride);
    a.in_slope = in_slope; a.out_scale = 1.0f;
    if (int rc = launch_conv_f32(a, B, st)) return rc;
  }
  return NVSE_OK;
}

// tc = false: fp32 CUDA cores everywhere.  tc = true: tcgen05 for the upsamplers and the MRF
// convolutions (bf16 operands, fp32 accumulate, fp32 residual stream; the c1 -> c2 intermediate of
// a ResBlock1 pair is stored activated, in bf16); conv_pre / conv_post stay fp32.
// Small jobs (B * frames <= kConcurrentFrames): a ResBlock launch fills only a fraction of the SMs, and a forward
// is a chain of ~25 dependent launches.  The ResBlocks of one MRF are independent, so they run concurrently on
// the caller's stream and two side streams, each into its own buffer, and are summed afterwards.
constexpr int64_t kConcurrentFrames = 1024;   // ~12 s of audio in total
constexpr int64_t kConcurrentRows = 48 * 1024;  // per stage: B * T rows below which a launch is < 1 wave
static int workspace_buffers(const nvse_generator* g, int64_t B, int64_t frames, int precision) {
  static const bool on = [] { const char* e = std::getenv("NVSE_CONCURRENT"); return !(e && e[0] == '0'); }();
  const bool small = on && precision == NVSE_PRECISION_BF16 && g->cfg.resblock_type == 1 && g->cfg.num_kernels >= 2 &&
                     g->cfg.num_kernels <= 3 && B * frames <= kConcurrentFrames;
  return small ? 12 : 4;
}

int ensure_side_streams(nvse_generator* g) {
  if (g->ev_fork) return NVSE_OK;
  NVSE_CUDA_CHECK(cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming));
  for (int q = 0; q < 2; ++q) {
    NVSE_CUDA_CHECK(cudaStreamCreateWithFlags(&g->side[q], cudaStreamNonBlocking));
    NVSE_CUDA_CHECK(cudaEventCreateWithFlags(&g->ev_join[q], cudaEventDisableTiming));
  }
  return NVSE_OK;
}

// out_i16 != null: the waveform is wanted as PCM_16 (int16) INSTEAD of float: fused into conv_post where the last
// kernel is the T32 conv_post kernel (HiFiGAN on the tensor-core plan), a separate quantisation pass otherwise.
// z[b][t][co] += sum_ci w[pad - t][ci][co] * lrelu(x[b][1][ci], 0.01) for t = 0 .. pad: the contribution of row 0 of
// ReflectionPad1d((1, 0)) (= x row 1; istftnet.py:296,312) to the first output rows of conv_post, which the shifted plain
// convolution of the tensor-core path reads as zero.  x in the T32 layout, z rows `pitch` floats apart, w = Layer::w [k][Cin][Cout].
__global__ void post_reflect_fix_kernel(const float* __restrict__ x, int64_t x_bstride, const float* __restrict__ w, float* __restrict__ z,
                                        int64_t z_bstride, int pitch, int Cin, int Cout, int pad) {
  const int64_t b = blockIdx.x;
  for (int e = threadIdx.x; e < (pad + 1) * Cout; e += blockDim.x) {
    const int t = e / Cout, co = e - t * Cout;
    float acc = 0.0f;
    for (int ci = 0; ci < Cin; ++ci) {
      const float v = x[b * x_bstride + t32_off(1, ci, Cin)];
      acc = fmaf(w[((int64_t)(pad - t) * Cin + ci) * Cout + co], v >= 0.0f ? v : v * 0.01f, acc);
    }
    z[b * z_bstride + (int64_t)t * pitch + co] += acc;
  }
}
static int launch_post_reflect_fix(const float* x_t32, const Layer& post, float* z, int pitch, int64_t B, int64_t T, cudaStream_t st) {
  const int pad = (post.k - 1) / 2;
  NVSE_REQUIRE(T >= pad + 1, NVSE_ERR_UNSUPPORTED, "iSTFTNet conv_post on the tensor cores needs at least %d rows", pad + 1);
  post_reflect_fix_kernel<<<(unsigned)B, 128, 0, st>>>(x_t32, t32_rows(T) * post.Cin, post.w, z, (T + 1) * pitch, pitch, post.Cin, post.Cout, pad);
  NVSE_LAUNCH_CHECK("post_reflect_fix_kernel");
  return NVSE_OK;
}

// frames_dev != null (ragged batch): mel frames per utterance on the device; utterance b is computed exactly as if it were
// alone with frames_dev[b] frames (every kernel treats its rows beyond that like rows beyond the end of the sequence); its
// output samples beyond its own length are undefined.  Tensor-core HiFiGAN plan only.
static int forward_impl(nvse_generator* g, bool tc, const float* mel, int64_t B, int64_t F, float* out, float* ws,
                        int64_t buf_elems, int nbuf, cudaStream_t st, int16_t* out_i16 = nullptr, const int* frames_dev = nullptr,
                        const float* mel_cl = nullptr) {  // mel_cl: the log-mel channels-last [B, F, conv_pre.pre_cin] (fused call)
  const nvse_generator_config& c = g->cfg;
  float* bufA = ws;  // conv_pre output, then the MRF accumulator of every stage
  float* bufU = ws + buf_elems;
  float* bufR = ws + 2 * buf_elems;
  float* bufT = ws + 3 * buf_elems;
  const float slope = 0.1f;  // LRELU_SLOPE, hifigan.py:7

  // [B, 80, F] -> channels-last, then conv_pre (hifigan.py:109).  16-bit path: a tensor-core GEMM with IEEE-half operands
  // (10 mantissa bits, like TF32: ~70 dB at this layer against the path's ~50 dB), the mel channels zero-padded to a
  // supported K and the output channels in slices of <= 256; fp32 path: CUDA cores.
  {
    const Layer& pre = g->layer("conv_pre");
    static const bool pre_tc_env = [] { const char* e = std::getenv("NVSE_PRE_TC"); return !(e && e[0] == '0'); }();
    NVSE_REQUIRE(!mel_cl || (tc && pre_tc_env && pre.pre_n > 0), NVSE_ERR_UNSUPPORTED, "the fused wav -> wav call needs conv_pre on the tensor cores");
    if (tc && pre_tc_env && pre.pre_n > 0) {
      if (!mel_cl)
        if (int rc = launch_transpose_pad(mel, bufR, B, c.in_channels, F, pre.pre_cin, st)) return rc;
      for (int sl = 0; sl < pre.pre_n; ++sl) {
        ConvTcArgs a{};
        a.x = mel_cl ? mel_cl : bufR; a.x_bstride = F * pre.pre_cin; a.Tin = (int)F; a.Cin = pre.pre_cin; a.Cout = pre.pre_cout;
        a.wimg = reinterpret_cast<const __nv_bfloat16*>(pre.w_pre[sl]); a.bias = pre.bias + sl * pre.pre_cout; a.ops_f16 = 1;
        a.y = bufA + sl * pre.pre_cout; a.y_bstride = F * pre.Cout; a.y_ld = pre.Cout; a.Tout = (int)F;
        conv1d_taps(pre.k, 1, &a.taps);
        a.out_mul = 1; a.out_add = 0; a.Trows = (int)F;
        a.in_slope = 1.0f; a.out_slope = 1.0f; a.out_scale = 1.0f;
        a.in_lens = RowLens{frames_dev, 1, 0};
        if (int rc = launch_conv_tc(a, B, st)) return rc;
      }
    } else {
      NVSE_REQUIRE(!frames_dev, NVSE_ERR_UNSUPPORTED, "ragged batches need conv_pre on the tensor cores");
      if (int rc = launch_transpose(mel, bufR, B, c.in_channels, F, st)) return rc;
      const ConvIO io{bufR, false, nullptr, bufA, false, 1.0f, 1.0f, 1.0f, 0};
      if (int rc = run_conv(pre, false, io, B, F, st)) return rc;
    }
  }
  // Fused tensor-core plan: every ResBlock1 as one launch (or one launch per pair, whichever the cost
  // model of resblock_tc.cu prefers).  When every stage can be fused the activations between the
  // kernels are kept in the T32 layout (common.cuh); otherwise the per-layer channels-last path runs.
  static const bool fuse_env = [] { const char* e = std::getenv("NVSE_RB_FUSE"); return !(e && e[0] == '0'); }();
  std::vector<int> per_launch((size_t)c.num_upsamples * c.num_kernels, 0);
  bool t32 = tc && fuse_env && c.resblock_type == 1;
  for (int i = 0; i < c.num_upsamples && t32; ++i) {
    t32 = g->layer("ups." + std::to_string(i)).w_bf16 != nullptr;
    for (int j = 0; j < c.num_kernels && t32; ++j) {
      const std::string p = "resblocks." + std::to_string(i * c.num_kernels + j);
      const int nd = c.num_dilations[j];
      const Layer& l0 = g->layer(p + ".convs1.0");
      for (int m = 0; m < nd && t32; ++m)
        t32 = g->layer(p + ".convs1." + std::to_string(m)).w_bf16 && g->layer(p + ".convs2." + std::to_string(m)).w_bf16;
      if (!t32) break;
      const int* dil = c.resblock_dilations[j];
      double whole = nd <= kRbMaxPairs ? rb_cost_per_row(l0.Cin, l0.k, dil, nd) : -1.0, single = 0.0;
      for (int m = 0; m < nd; ++m) {
        const double v = rb_cost_per_row(l0.Cin, l0.k, dil + m, 1);
        single = (v < 0.0 || single < 0.0) ? -1.0 : single + v;
      }
      if (whole < 0.0 && single < 0.0) t32 = false;
      // C <= 64: the whole block always (measured); above, the cost model decides, then the measured overrides below
      int per = (single < 0.0 || (whole >= 0.0 && (whole <= single || l0.Cin <= 64))) ? nd : 1;
      // measured (tools/rb_bench.py, C = 128, 32 x 55168 rows): the pipelined pair kernel beats the whole-block
      // launch from k = 5 up (k = 7: 3 x 0.86 ms vs 2.9 ms), the two-CTA whole-block launch wins at k = 3
      static const bool pairpipe_plan = [] { const char* e = std::getenv("NVSE_PAIRPIPE"); return !(e && e[0] == '0'); }();
      if (pairpipe_plan && single >= 0.0 && l0.k >= 5) {
        bool all = true;
        for (int m = 0; m < nd; ++m) all = all && pair_supported(l0.Cin, l0.k, dil[m]);
        if (all) per = 1;
      }
      per_launch[(size_t)i * c.num_kernels + j] = per;
    }
  }
  NVSE_REQUIRE(!frames_dev || t32, NVSE_ERR_UNSUPPORTED, "ragged batches are implemented on the fused tensor-core plan only");
  RowLens lens{frames_dev, 1, 0};  // valid rows per utterance at the current stage (affine in the frame count)
  int64_t T = F;
  for (int i = 0; i < c.num_upsamples; ++i) {
    const Layer& up = g->layer("ups." + std::to_string(i));
    if (int rc = run_conv_transpose(up, tc, bufA, B, T, bufU, slope, st, t32 && i > 0, t32, lens)) return rc;  // hifigan.py:111-112
    T = (T - 1) * up.stride - 2 * up.padding + up.k;
    lens.mul *= up.stride;
    lens.add = lens.add * up.stride + (up.k - up.stride - 2 * up.padding);
    const bool stage_concurrent = t32 && nbuf >= 12 && B * T <= kConcurrentRows;
    if (stage_concurrent) {
      if (int rc = ensure_side_streams(g)) return rc;
      NVSE_CUDA_CHECK(cudaEventRecord(g->ev_fork, st));
    }
    const float inv = 1.0f / (float)c.num_kernels;  // hifigan.py:119
    for (int j = 0; j < c.num_kernels; ++j) {
      const std::string p = "resblocks." + std::to_string(i * c.num_kernels + j);
      const int nd = c.num_dilations[j];
      const float* src = bufU;
      if (t32) {
        const int per = per_launch[(size_t)i * c.num_kernels + j];
        const Layer& l0 = g->layer(p + ".convs1.0");
        // concurrent mode: ResBlock j on its own stream, with its own temporaries and its own output buffer
        const bool conc = stage_concurrent;
        cudaStream_t sj = (conc && j > 0) ? g->side[j - 1] : st;
        float* tmpR = conc && j > 0 ? ws + (4 + 2 * j) * buf_elems : bufR;
        float* tmpT = conc && j > 0 ? ws + (5 + 2 * j) * buf_elems : bufT;
        float* outj = conc && j > 0 ? ws + (9 + j) * buf_elems : bufA;
        if (conc && j > 0) NVSE_CUDA_CHECK(cudaStreamWaitEvent(sj, g->ev_fork, 0));
        for (int m0 = 0; m0 < nd; m0 += per) {
          const bool last = (m0 + per == nd);
          float* dst = last ? outj : (src == tmpR ? tmpT : tmpR);
          ResblockTcArgs ra{};
          ra.x = src; ra.y = dst; ra.T = (int)T; ra.C = l0.Cin; ra.k = l0.k; ra.npairs = per;
          ra.t32 = 1; ra.bstride = t32_rows(T) * l0.Cin;
          ra.slope = slope; ra.out_scale = last ? inv : 1.0f; ra.accumulate = last && j > 0 && !conc;
          ra.h_fp16 = l0.Cin <= 32;  // the c1 -> c2 intermediate and w2 in IEEE half (see finalize_bf16)
          ra.lens = lens;
          for (int q = 0; q < per; ++q) {
            const Layer& c1 = g->layer(p + ".convs1." + std::to_string(m0 + q));
            const Layer& c2 = g->layer(p + ".convs2." + std::to_string(m0 + q));
            ra.pair[q] = RbPair{reinterpret_cast<const __nv_bfloat16*>(c1.w_bf16),
                                reinterpret_cast<const __nv_bfloat16*>(ra.h_fp16 ? c2.w_f16 : c2.w_bf16), c1.bias, c2.bias, c1.dilation};
          }
          static const bool pairpipe = [] { const char* e = std::getenv("NVSE_PAIRPIPE"); return !(e && e[0] == '0'); }();
          if (pairpipe && per == 1 && !ra.h_fp16 && pair_supported(ra.C, ra.k, ra.pair[0].dil)) {
            if (int rc = launch_pair_tc(ra, B, sj)) return rc;
          } else if (int rc = launch_resblock_tc(ra, B, sj)) {
            return rc;
          }
          src = dst;
        }
        if (conc && j > 0) NVSE_CUDA_CHECK(cudaEventRecord(g->ev_join[j - 1], sj));
        if (conc && j == c.num_kernels - 1) {  // join: bufA = (rb0 + rb1) + rb2, the order of the sequential path
          for (int q = 1; q < c.num_kernels; ++q) NVSE_CUDA_CHECK(cudaStreamWaitEvent(st, g->ev_join[q - 1], 0));
          if (int rc = launch_add3(bufA, ws + 10 * buf_elems, c.num_kernels > 2 ? ws + 11 * buf_elems : nullptr,
                                   B * t32_rows(T) * l0.Cin, st))
            return rc;
        }
        continue;
      }
      for (int m = 0; m < nd; ++m) {
        const bool last = (m == nd - 1);
        float* dst = last ? bufA : (c.resblock_type == 1 ? bufR : (src == bufR ? bufT : bufR));
        const float scale = last ? inv : 1.0f;
        const int accum = last && j > 0;
        if (c.resblock_type == 1) {  // hifigan.py:43-50
          const Layer& c1 = g->layer(p + ".convs1." + std::to_string(m));
          const Layer& c2 = g->layer(p + ".convs2." + std::to_string(m));
          const bool on_tc = tc && c1.w_bf16 && c2.w_bf16;
          const bool mid_bf16 = on_tc && !c2.tc_split;  // split layers keep the intermediate in fp32
          // c1: xt = conv(lrelu(x));  on the tensor-core path xt is stored already activated for c2
          const ConvIO io1{src, false, nullptr, bufT, mid_bf16, slope, on_tc ? slope : 1.0f, 1.0f, 0};
          if (int rc = run_conv(c1, tc, io1, B, T, st)) return rc;
          // c2: x' = conv(lrelu(xt)) + x, MRF scale/accumulate on the last pair
          const ConvIO io2{bufT, mid_bf16, src, dst, false, on_tc ? 1.0f : slope, 1.0f, scale, accum};
          if (int rc = run_conv(c2, tc, io2, B, T, st)) return rc;
        } else {  // hifigan.py:71-76
          const ConvIO io1{src, false, src, dst, false, slope, 1.0f, scale, accum};
          if (int rc = run_conv(g->layer(p + ".convs." + std::to_string(m)), tc, io1, B, T, st)) return rc;
        }
        src = dst;
      }
    }
  }
  const Layer& post = g->layer("conv_post");
  ConvF32Args a{};
  a.x = bufA; a.x_bstride = (t32 ? t32_rows(T) : T) * post.Cin; a.Cin = post.Cin; a.x_t32 = t32;
  a.w = post.w; a.bias = post.bias; a.Cout = post.Cout;
  conv1d_taps(post.k, 1, &a.taps);
  a.out_mul = 1; a.out_add = 0; a.in_slope = 0.01f; a.out_scale = 1.0f;  // F.leaky_relu default slope, hifigan.py:120
  a.in_lens = lens;
  if (c.kind == NVSE_GEN_HIFIGAN) {  // hifigan.py:120-124
    a.Tin = a.Tout = a.Trows = (int)T; a.y = out; a.y_bstride = T * post.Cout; a.out_act = 1;
    if (out_i16) {
      if (t32 && post.Cout == 1 && (post.Cin == 16 || post.Cin == 32 || post.Cin == 64)) {
        a.y_pcm16 = out_i16;  // PCM_16 straight out of the last kernel
        return launch_conv_f32(a, B, st);
      }
      a.y = bufU;
      if (int rc = launch_conv_f32(a, B, st)) return rc;
      return launch_pcm16(bufU, out_i16, B * T * post.Cout, st);
    }
    return launch_conv_f32(a, B, st);
  }
  // istftnet.py:311-318: lrelu(0.01) -> ReflectionPad1d((1,0)) -> conv_post -> exp / sin -> iSTFT
  static const bool post_tc_env = [] { const char* e = std::getenv("NVSE_POST_TC"); return !(e && e[0] == '0'); }();
  if (tc && t32 && post.w_post && post_tc_env && T >= 2) {
    // 16-bit path: conv_post as a tensor-core GEMM (IEEE-half operands like conv_pre; 26 % of an iSTFTNet forward on the fp32
    // CUDA cores otherwise).  Row v of the reflection-padded input is row v - 1 of x (v >= 1), so the layer is a plain
    // convolution with every tap offset shifted by -1 ... except for padded row 0 (= x row 1), which only the first
    // (k - 1) / 2 + 1 output rows read: post_reflect_fix_kernel adds that term.
    const int pitch = post.post_cout;
    ConvTcArgs ta{};
    ta.x = bufA; ta.x_bstride = t32_rows(T) * post.Cin; ta.Tin = (int)T; ta.Cin = post.Cin; ta.Cout = pitch; ta.x_t32 = 1;
    ta.wimg = reinterpret_cast<const __nv_bfloat16*>(post.w_post); ta.bias = post.bias_post; ta.ops_f16 = 1;
    ta.y = bufU; ta.y_bstride = (T + 1) * pitch; ta.y_ld = pitch; ta.Tout = (int)T + 1;
    conv1d_taps(post.k, 1, &ta.taps);
    for (int i = 0; i < ta.taps.ntaps; ++i) ta.taps.off[i] -= 1;
    ta.out_mul = 1; ta.out_add = 0; ta.Trows = (int)T + 1;
    ta.in_slope = 0.01f; ta.out_slope = 1.0f; ta.out_scale = 1.0f;
    ta.in_lens = lens;
    ta.ntile_hint = 1;  // HBM-bound (128 -> 32 channels): 0.70 -> 0.45 ms per 32 x 44161 rows with several 128-row CTAs per SM
    if (int rc = launch_conv_tc(ta, B, st)) return rc;
    if (int rc = launch_post_reflect_fix(bufA, post, bufU, pitch, B, T, st)) return rc;
    float* wav = out_i16 ? bufR : out;
    if (int rc = launch_istft_head(bufU, wav, B, T + 1, c.istft_n_fft, c.istft_hop, st, lens, pitch)) return rc;
    if (out_i16) return launch_pcm16(bufR, out_i16, B * nvse_generator_out_samples(g, F), st);
    return NVSE_OK;
  }
  a.reflect_left = 1;
  a.Tin = a.Tout = a.Trows = (int)T + 1; a.y = bufU; a.y_bstride = (T + 1) * post.Cout;
  if (int rc = launch_conv_f32(a, B, st)) return rc;
  if (out_i16) {  // the iSTFT head writes float; quantise from a workspace buffer
    if (int rc = launch_istft_head(bufU, bufR, B, T + 1, c.istft_n_fft, c.istft_hop, st, lens)) return rc;
    return launch_pcm16(bufR, out_i16, B * nvse_generator_out_samples(g, F), st);
  }
  return launch_istft_head(bufU, out, B, T + 1, c.istft_n_fft, c.istft_hop, st, lens);
}

static bool wants_tc(const Layer& L) { return L.name != "conv_pre" && L.name != "conv_post" && tc_supported(L.Cin, L.Cout); }
static bool wants_f16_copy(const Layer& L) { return !L.transposed && L.Cout <= 32 && L.name.find(".convs2.") != std::string::npos; }

// Buffers and precision flags of the tensor-core path (no launches).  Where bf16 rounding costs the most SNR:
//   * upsamplers: IEEE-half activations AND weights.  tests/bf16_budget.py: bf16 rounding of the ups weights is the
//     largest single error of the bf16 path (42.7 dB de-meaned SNR; hi+lo split activations with bf16 weights
//     46.1 dB; half operands 54.3 dB) -- and half costs one MMA per step where the split cost two;
//   * the <= 32-channel MRF stage (HBM-bound anyway): activations as hi + lo bf16 pairs on the per-layer path
//     (+5..6 dB at random init), an IEEE-half c1 -> c2 intermediate and w2 image in the fused kernels.
int finalize_plan(nvse_generator* g) {
  for (Layer& L : g->layers) {
    if (L.name == "conv_pre" && L.pre_n == 0) {
      // smallest supported K >= Cin, output channels in equal slices of <= 256
      int cin = 0, nsl = 1;
      for (int cand : {32, 64, 128}) if (!cin && L.Cin <= cand) cin = cand;
      while (L.Cout / nsl > 256 && nsl < 4) nsl *= 2;
      if (cin && L.Cout % nsl == 0 && tc_supported(cin, L.Cout / nsl) && L.k <= kMaxTaps) {
        L.pre_cin = cin; L.pre_cout = L.Cout / nsl;
        for (int sl = 0; sl < nsl; ++sl)
          NVSE_CUDA_CHECK(cudaMalloc(&L.w_pre[sl], sizeof(__nv_bfloat16) * tc_weight_image_elems(cin, L.pre_cout, L.k)));
        L.pre_n = nsl;
      }
    }
    if (L.name == "conv_post" && g->cfg.kind == NVSE_GEN_ISTFTNET && !L.w_post && L.Cout <= 32 && tc_supported(L.Cin, 32) &&
        L.k <= kMaxTaps && (L.k & 1)) {
      L.post_cout = 32;
      NVSE_CUDA_CHECK(cudaMalloc(&L.w_post, sizeof(__nv_bfloat16) * tc_weight_image_elems(L.Cin, L.post_cout, L.k)));
      NVSE_CUDA_CHECK(cudaMalloc(&L.bias_post, sizeof(float) * L.post_cout));
    }
    if (!wants_tc(L)) continue;
    const size_t bytes = sizeof(__nv_bfloat16) * tc_weight_image_elems(L.Cin, L.Cout, L.k);
    if (!L.w_bf16) NVSE_CUDA_CHECK(cudaMalloc(&L.w_bf16, bytes));
    if ((wants_f16_copy(L) || L.transposed) && !L.w_f16) NVSE_CUDA_CHECK(cudaMalloc(&L.w_f16, bytes));
    if (L.transposed) {
      L.tc_f16 = true;
      if (!L.w_ups && ups_tc_supported(L.Cin, L.Cout, L.k, L.stride, L.padding))
        NVSE_CUDA_CHECK(cudaMalloc(&L.w_ups, sizeof(__nv_bfloat16) * ups_tc_image_elems(L.Cin, L.Cout, L.k)));
    }
    else L.tc_split = L.Cout <= 32 && tc_split_fits(L.Cin, L.Cout, (L.k - 1) * L.dilation);
  }
  return NVSE_OK;
}

int build_extra_images(nvse_generator* g, cudaStream_t st) {
  for (Layer& L : g->layers)
    if (L.w_ups)
      if (int rc = launch_pack_weight_ups(L.w, L.w_ups, L.Cin, L.Cout, L.stride, st)) return rc;
  for (Layer& L : g->layers)
    for (int sl = 0; sl < L.pre_n; ++sl)
      if (int rc = launch_pack_weight_tc_slice(L.w, L.Cin, L.Cout, sl * L.pre_cout, reinterpret_cast<__nv_bfloat16*>(L.w_pre[sl]),
                                               L.pre_cin, L.pre_cout, L.k, st, true))
        return rc;
  for (Layer& L : g->layers)
    if (L.w_post) {
      if (int rc = launch_pack_weight_tc_slice(L.w, L.Cin, L.Cout, 0, reinterpret_cast<__nv_bfloat16*>(L.w_post), L.Cin, L.post_cout, L.k, st, true))
        return rc;
      NVSE_CUDA_CHECK(cudaMemsetAsync(L.bias_post, 0, sizeof(float) * L.post_cout, st));
      NVSE_CUDA_CHECK(cudaMemcpyAsync(L.bias_post, L.bias, sizeof(float) * L.Cout, cudaMemcpyDeviceToDevice, st));
    }
  return NVSE_OK;
}

int finalize_bf16(nvse_generator* g, cudaStream_t st) {
  if (int rc = finalize_plan(g)) return rc;
  if (int rc = build_extra_images(g, st)) return rc;
  for (Layer& L : g->layers) {
    if (!wants_tc(L)) continue;
    if (int rc = launch_pack_weight_tc(L.w, reinterpret_cast<__nv_bfloat16*>(L.w_bf16), L.Cin, L.Cout, L.k, st)) return rc;
    if (L.w_f16)
      if (int rc = launch_pack_weight_tc(L.w, reinterpret_cast<__nv_bfloat16*>(L.w_f16), L.Cin, L.Cout, L.k, st, true)) return rc;
  }
  return NVSE_OK;
}

}  // namespace nvse

using namespace nvse;

extern "C" int nvse_generator_create(const nvse_generator_config* cfg, nvse_generator** out) {
  NVSE_REQUIRE(cfg && out, NVSE_ERR_INVALID, "nvse_generator_create: null argument");
  NVSE_REQUIRE(cfg->kind == NVSE_GEN_HIFIGAN || cfg->kind == NVSE_GEN_ISTFTNET, NVSE_ERR_INVALID, "bad generator kind %d", cfg->kind);
  NVSE_REQUIRE(cfg->num_upsamples >= 1 && cfg->num_upsamples <= NVSE_MAX_UPS, NVSE_ERR_INVALID, "bad num_upsamples");
  NVSE_REQUIRE(cfg->num_kernels >= 1 && cfg->num_kernels <= NVSE_MAX_KERNELS, NVSE_ERR_INVALID, "bad num_kernels");
  NVSE_REQUIRE(cfg->resblock_type == 1 || cfg->resblock_type == 2, NVSE_ERR_INVALID, "resblock must be 1 or 2");
  NVSE_REQUIRE(cfg->in_channels > 0 && cfg->in_channels % 16 == 0, NVSE_ERR_UNSUPPORTED, "in_channels must be a multiple of 16");
  NVSE_REQUIRE(cfg->initial_channel > 0 && (cfg->initial_channel >> cfg->num_upsamples) >= 1 &&
                   (cfg->initial_channel % (1 << cfg->num_upsamples)) == 0,
               NVSE_ERR_INVALID, "upsample_initial_channel=%d is not divisible by 2^%d", cfg->initial_channel,
               cfg->num_upsamples);
  for (int i = 0; i < cfg->num_upsamples; ++i) {
    const int u = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
    NVSE_REQUIRE(u >= 1 && k >= u && (k - u) % 2 == 0 && (k + u - 1) / u <= kMaxTaps, NVSE_ERR_UNSUPPORTED,
                 "upsample stage %d: rate %d / kernel %d unsupported", i, u, k);
  }
  for (int j = 0; j < cfg->num_kernels; ++j) {
    const int k = cfg->resblock_kernel_sizes[j];
    NVSE_REQUIRE(k >= 1 && (k & 1) && k <= kMaxTaps, NVSE_ERR_UNSUPPORTED, "resblock kernel size %d unsupported (odd, <= %d)", k, kMaxTaps);
    NVSE_REQUIRE(cfg->num_dilations[j] >= 1 && cfg->num_dilations[j] <= NVSE_MAX_DILATIONS, NVSE_ERR_INVALID, "bad dilation count");
    for (int m = 0; m < cfg->num_dilations[j]; ++m)
      NVSE_REQUIRE(cfg->resblock_dilations[j][m] >= 1, NVSE_ERR_INVALID, "bad dilation");
  }
  if (cfg->kind == NVSE_GEN_ISTFTNET)
    NVSE_REQUIRE(cfg->istft_n_fft >= 4 && cfg->istft_hop >= 1 && cfg->istft_n_fft % cfg->istft_hop == 0, NVSE_ERR_UNSUPPORTED,
                 "istft n_fft=%d hop=%d unsupported", cfg->istft_n_fft, cfg->istft_hop);
  int ndev = 0;
  NVSE_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  NVSE_REQUIRE(ndev > 0, NVSE_ERR_CUDA, "nvse_generator_create: no CUDA device (there is no CPU fallback)");
  nvse_generator* g = new nvse_generator();
  g->cfg = *cfg;
  build_layers(g);
  *out = g;
  return NVSE_OK;
}

extern "C" int nvse_generator_destroy(nvse_generator* g) {
  if (!g) return NVSE_OK;
  if (g->ev_fork) cudaEventDestroy(g->ev_fork);
  for (int q = 0; q < 2; ++q) {
    if (g->ev_join[q]) cudaEventDestroy(g->ev_join[q]);
    if (g->side[q]) cudaStreamDestroy(g->side[q]);
  }
  destroy_weight_loader(g->loader);
  for (Layer& L : g->layers) {
    cudaFree(L.w);
    cudaFree(L.bias);
    cudaFree(L.w_bf16);
    cudaFree(L.w_f16);
    cudaFree(L.wT);
    cudaFree(L.wT_bf16);
    for (void* q : L.w_pre) cudaFree(q);
    cudaFree(L.w_post);
    cudaFree(L.bias_post);
    cudaFree(L.w_ups);
  }
  delete g;
  return NVSE_OK;
}

extern "C" int nvse_generator_set_weight(nvse_generator* g, const char* name, const float* data, const int64_t* shape,
                                         int ndim, void* stream) {
  NVSE_REQUIRE(g && name && data && shape, NVSE_ERR_INVALID, "nvse_generator_set_weight: null argument");
  const std::string full(name);
  const size_t dot = full.rfind('.');
  NVSE_REQUIRE(dot != std::string::npos, NVSE_ERR_INVALID, "unknown tensor name '%s'", name);
  const std::string prefix = full.substr(0, dot), leaf = full.substr(dot + 1);
  auto it = g->index.find(prefix);
  NVSE_REQUIRE(it != g->index.end(), NVSE_ERR_INVALID, "unknown tensor name '%s'", name);
  Layer& L = g->layers[it->second];
  cudaStream_t st = as_stream(stream);
  g->finalized = false;
  g->train_ready = false;
  if (leaf == "bias") {
    NVSE_REQUIRE(ndim == 1 && shape[0] == L.Cout, NVSE_ERR_INVALID, "%s: expected shape [%d]", name, L.Cout);
    if (!L.bias) NVSE_CUDA_CHECK(cudaMalloc(&L.bias, sizeof(float) * L.Cout));
    NVSE_CUDA_CHECK(cudaMemcpyAsync(L.bias, data, sizeof(float) * L.Cout, cudaMemcpyDeviceToDevice, st));
    L.have_bias = true;
    return NVSE_OK;
  }
  NVSE_REQUIRE(leaf == "weight", NVSE_ERR_INVALID,
               "%s: expected a folded '.weight' or '.bias' tensor (fold weight_g/weight_v first)", name);
  const int64_t d0 = L.transposed ? L.Cin : L.Cout, d1 = L.transposed ? L.Cout : L.Cin;
  NVSE_REQUIRE(ndim == 3 && shape[0] == d0 && shape[1] == d1 && shape[2] == L.k, NVSE_ERR_INVALID,
               "%s: expected shape [%lld, %lld, %d]", name, (long long)d0, (long long)d1, L.k);
  if (!L.w) NVSE_CUDA_CHECK(cudaMalloc(&L.w, sizeof(float) * (size_t)L.Cin * L.Cout * L.k));
  if (int rc = launch_repack_weight(data, L.w, L.Cin, L.Cout, L.k, L.transposed, st)) return rc;
  L.have_w = true;
  return NVSE_OK;
}

extern "C" int nvse_generator_finalize(nvse_generator* g, void* stream) {
  NVSE_REQUIRE(g, NVSE_ERR_INVALID, "nvse_generator_finalize: null handle");
  for (const Layer& L : g->layers)
    NVSE_REQUIRE(L.have_w && L.have_bias, NVSE_ERR_STATE, "layer '%s' is missing its %s", L.name.c_str(),
                 L.have_w ? "bias" : "weight");
  if (int rc = finalize_bf16(g, as_stream(stream))) return rc;
  g->finalized = true;
  return NVSE_OK;
}

extern "C" int64_t nvse_generator_out_samples(const nvse_generator* g, int64_t frames) {
  if (!g || frames < 0) return -1;
  int64_t T = frames;
  for (int i = 0; i < g->cfg.num_upsamples; ++i)
    T = (T - 1) * g->cfg.upsample_rates[i] - 2 * ((g->cfg.upsample_kernel_sizes[i] - g->cfg.upsample_rates[i]) / 2) +
        g->cfg.upsample_kernel_sizes[i];
  if (g->cfg.kind == NVSE_GEN_ISTFTNET) T = T * g->cfg.istft_hop;  // hop * ((T + 1) - 1)
  return T;
}

extern "C" size_t nvse_generator_workspace_bytes(const nvse_generator* g, int64_t B, int64_t frames, int precision) {
  if (!g || B < 0 || frames < 0) return 0;
  const size_t buf = align_up((size_t)B * (size_t)max_activation_elems(g, frames) * sizeof(float), 256);
  return (size_t)workspace_buffers(g, B, frames, precision) * buf + 256;
}

extern "C" int nvse_generator_forward(nvse_generator* g, const float* mel, int64_t B, int64_t frames, float* out,
                                      void* workspace, size_t workspace_bytes, int precision, void* stream) {
  NVSE_REQUIRE(g && mel && out, NVSE_ERR_INVALID, "nvse_generator_forward: null argument");
  NVSE_REQUIRE(g->finalized, NVSE_ERR_STATE, "nvse_generator_forward: call nvse_generator_finalize first");
  NVSE_REQUIRE(B >= 0 && frames >= 1, NVSE_ERR_INVALID, "nvse_generator_forward: bad B=%lld / frames=%lld", (long long)B, (long long)frames);
  NVSE_REQUIRE(precision == NVSE_PRECISION_F32 || precision == NVSE_PRECISION_BF16, NVSE_ERR_INVALID, "bad precision %d", precision);
  if (B == 0) return NVSE_OK;
  if (int rc = tc_abort_poll(as_stream(stream))) return rc;  // a timed-out kernel of an earlier call is an error, not garbage
  const size_t need = nvse_generator_workspace_bytes(g, B, frames, precision);
  NVSE_REQUIRE(workspace && workspace_bytes >= need, NVSE_ERR_INVALID, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
  NVSE_REQUIRE(max_activation_elems(g, frames) * 4 < (int64_t)1 << 40, NVSE_ERR_INVALID, "utterance too long");
  const int64_t buf_elems = (int64_t)(align_up((size_t)B * (size_t)max_activation_elems(g, frames) * sizeof(float), 256) / sizeof(float));
  float* ws = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  return forward_impl(g, precision == NVSE_PRECISION_BF16, mel, B, frames, out, ws, buf_elems,
                      workspace_buffers(g, B, frames, precision), as_stream(stream));
}

extern "C" int nvse_generator_forward_ragged(nvse_generator* g, const float* mel, int64_t B, int64_t frames, const int32_t* frames_dev,
                                             float* out, void* workspace, size_t workspace_bytes, int precision, void* stream) {
  NVSE_REQUIRE(g && mel && out && frames_dev, NVSE_ERR_INVALID, "nvse_generator_forward_ragged: null argument");
  NVSE_REQUIRE(g->finalized, NVSE_ERR_STATE, "nvse_generator_forward_ragged: call nvse_generator_finalize first");
  NVSE_REQUIRE(B >= 0 && frames >= 1, NVSE_ERR_INVALID, "nvse_generator_forward_ragged: bad B=%lld / frames=%lld", (long long)B, (long long)frames);
  NVSE_REQUIRE(precision == NVSE_PRECISION_BF16, NVSE_ERR_UNSUPPORTED, "nvse_generator_forward_ragged: the 16-bit tensor-core path only");
  if (B == 0) return NVSE_OK;
  if (int rc = tc_abort_poll(as_stream(stream))) return rc;
  const size_t need = nvse_generator_workspace_bytes(g, B, frames, precision);
  NVSE_REQUIRE(workspace && workspace_bytes >= need, NVSE_ERR_INVALID, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
  NVSE_REQUIRE(max_activation_elems(g, frames) * 4 < (int64_t)1 << 40, NVSE_ERR_INVALID, "utterance too long");
  const int64_t buf_elems = (int64_t)(align_up((size_t)B * (size_t)max_activation_elems(g, frames) * sizeof(float), 256) / sizeof(float));
  float* ws = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  return forward_impl(g, true, mel, B, frames, out, ws, buf_elems, workspace_buffers(g, B, frames, precision), as_stream(stream), nullptr,
                      frames_dev);
}

extern "C" int nvse_generator_forward_pcm16(nvse_generator* g, const float* mel, int64_t B, int64_t frames, int16_t* out,
                                            void* workspace, size_t workspace_bytes, int precision, void* stream) {
  NVSE_REQUIRE(g && mel && out, NVSE_ERR_INVALID, "nvse_generator_forward_pcm16: null argument");
  NVSE_REQUIRE(g->finalized, NVSE_ERR_STATE, "nvse_generator_forward_pcm16: call nvse_generator_finalize first");
  NVSE_REQUIRE(B >= 0 && frames >= 1, NVSE_ERR_INVALID, "nvse_generator_forward_pcm16: bad B=%lld / frames=%lld", (long long)B, (long long)frames);
  NVSE_REQUIRE(precision == NVSE_PRECISION_F32 || precision == NVSE_PRECISION_BF16, NVSE_ERR_INVALID, "bad precision %d", precision);
  if (B == 0) return NVSE_OK;
  if (int rc = tc_abort_poll(as_stream(stream))) return rc;
  const size_t need = nvse_generator_workspace_bytes(g, B, frames, precision);
  NVSE_REQUIRE(workspace && workspace_bytes >= need, NVSE_ERR_INVALID, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
  NVSE_REQUIRE(max_activation_elems(g, frames) * 4 < (int64_t)1 << 40, NVSE_ERR_INVALID, "utterance too long");
  const int64_t buf_elems = (int64_t)(align_up((size_t)B * (size_t)max_activation_elems(g, frames) * sizeof(float), 256) / sizeof(float));
  float* ws = reinterpret_cast<float*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  return forward_impl(g, precision == NVSE_PRECISION_BF16, mel, B, frames, nullptr, ws, buf_elems,
                      workspace_buffers(g, B, frames, precision), as_stream(stream), out);
}

// ---- fused wav -> wav call ---------------------------------------------------------------------------
extern "C" int nvse_vocoder_mel_pitch(const nvse_generator* g) {
  if (!g || !g->finalized) return 0;
  const Layer& pre = g->layer("conv_pre");
  return pre.pre_n > 0 ? pre.pre_cin : 0;
}

extern "C" size_t nvse_vocoder_workspace_bytes(const nvse_frontend* fe, const nvse_generator* g, int64_t B, int64_t T) {
  if (!fe || !g || B < 0 || T < 0) return 0;
  const int64_t frames = nvse_frontend_num_frames(fe, T);
  const int pitch = nvse_vocoder_mel_pitch(g);
  if (frames < 1 || pitch == 0) return 0;
  return nvse_generator_workspace_bytes(g, B, frames, NVSE_PRECISION_BF16) + align_up((size_t)B * frames * pitch * sizeof(float), 256) +
         align_up((size_t)B * sizeof(int), 256) + 256;
}

extern "C" int nvse_vocoder_forward(const nvse_frontend* fe, nvse_generator* g, const float* wav, int64_t B, int64_t T,
                                    int64_t wav_row_stride, const int32_t* samples_dev, float* out, int16_t* out_pcm16, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  NVSE_REQUIRE(fe && g && wav && ((out != nullptr) != (out_pcm16 != nullptr)), NVSE_ERR_INVALID,
               "nvse_vocoder_forward: null argument (exactly one of out / out_pcm16)");
  NVSE_REQUIRE(g->finalized, NVSE_ERR_STATE, "nvse_vocoder_forward: the generator has no weights yet");
  NVSE_REQUIRE(!(samples_dev && out_pcm16), NVSE_ERR_UNSUPPORTED, "nvse_vocoder_forward: ragged batches deliver float output");
  NVSE_REQUIRE(B >= 0 && T >= 1 && wav_row_stride >= T, NVSE_ERR_INVALID, "nvse_vocoder_forward: bad shape");
  if (B == 0) return NVSE_OK;
  cudaStream_t st = as_stream(stream);
  if (int rc = tc_abort_poll(st)) return rc;
  const int pitch = nvse_vocoder_mel_pitch(g);
  NVSE_REQUIRE(pitch > 0 && g->cfg.in_channels <= pitch, NVSE_ERR_UNSUPPORTED, "nvse_vocoder_forward: conv_pre is not on the tensor cores");
  const int64_t frames = nvse_frontend_num_frames(fe, T);
  const size_t need = nvse_vocoder_workspace_bytes(fe, g, B, T);
  NVSE_REQUIRE(workspace && need > 0 && workspace_bytes >= need, NVSE_ERR_INVALID, "nvse_vocoder_forward: workspace too small: %zu < %zu bytes",
               workspace_bytes, need);
  NVSE_REQUIRE(max_activation_elems(g, frames) * 4 < (int64_t)1 << 40, NVSE_ERR_INVALID, "utterance too long");
  const size_t gen_bytes = nvse_generator_workspace_bytes(g, B, frames, NVSE_PRECISION_BF16);
  const int64_t buf_elems = (int64_t)(align_up((size_t)B * (size_t)max_activation_elems(g, frames) * sizeof(float), 256) / sizeof(float));
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  float* ws = reinterpret_cast<float*>(base);
  float* mel_cl = reinterpret_cast<float*>(base + align_up(gen_bytes, 256));
  int* frames_dev = reinterpret_cast<int*>(reinterpret_cast<char*>(mel_cl) + align_up((size_t)B * frames * pitch * sizeof(float), 256));
  // front-end: the log-mel goes straight into the channels-last, zero-padded layout conv_pre's tensor-core launch stages from
  if (int rc = frontend_mel_cl(fe, wav, B, T, wav_row_stride, samples_dev, pitch, mel_cl, st)) return rc;
  if (samples_dev)
    if (int rc = launch_frames_from_samples(fe, samples_dev, frames_dev, B, st)) return rc;
  return forward_impl(g, true, nullptr, B, frames, out, ws, buf_elems, workspace_buffers(g, B, frames, NVSE_PRECISION_BF16), st, out_pcm16,
                      samples_dev ? frames_dev : nullptr, mel_cl);
}
