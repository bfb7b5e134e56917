// Training path of the generator handle (fp32, or tensor cores for the MRF convolutions): a forward that keeps the input of every convolution
// on a caller-owned tape, and the backward that walks it.  Reference: the generator step of
// train_time_wi_inv.py:222-236 (loss.backward() through HiFiGAN.forward, Models/hifigan.py:108-124;
// ResBlock1/2 hifigan.py:43-50,71-76).  Gradients are produced for the FOLDED weights (weight, bias of
// every layer, PyTorch layouts) in one flat buffer; weight_norm's own backward is nvse_weight_norm_backward_f32.
//
// Per pair of a ResBlock1 (x' = x + c2(lrelu(c1(lrelu(x))))), with g = dL/dx':
//   gh = lrelu'(h) * c2^T(g)             dW2 = corr(lrelu(h), g)     db2 = sum g
//   gx = lrelu'(x) * c1^T(gh) + g        dW1 = corr(lrelu(x), gh)    db1 = sum gh
// where c^T is the forward tap-list kernel with per-tap transposed weights and mirrored taps, the
// derivative mask and the "+ g" are its epilogue (conv_f32.cu), and corr is grad.cu's split reduction.
#include "generator.cuh"
#include "grad.cuh"
#include "conv_tc.cuh"

#include <cstdlib>
#include <cstring>
#include <utility>

namespace nvse {

namespace {

inline int64_t align64(int64_t v) { return (v + 63) / 64 * 64; }

struct TapePlan {
  int nstage = 0;
  std::vector<int64_t> T;                               // rows per utterance after upsampler i
  int64_t melT = 0, x_pre = 0;                          // [B, F, 80], [B, F, C0]
  std::vector<int64_t> xu, xs;                          // upsampler output / MRF output of stage i
  std::vector<std::vector<std::vector<int64_t>>> h;     // [i][j][m]: c1 output of pair m (ResBlock1)
  std::vector<std::vector<std::vector<int64_t>>> xin;   // [i][j][m]: input of pair m, m >= 1 (m = 0 reads xu)
  int64_t tmp[2] = {-1, -1};                            // outputs of ResBlocks 1, 2 of a stage while they run concurrently
  int64_t xpad = -1, z = -1;                            // iSTFTNet: reflect-padded MRF output [B, T+1, C], conv_post output [B, T+1, n_fft+2]
  int64_t total = 0;
  int64_t max_act = 0;  // floats of the largest activation of the whole batch
};

TapePlan make_tape(const nvse_generator* g, int64_t B, int64_t F) {
  const nvse_generator_config& c = g->cfg;
  TapePlan p;
  p.nstage = c.num_upsamples;
  int64_t at = 0;
  auto take = [&](int64_t n) { const int64_t o = at; at += align64(n); p.max_act = std::max(p.max_act, n); return o; };
  p.melT = take(B * F * c.in_channels);
  p.x_pre = take(B * F * c.initial_channel);
  int64_t T = F;
  p.h.resize(p.nstage); p.xin.resize(p.nstage);
  for (int i = 0; i < p.nstage; ++i) {
    const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i];
    T = (T - 1) * u - 2 * ((k - u) / 2) + k;
    p.T.push_back(T);
    const int64_t n = B * T * (c.initial_channel >> (i + 1));
    p.xu.push_back(take(n));
    p.h[i].resize(c.num_kernels); p.xin[i].resize(c.num_kernels);
    for (int j = 0; j < c.num_kernels; ++j)
      for (int m = 0; m < c.num_dilations[j]; ++m) {
        p.h[i][j].push_back(c.resblock_type == 1 ? take(n) : -1);
        p.xin[i][j].push_back(m > 0 ? take(n) : -1);
      }
    p.xs.push_back(take(n));
  }
  {
    int64_t big = 0;
    for (int i = 0; i < p.nstage; ++i) big = std::max(big, B * p.T[i] * (c.initial_channel >> (i + 1)));
    p.tmp[0] = take(big);
    p.tmp[1] = take(big);
  }
  if (c.kind == NVSE_GEN_ISTFTNET) {
    p.xpad = take(B * (T + 1) * (c.initial_channel >> p.nstage));
    p.z = take(B * (T + 1) * (c.istft_n_fft + 2));
    p.max_act = std::max(p.max_act, B * (T + 1) * 32);  // the 32-column padded gradient of z
  }
  p.max_act = std::max(p.max_act, B * T * (c.kind == NVSE_GEN_ISTFTNET ? c.istft_hop : 1));
  p.total = at;
  return p;
}

// fp32 Conv1d layer:  y = [accumulate ? y : 0] + out_scale * (conv(lrelu(x, in_slope)) + bias [+ residual])
int conv_fwd(const Layer& L, const float* x, const float* residual, float* y, int64_t B, int64_t T, float in_slope,
             float out_scale, int accumulate, int out_act, cudaStream_t st, bool tc = false) {
  if (tc && L.w_bf16 && !L.transposed && out_act == 0) {  // MRF convolutions on the tensor cores (bf16 operands, fp32 in / out)
    ConvTcArgs t{};
    t.x = x; t.x_bstride = T * L.Cin; t.Tin = (int)T; t.Cin = L.Cin; t.Cout = L.Cout;
    t.wimg = reinterpret_cast<const __nv_bfloat16*>(L.w_bf16); t.bias = L.bias; t.residual = residual;
    t.y = y; t.y_bstride = T * L.Cout; t.Tout = (int)T;
    conv1d_taps(L.k, L.dilation, &t.taps);
    t.out_mul = 1; t.Trows = (int)T;
    t.in_slope = in_slope; t.out_slope = 1.0f; t.out_scale = out_scale; t.accumulate = accumulate;
    t.split_act = L.tc_split;
    return launch_conv_tc(t, B, st);
  }
  ConvF32Args a{};
  a.x = x; a.x_bstride = T * L.Cin; a.Tin = (int)T; a.Cin = L.Cin;
  a.w = L.w; a.bias = L.bias; a.residual = residual;
  a.y = y; a.y_bstride = T * L.Cout; a.Tout = (int)T; a.Cout = L.Cout;
  conv1d_taps(L.k, L.dilation, &a.taps);
  a.out_mul = 1; a.Trows = (int)T;
  a.in_slope = in_slope; a.out_scale = out_scale; a.accumulate = accumulate; a.out_act = out_act;
  return launch_conv_f32(a, B, st);
}

int convT_fwd(const Layer& L, const float* x, float* y, int64_t B, int64_t Tin, float in_slope, cudaStream_t st) {
  const int64_t Tout = (Tin - 1) * L.stride - 2 * L.padding + L.k;
  for (int r = 0; r < L.stride && r < Tout; ++r) {
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

// data gradient of a Conv1d layer:  dx = [accumulate ? dx : 0] + out_scale * (lrelu'(x_in) * conv^T(dy) [+ dres])
int conv_dgrad(const Layer& L, const float* dy, const float* x_in, float mask_slope, const float* dres, float* dx, int64_t B,
               int64_t T, float out_scale, int accumulate, cudaStream_t st, bool tc = false) {
  if (tc && L.wT_bf16) {
    ConvTcArgs t{};
    t.x = dy; t.x_bstride = T * L.Cout; t.Tin = (int)T; t.Cin = L.Cout; t.Cout = L.Cin;
    t.wimg = reinterpret_cast<const __nv_bfloat16*>(L.wT_bf16); t.residual = dres;
    t.y = dx; t.y_bstride = T * L.Cin; t.Tout = (int)T;
    t.taps.ntaps = L.k;
    for (int j = 0; j < L.k; ++j) { t.taps.off[j] = L.padding - j * L.dilation; t.taps.widx[j] = j; }
    t.out_mul = 1; t.Trows = (int)T;
    t.in_slope = 1.0f; t.out_slope = 1.0f; t.out_scale = out_scale; t.accumulate = accumulate;
    if (x_in && mask_slope != 1.0f) { t.mask = x_in; t.mask_slope = mask_slope; }
    return launch_conv_tc(t, B, st);
  }
  ConvF32Args a{};
  a.x = dy; a.x_bstride = T * L.Cout; a.Tin = (int)T; a.Cin = L.Cout;
  a.w = L.wT; a.residual = dres;
  a.y = dx; a.y_bstride = T * L.Cin; a.Tout = (int)T; a.Cout = L.Cin;
  a.taps.ntaps = L.k;
  for (int j = 0; j < L.k; ++j) { a.taps.off[j] = L.padding - j * L.dilation; a.taps.widx[j] = j; }
  a.out_mul = 1; a.Trows = (int)T; a.in_slope = 1.0f; a.out_scale = out_scale; a.accumulate = accumulate;
  if (x_in && mask_slope != 1.0f) { a.mask = x_in; a.mask_slope = mask_slope; }
  return launch_conv_f32(a, B, st);
}

// data gradient of a ConvTranspose1d layer (a strided convolution over dy)
int convT_dgrad(const Layer& L, const float* dy, const float* x_in, float mask_slope, float* dx, int64_t B, int64_t Tin,
                cudaStream_t st) {
  const int64_t Tout = (Tin - 1) * L.stride - 2 * L.padding + L.k;
  ConvF32Args a{};
  a.x = dy; a.x_bstride = Tout * L.Cout; a.Tin = (int)Tout; a.Cin = L.Cout;
  a.w = L.wT;
  a.y = dx; a.y_bstride = Tin * L.Cin; a.Tout = (int)Tin; a.Cout = L.Cin;
  a.taps.ntaps = L.k;
  for (int j = 0; j < L.k; ++j) { a.taps.off[j] = j - L.padding; a.taps.widx[j] = j; }
  a.in_stride = L.stride;
  a.out_mul = 1; a.Trows = (int)Tin; a.in_slope = 1.0f; a.out_scale = 1.0f;
  if (x_in && mask_slope != 1.0f) { a.mask = x_in; a.mask_slope = mask_slope; }
  return launch_conv_f32(a, B, st);
}

struct GradSink {
  const nvse_generator* g;
  float* grads;
  float* scratch;
  int64_t B;
  cudaStream_t st;
  int tc;  // weight gradients of the MRF convolutions on the tensor cores (bf16 operands)
  float* dw(const Layer& L) const { return grads + L.grad_off; }
  float* db(const Layer& L) const { return grads + L.grad_off + (int64_t)L.Cin * L.Cout * L.k; }
  // Conv1d: dW = scale * corr(lrelu(x_in, in_slope), dy), db = scale * sum dy
  int conv(const Layer& L, const float* x_in, float in_slope, const float* dy, int64_t T, float scale) const {
    WgradArgs w{};
    w.U = x_in; w.u_bstride = T * L.Cin; w.Tu = (int)T; w.Ca = L.Cin; w.u_slope = in_slope;
    w.V = dy; w.v_bstride = T * L.Cout; w.Tv = (int)T; w.Cb = L.Cout; w.v_slope = 1.0f;
    w.u_stride = 1; w.ntaps = L.k;
    for (int j = 0; j < L.k; ++j) w.off[j] = j * L.dilation - L.padding;
    w.dst = dw(L); w.scale = scale; w.tc = tc;
    if (int rc = launch_wgrad(w, B, scratch, st)) return rc;
    return launch_colsum(dy, B * T, L.Cout, db(L), scale, scratch, st);
  }
  int convT(const Layer& L, const float* x_in, float in_slope, const float* dy, int64_t Tin) const {
    const int64_t Tout = (Tin - 1) * L.stride - 2 * L.padding + L.k;
    WgradArgs w{};
    w.U = dy; w.u_bstride = Tout * L.Cout; w.Tu = (int)Tout; w.Ca = L.Cout; w.u_slope = 1.0f;
    w.V = x_in; w.v_bstride = Tin * L.Cin; w.Tv = (int)Tin; w.Cb = L.Cin; w.v_slope = in_slope;
    w.u_stride = L.stride; w.ntaps = L.k;
    for (int j = 0; j < L.k; ++j) w.off[j] = j - L.padding;
    w.dst = dw(L); w.scale = 1.0f;
    if (int rc = launch_wgrad(w, B, scratch, st)) return rc;
    return launch_colsum(dy, B * Tout, L.Cout, db(L), 1.0f, scratch, st);
  }
};

size_t train_scratch_elems(const nvse_generator* g, int64_t B, int64_t F) {
  const TapePlan p = make_tape(g, B, F);
  size_t m = 64;
  for (const Layer& L : g->layers) {
    int64_t T = F;  // rows of the layer's V operand (its output rows for Conv1d, its input rows for ConvTranspose1d)
    if (L.name.rfind("ups.", 0) == 0) {
      const int i = std::atoi(L.name.c_str() + 4);
      T = i == 0 ? F : p.T[i - 1];
      m = std::max(m, wgrad_scratch_elems(L.Cout, L.Cin, L.k, B, (int)T));
      m = std::max(m, colsum_scratch_elems(L.Cout, B * p.T[i]));
      continue;
    }
    if (L.name.rfind("resblocks.", 0) == 0) T = p.T[std::atoi(L.name.c_str() + 10) / g->cfg.num_kernels];
    if (L.name == "conv_post") {
      T = p.T.back() + (g->cfg.kind == NVSE_GEN_ISTFTNET ? 1 : 0);
      m = std::max(m, wgrad_scratch_elems(L.Cin, 32, L.k, B, (int)T));  // iSTFTNet: gradient rows padded to 32 columns
      m = std::max(m, colsum_scratch_elems(32, B * T));
    }
    m = std::max(m, wgrad_scratch_elems(L.Cin, L.Cout, L.k, B, (int)T));
    m = std::max(m, colsum_scratch_elems(L.Cout, B * T));
  }
  return m;
}

int prepare_train(nvse_generator* g, cudaStream_t st) {
  if (g->train_ready) return NVSE_OK;
  for (Layer& L : g->layers) {
    if (!L.wT) NVSE_CUDA_CHECK(cudaMalloc(&L.wT, sizeof(float) * (size_t)L.Cin * L.Cout * L.k));
    if (int rc = launch_transpose_taps(L.w, L.wT, L.k, L.Cin, L.Cout, st)) return rc;
    if (L.w_bf16 && !L.transposed && tc_supported(L.Cout, L.Cin)) {  // the dgrad is a layer with Cin <-> Cout
      if (!L.wT_bf16) NVSE_CUDA_CHECK(cudaMalloc(&L.wT_bf16, sizeof(__nv_bfloat16) * tc_weight_image_elems(L.Cout, L.Cin, L.k)));
      if (int rc = launch_pack_weight_tc(L.wT, reinterpret_cast<__nv_bfloat16*>(L.wT_bf16), L.Cout, L.Cin, L.k, st)) return rc;
    }
  }
  g->train_ready = true;
  return NVSE_OK;
}

constexpr int kBwdBuffers = 14;  // dA, dz, then per concurrent ResBlock j: dU_j, gT_j, gR_j, gS_j
inline bool concurrent_ok(const nvse_generator_config& c) {
  static const bool on = [] { const char* e = std::getenv("NVSE_CONCURRENT"); return !(e && e[0] == '0'); }();
  return on && c.num_kernels >= 2 && c.num_kernels <= 3;
}

// (block j, pair m) steps of one MRF.  Sequential: block by block (the blocks accumulate into one buffer in order).
// Concurrent: pair by pair across the blocks, so that all three streams get work from the first enqueued launches on
// (the host enqueues ~5 us per launch: block-major order would start stream 2 a millisecond late).
inline std::vector<std::pair<int, int>> mrf_schedule(const nvse_generator_config& c, bool conc, bool backward) {
  std::vector<std::pair<int, int>> steps;
  int max_nd = 0;
  for (int j = 0; j < c.num_kernels; ++j) max_nd = std::max(max_nd, c.num_dilations[j]);
  if (!conc) {
    for (int j = 0; j < c.num_kernels; ++j)
      for (int q = 0; q < c.num_dilations[j]; ++q) steps.emplace_back(j, backward ? c.num_dilations[j] - 1 - q : q);
  } else {
    for (int q = 0; q < max_nd; ++q)
      for (int j = 0; j < c.num_kernels; ++j)
        if (q < c.num_dilations[j]) steps.emplace_back(j, backward ? c.num_dilations[j] - 1 - q : q);
  }
  return steps;
}

constexpr int64_t kBwdExtraElems = 1 << 18;  // iSTFTNet conv_post: 32-row padded wT and the padded weight gradient

}  // namespace

}  // namespace nvse

using namespace nvse;

extern "C" size_t nvse_generator_tape_bytes(const nvse_generator* g, int64_t B, int64_t frames) {
  if (!g || B < 0 || frames < 1) return 0;
  return (size_t)make_tape(g, B, frames).total * sizeof(float) + 256;
}

extern "C" size_t nvse_generator_backward_workspace_bytes(const nvse_generator* g, int64_t B, int64_t frames) {
  if (!g || B < 0 || frames < 1) return 0;
  const TapePlan p = make_tape(g, B, frames);
  return ((size_t)kBwdBuffers * (size_t)align64(p.max_act) + (size_t)kBwdExtraElems + 3 * (size_t)align64((int64_t)train_scratch_elems(g, B, frames))) * sizeof(float) + 256;
}

extern "C" int64_t nvse_generator_grad_elems(const nvse_generator* g) {
  if (!g) return -1;
  int64_t n = 0;
  for (const Layer& L : g->layers) n += (int64_t)L.Cin * L.Cout * L.k + L.Cout;
  return n;
}

extern "C" int nvse_generator_grad_offset(const nvse_generator* g, const char* name, int64_t* offset, int64_t* numel) {
  NVSE_REQUIRE(g && name && offset && numel, NVSE_ERR_INVALID, "nvse_generator_grad_offset: null argument");
  const std::string full(name);
  const size_t dot = full.rfind('.');
  NVSE_REQUIRE(dot != std::string::npos, NVSE_ERR_INVALID, "unknown tensor name '%s'", name);
  auto it = g->index.find(full.substr(0, dot));
  NVSE_REQUIRE(it != g->index.end(), NVSE_ERR_INVALID, "unknown tensor name '%s'", name);
  const Layer& L = g->layers[it->second];
  const std::string leaf = full.substr(dot + 1);
  const int64_t wn = (int64_t)L.Cin * L.Cout * L.k;
  if (leaf == "weight") { *offset = L.grad_off; *numel = wn; return NVSE_OK; }
  if (leaf == "bias") { *offset = L.grad_off + wn; *numel = L.Cout; return NVSE_OK; }
  return fail(NVSE_ERR_INVALID, "unknown tensor name '%s'", name);
}

extern "C" int nvse_generator_forward_train(nvse_generator* g, const float* mel, int64_t B, int64_t frames, float* out,
                                            void* tape, size_t tape_bytes, int precision, void* stream) {
  NVSE_REQUIRE(g && mel && out && tape, NVSE_ERR_INVALID, "nvse_generator_forward_train: null argument");
  NVSE_REQUIRE(g->finalized, NVSE_ERR_STATE, "nvse_generator_forward_train: call nvse_generator_finalize first");
  NVSE_REQUIRE(B >= 1 && frames >= 1, NVSE_ERR_INVALID, "nvse_generator_forward_train: bad B=%lld / frames=%lld", (long long)B, (long long)frames);
  NVSE_REQUIRE(tape_bytes >= nvse_generator_tape_bytes(g, B, frames), NVSE_ERR_INVALID, "tape too small");
  NVSE_REQUIRE(precision == NVSE_PRECISION_F32 || precision == NVSE_PRECISION_BF16, NVSE_ERR_INVALID, "bad precision %d", precision);
  const nvse_generator_config& c = g->cfg;
  const TapePlan p = make_tape(g, B, frames);
  float* tp = reinterpret_cast<float*>((reinterpret_cast<size_t>(tape) + 255) / 256 * 256);
  cudaStream_t st = as_stream(stream);
  if (int rc = tc_abort_poll(st)) return rc;
  const float slope = 0.1f;  // LRELU_SLOPE, hifigan.py:7
  const bool tc = precision == NVSE_PRECISION_BF16;

  if (int rc = launch_transpose(mel, tp + p.melT, B, c.in_channels, frames, st)) return rc;
  if (int rc = conv_fwd(g->layer("conv_pre"), tp + p.melT, nullptr, tp + p.x_pre, B, frames, 1.0f, 1.0f, 0, 0, st)) return rc;
  const float* prev = tp + p.x_pre;
  int64_t Tprev = frames;
  const float inv = 1.0f / (float)c.num_kernels;  // hifigan.py:119
  for (int i = 0; i < c.num_upsamples; ++i) {
    const Layer& up = g->layer("ups." + std::to_string(i));
    float* xu = tp + p.xu[i];
    float* xs = tp + p.xs[i];
    if (int rc = (tc ? run_conv_transpose(up, true, prev, B, Tprev, xu, slope, st) : convT_fwd(up, prev, xu, B, Tprev, slope, st))) return rc;  // hifigan.py:111-112
    const int64_t T = p.T[i];
    const bool conc = concurrent_ok(c);  // the ResBlocks of the MRF are independent: caller's stream + two side streams
    if (conc) {
      if (int rc = ensure_side_streams(g)) return rc;
      NVSE_CUDA_CHECK(cudaEventRecord(g->ev_fork, st));
    }
    const float* srcs[NVSE_MAX_KERNELS];
    for (const auto& jm : mrf_schedule(c, conc, false)) {  // concurrent: pair by pair across the blocks, so every stream has work early
      const int j = jm.first, m = jm.second;
      const std::string pj = "resblocks." + std::to_string(i * c.num_kernels + j);
      const int nd = c.num_dilations[j];
      cudaStream_t sj = (conc && j > 0) ? g->side[j - 1] : st;
      if (m == 0) {
        if (conc && j > 0) NVSE_CUDA_CHECK(cudaStreamWaitEvent(sj, g->ev_fork, 0));
        srcs[j] = xu;
      }
      float* outj = (conc && j > 0) ? tp + p.tmp[j - 1] : xs;
      const float* src = srcs[j];
      const bool last = (m == nd - 1);
      float* dst = last ? outj : tp + p.xin[i][j][m + 1];
      const float scale = last ? inv : 1.0f;
      const int accum = last && j > 0 && !conc;
      if (c.resblock_type == 1) {  // hifigan.py:43-50
        float* h = tp + p.h[i][j][m];
        if (int rc = conv_fwd(g->layer(pj + ".convs1." + std::to_string(m)), src, nullptr, h, B, T, slope, 1.0f, 0, 0, sj, tc)) return rc;
        if (int rc = conv_fwd(g->layer(pj + ".convs2." + std::to_string(m)), h, src, dst, B, T, slope, scale, accum, 0, sj, tc)) return rc;
      } else {  // hifigan.py:71-76
        if (int rc = conv_fwd(g->layer(pj + ".convs." + std::to_string(m)), src, src, dst, B, T, slope, scale, accum, 0, sj, tc)) return rc;
      }
      srcs[j] = dst;
      if (last && conc && j > 0) NVSE_CUDA_CHECK(cudaEventRecord(g->ev_join[j - 1], sj));
    }
    if (conc) {  // xs = (rb0 + rb1) + rb2: the summation order of the sequential path
      for (int q = 1; q < c.num_kernels; ++q) NVSE_CUDA_CHECK(cudaStreamWaitEvent(st, g->ev_join[q - 1], 0));
      if (int rc = launch_add3(xs, tp + p.tmp[0], c.num_kernels > 2 ? tp + p.tmp[1] : nullptr, B * T * (c.initial_channel >> (i + 1)), st)) return rc;
    }
    prev = xs;
    Tprev = T;
  }
  const Layer& post = g->layer("conv_post");
  if (c.kind == NVSE_GEN_HIFIGAN)  // hifigan.py:120-122: leaky_relu (default slope 0.01) -> conv_post -> tanh
    return conv_fwd(post, prev, nullptr, out, B, Tprev, 0.01f, 1.0f, 0, 1, st);
  // istftnet.py:311-318: leaky_relu (0.01) -> ReflectionPad1d((1, 0)) -> conv_post -> exp / sin -> iSTFT
  NVSE_REQUIRE(post.Cout <= 32, NVSE_ERR_UNSUPPORTED, "iSTFTNet training: gen_istft_n_fft + 2 = %d > 32", post.Cout);
  if (int rc = launch_pad_reflect_left(prev, tp + p.xpad, B, Tprev, post.Cin, st)) return rc;
  if (int rc = conv_fwd(post, tp + p.xpad, nullptr, tp + p.z, B, Tprev + 1, 0.01f, 1.0f, 0, 0, st)) return rc;
  return launch_istft_head(tp + p.z, out, B, Tprev + 1, c.istft_n_fft, c.istft_hop, st);
}

extern "C" int nvse_generator_backward(nvse_generator* g, int64_t B, int64_t frames, const float* out, const float* dout,
                                       const void* tape, size_t tape_bytes, float* grads, float* dmel, void* workspace,
                                       size_t workspace_bytes, int precision, void* stream) {
  NVSE_REQUIRE(g && out && dout && tape && grads && workspace, NVSE_ERR_INVALID, "nvse_generator_backward: null argument");
  NVSE_REQUIRE(g->finalized, NVSE_ERR_STATE, "nvse_generator_backward: call nvse_generator_finalize first");
  NVSE_REQUIRE(B >= 1 && frames >= 1, NVSE_ERR_INVALID, "nvse_generator_backward: bad B=%lld / frames=%lld", (long long)B, (long long)frames);
  NVSE_REQUIRE(tape_bytes >= nvse_generator_tape_bytes(g, B, frames), NVSE_ERR_INVALID, "tape too small");
  NVSE_REQUIRE(workspace_bytes >= nvse_generator_backward_workspace_bytes(g, B, frames), NVSE_ERR_INVALID, "workspace too small");
  NVSE_REQUIRE(precision == NVSE_PRECISION_F32 || precision == NVSE_PRECISION_BF16, NVSE_ERR_INVALID, "bad precision %d", precision);
  const nvse_generator_config& c = g->cfg;
  const TapePlan p = make_tape(g, B, frames);
  const float* tp = reinterpret_cast<const float*>((reinterpret_cast<size_t>(tape) + 255) / 256 * 256);
  float* ws = reinterpret_cast<float*>((reinterpret_cast<size_t>(workspace) + 255) / 256 * 256);
  cudaStream_t st = as_stream(stream);
  if (int rc = tc_abort_poll(st)) return rc;
  if (int rc = prepare_train(g, st)) return rc;
  const int64_t be = align64(p.max_act);
  float* dA = ws;            // gradient w.r.t. the MRF output of the current stage (then w.r.t. the previous stage's)
  float* dz = ws + be;       // gradient w.r.t. the conv_post output
  // per concurrent ResBlock j: dU_j (gradient w.r.t. the upsampler output through block j), gT (w.r.t. the c1 output of the
  // current pair), gR / gS (running gradient of the residual stream, ping / pong)
  auto jbuf = [&](int j, int which) { return ws + (2 + 4 * j + which) * be; };
  float* dU = jbuf(0, 0);
  float* gT = jbuf(0, 1);
  const bool tc = precision == NVSE_PRECISION_BF16;
  float* scratch0 = ws + kBwdBuffers * be + kBwdExtraElems;
  const int64_t scratch_stride = align64((int64_t)train_scratch_elems(g, B, frames));
  const GradSink sink{g, grads, scratch0, B, st, tc ? 1 : 0};
  const bool conc = concurrent_ok(c);
  if (conc)
    if (int rc = ensure_side_streams(g)) return rc;
  const float slope = 0.1f;
  const float inv = 1.0f / (float)c.num_kernels;

  const int64_t Tl = p.T.back();
  const Layer& post = g->layer("conv_post");
  const float* xs_last = tp + p.xs.back();
  if (c.kind == NVSE_GEN_HIFIGAN) {  // tanh, conv_post (hifigan.py:120-122)
    if (int rc = launch_tanh_bwd(out, dout, dz, B * Tl * post.Cout, st)) return rc;
    if (int rc = sink.conv(post, xs_last, 0.01f, dz, Tl, 1.0f)) return rc;
    if (int rc = conv_dgrad(post, dz, xs_last, 0.01f, nullptr, dA, B, Tl, 1.0f, 0, st)) return rc;
  } else {
    // iSTFT head, conv_post over the reflect-padded rows (istftnet.py:311-318).  The gradient of z is kept with its
    // n_fft + 2 columns padded to 32 (zeros): the data gradient then runs on the wide kernel (its K must be a multiple
    // of 16) against a 32-row zero-padded wT, and the weight gradient comes out as the first Cout rows of a 32-row one.
    NVSE_REQUIRE(post.Cout <= 32 && (int64_t)post.k * 32 * post.Cin + 32 <= kBwdExtraElems / 2, NVSE_ERR_UNSUPPORTED,
                 "iSTFTNet training: conv_post %d -> %d unsupported", post.Cin, post.Cout);
    const int64_t Tp = Tl + 1;
    const float* xpad = tp + p.xpad;
    float* extra = ws + kBwdBuffers * be;          // [k][32][Cin] padded wT, then [32][Cin][k] + [32] padded gradients
    float* wT_pad = extra;
    float* dw_pad = extra + kBwdExtraElems / 2;
    float* db_pad = dw_pad + (int64_t)32 * post.Cin * post.k;
    float* scratch = ws + kBwdBuffers * be + kBwdExtraElems;
    if (int rc = launch_istft_head_bwd(tp + p.z, dout, dz, B, Tp, c.istft_n_fft, c.istft_hop, 32, st)) return rc;
    NVSE_CUDA_CHECK(cudaMemsetAsync(wT_pad, 0, sizeof(float) * post.k * 32 * post.Cin, st));
    NVSE_CUDA_CHECK(cudaMemcpy2DAsync(wT_pad, sizeof(float) * 32 * post.Cin, post.wT, sizeof(float) * post.Cout * post.Cin,
                                      sizeof(float) * post.Cout * post.Cin, post.k, cudaMemcpyDeviceToDevice, st));
    {
      WgradArgs w{};
      w.U = xpad; w.u_bstride = Tp * post.Cin; w.Tu = (int)Tp; w.Ca = post.Cin; w.u_slope = 0.01f;
      w.V = dz; w.v_bstride = Tp * 32; w.Tv = (int)Tp; w.Cb = 32; w.v_slope = 1.0f;
      w.u_stride = 1; w.ntaps = post.k;
      for (int j = 0; j < post.k; ++j) w.off[j] = j - post.padding;
      w.dst = dw_pad; w.scale = 1.0f; w.tc = 0;
      if (int rc = launch_wgrad(w, B, scratch, st)) return rc;
      if (int rc = launch_colsum(dz, B * Tp, 32, db_pad, 1.0f, scratch, st)) return rc;
      NVSE_CUDA_CHECK(cudaMemcpyAsync(sink.dw(post), dw_pad, sizeof(float) * post.Cout * post.Cin * post.k, cudaMemcpyDeviceToDevice, st));
      NVSE_CUDA_CHECK(cudaMemcpyAsync(sink.db(post), db_pad, sizeof(float) * post.Cout, cudaMemcpyDeviceToDevice, st));
    }
    {
      ConvF32Args a{};
      a.x = dz; a.x_bstride = Tp * 32; a.Tin = (int)Tp; a.Cin = 32;
      a.w = wT_pad;
      a.y = gT; a.y_bstride = Tp * post.Cin; a.Tout = (int)Tp; a.Cout = post.Cin;
      a.taps.ntaps = post.k;
      for (int j = 0; j < post.k; ++j) { a.taps.off[j] = post.padding - j; a.taps.widx[j] = j; }
      a.out_mul = 1; a.Trows = (int)Tp; a.in_slope = 1.0f; a.out_scale = 1.0f;
      a.mask = xpad; a.mask_slope = 0.01f;
      if (int rc = launch_conv_f32(a, B, st)) return rc;
    }
    if (int rc = launch_unpad_reflect_left(gT, dA, B, Tl, post.Cin, st)) return rc;
  }

  for (int i = c.num_upsamples - 1; i >= 0; --i) {
    const int64_t T = p.T[i];
    const float* xu = tp + p.xu[i];
    if (conc) NVSE_CUDA_CHECK(cudaEventRecord(g->ev_fork, st));
    const float* gcurs[NVSE_MAX_KERNELS];
    for (const auto& jm : mrf_schedule(c, conc, true)) {
      const int j = jm.first, m = jm.second;
      const std::string pj = "resblocks." + std::to_string(i * c.num_kernels + j);
      const int nd = c.num_dilations[j];
      const int jb = conc ? j : 0;
      cudaStream_t sj = (conc && j > 0) ? g->side[j - 1] : st;
      if (m == nd - 1) {
        if (conc && j > 0) NVSE_CUDA_CHECK(cudaStreamWaitEvent(sj, g->ev_fork, 0));
        gcurs[j] = dA;  // unscaled: the 1/num_kernels of the MRF average is applied where gradients leave the block
      }
      const GradSink sj_sink{g, grads, scratch0 + jb * scratch_stride, B, sj, tc ? 1 : 0};
      float* dUj = jbuf(jb, 0);
      float* gTj = jbuf(jb, 1);
      float* gRj = jbuf(jb, 2);
      float* gSj = jbuf(jb, 3);
      const float* gcur = gcurs[j];
      const float* x_m = m == 0 ? xu : tp + p.xin[i][j][m];
      float* gnext = m == 0 ? dUj : (gcur == gRj ? gSj : gRj);
      const float oscale = m == 0 ? inv : 1.0f;
      const int accum = m == 0 && j > 0 && !conc;
      if (c.resblock_type == 1) {
        const Layer& c1 = g->layer(pj + ".convs1." + std::to_string(m));
        const Layer& c2 = g->layer(pj + ".convs2." + std::to_string(m));
        const float* h = tp + p.h[i][j][m];
        if (int rc = conv_dgrad(c2, gcur, h, slope, nullptr, gTj, B, T, 1.0f, 0, sj, tc)) return rc;
        if (int rc = sj_sink.conv(c2, h, slope, gcur, T, inv)) return rc;
        if (int rc = conv_dgrad(c1, gTj, x_m, slope, gcur, gnext, B, T, oscale, accum, sj, tc)) return rc;
        if (int rc = sj_sink.conv(c1, x_m, slope, gTj, T, inv)) return rc;
      } else {
        const Layer& cv = g->layer(pj + ".convs." + std::to_string(m));
        if (int rc = conv_dgrad(cv, gcur, x_m, slope, gcur, gnext, B, T, oscale, accum, sj, tc)) return rc;
        if (int rc = sj_sink.conv(cv, x_m, slope, gcur, T, inv)) return rc;
      }
      gcurs[j] = gnext;
      if (m == 0 && conc && j > 0) NVSE_CUDA_CHECK(cudaEventRecord(g->ev_join[j - 1], sj));
    }
    if (conc) {  // dU = (dU_0 + dU_1) + dU_2, the order of the sequential path
      for (int q = 1; q < c.num_kernels; ++q) NVSE_CUDA_CHECK(cudaStreamWaitEvent(st, g->ev_join[q - 1], 0));
      if (int rc = launch_add3(dU, jbuf(1, 0), c.num_kernels > 2 ? jbuf(2, 0) : nullptr, B * T * (c.initial_channel >> (i + 1)), st)) return rc;
    }
    // upsampler i (hifigan.py:111-112): input = lrelu(previous stage output, 0.1)
    const Layer& up = g->layer("ups." + std::to_string(i));
    const float* prev = i == 0 ? tp + p.x_pre : tp + p.xs[i - 1];
    const int64_t Tprev = i == 0 ? frames : p.T[i - 1];
    if (conc) {  // the upsampler's weight gradient (side stream) overlaps its data gradient: both only read dU
      const GradSink side_sink{g, grads, scratch0 + scratch_stride, B, g->side[0], tc ? 1 : 0};
      NVSE_CUDA_CHECK(cudaEventRecord(g->ev_fork, st));
      NVSE_CUDA_CHECK(cudaStreamWaitEvent(g->side[0], g->ev_fork, 0));
      if (int rc = side_sink.convT(up, prev, slope, dU, Tprev)) return rc;
      NVSE_CUDA_CHECK(cudaEventRecord(g->ev_join[0], g->side[0]));
      if (int rc = convT_dgrad(up, dU, prev, slope, dA, B, Tprev, st)) return rc;
      NVSE_CUDA_CHECK(cudaStreamWaitEvent(st, g->ev_join[0], 0));  // dU is rewritten by the next stage
    } else {
      if (int rc = sink.convT(up, prev, slope, dU, Tprev)) return rc;
      if (int rc = convT_dgrad(up, dU, prev, slope, dA, B, Tprev, st)) return rc;
    }
  }
  // conv_pre (hifigan.py:109): no activation in front
  const Layer& pre = g->layer("conv_pre");
  if (int rc = sink.conv(pre, tp + p.melT, 1.0f, dA, frames, 1.0f)) return rc;
  if (dmel) {
    if (int rc = conv_dgrad(pre, dA, nullptr, 1.0f, nullptr, gT, B, frames, 1.0f, 0, st)) return rc;
    if (int rc = launch_transpose(gT, dmel, B, frames, c.in_channels, st)) return rc;
  }
  return NVSE_OK;
}
