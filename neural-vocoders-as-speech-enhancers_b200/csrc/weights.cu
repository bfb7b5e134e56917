// Batched weight load: every layer of a generator folded (weight_norm), repacked and converted in TWO launches.
//
// A training step changes every parameter, so the handle's copies -- fp32 [k][Cin][Cout], the per-tap transposed
// copy of the backward, and the bf16 / IEEE-half tensor-core images of both -- are rebuilt once per step.  Layer
// by layer that is ~7 small launches x 78 layers (fold, repack, bias copy, 2-3 image packs, transpose, image of
// the transpose); here one kernel computes all row scales g / ||v|| and one kernel writes every derived copy of
// every layer from the source tensors directly.  The arithmetic (summation order of the norm, the single multiply
// v * scale, round-to-nearest conversions) is that of the per-layer kernels, so both paths give identical bits.
#include "generator.cuh"
#include "conv_tc.cuh"

namespace nvse {

struct LayerDev {
  const float* v;         // weight_v, or the folded weight when g is null (PyTorch layout)
  const float* g;         // weight_g [rows] or null
  const float* bias_src;
  float* w;               // [k][Cin][Cout]
  float* wT;              // [k][Cout][Cin] or null
  float* bias;
  __nv_bfloat16* img;     // bf16 tensor-core image of w, or null
  __nv_bfloat16* img16;   // IEEE-half image of w, or null
  __nv_bfloat16* imgT;    // bf16 image of wT (a layer with Cin <-> Cout), or null
  float* scale;           // [rows]
  int Cin, Cout, k, transposed, kc, kcT, rows;
  int64_t cols;
};

struct WnBwdDev {
  const float* v; const float* g; const float* dw;
  float* dv; float* dg;
  int64_t cols;
};

namespace {

constexpr int kElemsPerBlock = 1024;

__device__ __forceinline__ float wl_block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// weight_norm backward of every layer in one launch: one CTA per row (same arithmetic as grad.cu's per-layer kernel)
__global__ void __launch_bounds__(256) wn_bwd_all_kernel(const WnBwdDev* __restrict__ layers, const int* __restrict__ row_layer,
                                                         const int* __restrict__ row_idx) {
  __shared__ float red[8];
  const WnBwdDev L = layers[row_layer[blockIdx.x]];
  if (!L.g) return;
  const int64_t r = row_idx[blockIdx.x];
  const float* vr = L.v + r * L.cols;
  const float* wr = L.dw + r * L.cols;
  float ss = 0.0f, dot = 0.0f;
  for (int64_t c = threadIdx.x; c < L.cols; c += blockDim.x) {
    const float x = vr[c];
    ss = fmaf(x, x, ss);
    dot = fmaf(wr[c], x, dot);
  }
  ss = wl_block_sum(ss, red);
  dot = wl_block_sum(dot, red);
  const float nrm = sqrtf(ss), inv = 1.0f / nrm, gr = L.g[r];
  if (threadIdx.x == 0) L.dg[r] = dot * inv;
  const float s1 = gr * inv, s2 = dot * inv * inv;
  for (int64_t c = threadIdx.x; c < L.cols; c += blockDim.x) L.dv[r * L.cols + c] = s1 * (wr[c] - vr[c] * s2);
}


__global__ void __launch_bounds__(256) load_scales_kernel(const LayerDev* __restrict__ layers, const int* __restrict__ row_layer,
                                                          const int* __restrict__ row_idx) {
  __shared__ float red[8];
  const LayerDev& L = layers[row_layer[blockIdx.x]];
  const int r = row_idx[blockIdx.x];
  if (!L.g) {
    if (threadIdx.x == 0) L.scale[r] = 1.0f;
    return;
  }
  const float* vr = L.v + (int64_t)r * L.cols;
  float acc = 0.0f;
  for (int64_t c = threadIdx.x; c < L.cols; c += blockDim.x) acc = fmaf(vr[c], vr[c], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int i = 0; i < 8; ++i) s += red[i];
    L.scale[r] = L.g[r] / sqrtf(s);
  }
}

__device__ __forceinline__ int64_t tc_image_index(int j, int ci, int co, int Cin, int Cout, int kc) {
  return ((((int64_t)j * (Cin / kc) + ci / kc) * (kc / 8) + (ci % kc) / 8) * Cout + co) * 8 + (ci % 8);
}

__global__ void __launch_bounds__(256) load_weights_kernel(const LayerDev* __restrict__ layers, const int* __restrict__ blk_layer,
                                                           const int* __restrict__ blk_first) {
  const LayerDev L = layers[blk_layer[blockIdx.x]];
  const int64_t n = (int64_t)L.Cin * L.Cout * L.k;
  const int64_t e0 = (int64_t)blk_first[blockIdx.x] * kElemsPerBlock;
#pragma unroll
  for (int u = 0; u < kElemsPerBlock / 256; ++u) {
    const int64_t e = e0 + u * 256 + threadIdx.x;
    if (e >= n) break;
    const int co = (int)(e % L.Cout), ci = (int)((e / L.Cout) % L.Cin), j = (int)(e / ((int64_t)L.Cout * L.Cin));
    const int64_t src = L.transposed ? ((int64_t)ci * L.Cout + co) * L.k + j : ((int64_t)co * L.Cin + ci) * L.k + j;
    const int row = L.transposed ? ci : co;
    float val = L.v[src];
    if (L.g) val *= L.scale[row];
    L.w[e] = val;
    if (L.wT) L.wT[((int64_t)j * L.Cout + co) * L.Cin + ci] = val;
    if (L.img) L.img[tc_image_index(j, ci, co, L.Cin, L.Cout, L.kc)] = __float2bfloat16_rn(val);
    if (L.img16) reinterpret_cast<__half*>(L.img16)[tc_image_index(j, ci, co, L.Cin, L.Cout, L.kc)] = __float2half_rn(val);
    if (L.imgT) L.imgT[tc_image_index(j, co, ci, L.Cout, L.Cin, L.kcT)] = __float2bfloat16_rn(val);
    if (e < L.Cout) L.bias[e] = L.bias_src[e];
  }
}

}  // namespace

struct WeightLoader {
  std::vector<LayerDev> host;
  LayerDev* d_layers = nullptr;
  int* d_row_layer = nullptr;
  int* d_row_idx = nullptr;
  int* d_blk_layer = nullptr;
  int* d_blk_first = nullptr;
  float* d_scale = nullptr;
  std::vector<WnBwdDev> wn_host;
  WnBwdDev* d_wn = nullptr;
  int total_rows = 0, total_blocks = 0;
  bool with_train = false;
  ~WeightLoader() {
    cudaFree(d_layers); cudaFree(d_row_layer); cudaFree(d_row_idx); cudaFree(d_blk_layer); cudaFree(d_blk_first); cudaFree(d_scale); cudaFree(d_wn);
  }
};

void destroy_weight_loader(WeightLoader* p) { delete p; }

// allocate every per-layer buffer the forward (and, with_train, the backward) reads; build the static tables
static int prepare_loader(nvse_generator* g, bool with_train) {
  if (int rc = finalize_plan(g)) return rc;  // tensor-core image buffers + precision flags
  for (Layer& L : g->layers) {
    const size_t n = (size_t)L.Cin * L.Cout * L.k;
    if (!L.w) NVSE_CUDA_CHECK(cudaMalloc(&L.w, sizeof(float) * n));
    if (!L.bias) NVSE_CUDA_CHECK(cudaMalloc(&L.bias, sizeof(float) * L.Cout));
    if (with_train) {
      if (!L.wT) NVSE_CUDA_CHECK(cudaMalloc(&L.wT, sizeof(float) * n));
      if (L.w_bf16 && !L.transposed && tc_supported(L.Cout, L.Cin) && !L.wT_bf16)
        NVSE_CUDA_CHECK(cudaMalloc(&L.wT_bf16, sizeof(__nv_bfloat16) * tc_weight_image_elems(L.Cout, L.Cin, L.k)));
    }
  }
  if (g->loader && g->loader->with_train >= with_train) return NVSE_OK;
  delete g->loader;
  WeightLoader* ld = g->loader = new WeightLoader();
  ld->with_train = with_train;
  std::vector<int> row_layer, row_idx, blk_layer, blk_first;
  for (size_t i = 0; i < g->layers.size(); ++i) {
    const Layer& L = g->layers[i];
    const int rows = L.transposed ? L.Cin : L.Cout;
    const int64_t n = (int64_t)L.Cin * L.Cout * L.k;
    for (int r = 0; r < rows; ++r) { row_layer.push_back((int)i); row_idx.push_back(r); }
    for (int64_t b = 0; b * kElemsPerBlock < n; ++b) { blk_layer.push_back((int)i); blk_first.push_back((int)b); }
  }
  ld->total_rows = (int)row_layer.size();
  ld->total_blocks = (int)blk_layer.size();
  NVSE_CUDA_CHECK(cudaMalloc(&ld->d_layers, sizeof(LayerDev) * g->layers.size()));
  NVSE_CUDA_CHECK(cudaMalloc(&ld->d_row_layer, sizeof(int) * row_layer.size()));
  NVSE_CUDA_CHECK(cudaMalloc(&ld->d_row_idx, sizeof(int) * row_idx.size()));
  NVSE_CUDA_CHECK(cudaMalloc(&ld->d_blk_layer, sizeof(int) * blk_layer.size()));
  NVSE_CUDA_CHECK(cudaMalloc(&ld->d_blk_first, sizeof(int) * blk_first.size()));
  NVSE_CUDA_CHECK(cudaMalloc(&ld->d_scale, sizeof(float) * row_layer.size()));
  NVSE_CUDA_CHECK(cudaMemcpy(ld->d_row_layer, row_layer.data(), sizeof(int) * row_layer.size(), cudaMemcpyHostToDevice));
  NVSE_CUDA_CHECK(cudaMemcpy(ld->d_row_idx, row_idx.data(), sizeof(int) * row_idx.size(), cudaMemcpyHostToDevice));
  NVSE_CUDA_CHECK(cudaMemcpy(ld->d_blk_layer, blk_layer.data(), sizeof(int) * blk_layer.size(), cudaMemcpyHostToDevice));
  NVSE_CUDA_CHECK(cudaMemcpy(ld->d_blk_first, blk_first.data(), sizeof(int) * blk_first.size(), cudaMemcpyHostToDevice));
  NVSE_CUDA_CHECK(cudaMalloc(&ld->d_wn, sizeof(WnBwdDev) * g->layers.size()));
  ld->host.resize(g->layers.size());
  ld->wn_host.resize(g->layers.size());
  return NVSE_OK;
}

}  // namespace nvse

using namespace nvse;

extern "C" int nvse_generator_load_weights(nvse_generator* g, const float* const* weight, const float* const* weight_g,
                                           const float* const* bias, int n_layers, int with_train, void* stream) {
  NVSE_REQUIRE(g && weight && weight_g && bias, NVSE_ERR_INVALID, "nvse_generator_load_weights: null argument");
  NVSE_REQUIRE(n_layers == (int)g->layers.size(), NVSE_ERR_INVALID, "nvse_generator_load_weights: expected %d layers, got %d",
               (int)g->layers.size(), n_layers);
  for (int i = 0; i < n_layers; ++i)
    NVSE_REQUIRE(weight[i] && bias[i], NVSE_ERR_INVALID, "nvse_generator_load_weights: layer '%s' is missing its weight or bias",
                 g->layers[i].name.c_str());
  cudaStream_t st = as_stream(stream);
  g->finalized = false;
  g->train_ready = false;
  if (int rc = prepare_loader(g, with_train != 0)) return rc;
  WeightLoader* ld = g->loader;
  int row_off = 0;
  for (int i = 0; i < n_layers; ++i) {
    Layer& L = g->layers[i];
    LayerDev& d = ld->host[i];
    d.v = weight[i]; d.g = weight_g[i]; d.bias_src = bias[i];
    d.w = L.w; d.bias = L.bias;
    d.wT = with_train ? L.wT : nullptr;
    d.img = reinterpret_cast<__nv_bfloat16*>(L.w_bf16);
    d.img16 = reinterpret_cast<__nv_bfloat16*>(L.w_f16);
    d.imgT = with_train ? reinterpret_cast<__nv_bfloat16*>(L.wT_bf16) : nullptr;
    d.Cin = L.Cin; d.Cout = L.Cout; d.k = L.k; d.transposed = L.transposed ? 1 : 0;
    d.kc = tc_kchunk(L.Cin); d.kcT = tc_kchunk(L.Cout);
    d.rows = L.transposed ? L.Cin : L.Cout;
    d.cols = (int64_t)L.Cin * L.Cout * L.k / d.rows;
    d.scale = ld->d_scale + row_off;
    row_off += d.rows;
  }
  NVSE_CUDA_CHECK(cudaMemcpyAsync(ld->d_layers, ld->host.data(), sizeof(LayerDev) * n_layers, cudaMemcpyHostToDevice, st));
  load_scales_kernel<<<(unsigned)ld->total_rows, 256, 0, st>>>(ld->d_layers, ld->d_row_layer, ld->d_row_idx);
  NVSE_LAUNCH_CHECK("load_scales_kernel");
  load_weights_kernel<<<(unsigned)ld->total_blocks, 256, 0, st>>>(ld->d_layers, ld->d_blk_layer, ld->d_blk_first);
  NVSE_LAUNCH_CHECK("load_weights_kernel");
  if (int rc = build_extra_images(g, st)) return rc;
  for (Layer& L : g->layers) L.have_w = L.have_bias = true;
  g->finalized = true;
  g->train_ready = with_train != 0;
  return NVSE_OK;
}

extern "C" int nvse_generator_num_layers(const nvse_generator* g) { return g ? (int)g->layers.size() : -1; }

extern "C" int nvse_generator_layer_name(const nvse_generator* g, int index, char* out, size_t capacity) {
  NVSE_REQUIRE(g && out && index >= 0 && index < (int)g->layers.size(), NVSE_ERR_INVALID, "nvse_generator_layer_name: bad argument");
  const std::string& s = g->layers[index].name;
  NVSE_REQUIRE(s.size() + 1 <= capacity, NVSE_ERR_INVALID, "nvse_generator_layer_name: buffer too small");
  memcpy(out, s.c_str(), s.size() + 1);
  return NVSE_OK;
}

extern "C" int64_t nvse_generator_total_rows(const nvse_generator* g) {
  if (!g) return -1;
  int64_t n = 0;
  for (const Layer& L : g->layers) n += L.transposed ? L.Cin : L.Cout;
  return n;
}

extern "C" int nvse_generator_weight_norm_backward(nvse_generator* g, const float* const* weight_v, const float* const* weight_g,
                                                   const float* grads, float* dv_flat, float* dg_flat, int n_layers,
                                                   void* stream) {
  NVSE_REQUIRE(g && weight_v && weight_g && grads && dv_flat && dg_flat, NVSE_ERR_INVALID,
               "nvse_generator_weight_norm_backward: null argument");
  NVSE_REQUIRE(n_layers == (int)g->layers.size(), NVSE_ERR_INVALID, "nvse_generator_weight_norm_backward: expected %d layers",
               (int)g->layers.size());
  NVSE_REQUIRE(g->loader, NVSE_ERR_STATE, "nvse_generator_weight_norm_backward: load the weights with nvse_generator_load_weights first");
  WeightLoader* ld = g->loader;
  cudaStream_t st = as_stream(stream);
  int64_t row_off = 0;
  for (int i = 0; i < n_layers; ++i) {
    const Layer& L = g->layers[i];
    const int rows = L.transposed ? L.Cin : L.Cout;
    WnBwdDev& d = ld->wn_host[i];
    d.v = weight_v[i]; d.g = weight_g[i];
    d.dw = grads + L.grad_off;
    d.dv = dv_flat + L.grad_off;
    d.dg = dg_flat + row_off;
    d.cols = (int64_t)L.Cin * L.Cout * L.k / rows;
    NVSE_REQUIRE(!d.g || d.v, NVSE_ERR_INVALID, "layer '%s': weight_g without weight_v", L.name.c_str());
    row_off += rows;
  }
  NVSE_CUDA_CHECK(cudaMemcpyAsync(ld->d_wn, ld->wn_host.data(), sizeof(WnBwdDev) * n_layers, cudaMemcpyHostToDevice, st));
  wn_bwd_all_kernel<<<(unsigned)ld->total_rows, 256, 0, st>>>(ld->d_wn, ld->d_row_layer, ld->d_row_idx);
  NVSE_LAUNCH_CHECK("wn_bwd_all_kernel");
  return NVSE_OK;
}
