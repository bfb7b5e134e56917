// Error plumbing, launch accounting and the small layout / weight-norm kernels.
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace nvse {

std::atomic<uint64_t> g_launches{0};

std::string& last_error_slot() {
  static thread_local std::string slot;
  return slot;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_slot() = buf;
  return code;
}

// ---- per-launch profiler ---------------------------------------------------------------------
namespace {
struct ProfRecord {
  std::string key;
  double flops, bytes;
  cudaEvent_t start, stop;
};
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRecord> g_prof;
std::vector<cudaEvent_t> g_event_pool;

cudaEvent_t take_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

ProfScope::ProfScope(const char* kernel, int c_in, int c_out, double flops, double bytes, cudaStream_t stream) : st(stream) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRecord r;
  char buf[96];
  snprintf(buf, sizeof(buf), "%s[%d->%d]", kernel, c_in, c_out);
  r.key = buf;
  r.flops = flops;
  r.bytes = bytes;
  r.start = take_event();
  r.stop = take_event();
  cudaEventRecord(r.start, st);
  slot = (int)g_prof.size();
  g_prof.push_back(r);
}

ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].stop, st);
}

namespace {

// w[r,:] = g[r] * v[r,:] / ||v[r,:]||   (one CTA per row)
__global__ void __launch_bounds__(256) weight_norm_fold_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                                float* __restrict__ w, int64_t cols) {
  __shared__ float red[8];
  __shared__ float scale;
  const int64_t r = blockIdx.x;
  const float* vr = v + r * cols;
  float acc = 0.0f;
  for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) acc = fmaf(vr[c], vr[c], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int i = 0; i < 8; ++i) s += red[i];
    scale = g[r] / sqrtf(s);
  }
  __syncthreads();
  const float sc = scale;
  for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) w[r * cols + c] = vr[c] * sc;
}

// y[b, j, i] = x[b, i, j] for x: [B, R, C]  (32x32 tiles through padded smem); y rows have Rpad >= R entries, zero beyond R
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t R,
                                                         int64_t C, int64_t Rpad) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* xb = x + b * R * C;
  float* yb = y + b * Rpad * C;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < R && c < C) ? xb[r * C + c] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t c = c0 + ty + i, r = r0 + tx;
    if (r < Rpad && c < C) yb[c * Rpad + r] = tile[tx][ty + i];
  }
}

}  // namespace

namespace {
// channels-last [B, T, C] <-> T32 (common.cuh); element-wise gather, used by tests / the layer-level entry points
__global__ void __launch_bounds__(256) relayout_t32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t T, int C,
                                                           int to_t32) {
  const int64_t b = blockIdx.y, n = T * C, tstride = t32_rows(T) * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = e / C;
    const int c = (int)(e - t * C);
    if (to_t32) y[b * tstride + t32_off(t, c, C)] = x[b * n + e];
    else y[b * n + e] = x[b * tstride + t32_off(t, c, C)];
  }
}
}  // namespace

int launch_relayout_t32(const float* x, float* y, int64_t B, int64_t T, int C, bool to_t32, cudaStream_t st) {
  if (B == 0 || T == 0) return NVSE_OK;
  NVSE_REQUIRE(B <= 65535 && C % 4 == 0, NVSE_ERR_INVALID, "relayout: bad shape");
  dim3 grid((unsigned)std::min<int64_t>((T * C + 255) / 256, 8192), (unsigned)B);
  relayout_t32_kernel<<<grid, 256, 0, st>>>(x, y, T, C, to_t32 ? 1 : 0);
  NVSE_LAUNCH_CHECK("relayout_t32_kernel");
  return NVSE_OK;
}

namespace {
// float waveform -> PCM_16 as libsndfile writes it for the reference (sf.write(..., 'PCM_16'),
// infers/inference_hifigan.py:93): lrint(x * 0x7FFF), clipped to the int16 range
__global__ void __launch_bounds__(256) pcm16_kernel(const float* __restrict__ x, int16_t* __restrict__ y, int64_t n) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const float v = fminf(fmaxf(x[e] * 32767.0f, -32768.0f), 32767.0f);
    y[e] = (int16_t)__float2int_rn(v);
  }
}
}  // namespace

namespace {
__global__ void __launch_bounds__(256) add3_kernel(float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c, int64_t n4) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
    float4 va = reinterpret_cast<float4*>(a)[e];
    const float4 vb = reinterpret_cast<const float4*>(b)[e];
    va.x += vb.x; va.y += vb.y; va.z += vb.z; va.w += vb.w;
    if (c) {
      const float4 vc = reinterpret_cast<const float4*>(c)[e];
      va.x += vc.x; va.y += vc.y; va.z += vc.z; va.w += vc.w;
    }
    reinterpret_cast<float4*>(a)[e] = va;
  }
}
}  // namespace

// a = (a + b) + c  (c may be null); n a multiple of 4, 16-byte aligned buffers
int launch_add3(float* a, const float* b, const float* c, int64_t n, cudaStream_t st) {
  if (n == 0) return NVSE_OK;
  NVSE_REQUIRE(n % 4 == 0, NVSE_ERR_INVALID, "add3: length must be a multiple of 4");
  add3_kernel<<<(unsigned)std::min<int64_t>((n / 4 + 255) / 256, 148 * 8), 256, 0, st>>>(a, b, c, n / 4);
  NVSE_LAUNCH_CHECK("add3_kernel");
  return NVSE_OK;
}

int launch_pcm16(const float* x, int16_t* y, int64_t n, cudaStream_t st) {
  if (n == 0) return NVSE_OK;
  pcm16_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(x, y, n);
  NVSE_LAUNCH_CHECK("pcm16_kernel");
  return NVSE_OK;
}

int launch_transpose_pad(const float* x, float* y, int64_t B, int64_t R, int64_t C, int64_t Rpad, cudaStream_t st) {
  if (B == 0 || R == 0 || C == 0) return NVSE_OK;
  NVSE_REQUIRE(B <= 65535 && Rpad >= R, NVSE_ERR_INVALID, "transpose: batch %lld exceeds 65535 or bad padding", (long long)B);
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)((Rpad + 31) / 32), (unsigned)B);
  NVSE_REQUIRE((Rpad + 31) / 32 <= 65535, NVSE_ERR_INVALID, "transpose: too many rows");
  transpose_kernel<<<grid, 256, 0, st>>>(x, y, R, C, Rpad);
  NVSE_LAUNCH_CHECK("transpose_kernel");
  return NVSE_OK;
}
int launch_transpose(const float* x, float* y, int64_t B, int64_t R, int64_t C, cudaStream_t st) {
  return launch_transpose_pad(x, y, B, R, C, R, st);
}

}  // namespace nvse

namespace nvse {
int device_sm_count() {
  static std::atomic<int> cache[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}
}  // namespace nvse

extern "C" int nvse_abi_version(void) { return NVSE_ABI_VERSION; }
extern "C" const char* nvse_last_error(void) { return nvse::last_error_slot().c_str(); }
extern "C" uint64_t nvse_launch_count(void) { return nvse::g_launches.load(); }

extern "C" int nvse_profile_begin(void) {
  using namespace nvse;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (ProfRecord& r : g_prof) {
    g_event_pool.push_back(r.start);
    g_event_pool.push_back(r.stop);
  }
  g_prof.clear();
  g_prof_on = true;
  return NVSE_OK;
}

extern "C" int nvse_profile_end(char* json_out, size_t capacity) {
  using namespace nvse;
  NVSE_REQUIRE(json_out && capacity > 2, NVSE_ERR_INVALID, "nvse_profile_end: bad buffer");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = false;
  struct Agg { uint64_t n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (ProfRecord& r : g_prof) {
    NVSE_CUDA_CHECK(cudaEventSynchronize(r.stop));
    float ms = 0.0f;
    NVSE_CUDA_CHECK(cudaEventElapsedTime(&ms, r.start, r.stop));
    Agg& a = agg[r.key];
    a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
    g_event_pool.push_back(r.start);
    g_event_pool.push_back(r.stop);
  }
  g_prof.clear();
  std::string js = "[";
  for (auto& kv : agg) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s{\"kernel\": \"%s\", \"launches\": %llu, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}",
             js.size() > 1 ? ", " : "", kv.first.c_str(), (unsigned long long)kv.second.n, kv.second.ms, kv.second.flops,
             kv.second.bytes);
    js += buf;
  }
  js += "]";
  NVSE_REQUIRE(js.size() + 1 <= capacity, NVSE_ERR_INVALID, "nvse_profile_end: buffer too small (%zu needed)", js.size() + 1);
  memcpy(json_out, js.c_str(), js.size() + 1);
  return NVSE_OK;
}

extern "C" int nvse_weight_norm_fold_f32(const float* v, const float* g, float* w, int64_t rows, int64_t cols,
                                         void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(v && g && w && rows > 0 && cols > 0, NVSE_ERR_INVALID, "nvse_weight_norm_fold_f32: bad argument");
  NVSE_REQUIRE(rows <= 0x7fffffff, NVSE_ERR_INVALID, "nvse_weight_norm_fold_f32: too many rows");
  weight_norm_fold_kernel<<<(unsigned)rows, 256, 0, as_stream(stream)>>>(v, g, w, cols);
  NVSE_LAUNCH_CHECK("weight_norm_fold_kernel");
  return NVSE_OK;
}

extern "C" int nvse_pcm16_from_f32(const float* x, int16_t* y, int64_t n, void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(n >= 0 && (n == 0 || (x && y)), NVSE_ERR_INVALID, "nvse_pcm16_from_f32: bad argument");
  return launch_pcm16(x, y, n, as_stream(stream));
}

extern "C" int nvse_transpose_bct_to_btc_f32(const float* x, float* y, int64_t B, int64_t C, int64_t T, void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(x && y && B >= 0 && C >= 0 && T >= 0, NVSE_ERR_INVALID, "nvse_transpose_bct_to_btc_f32: bad argument");
  return launch_transpose(x, y, B, C, T, as_stream(stream));
}

extern "C" int nvse_transpose_btc_to_bct_f32(const float* x, float* y, int64_t B, int64_t T, int64_t C, void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(x && y && B >= 0 && C >= 0 && T >= 0, NVSE_ERR_INVALID, "nvse_transpose_btc_to_bct_f32: bad argument");
  return launch_transpose(x, y, B, T, C, as_stream(stream));
}
