// Backward-pass kernels of the generator (fp32, CUDA cores): the training step of the reference
// (train_time_wi_inv.py:173-236) back-propagates through HiFiGAN.forward; autograd there dispatches
// cuDNN wgrad/dgrad per layer.  Here
//   * a data gradient is the forward tap-list kernel (conv_f32.cu) run with per-tap transposed
//     weights, mirrored taps and the leaky_relu derivative + residual-branch gradient in the epilogue;
//   * a weight gradient is the correlation below: an implicit GEMM whose reduction dimension is
//     (batch x time), split over CTAs into partial sums that a second kernel adds in a fixed order
//     (bit-reproducible gradients, no atomics).
#include "grad.cuh"

namespace nvse {

namespace {

__device__ __forceinline__ float lrelu(float v, float slope) { return v >= 0.0f ? v : v * slope; }

constexpr int WG_BK = 16;

// TILE x WG_BK channel slab of one operand: row `r` of the chunk from `src_row` (null: zeros), channels c0 .. c0+TILE-1
template <int TILE>
__device__ __forceinline__ void wg_load(float (*S)[TILE], int r, int c4, const float* __restrict__ src_row, int c0, int C,
                                        float slope) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  const int c = c0 + c4 * 4;
  if (src_row) {
    if ((C & 3) == 0 && c + 3 < C) {
      v = *reinterpret_cast<const float4*>(src_row + c);
    } else {
      if (c + 0 < C) v.x = src_row[c + 0];
      if (c + 1 < C) v.y = src_row[c + 1];
      if (c + 2 < C) v.z = src_row[c + 2];
      if (c + 3 < C) v.w = src_row[c + 3];
    }
    v.x = lrelu(v.x, slope); v.y = lrelu(v.y, slope); v.z = lrelu(v.z, slope); v.w = lrelu(v.w, slope);
  }
  *reinterpret_cast<float4*>(&S[r][c4 * 4]) = v;
}

// grid: (tiles_a * tiles_b, ntaps, nsplit); one TILE x TILE tile of G[tap] over rows [split * rows_per_split, +rows_per_split);
// (TILE/4)^2 threads, 4 x 4 outputs each (TILE = 32 for the <= 32-channel layers: a 64-wide tile would idle 3/4 of the FMAs)
template <int TILE>
__global__ void __launch_bounds__((TILE / 4) * (TILE / 4)) wgrad_kernel(const WgradArgs a, int64_t rows_total, int rows_per_split,
                                                                        int tiles_b, float* __restrict__ partial) {
  constexpr int TQ = TILE / 4, NT = TQ * TQ;
  __shared__ __align__(16) float As[WG_BK][TILE];  // As[row][ca]
  __shared__ __align__(16) float Bs[WG_BK][TILE];  // Bs[row][cb]
  const int tid = threadIdx.x;
  const int tx = tid % TQ, ty = tid / TQ;
  const int ca0 = ((int)blockIdx.x / tiles_b) * TILE, cb0 = ((int)blockIdx.x % tiles_b) * TILE;
  const int tap = blockIdx.y, off = a.off[tap];
  const int64_t r_lo = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_hi = r_lo + rows_per_split < rows_total ? r_lo + rows_per_split : rows_total;
  const int u_stride = a.u_stride > 1 ? a.u_stride : 1;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int64_t r0 = r_lo; r0 < r_hi; r0 += WG_BK) {
#pragma unroll
    for (int e = tid; e < WG_BK * TQ; e += NT) {  // WG_BK rows x TQ float4 per operand
      const int l_row = e / TQ, l_c4 = e % TQ;
      const int64_t r = r0 + l_row;
      const float* urow = nullptr;
      const float* vrow = nullptr;
      if (r < r_hi) {
        const int64_t b = r / a.Tv;
        const int t = (int)(r - b * a.Tv);
        const int64_t tu = (int64_t)u_stride * t + off;
        vrow = a.V + b * a.v_bstride + (int64_t)t * a.Cb;
        if (tu >= 0 && tu < a.Tu) urow = a.U + b * a.u_bstride + tu * a.Ca;
      }
      wg_load<TILE>(As, l_row, l_c4, urow, ca0, a.Ca, a.u_slope);
      wg_load<TILE>(Bs, l_row, l_c4, urow ? vrow : nullptr, cb0, a.Cb, a.v_slope);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < WG_BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* pt = partial + ((int64_t)blockIdx.z * a.ntaps + tap) * a.Ca * a.Cb;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ca = ca0 + ty * 4 + i;
    if (ca >= a.Ca) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cb = cb0 + tx * 4 + j;
      if (cb < a.Cb) pt[(int64_t)ca * a.Cb + cb] = acc[i][j];
    }
  }
}

// dst[(cb*Ca + ca)*ntaps + j] = scale * sum_s partial[s][j][ca][cb].  32 elements x 8 split lanes per CTA: lane q adds
// the partials s = q, q + 8, ... in ascending order, then the 8 lane sums are added in ascending q -- a fixed order,
// so the result is bit-reproducible, with nsplit / 8 dependent loads per thread instead of nsplit.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, int ntaps, int Ca,
                                                            int Cb, float* __restrict__ dst, float scale) {
  __shared__ float red[8][33];
  const int64_t n = (int64_t)ntaps * Ca * Cb;
  const int el = threadIdx.x & 31, q = threadIdx.x >> 5;
  for (int64_t e0 = (int64_t)blockIdx.x * 32; e0 < n; e0 += (int64_t)gridDim.x * 32) {
    const int64_t e = e0 + el;
    float s = 0.0f;
    if (e < n)
      for (int k = q; k < nsplit; k += 8) s += partial[(int64_t)k * n + e];
    red[q][el] = s;
    __syncthreads();
    if (q == 0 && e < n) {
      float t = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][el];
      const int cb = (int)(e % Cb), ca = (int)((e / Cb) % Ca), j = (int)(e / ((int64_t)Ca * Cb));
      dst[((int64_t)cb * Ca + ca) * ntaps + j] = scale * t;
    }
    __syncthreads();
  }
}

// partial[split][c] = sum of V[row][c] over the rows of the split; 32 channels x 8 row lanes per CTA
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ V, int64_t rows, int C, int64_t rows_per_split,
                                                     float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  const int64_t r_lo = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_hi = r_lo + rows_per_split < rows ? r_lo + rows_per_split : rows;
  float s = 0.0f;
  if (c < C)
    for (int64_t r = r_lo + rl; r < r_hi; r += 8) s += V[r * C + c];
  red[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < C) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    partial[(int64_t)blockIdx.y * C + c] = t;
  }
}

__global__ void __launch_bounds__(256) tanh_bwd_kernel(const float* __restrict__ out, const float* __restrict__ dout,
                                                       float* __restrict__ dz, int64_t n) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const float o = out[e];
    dz[e] = dout[e] * (1.0f - o * o);
  }
}

__global__ void __launch_bounds__(256) transpose_taps_kernel(const float* __restrict__ src, float* __restrict__ dst, int k,
                                                             int Cin, int Cout) {
  const int64_t n = (int64_t)k * Cin * Cout;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(e % Cin), co = (int)((e / Cin) % Cout), j = (int)(e / ((int64_t)Cin * Cout));
    dst[e] = src[((int64_t)j * Cin + ci) * Cout + co];
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.0f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// one CTA per row (torch.nn.utils.weight_norm, dim = 0; the backward of the fold in core.cu)
__global__ void __launch_bounds__(256) weight_norm_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                              const float* __restrict__ dw, float* __restrict__ dv,
                                                              float* __restrict__ dg, int64_t cols) {
  __shared__ float red[8];
  const int64_t r = blockIdx.x;
  const float* vr = v + r * cols;
  const float* wr = dw + r * cols;
  float ss = 0.0f, dot = 0.0f;
  for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) {
    const float x = vr[c];
    ss = fmaf(x, x, ss);
    dot = fmaf(wr[c], x, dot);
  }
  ss = block_sum(ss, red);
  dot = block_sum(dot, red);
  const float nrm = sqrtf(ss), inv = 1.0f / nrm, gr = g[r];
  if (threadIdx.x == 0) dg[r] = dot * inv;
  const float s1 = gr * inv, s2 = dot * inv * inv;
  for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) dv[r * cols + c] = s1 * (wr[c] - vr[c] * s2);
}

struct WgSplit { int nsplit; int rows_per_split; };
inline int wg_tile(int Ca, int Cb) { return (Ca <= 32 && Cb <= 32) ? 32 : 64; }
WgSplit wg_split(int Ca, int Cb, int ntaps, int64_t rows) {
  const int tile = wg_tile(Ca, Cb);
  const int64_t tiles = (int64_t)((Ca + tile - 1) / tile) * ((Cb + tile - 1) / tile) * ntaps;
  int64_t want = ((tile == 32 ? 2368 : 592) + tiles - 1) / tiles;  // ~1024 threads per SM in flight
  const int64_t cap = (rows + 127) / 128;    // at least 128 rows per split
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  int64_t per = (rows + want - 1) / want;
  per = (per + WG_BK - 1) / WG_BK * WG_BK;
  if (per < WG_BK) per = WG_BK;
  WgSplit s;
  s.rows_per_split = (int)per;
  s.nsplit = (int)std::max<int64_t>(1, (rows + per - 1) / per);
  return s;
}

struct CsSplit { int nsplit; int64_t rows_per_split; };
CsSplit cs_split(int C, int64_t rows) {
  const int64_t groups = (C + 31) / 32;
  int64_t want = std::max<int64_t>(1, std::min<int64_t>((592 + groups - 1) / groups, (rows + 255) / 256));
  CsSplit s;
  s.rows_per_split = (rows + want - 1) / want;
  if (s.rows_per_split < 1) s.rows_per_split = 1;
  s.nsplit = (int)std::max<int64_t>(1, (rows + s.rows_per_split - 1) / s.rows_per_split);
  return s;
}

}  // namespace

size_t wgrad_scratch_elems(int Ca, int Cb, int ntaps, int64_t B, int Tv) {
  const WgSplit s = wg_split(Ca, Cb, ntaps, B * Tv);
  return std::max((size_t)s.nsplit * ntaps * Ca * Cb, wgrad_tc_scratch_elems(Ca, Cb, ntaps, B, Tv));
}

int launch_wgrad_reduce(const float* partial, int nsplit, int ntaps, int Ca, int Cb, float* dst, float scale, cudaStream_t st) {
  const int64_t n = (int64_t)ntaps * Ca * Cb;
  wgrad_reduce_kernel<<<(unsigned)std::min<int64_t>((n + 31) / 32, 8192), 256, 0, st>>>(partial, nsplit, ntaps, Ca, Cb, dst, scale);
  NVSE_LAUNCH_CHECK("wgrad_reduce_kernel");
  return NVSE_OK;
}

int launch_wgrad(const WgradArgs& a, int64_t B, float* scratch, cudaStream_t st) {
  NVSE_REQUIRE(a.ntaps >= 1 && a.ntaps <= kMaxTaps, NVSE_ERR_INVALID, "wgrad: %d taps unsupported", a.ntaps);
  NVSE_REQUIRE(a.U && a.V && a.dst && scratch, NVSE_ERR_INVALID, "wgrad: null argument");
  const int64_t rows = B * a.Tv;
  if (rows <= 0) {
    NVSE_CUDA_CHECK(cudaMemsetAsync(a.dst, 0, sizeof(float) * (size_t)a.ntaps * a.Ca * a.Cb, st));
    return NVSE_OK;
  }
  if (a.tc && wgrad_tc_supported(a.Ca, a.Cb, a.ntaps, a.off, a.u_stride, B, a.Tv)) return launch_wgrad_tc(a, B, scratch, st);
  const WgSplit s = wg_split(a.Ca, a.Cb, a.ntaps, rows);
  NVSE_REQUIRE(s.nsplit <= 65535, NVSE_ERR_INVALID, "wgrad: too many splits");
  const int tile = wg_tile(a.Ca, a.Cb);
  const int tiles_a = (a.Ca + tile - 1) / tile, tiles_b = (a.Cb + tile - 1) / tile;
  {
    ProfScope prof("wgrad_f32", a.Ca, a.Cb, 2.0 * (double)rows * a.Ca * a.Cb * a.ntaps,
                   (double)rows * 4.0 * (a.Ca + a.Cb) * a.ntaps, st);
    dim3 grid((unsigned)(tiles_a * tiles_b), (unsigned)a.ntaps, (unsigned)s.nsplit);
    if (tile == 32) wgrad_kernel<32><<<grid, 64, 0, st>>>(a, rows, s.rows_per_split, tiles_b, scratch);
    else wgrad_kernel<64><<<grid, 256, 0, st>>>(a, rows, s.rows_per_split, tiles_b, scratch);
    NVSE_LAUNCH_CHECK("wgrad_kernel");
  }
  return launch_wgrad_reduce(scratch, s.nsplit, a.ntaps, a.Ca, a.Cb, a.dst, a.scale, st);
}

size_t colsum_scratch_elems(int C, int64_t rows) { return (size_t)cs_split(C, rows).nsplit * C; }

int launch_colsum(const float* V, int64_t rows, int C, float* dst, float scale, float* scratch, cudaStream_t st) {
  NVSE_REQUIRE(V && dst && scratch && C >= 1, NVSE_ERR_INVALID, "colsum: bad argument");
  if (rows <= 0) {
    NVSE_CUDA_CHECK(cudaMemsetAsync(dst, 0, sizeof(float) * C, st));
    return NVSE_OK;
  }
  const CsSplit s = cs_split(C, rows);
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)s.nsplit);
  colsum_kernel<<<grid, 256, 0, st>>>(V, rows, C, s.rows_per_split, scratch);
  NVSE_LAUNCH_CHECK("colsum_kernel");
  wgrad_reduce_kernel<<<(unsigned)((C + 31) / 32), 256, 0, st>>>(scratch, s.nsplit, 1, 1, C, dst, scale);
  NVSE_LAUNCH_CHECK("wgrad_reduce_kernel");
  return NVSE_OK;
}

int launch_tanh_bwd(const float* out, const float* dout, float* dz, int64_t n, cudaStream_t st) {
  if (n <= 0) return NVSE_OK;
  tanh_bwd_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, st>>>(out, dout, dz, n);
  NVSE_LAUNCH_CHECK("tanh_bwd_kernel");
  return NVSE_OK;
}

int launch_transpose_taps(const float* src, float* dst, int k, int Cin, int Cout, cudaStream_t st) {
  const int64_t n = (int64_t)k * Cin * Cout;
  transpose_taps_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 4096), 256, 0, st>>>(src, dst, k, Cin, Cout);
  NVSE_LAUNCH_CHECK("transpose_taps_kernel");
  return NVSE_OK;
}

int launch_weight_norm_bwd(const float* v, const float* g, const float* dw, float* dv, float* dg, int64_t rows, int64_t cols,
                           cudaStream_t st) {
  if (rows <= 0) return NVSE_OK;
  NVSE_REQUIRE(rows <= 0x7fffffff, NVSE_ERR_INVALID, "weight_norm_bwd: too many rows");
  weight_norm_bwd_kernel<<<(unsigned)rows, 256, 0, st>>>(v, g, dw, dv, dg, cols);
  NVSE_LAUNCH_CHECK("weight_norm_bwd_kernel");
  return NVSE_OK;
}

}  // namespace nvse

// ---------------------------------------------------------------------------------------------
// C ABI (layer-level entry points of the backward pass; parity-tested against torch autograd)
// ---------------------------------------------------------------------------------------------
namespace {
struct Scratch {
  float* p = nullptr;
  cudaStream_t st;
  explicit Scratch(cudaStream_t s) : st(s) {}
  cudaError_t alloc(size_t n) { return cudaMallocAsync(reinterpret_cast<void**>(&p), std::max<size_t>(n, 1) * sizeof(float), st); }
  ~Scratch() {
    if (p) cudaFreeAsync(p, st);
  }
};
}  // namespace

using namespace nvse;

extern "C" int nvse_conv1d_backward_f32(const float* x, const float* w, const float* dy, const float* dresidual_in,
                                        float* dx, float* dw, float* dbias, int64_t B, int64_t T, int Cin, int Cout, int k,
                                        int dilation, float in_slope, int precision, void* stream) {
  NVSE_REQUIRE(x && w && dy, NVSE_ERR_INVALID, "nvse_conv1d_backward_f32: null argument");
  NVSE_REQUIRE(B >= 0 && T >= 0 && Cin > 0 && Cout > 0 && dilation >= 1, NVSE_ERR_INVALID, "nvse_conv1d_backward_f32: bad shape");
  NVSE_REQUIRE(k >= 1 && (k & 1) && k <= kMaxTaps, NVSE_ERR_UNSUPPORTED, "nvse_conv1d_backward_f32: k=%d (odd k <= %d only)", k, kMaxTaps);
  NVSE_REQUIRE(T <= 0x7fffffff, NVSE_ERR_INVALID, "nvse_conv1d_backward_f32: T too large");
  cudaStream_t st = as_stream(stream);
  const int pad = (k * dilation - dilation) / 2;
  if (dx) {
    Scratch wp(st), wt(st);
    NVSE_CUDA_CHECK(wp.alloc((size_t)Cin * Cout * k));
    NVSE_CUDA_CHECK(wt.alloc((size_t)Cin * Cout * k));
    if (int rc = launch_repack_weight(w, wp.p, Cin, Cout, k, false, st)) return rc;
    if (int rc = launch_transpose_taps(wp.p, wt.p, k, Cin, Cout, st)) return rc;
    ConvF32Args a{};
    a.x = dy; a.x_bstride = T * Cout; a.Tin = (int)T; a.Cin = Cout;
    a.w = wt.p; a.residual = dresidual_in;
    a.y = dx; a.y_bstride = T * Cin; a.Tout = (int)T; a.Cout = Cin;
    a.taps.ntaps = k;
    for (int j = 0; j < k; ++j) { a.taps.off[j] = pad - j * dilation; a.taps.widx[j] = j; }
    a.out_mul = 1; a.Trows = (int)T; a.in_slope = 1.0f; a.out_scale = 1.0f;
    if (in_slope != 1.0f) { a.mask = x; a.mask_slope = in_slope; }
    if (int rc = launch_conv_f32(a, B, st)) return rc;
  }
  if (dw) {
    WgradArgs g{};
    g.U = x; g.u_bstride = T * Cin; g.Tu = (int)T; g.Ca = Cin; g.u_slope = in_slope;
    g.V = dy; g.v_bstride = T * Cout; g.Tv = (int)T; g.Cb = Cout; g.v_slope = 1.0f;
    g.u_stride = 1; g.ntaps = k;
    for (int j = 0; j < k; ++j) g.off[j] = j * dilation - pad;
    g.dst = dw; g.scale = 1.0f; g.tc = precision == NVSE_PRECISION_BF16;
    Scratch sc(st);
    NVSE_CUDA_CHECK(sc.alloc(wgrad_scratch_elems(Cin, Cout, k, B, (int)T)));
    if (int rc = launch_wgrad(g, B, sc.p, st)) return rc;
  }
  if (dbias) {
    Scratch sc(st);
    NVSE_CUDA_CHECK(sc.alloc(colsum_scratch_elems(Cout, B * T)));
    if (int rc = launch_colsum(dy, B * T, Cout, dbias, 1.0f, sc.p, st)) return rc;
  }
  return NVSE_OK;
}

extern "C" int nvse_conv_transpose1d_backward_f32(const float* x, const float* w, const float* dy, float* dx, float* dw,
                                                  float* dbias, int64_t B, int64_t T, int Cin, int Cout, int k, int stride,
                                                  int padding, float in_slope, void* stream) {
  NVSE_REQUIRE(x && w && dy, NVSE_ERR_INVALID, "nvse_conv_transpose1d_backward_f32: null argument");
  NVSE_REQUIRE(B >= 0 && T >= 1 && Cin > 0 && Cout > 0 && stride >= 1 && padding >= 0 && k >= 1 && k <= kMaxTaps,
               NVSE_ERR_INVALID, "nvse_conv_transpose1d_backward_f32: bad shape");
  const int64_t Tout = (T - 1) * stride - 2 * padding + k;
  NVSE_REQUIRE(Tout > 0 && Tout <= 0x7fffffff, NVSE_ERR_INVALID, "nvse_conv_transpose1d_backward_f32: bad output length");
  cudaStream_t st = as_stream(stream);
  if (dx) {
    Scratch wp(st), wt(st);
    NVSE_CUDA_CHECK(wp.alloc((size_t)Cin * Cout * k));
    NVSE_CUDA_CHECK(wt.alloc((size_t)Cin * Cout * k));
    if (int rc = launch_repack_weight(w, wp.p, Cin, Cout, k, true, st)) return rc;
    if (int rc = launch_transpose_taps(wp.p, wt.p, k, Cin, Cout, st)) return rc;
    ConvF32Args a{};
    a.x = dy; a.x_bstride = Tout * Cout; a.Tin = (int)Tout; a.Cin = Cout;
    a.w = wt.p;
    a.y = dx; a.y_bstride = T * Cin; a.Tout = (int)T; a.Cout = Cin;
    a.taps.ntaps = k;
    for (int j = 0; j < k; ++j) { a.taps.off[j] = j - padding; a.taps.widx[j] = j; }
    a.in_stride = stride;
    a.out_mul = 1; a.Trows = (int)T; a.in_slope = 1.0f; a.out_scale = 1.0f;
    if (in_slope != 1.0f) { a.mask = x; a.mask_slope = in_slope; }
    if (int rc = launch_conv_f32(a, B, st)) return rc;
  }
  if (dw) {
    WgradArgs g{};
    g.U = dy; g.u_bstride = Tout * Cout; g.Tu = (int)Tout; g.Ca = Cout; g.u_slope = 1.0f;
    g.V = x; g.v_bstride = T * Cin; g.Tv = (int)T; g.Cb = Cin; g.v_slope = in_slope;
    g.u_stride = stride; g.ntaps = k;
    for (int j = 0; j < k; ++j) g.off[j] = j - padding;
    g.dst = dw; g.scale = 1.0f;
    Scratch sc(st);
    NVSE_CUDA_CHECK(sc.alloc(wgrad_scratch_elems(Cout, Cin, k, B, (int)T)));
    if (int rc = launch_wgrad(g, B, sc.p, st)) return rc;
  }
  if (dbias) {
    Scratch sc(st);
    NVSE_CUDA_CHECK(sc.alloc(colsum_scratch_elems(Cout, B * Tout)));
    if (int rc = launch_colsum(dy, B * Tout, Cout, dbias, 1.0f, sc.p, st)) return rc;
  }
  return NVSE_OK;
}

extern "C" int nvse_weight_norm_backward_f32(const float* v, const float* g, const float* dw, float* dv, float* dg,
                                             int64_t rows, int64_t cols, void* stream) {
  NVSE_REQUIRE(v && g && dw && dv && dg, NVSE_ERR_INVALID, "nvse_weight_norm_backward_f32: null argument");
  NVSE_REQUIRE(rows >= 0 && cols >= 1, NVSE_ERR_INVALID, "nvse_weight_norm_backward_f32: bad shape");
  return launch_weight_norm_bwd(v, g, dw, dv, dg, rows, cols, as_stream(stream));
}
