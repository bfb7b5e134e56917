// fp32 CUDA-core convolutions over channels-last activations [B, T, C].
//
// These are (a) the fp32 path of the generator (parity gate: max-abs <= 1e-4 vs the
// reference), (b) the on-device reference the tcgen05 kernels are diffed against, and
// (c) the numerically sensitive edge layers (conv_pre, conv_post) of the bf16 path.
//
// One kernel shape covers Conv1d and ConvTranspose1d: a "tap list" (input row offset +
// weight slice per tap) and an affine output row map (row = out_mul*t + out_add).  A
// dilated Conv1d has taps j*d - pad; phase r of a stride-u ConvTranspose1d has taps
// delta - i with weight slices j0 + i*u and rows u*t + r.
#include "conv_f32.cuh"

namespace nvse {

namespace {

__device__ __forceinline__ float lrelu(float v, float slope) { return v >= 0.0f ? v : v * slope; }

// ------------------------------------------------------------------------------------------
// Wide kernel: implicit GEMM  M = time (64 rows / CTA), N = Cout (BN / CTA), K = taps x Cin.
// ------------------------------------------------------------------------------------------
constexpr int BM = 64, BK = 16, A_STRIDE = BM + 4;

template <int BN>
__global__ void __launch_bounds__(256) conv_taps_f32_kernel(const ConvF32Args a) {
  constexpr int TN = BN / 16;  // outputs per thread along Cout
  __shared__ __align__(16) float As[BK][A_STRIDE];  // As[ci][t]
  __shared__ __align__(16) float Bs[BK][BN];        // Bs[ci][co]

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t b = blockIdx.z;
  const int t0 = blockIdx.x * BM;
  const int co0 = blockIdx.y * BN;
  const float* __restrict__ xb = a.x + b * a.x_bstride;

  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  const int a_row = tid >> 2, a_c4 = tid & 3;  // 64 rows x 4 float4 (16 channels)
  const bool vecB = (a.Cout % 4) == 0;
  const int in_stride = a.in_stride > 1 ? a.in_stride : 1;

  for (int tap = 0; tap < a.taps.ntaps; ++tap) {
    const int off = a.taps.off[tap];
    const float* __restrict__ wt = a.w + (int64_t)a.taps.widx[tap] * a.Cin * a.Cout;
    const int trow = in_stride * (t0 + a_row) + off;
    const bool row_ok = (t0 + a_row) < a.Trows && trow >= 0 && trow < a.Tin;
    for (int ci0 = 0; ci0 < a.Cin; ci0 += BK) {
      // A tile: activation rows (with the input leaky_relu fused), stored transposed
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok) v = *reinterpret_cast<const float4*>(xb + (int64_t)trow * a.Cin + ci0 + a_c4 * 4);
      As[a_c4 * 4 + 0][a_row] = lrelu(v.x, a.in_slope);
      As[a_c4 * 4 + 1][a_row] = lrelu(v.y, a.in_slope);
      As[a_c4 * 4 + 2][a_row] = lrelu(v.z, a.in_slope);
      As[a_c4 * 4 + 3][a_row] = lrelu(v.w, a.in_slope);
      // B tile: weight slice [ci][co]
      for (int e = tid; e < BK * BN / 4; e += 256) {
        const int kk = e / (BN / 4), c4 = e % (BN / 4);
        const int co = co0 + c4 * 4;
        const float* src = wt + (int64_t)(ci0 + kk) * a.Cout + co;
        float4 wv;
        if (vecB && co + 3 < a.Cout) {
          wv = *reinterpret_cast<const float4*>(src);
        } else {
          wv.x = co + 0 < a.Cout ? src[0] : 0.f;
          wv.y = co + 1 < a.Cout ? src[1] : 0.f;
          wv.z = co + 2 < a.Cout ? src[2] : 0.f;
          wv.w = co + 3 < a.Cout ? src[3] : 0.f;
        }
        *reinterpret_cast<float4*>(&Bs[kk][c4 * 4]) = wv;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float bv[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
        const float ar[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // epilogue: bias, residual, MRF scale / accumulate, optional tanh
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= a.Trows) continue;
    const int64_t orow = (int64_t)a.out_mul * t + a.out_add;
    if (orow >= a.Tout) continue;
    float* yr = a.y + b * a.y_bstride + orow * a.Cout;
    const float* rr = a.residual ? a.residual + b * a.y_bstride + orow * a.Cout : nullptr;
    const float* mr = a.mask ? a.mask + b * a.y_bstride + orow * a.Cout : nullptr;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = co0 + tx * TN + j;
      if (co >= a.Cout) continue;
      float v = acc[i][j] + (a.bias ? a.bias[co] : 0.0f);
      if (mr && !(mr[co] > 0.0f)) v *= a.mask_slope;
      if (rr) v += rr[co];
      v *= a.out_scale;
      if (a.accumulate) v += yr[co];
      if (a.out_act == 1) v = tanhf(v);
      yr[co] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Thin kernel for few output channels (conv_post: Cout = 1 for HiFiGAN, n_fft+2 for iSTFTNet).
// One thread per output row, all NOUT outputs in registers; the activation tile (with halo
// and the fused leaky_relu) and the weight chunk are staged in shared memory.
// ------------------------------------------------------------------------------------------
constexpr int THIN_ROWS = 128, THIN_CK = 32, THIN_XSTRIDE = THIN_CK + 1, THIN_MAX_SPAN = 64;

template <int NOUT>
__global__ void __launch_bounds__(THIN_ROWS) conv_thin_f32_kernel(const ConvF32Args a, int min_off, int span) {
  extern __shared__ float sm[];
  float* xs = sm;                                             // [(THIN_ROWS + span)][33]
  float* ws = sm + (THIN_ROWS + THIN_MAX_SPAN) * THIN_XSTRIDE;  // [ntaps][32][NOUT]
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.y;
  const int t0 = blockIdx.x * THIN_ROWS;
  const float* __restrict__ xb = a.x + b * a.x_bstride;
  const int nrows = THIN_ROWS + span;
  // ragged batch: rows of this utterance beyond its own length read as zero, like rows beyond the end of the sequence
  const int tin = a.in_lens.lens ? valid_rows(a.in_lens, b, a.Tin - a.reflect_left) + a.reflect_left : a.Tin;

  float acc[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; ++j) acc[j] = 0.0f;

  for (int ci0 = 0; ci0 < a.Cin; ci0 += THIN_CK) {
    const int ck = min(THIN_CK, a.Cin - ci0);
    for (int e = tid; e < nrows * THIN_CK; e += THIN_ROWS) {
      // channels-last: lanes run along the channels of a row; T32: along (4 channels, consecutive rows)
      const int r = a.x_t32 ? (e >> 2) % nrows : e / THIN_CK;
      const int c = a.x_t32 ? ((e & 3) | (((e >> 2) / nrows) << 2)) : e % THIN_CK;
      const int vrow = t0 + min_off + r;  // row in the (virtually reflection-padded) input
      float v = 0.0f;
      if (c < ck && vrow >= 0 && vrow < tin) {
        const int arow = vrow < a.reflect_left ? a.reflect_left - vrow : vrow - a.reflect_left;
        v = lrelu(xb[a.x_t32 ? t32_off(arow, ci0 + c, a.Cin) : (int64_t)arow * a.Cin + ci0 + c], a.in_slope);
      }
      xs[r * THIN_XSTRIDE + c] = v;
    }
    for (int e = tid; e < a.taps.ntaps * THIN_CK * NOUT; e += THIN_ROWS) {
      const int co = e % NOUT, c = (e / NOUT) % THIN_CK, tap = e / (NOUT * THIN_CK);
      float v = 0.0f;
      if (c < ck && co < a.Cout) v = a.w[((int64_t)a.taps.widx[tap] * a.Cin + ci0 + c) * a.Cout + co];
      ws[e] = v;
    }
    __syncthreads();
    for (int tap = 0; tap < a.taps.ntaps; ++tap) {
      const float* xr = xs + (tid + a.taps.off[tap] - min_off) * THIN_XSTRIDE;
      const float* wr = ws + tap * THIN_CK * NOUT;
#pragma unroll 8
      for (int c = 0; c < THIN_CK; ++c) {
        const float xv = xr[c];
#pragma unroll
        for (int j = 0; j < NOUT; ++j) acc[j] = fmaf(xv, wr[c * NOUT + j], acc[j]);
      }
    }
    __syncthreads();
  }

  const int t = t0 + tid;
  if (t >= a.Trows) return;
  const int64_t orow = (int64_t)a.out_mul * t + a.out_add;
  if (orow >= a.Tout) return;
  float* yr = a.y + b * a.y_bstride + orow * a.Cout;
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
    if (j >= a.Cout) break;
    float v = acc[j] + (a.bias ? a.bias[j] : 0.0f);
    if (a.mask && !(a.mask[b * a.y_bstride + orow * a.Cout + j] > 0.0f)) v *= a.mask_slope;
    if (a.residual) v += a.residual[b * a.y_bstride + orow * a.Cout + j];
    v *= a.out_scale;
    if (a.accumulate) v += yr[j];
    if (a.out_act == 1) v = tanhf(v);
    yr[j] = v;
  }
}

// ------------------------------------------------------------------------------------------
// conv_post of the tensor-core path (HiFi-GAN: Cin -> 1 channel, hifigan.py:120-122): the input is the
// T32-layout MRF output.  Memory-bound: one pass over x with fully coalesced 16-byte loads, one thread
// per input row (see the kernel).
// ------------------------------------------------------------------------------------------
constexpr int POST_ROWS = 256;

// conv_post of the tensor-core path: Cout = 1, T32 input.  One thread per INPUT row: the row's CIN values are loaded straight
// into registers (CIN / 4 independent 16-byte loads, coalesced 512-byte warp requests), activated, and multiplied against every
// tap's weight vector (broadcast shared-memory reads): ntaps partial dot products per row.  The partials go through a small
// shared array and output o sums part[j][o + off_j - min_off].  A CTA of 256 input rows produces 256 - span outputs.  The
// round-1 kernel staged the activated rows in shared memory and read each of them ntaps times: 4 x the shared-memory
// wavefronts, which (not HBM) bounded it at 2.5 TB/s.
template <int CIN>
__global__ void __launch_bounds__(POST_ROWS) conv_post1_t32_kernel(const ConvF32Args a, int min_off, int span) {
  extern __shared__ __align__(16) float sm[];
  constexpr int c4n = CIN >> 2;
  const int ntaps = a.taps.ntaps;
  float* ws = sm;                    // [ntaps][CIN]
  float* part = sm + ntaps * CIN;    // [ntaps][POST_ROWS]
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.y;
  const int nout = POST_ROWS - span;
  const int o0 = blockIdx.x * nout;   // first output row of this CTA
  const int t = o0 + min_off + tid;   // this thread's input row
  const float* __restrict__ xb = a.x + b * a.x_bstride;
  float4 v[c4n];
  const bool in = t >= 0 && t < valid_rows(a.in_lens, b, a.Tin);
  const float4* src = reinterpret_cast<const float4*>(xb + t32_off(in ? t : 0, 0, CIN));
#pragma unroll
  for (int c4 = 0; c4 < c4n; ++c4) v[c4] = in ? __ldg(src + 32 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);  // T32: 4-channel groups 128 floats apart
  for (int e = tid; e < ntaps * CIN; e += POST_ROWS) {
    const int tap = e / CIN, c = e - tap * CIN;
    ws[e] = a.w[(int64_t)a.taps.widx[tap] * CIN + c];
  }
#pragma unroll
  for (int c4 = 0; c4 < c4n; ++c4) {
    v[c4].x = lrelu(v[c4].x, a.in_slope); v[c4].y = lrelu(v[c4].y, a.in_slope);
    v[c4].z = lrelu(v[c4].z, a.in_slope); v[c4].w = lrelu(v[c4].w, a.in_slope);
  }
  __syncthreads();
  for (int tap = 0; tap < ntaps; ++tap) {
    const float4* wr = reinterpret_cast<const float4*>(ws + tap * CIN);
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < c4n; ++c4) {
      const float4 wv = wr[c4];
      p0 = fmaf(v[c4].x, wv.x, p0); p1 = fmaf(v[c4].y, wv.y, p1); p2 = fmaf(v[c4].z, wv.z, p2); p3 = fmaf(v[c4].w, wv.w, p3);
    }
    part[tap * POST_ROWS + tid] = (p0 + p1) + (p2 + p3);
  }
  __syncthreads();
  const int o = o0 + tid;
  if (tid >= nout || o >= a.Trows) return;
  float acc = a.bias ? a.bias[0] : 0.0f;
  for (int tap = 0; tap < ntaps; ++tap) acc += part[tap * POST_ROWS + tid + a.taps.off[tap] - min_off];
  acc *= a.out_scale;
  if (a.y_pcm16) {  // same arithmetic as pcm16_kernel (core.cu) on the float result
    const float ov = a.out_act == 1 ? tanhf(acc) : acc;
    a.y_pcm16[b * a.y_bstride + o] = (int16_t)__float2int_rn(fminf(fmaxf(ov * 32767.0f, -32768.0f), 32767.0f));
    return;
  }
  float* yr = a.y + b * a.y_bstride + o;
  if (a.accumulate) acc += *yr;
  *yr = a.out_act == 1 ? tanhf(acc) : acc;
}

// [Cout, Cin, k] (Conv1d) or [Cin, Cout, k] (ConvTranspose1d) -> [k][Cin][Cout]
__global__ void __launch_bounds__(256) repack_weight_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                             int Cin, int Cout, int k, int transposed) {
  const int64_t n = (int64_t)Cin * Cout * k;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int co = e % Cout, ci = (e / Cout) % Cin, j = e / ((int64_t)Cout * Cin);
    const int64_t s = transposed ? ((int64_t)ci * Cout + co) * k + j : ((int64_t)co * Cin + ci) * k + j;
    dst[e] = src[s];
  }
}

template <int NOUT>
int launch_thin(const ConvF32Args& a, int64_t B, int min_off, int span, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((THIN_ROWS + THIN_MAX_SPAN) * THIN_XSTRIDE + (size_t)a.taps.ntaps * THIN_CK * NOUT);
  NVSE_CUDA_CHECK(cudaFuncSetAttribute(conv_thin_f32_kernel<NOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((a.Trows + THIN_ROWS - 1) / THIN_ROWS), (unsigned)B);
  conv_thin_f32_kernel<NOUT><<<grid, THIN_ROWS, smem, st>>>(a, min_off, span);
  NVSE_LAUNCH_CHECK("conv_thin_f32_kernel");
  return NVSE_OK;
}

}  // namespace

int launch_repack_weight(const float* src, float* dst, int Cin, int Cout, int k, bool transposed, cudaStream_t st) {
  const int64_t n = (int64_t)Cin * Cout * k;
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 4096);
  repack_weight_kernel<<<grid, 256, 0, st>>>(src, dst, Cin, Cout, k, transposed ? 1 : 0);
  NVSE_LAUNCH_CHECK("repack_weight_kernel");
  return NVSE_OK;
}

int launch_conv_f32(const ConvF32Args& a, int64_t B, cudaStream_t st) {
  NVSE_REQUIRE(a.taps.ntaps >= 1 && a.taps.ntaps <= kMaxTaps, NVSE_ERR_INVALID, "conv: %d taps unsupported", a.taps.ntaps);
  NVSE_REQUIRE(B <= 65535, NVSE_ERR_INVALID, "conv: batch %lld exceeds 65535 per launch", (long long)B);
  if (B == 0 || a.Trows <= 0) return NVSE_OK;
  const bool post_t32 = a.x_t32 && a.Cout == 1 && !a.reflect_left && !a.residual && !a.mask && a.out_mul == 1 && a.out_add == 0 &&
                        (a.Cin == 16 || a.Cin == 32 || a.Cin == 64);
  const bool thin_kernel = a.Cout < 32 || a.reflect_left || (a.Cin % BK != 0 && a.Cout <= 32);
  NVSE_REQUIRE(!a.in_lens.lens || post_t32 || (thin_kernel && a.in_stride <= 1), NVSE_ERR_UNSUPPORTED,
               "fp32 conv: per-utterance lengths are only implemented by the thin (conv_post) kernels");
  int min_off = a.taps.off[0], max_off = a.taps.off[0];
  for (int i = 1; i < a.taps.ntaps; ++i) {
    min_off = std::min(min_off, a.taps.off[i]);
    max_off = std::max(max_off, a.taps.off[i]);
  }
  const double prows = (double)B * a.Trows;
  ProfScope prof("conv_f32", a.Cin, a.Cout, 2.0 * prows * a.Cin * a.Cout * a.taps.ntaps,
                 prows * 4.0 * (a.Cin + a.Cout * (a.accumulate ? 2.0 : 1.0) + (a.residual ? a.Cout : 0.0)), st);
  const bool thin = a.Cout < 32 || a.reflect_left || (a.Cin % BK != 0 && a.Cout <= 32);
  NVSE_REQUIRE(!a.x_t32 || (thin && a.Cin % 4 == 0), NVSE_ERR_UNSUPPORTED, "fp32 conv: T32 input is only read by the thin kernel");
  if (thin) {
    NVSE_REQUIRE(a.Cout <= 32, NVSE_ERR_UNSUPPORTED, "thin conv: Cout=%d > 32", a.Cout);
    NVSE_REQUIRE(a.in_stride <= 1, NVSE_ERR_UNSUPPORTED, "thin conv: strided input rows are not supported");
    NVSE_REQUIRE(max_off - min_off <= THIN_MAX_SPAN, NVSE_ERR_UNSUPPORTED, "thin conv: tap span %d too wide", max_off - min_off);
    const int span = max_off - min_off;
    NVSE_REQUIRE(!a.y_pcm16 || (a.x_t32 && a.Cout == 1 && !a.reflect_left && !a.residual && !a.mask && !a.accumulate && a.out_mul == 1 &&
                                a.out_add == 0 && (a.Cin == 16 || a.Cin == 32 || a.Cin == 64)),
                 NVSE_ERR_UNSUPPORTED, "fp32 conv: fused PCM_16 output is only implemented by the T32 conv_post kernel");
    if (a.x_t32 && a.Cout == 1 && !a.reflect_left && !a.residual && !a.mask && a.out_mul == 1 && a.out_add == 0 &&
        (a.Cin == 16 || a.Cin == 32 || a.Cin == 64)) {
      const size_t smem = sizeof(float) * ((size_t)a.taps.ntaps * POST_ROWS + (size_t)a.taps.ntaps * a.Cin);
      const int nout = POST_ROWS - span;  // outputs per CTA (one thread per input row)
      dim3 grid((unsigned)((a.Trows + nout - 1) / nout), (unsigned)B);
#define POST_LAUNCH(CV)                                                                                                          \
  if (a.Cin == CV) {                                                                                                             \
    NVSE_CUDA_CHECK(cudaFuncSetAttribute(conv_post1_t32_kernel<CV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    conv_post1_t32_kernel<CV><<<grid, POST_ROWS, smem, st>>>(a, min_off, span);                                                  \
  }
      POST_LAUNCH(16) POST_LAUNCH(32) POST_LAUNCH(64)
#undef POST_LAUNCH
      NVSE_LAUNCH_CHECK("conv_post1_t32_kernel");
      return NVSE_OK;
    }
    if (a.Cout == 1) return launch_thin<1>(a, B, min_off, span, st);
    if (a.Cout <= 4) return launch_thin<4>(a, B, min_off, span, st);
    if (a.Cout <= 8) return launch_thin<8>(a, B, min_off, span, st);
    if (a.Cout <= 20) return launch_thin<20>(a, B, min_off, span, st);
    return launch_thin<32>(a, B, min_off, span, st);
  }
  NVSE_REQUIRE(a.Cin % BK == 0, NVSE_ERR_UNSUPPORTED, "conv: Cin=%d must be a multiple of %d", a.Cin, BK);
  const unsigned gx = (unsigned)((a.Trows + BM - 1) / BM);
  if (a.Cout % 64 == 0 || a.Cout > 32) {
    dim3 grid(gx, (unsigned)((a.Cout + 63) / 64), (unsigned)B);
    conv_taps_f32_kernel<64><<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid(gx, (unsigned)((a.Cout + 31) / 32), (unsigned)B);
    conv_taps_f32_kernel<32><<<grid, 256, 0, st>>>(a);
  }
  NVSE_LAUNCH_CHECK("conv_taps_f32_kernel");
  return NVSE_OK;
}

void conv1d_taps(int k, int dilation, ConvTaps* taps) {
  const int pad = (k * dilation - dilation) / 2;  // get_padding, hifigan.py:15-16
  taps->ntaps = k;
  for (int j = 0; j < k; ++j) {
    taps->off[j] = j * dilation - pad;
    taps->widx[j] = j;
  }
}

int conv_transpose_phase_taps(int k, int stride, int padding, int phase, ConvTaps* taps) {
  // y[u*t + j - p] += x[t] * W[j]  =>  for output q = u*t' + r:  j = j0 + i*u,  t = t' + delta - i
  const int j0 = (phase + padding) % stride, delta = (phase + padding) / stride;
  int n = 0;
  for (int j = j0, i = 0; j < k; j += stride, ++i) {
    if (n >= kMaxTaps) return -1;
    taps->off[n] = delta - i;
    taps->widx[n] = j;
    ++n;
  }
  taps->ntaps = n;
  return n;
}

}  // namespace nvse

// ---------------------------------------------------------------------------------------------
// C ABI (layer-level entry points)
// ---------------------------------------------------------------------------------------------
namespace {

struct ScratchF32 {  // stream-ordered scratch that is freed on the same stream
  float* p = nullptr;
  cudaStream_t st;
  explicit ScratchF32(cudaStream_t s) : st(s) {}
  cudaError_t alloc(size_t n) { return cudaMallocAsync(reinterpret_cast<void**>(&p), n * sizeof(float), st); }
  ~ScratchF32() {
    if (p) cudaFreeAsync(p, st);
  }
};

}  // namespace

extern "C" int nvse_conv1d_f32(const float* x, const float* w, const float* bias, const float* residual, float* y,
                               int64_t B, int64_t T, int Cin, int Cout, int k, int dilation, float in_slope,
                               float out_scale, int accumulate, void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(x && w && y, NVSE_ERR_INVALID, "nvse_conv1d_f32: null argument");
  NVSE_REQUIRE(B >= 0 && T >= 0 && Cin > 0 && Cout > 0 && dilation >= 1, NVSE_ERR_INVALID, "nvse_conv1d_f32: bad shape");
  NVSE_REQUIRE(k >= 1 && (k & 1) && k <= kMaxTaps, NVSE_ERR_UNSUPPORTED, "nvse_conv1d_f32: k=%d (odd k <= %d only)", k, kMaxTaps);
  NVSE_REQUIRE(T <= 0x7fffffff, NVSE_ERR_INVALID, "nvse_conv1d_f32: T too large");
  cudaStream_t st = as_stream(stream);
  ScratchF32 wp(st);
  NVSE_CUDA_CHECK(wp.alloc((size_t)Cin * Cout * k));
  if (int rc = launch_repack_weight(w, wp.p, Cin, Cout, k, false, st)) return rc;
  ConvF32Args a{};
  a.x = x; a.x_bstride = T * Cin; a.Tin = (int)T; a.Cin = Cin;
  a.w = wp.p; a.bias = bias; a.residual = residual;
  a.y = y; a.y_bstride = T * Cout; a.Tout = (int)T; a.Cout = Cout;
  conv1d_taps(k, dilation, &a.taps);
  a.out_mul = 1; a.out_add = 0; a.Trows = (int)T;
  a.in_slope = in_slope; a.out_scale = out_scale; a.accumulate = accumulate; a.out_act = 0; a.reflect_left = 0;
  return launch_conv_f32(a, B, st);
}

extern "C" int nvse_conv_transpose1d_f32(const float* x, const float* w, const float* bias, float* y, int64_t B,
                                         int64_t T, int Cin, int Cout, int k, int stride, int padding, float in_slope,
                                         void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(x && w && y, NVSE_ERR_INVALID, "nvse_conv_transpose1d_f32: null argument");
  NVSE_REQUIRE(B >= 0 && T >= 1 && Cin > 0 && Cout > 0 && stride >= 1 && padding >= 0 && k >= 1, NVSE_ERR_INVALID,
               "nvse_conv_transpose1d_f32: bad shape");
  const int64_t Tout = (T - 1) * stride - 2 * padding + k;
  NVSE_REQUIRE(Tout > 0 && Tout <= 0x7fffffff, NVSE_ERR_INVALID, "nvse_conv_transpose1d_f32: bad output length");
  cudaStream_t st = as_stream(stream);
  ScratchF32 wp(st);
  NVSE_CUDA_CHECK(wp.alloc((size_t)Cin * Cout * k));
  if (int rc = launch_repack_weight(w, wp.p, Cin, Cout, k, true, st)) return rc;
  for (int r = 0; r < stride && r < Tout; ++r) {
    ConvF32Args a{};
    a.x = x; a.x_bstride = T * Cin; a.Tin = (int)T; a.Cin = Cin;
    a.w = wp.p; a.bias = bias; a.residual = nullptr;
    a.y = y; a.y_bstride = Tout * Cout; a.Tout = (int)Tout; a.Cout = Cout;
    const int n = conv_transpose_phase_taps(k, stride, padding, r, &a.taps);
    NVSE_REQUIRE(n >= 0, NVSE_ERR_UNSUPPORTED, "nvse_conv_transpose1d_f32: more than %d taps per phase", kMaxTaps);
    a.out_mul = stride; a.out_add = r; a.Trows = (int)((Tout - r + stride - 1) / stride);
    a.in_slope = in_slope; a.out_scale = 1.0f; a.accumulate = 0; a.out_act = 0; a.reflect_left = 0;
    NVSE_REQUIRE(n > 0, NVSE_ERR_UNSUPPORTED, "nvse_conv_transpose1d_f32: k < stride is not supported");
    if (int rc = launch_conv_f32(a, B, st)) return rc;
  }
  return NVSE_OK;
}
