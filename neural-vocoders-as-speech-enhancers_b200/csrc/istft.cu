// iSTFT head of iSTFTNet (reference Models/istftnet.py:314-316 + TorchSTFT.inverse :183-188):
//
//   mag = exp(z[:nb]), phase = sin(z[nb:]), spec = mag * e^{i phase}
//   frames = irfft(spec, n_fft) * hann ; overlap-add at hop ; / sum(hann^2) ; trim n_fft/2
//
// One memory-bound kernel: each thread turns one conv_post output row (n_fft+2 floats)
// into a windowed n_fft-sample frame in shared memory (direct real inverse DFT from a
// shared cos/sin table: n_fft is tiny), then each thread gathers the n_fft/hop frames
// that overlap its `hop` output samples and divides by the window envelope, which is
// accumulated on the fly so the edge samples (where fewer frames overlap) are exact.
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace nvse {

namespace {

constexpr int kGroupsPerCta = 128;  // hop-groups (of `hop` output samples) per CTA

template <int NFFT>
__global__ void __launch_bounds__(kGroupsPerCta) istft_head_kernel(const float* __restrict__ z, float* __restrict__ out,
                                                                    int64_t Tp_all, int hop, const RowLens lens, int zpitch) {
  constexpr int NB = NFFT / 2 + 1;
  constexpr int CH = NFFT + 2;
  constexpr int FS = NFFT + 1;  // padded frame stride in smem
  extern __shared__ float sm[];
  float* ctab = sm;            // cos(2 pi j / NFFT)
  float* stab = ctab + NFFT;   // sin(2 pi j / NFFT)
  float* win = stab + NFFT;    // periodic Hann
  float* frames = win + NFFT;  // [nframes][FS]

  const int tid = threadIdx.x;
  const int64_t b = blockIdx.y;
  // ragged batch: utterance b has Tp frames of its own (conv_post rows incl. the reflected one); the frames beyond them exist in
  // memory but belong to the padding and must not reach the overlap-add of its last samples
  const int64_t Tp = lens.lens ? (int64_t)valid_rows(lens, b, (int)(Tp_all - 1)) + 1 : Tp_all;
  const int R = NFFT / hop;            // frames overlapping one output sample
  const int64_t n_groups = Tp - 1;     // output length = hop * (Tp - 1)
  const int64_t g0 = (int64_t)blockIdx.x * kGroupsPerCta;
  // output sample o (trimmed) = untrimmed q - NFFT/2; frames tau with hop*tau <= q < hop*tau + NFFT
  const int64_t q_lo = g0 * hop + NFFT / 2;
  const int64_t q_hi = min(g0 + kGroupsPerCta, n_groups) * hop + NFFT / 2 - 1;
  const int64_t tau_lo = max((int64_t)0, q_lo / hop - (R - 1));
  const int64_t tau_hi = min(Tp - 1, q_hi / hop);
  const int nframes = (int)(tau_hi - tau_lo + 1);

  for (int j = tid; j < NFFT; j += kGroupsPerCta) {
    float s, c;
    sincospif(2.0f * (float)j / (float)NFFT, &s, &c);
    ctab[j] = c;
    stab[j] = s;
    win[j] = 0.5f - 0.5f * c;
  }
  __syncthreads();

  const float* zb = z + b * Tp_all * zpitch;
  for (int fi = tid; fi < nframes; fi += kGroupsPerCta) {
    const float* zr = zb + (tau_lo + fi) * zpitch;
    float re[NB], im[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const float mag = expf(zr[k]);
      float s, c;
      sincosf(sinf(zr[NB + k]), &s, &c);
      re[k] = mag * c;
      im[k] = mag * s;
    }
#pragma unroll
    for (int n = 0; n < NFFT; ++n) {
      // C2R: imaginary parts of DC and Nyquist are ignored
      float acc = re[0] + ((n & 1) ? -re[NB - 1] : re[NB - 1]);
#pragma unroll
      for (int k = 1; k < NB - 1; ++k) {
        const int j = (k * n) % NFFT;
        acc += 2.0f * (re[k] * ctab[j] - im[k] * stab[j]);
      }
      frames[fi * FS + n] = acc * (1.0f / NFFT) * win[n];
    }
  }
  __syncthreads();

  const int64_t g = g0 + tid;
  if (g >= n_groups) return;
  float* ob = out + b * (Tp_all - 1) * hop + g * hop;
  for (int i = 0; i < hop; ++i) {
    const int64_t q = g * hop + i + NFFT / 2;
    const int64_t t_hi = min(Tp - 1, q / hop);
    const int64_t t_lo = max((int64_t)0, q / hop - (R - 1));
    float acc = 0.0f, env = 0.0f;
    for (int64_t tau = t_lo; tau <= t_hi; ++tau) {
      const int n = (int)(q - tau * hop);
      acc += frames[(tau - tau_lo) * FS + n];
      env += win[n] * win[n];
    }
    ob[i] = acc / env;
  }
}

// Backward of the head: one thread per (utterance, frame).  d frame[m] = win[m] * g[q - NFFT/2] / env(q) at q = hop*tau + m,
// d Re_k = c_k/N sum_m d frame[m] cos(2 pi k m / N), d Im_k = -c_k/N sum_m d frame[m] sin(2 pi k m / N)  (c_0 = c_N/2 = 1 and
// their imaginary parts get no gradient: C2R ignores them), then through mag = exp(z_k), theta = sin(z_{NB+k}):
//   dz_k = dRe * re + dIm * im,   dz_{NB+k} = (dIm * re - dRe * im) * cos(z_{NB+k}).
// dz rows have `pitch` >= NFFT + 2 floats; the columns above NFFT + 2 are written as zeros.
template <int NFFT>
__global__ void __launch_bounds__(128) istft_head_bwd_kernel(const float* __restrict__ z, const float* __restrict__ gout,
                                                             float* __restrict__ dz, int64_t Tp, int hop, int pitch) {
  constexpr int NB = NFFT / 2 + 1;
  constexpr int CH = NFFT + 2;
  __shared__ float ctab[NFFT], stab[NFFT], win[NFFT];
  for (int j = threadIdx.x; j < NFFT; j += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)j / (float)NFFT, &s, &c);
    ctab[j] = c;
    stab[j] = s;
    win[j] = 0.5f - 0.5f * c;
  }
  __syncthreads();
  const int64_t b = blockIdx.y;
  const int64_t tau = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tau >= Tp) return;
  const int R = NFFT / hop;
  const int64_t n_out = (Tp - 1) * hop;
  float df[NFFT];
#pragma unroll
  for (int m = 0; m < NFFT; ++m) {
    const int64_t q = tau * hop + m, o = q - NFFT / 2;
    float v = 0.0f;
    if (o >= 0 && o < n_out) {
      const int64_t t_hi = min(Tp - 1, q / hop), t_lo = max((int64_t)0, q / hop - (R - 1));
      float env = 0.0f;
      for (int64_t t = t_lo; t <= t_hi; ++t) {
        const float w = win[(int)(q - t * hop)];
        env += w * w;
      }
      v = gout[b * n_out + o] / env * win[m];
    }
    df[m] = v;
  }
  const float* zr = z + (b * Tp + tau) * CH;
  float* dr = dz + (b * Tp + tau) * pitch;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    float dre = 0.0f, dim = 0.0f;
#pragma unroll
    for (int m = 0; m < NFFT; ++m) {
      const int j = (k * m) % NFFT;
      dre = fmaf(df[m], ctab[j], dre);
      dim = fmaf(df[m], -stab[j], dim);
    }
    const bool edge = (k == 0 || k == NB - 1);
    const float c = (edge ? 1.0f : 2.0f) / NFFT;
    dre *= c;
    dim = edge ? 0.0f : dim * c;
    const float zp = zr[NB + k];
    const float mag = expf(zr[k]);
    float s, co;
    sincosf(sinf(zp), &s, &co);
    const float re = mag * co, im = mag * s;
    dr[k] = dre * re + dim * im;
    dr[NB + k] = (dim * re - dre * im) * cosf(zp);
  }
  for (int c = CH; c < pitch; ++c) dr[c] = 0.0f;
}

// ReflectionPad1d((1, 0)) over channels-last rows and its adjoint (istftnet.py:296,312)
__global__ void __launch_bounds__(256) pad_reflect_left_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t B, int64_t T, int C) {
  const int64_t n = B * (T + 1) * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const int64_t t = (e / C) % (T + 1), b = e / ((int64_t)C * (T + 1));
    y[e] = x[(b * T + (t == 0 ? 1 : t - 1)) * C + c];
  }
}
__global__ void __launch_bounds__(256) unpad_reflect_left_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t B, int64_t T, int C) {
  const int64_t n = B * T * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const int64_t t = (e / C) % T, b = e / ((int64_t)C * T);
    float v = dy[(b * (T + 1) + t + 1) * C + c];
    if (t == 1) v += dy[(b * (T + 1)) * C + c];
    dx[e] = v;
  }
}

template <int NFFT>
int launch_istft(const float* z, float* out, int64_t B, int64_t Tp, int hop, const RowLens& lens, int zpitch, cudaStream_t st) {
  const int R = NFFT / hop;
  const int max_frames = kGroupsPerCta + R;
  const size_t smem = sizeof(float) * (3 * NFFT + (size_t)max_frames * (NFFT + 1));
  NVSE_CUDA_CHECK(cudaFuncSetAttribute(istft_head_kernel<NFFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((Tp - 1 + kGroupsPerCta - 1) / kGroupsPerCta), (unsigned)B);
  ProfScope prof("istft_head", NFFT + 2, 1, 0.0, 4.0 * (double)B * ((double)Tp * (NFFT + 2) + (double)(Tp - 1) * hop), st);
  istft_head_kernel<NFFT><<<grid, kGroupsPerCta, smem, st>>>(z, out, Tp, hop, lens, zpitch > 0 ? zpitch : NFFT + 2);
  NVSE_LAUNCH_CHECK("istft_head_kernel");
  return NVSE_OK;
}

}  // namespace

int launch_istft_head(const float* z, float* out, int64_t B, int64_t Tp, int n_fft, int hop, cudaStream_t st, RowLens lens, int zpitch) {
  NVSE_REQUIRE(zpitch == 0 || zpitch >= n_fft + 2, NVSE_ERR_INVALID, "istft head: row pitch %d below n_fft + 2", zpitch);
  NVSE_REQUIRE(hop >= 1 && n_fft % hop == 0, NVSE_ERR_UNSUPPORTED, "istft head: hop=%d must divide n_fft=%d", hop, n_fft);
  NVSE_REQUIRE(B <= 65535, NVSE_ERR_INVALID, "istft head: batch too large");
  if (B == 0 || Tp <= 1) return NVSE_OK;
  switch (n_fft) {
    case 4: return launch_istft<4>(z, out, B, Tp, hop, lens, zpitch, st);
    case 8: return launch_istft<8>(z, out, B, Tp, hop, lens, zpitch, st);
    case 16: return launch_istft<16>(z, out, B, Tp, hop, lens, zpitch, st);
    case 32: return launch_istft<32>(z, out, B, Tp, hop, lens, zpitch, st);
    case 64: return launch_istft<64>(z, out, B, Tp, hop, lens, zpitch, st);
    default:
      return fail(NVSE_ERR_UNSUPPORTED, "istft head: n_fft=%d not supported (4, 8, 16, 32, 64)", n_fft);
  }
}

int launch_istft_head_bwd(const float* z, const float* gout, float* dz, int64_t B, int64_t Tp, int n_fft, int hop, int pitch,
                          cudaStream_t st) {
  NVSE_REQUIRE(hop >= 1 && n_fft % hop == 0 && pitch >= n_fft + 2, NVSE_ERR_UNSUPPORTED, "istft head backward: bad n_fft / hop / pitch");
  NVSE_REQUIRE(B <= 65535, NVSE_ERR_INVALID, "istft head backward: batch too large");
  if (B == 0 || Tp < 1) return NVSE_OK;
  dim3 grid((unsigned)((Tp + 127) / 128), (unsigned)B);
  switch (n_fft) {
    case 4: istft_head_bwd_kernel<4><<<grid, 128, 0, st>>>(z, gout, dz, Tp, hop, pitch); break;
    case 8: istft_head_bwd_kernel<8><<<grid, 128, 0, st>>>(z, gout, dz, Tp, hop, pitch); break;
    case 16: istft_head_bwd_kernel<16><<<grid, 128, 0, st>>>(z, gout, dz, Tp, hop, pitch); break;
    case 32: istft_head_bwd_kernel<32><<<grid, 128, 0, st>>>(z, gout, dz, Tp, hop, pitch); break;
    default:
      return fail(NVSE_ERR_UNSUPPORTED, "istft head backward: n_fft=%d not supported (4, 8, 16, 32)", n_fft);
  }
  NVSE_LAUNCH_CHECK("istft_head_bwd_kernel");
  return NVSE_OK;
}

int launch_pad_reflect_left(const float* x, float* y, int64_t B, int64_t T, int C, cudaStream_t st) {
  NVSE_REQUIRE(T >= 2, NVSE_ERR_INVALID, "ReflectionPad1d((1, 0)) needs at least two rows");
  const int64_t n = B * (T + 1) * C;
  pad_reflect_left_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 8192), 256, 0, st>>>(x, y, B, T, C);
  NVSE_LAUNCH_CHECK("pad_reflect_left_kernel");
  return NVSE_OK;
}

int launch_unpad_reflect_left(const float* dy, float* dx, int64_t B, int64_t T, int C, cudaStream_t st) {
  const int64_t n = B * T * C;
  unpad_reflect_left_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 8192), 256, 0, st>>>(dy, dx, B, T, C);
  NVSE_LAUNCH_CHECK("unpad_reflect_left_kernel");
  return NVSE_OK;
}

}  // namespace nvse

extern "C" int nvse_istft_head_backward_f32(const float* z, const float* dout, float* dz, int64_t B, int64_t Tp, int n_fft,
                                            int hop, void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(z && dout && dz && B >= 0 && Tp >= 1, NVSE_ERR_INVALID, "nvse_istft_head_backward_f32: bad argument");
  return launch_istft_head_bwd(z, dout, dz, B, Tp, n_fft, hop, n_fft + 2, as_stream(stream));
}

extern "C" int nvse_istft_head_f32(const float* z, float* out, int64_t B, int64_t Tp, int n_fft, int hop, void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(z && out && B >= 0 && Tp >= 1, NVSE_ERR_INVALID, "nvse_istft_head_f32: bad argument");
  return launch_istft_head(z, out, B, Tp, n_fft, hop, as_stream(stream), RowLens{nullptr, 1, 0}, 0);
}
