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

#include <cmath>

namespace nvse {

namespace {

constexpr int kGroupsPerCta = 128;  // hop-groups (of `hop` output samples) per CTA

template <int NFFT>
__global__ void __launch_bounds__(kGroupsPerCta) istft_head_kernel(const float* __restrict__ z, float* __restrict__ out,
                                                                    int64_t Tp, int hop) {
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

  const float* zb = z + b * Tp * CH;
  for (int fi = tid; fi < nframes; fi += kGroupsPerCta) {
    const float* zr = zb + (tau_lo + fi) * CH;
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
  float* ob = out + b * n_groups * hop + g * hop;
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

template <int NFFT>
int launch_istft(const float* z, float* out, int64_t B, int64_t Tp, int hop, cudaStream_t st) {
  const int R = NFFT / hop;
  const int max_frames = kGroupsPerCta + R;
  const size_t smem = sizeof(float) * (3 * NFFT + (size_t)max_frames * (NFFT + 1));
  NVSE_CUDA_CHECK(cudaFuncSetAttribute(istft_head_kernel<NFFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((Tp - 1 + kGroupsPerCta - 1) / kGroupsPerCta), (unsigned)B);
  ProfScope prof("istft_head", NFFT + 2, 1, 0.0, 4.0 * (double)B * ((double)Tp * (NFFT + 2) + (double)(Tp - 1) * hop), st);
  istft_head_kernel<NFFT><<<grid, kGroupsPerCta, smem, st>>>(z, out, Tp, hop);
  NVSE_LAUNCH_CHECK("istft_head_kernel");
  return NVSE_OK;
}

}  // namespace

int launch_istft_head(const float* z, float* out, int64_t B, int64_t Tp, int n_fft, int hop, cudaStream_t st) {
  NVSE_REQUIRE(hop >= 1 && n_fft % hop == 0, NVSE_ERR_UNSUPPORTED, "istft head: hop=%d must divide n_fft=%d", hop, n_fft);
  NVSE_REQUIRE(B <= 65535, NVSE_ERR_INVALID, "istft head: batch too large");
  if (B == 0 || Tp <= 1) return NVSE_OK;
  switch (n_fft) {
    case 4: return launch_istft<4>(z, out, B, Tp, hop, st);
    case 8: return launch_istft<8>(z, out, B, Tp, hop, st);
    case 16: return launch_istft<16>(z, out, B, Tp, hop, st);
    case 32: return launch_istft<32>(z, out, B, Tp, hop, st);
    case 64: return launch_istft<64>(z, out, B, Tp, hop, st);
    default:
      return fail(NVSE_ERR_UNSUPPORTED, "istft head: n_fft=%d not supported (4, 8, 16, 32, 64)", n_fft);
  }
}

}  // namespace nvse

extern "C" int nvse_istft_head_f32(const float* z, float* out, int64_t B, int64_t Tp, int n_fft, int hop, void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(z && out && B >= 0 && Tp >= 1, NVSE_ERR_INVALID, "nvse_istft_head_f32: bad argument");
  return launch_istft_head(z, out, B, Tp, n_fft, hop, as_stream(stream));
}
