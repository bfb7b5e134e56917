// Fused log-mel front-end: dataset.mel_spectrogram (reference dataset.py:53-91).
//
//   reflect-pad | frame | x Hann | 1024-pt real FFT | magnitude | mel projection | log-clamp
//
// One warp transforms TWO consecutive frames of one utterance at once: the two real frames
// are packed as z = a + i*b and pushed through ONE 1024-point complex FFT, factored
// 32 x 32 (Cooley-Tukey): a radix-32 DIF butterfly network entirely in registers over the
// strided samples each lane owns, a twiddle multiply, a 32x32 transpose through padded
// shared memory, and a second in-register radix-32 network.  The two spectra are separated
// with the Hermitian identities using warp shuffles (lane L pairs with lane 32-L), so no
// post-twiddle is needed.  Magnitudes go to shared memory and the (banded) mel basis is
// applied from a tap-major packed copy so that lanes read consecutive weights.
// Nothing but the waveform is read from HBM and nothing but the log-mel is written: the
// [B,513,F] complex / magnitude tensors of the reference never exist.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace nvse {

namespace {

constexpr int kNfft = 1024;
constexpr int kBins = kNfft / 2 + 1;
// 4 warps per CTA and a register budget for 5 CTAs per SM (96 registers, a few spilled values): the kernel is
// issue-latency bound, and 20 resident warps instead of 16 at 128 registers measured 1.6x faster (same-box A/B)
#ifndef NVSE_FE_WARPS
#define NVSE_FE_WARPS 4
#endif
constexpr int kWarpsPerCta = NVSE_FE_WARPS;
constexpr int kTransposeStride = 33;                       // 32 + 1 pad: conflict-free both ways
constexpr int kWarpSmemFloats = 2 * 32 * kTransposeStride;  // re plane + im plane

__host__ __device__ constexpr int brev5(int v) {
  return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

// cos(pi*j/16), j = 0..8
__device__ __forceinline__ constexpr float c16(int j) {
  switch (j) {
    case 0: return 1.0f;
    case 1: return 0.98078528040323044913f;
    case 2: return 0.92387953251128675613f;
    case 3: return 0.83146961230254523708f;
    case 4: return 0.70710678118654752440f;
    case 5: return 0.55557023301960222474f;
    case 6: return 0.38268343236508977173f;
    case 7: return 0.19509032201612826785f;
    default: return 0.0f;
  }
}
__device__ __forceinline__ constexpr float cos32(int j) { return j <= 8 ? c16(j) : -c16(16 - j); }  // cos(2*pi*j/32), j<16
__device__ __forceinline__ constexpr float sin32(int j) { return j <= 8 ? c16(8 - j) : c16(j - 8); }  // sin(2*pi*j/32), j<16

// In-place radix-2 decimation-in-frequency FFT of 32 complex values held in registers.
// Forward transform (e^{-i...}); X[k] ends up at index brev5(k).  Everything is unrolled so
// all indices and twiddles are compile-time constants.
__device__ __forceinline__ void fft32_dif(float (&re)[32], float (&im)[32]) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
#pragma unroll
    for (int g = 0; g < 32; g += 2 * half) {
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const int a = g + i, b = g + i + half;
        const int tw = i * (16 / half);  // W_32^tw
        const float ar = re[a], ai = im[a], br = re[b], bi = im[b];
        re[a] = ar + br;
        im[a] = ai + bi;
        const float dr = ar - br, di = ai - bi;
        if (tw == 0) {
          re[b] = dr;
          im[b] = di;
        } else if (tw == 8) {  // multiply by -i
          re[b] = di;
          im[b] = -dr;
        } else {
          const float c = cos32(tw), s = sin32(tw);  // W = c - i s
          re[b] = dr * c + di * s;
          im[b] = di * c - dr * s;
        }
      }
    }
  }
}

struct FrontendParams {
  const float* y;
  int64_t y_stride;
  int64_t T;
  int64_t B;
  int64_t F;
  int64_t pairs;  // ceil(F / 2)
  int hop;
  int n_mels;
  const float* window;   // [1024]
  const float2* twiddle;  // [32][32]: twiddle[k1*32 + l] = exp(-2*pi*i*l*k1/1024)
  const float* wpack;     // [max_band][n_mels] tap-major banded mel weights
  const int* band_lo;     // [n_mels]
  const int* band_len;    // [n_mels]
  float* out;             // [B, n_mels, F]
};

__device__ __forceinline__ int64_t reflect_index(int64_t i, int64_t T) {
  if (i < 0) i = -i;
  if (i >= T) i = 2 * (T - 1) - i;
  return i;
}

#ifndef NVSE_FE_MINB
#define NVSE_FE_MINB 5
#endif
__global__ void __launch_bounds__(kWarpsPerCta * 32, NVSE_FE_MINB) mel_frontend_kernel(const FrontendParams p) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t task = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  if (task >= p.B * p.pairs) return;  // warp-uniform; only __syncwarp below
  float* sre = smem + warp * kWarpSmemFloats;
  float* sim = sre + 32 * kTransposeStride;

  const int64_t b = task / p.pairs;
  const int64_t f0 = 2 * (task % p.pairs);
  const bool has_b = (f0 + 1) < p.F;
  const float* __restrict__ yrow = p.y + b * p.y_stride;
  const int64_t sa = f0 * p.hop - kNfft / 2;  // first sample of frame a in un-padded coordinates
  const int64_t sb = sa + p.hop;

  float re[32], im[32];
  if (sa >= 0 && sb + kNfft <= p.T && has_b) {
#pragma unroll
    for (int m = 0; m < 32; ++m) {
      const int n = lane + 32 * m;
      const float w = __ldg(p.window + n);
      re[m] = __ldg(yrow + sa + n) * w;
      im[m] = __ldg(yrow + sb + n) * w;
    }
  } else {
#pragma unroll
    for (int m = 0; m < 32; ++m) {
      const int n = lane + 32 * m;
      const float w = __ldg(p.window + n);
      re[m] = __ldg(yrow + reflect_index(sa + n, p.T)) * w;
      im[m] = has_b ? __ldg(yrow + reflect_index(sb + n, p.T)) * w : 0.0f;
    }
  }

  // pass 1: 32-point DFTs over m (samples lane + 32 m)
  fft32_dif(re, im);

  // twiddle by W_1024^(lane*k1), then transpose so that lane k1 owns all 32 values of column k1
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {
    const int r = brev5(k1);
    const float2 t = __ldg(p.twiddle + k1 * 32 + lane);
    sre[k1 * kTransposeStride + lane] = re[r] * t.x - im[r] * t.y;
    sim[k1 * kTransposeStride + lane] = re[r] * t.y + im[r] * t.x;
  }
  __syncwarp();
#pragma unroll
  for (int l = 0; l < 32; ++l) {
    re[l] = sre[lane * kTransposeStride + l];
    im[l] = sim[lane * kTransposeStride + l];
  }
  __syncwarp();

  // pass 2: 32-point DFTs over l;  Z[lane + 32*k2] is now at register brev5(k2)
  fft32_dif(re, im);

  // Separate the two real spectra:  A[k] = (Z[k] + conj Z[N-k]) / 2,  B[k] = (Z[k] - conj Z[N-k]) / 2i.
  // For lane L != 0 the partner bin N-k lives in lane 32-L at k2' = 31-k2; lane 0 pairs with itself.
  float* mag_a = sre;  // [513], reuses the transpose planes (each plane holds 1056 floats)
  float* mag_b = sim;
  const int src_lane = (32 - lane) & 31;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const float zr = re[brev5(k2)], zi = im[brev5(k2)];
    float pr = __shfl_sync(0xffffffffu, re[brev5(31 - k2)], src_lane);
    float pi = __shfl_sync(0xffffffffu, im[brev5(31 - k2)], src_lane);
    if (lane == 0) {
      pr = re[brev5((32 - k2) & 31)];
      pi = im[brev5((32 - k2) & 31)];
    }
    const float ar = zr + pr, ai = zi - pi;
    const float br = zr - pr, bi = zi + pi;
    mag_a[lane + 32 * k2] = 0.5f * sqrtf(ar * ar + ai * ai);
    mag_b[lane + 32 * k2] = 0.5f * sqrtf(br * br + bi * bi);
  }
  if (lane == 0) {  // Nyquist bin: Z[512] = A[512] + i B[512], both real
    mag_a[512] = fabsf(re[brev5(16)]);
    mag_b[512] = fabsf(im[brev5(16)]);
  }
  __syncwarp();

  // mel projection (banded) + log-clamp
  for (int m = lane; m < p.n_mels; m += 32) {
    const int lo = __ldg(p.band_lo + m), len = __ldg(p.band_len + m);
    float acc_a = 0.0f, acc_b = 0.0f;
    for (int t = 0; t < len; ++t) {
      const float w = __ldg(p.wpack + (int64_t)t * p.n_mels + m);
      acc_a = fmaf(w, mag_a[lo + t], acc_a);
      acc_b = fmaf(w, mag_b[lo + t], acc_b);
    }
    float* o = p.out + (b * p.n_mels + m) * p.F + f0;
    o[0] = logf(fmaxf(acc_a, 1e-5f));
    if (has_b) o[1] = logf(fmaxf(acc_b, 1e-5f));
  }
}


// ------------------------------------------------------------------------------------------------
// Forward kernel, second generation ("staged"): what nvse_frontend_mel_f32 launches.
//
// A persistent CTA of 4 warps walks over groups of 8 CONSECUTIVE frames of one utterance (one frame pair per warp):
//   * the 7 * hop + 1024 samples a group covers are staged in shared memory ONCE, by cp.async, while the previous
//     group is being transformed (the sample buffer is free again as soon as every warp holds its frames in
//     registers): frames overlap by 75 %, so every sample is read from L2/HBM once instead of four times and no warp
//     has a global load on its critical path.  Groups that touch the reflect padding are staged element by element;
//   * the packed 1024-point FFT is the one above, with the 32 x 32 transpose done on (re, im) pairs (64-bit shared
//     memory accesses, half the instructions);
//   * the factor 1/2 of the two-real-FFTs-in-one separation is folded into the window (exact: a power of two),
//     magnitudes use sqrt.approx (1 ulp; the mel sums are checked at 2e-5 relative), and only the bins that carry a
//     non-zero mel weight are separated at all (fmax 8000 at 22.05 kHz: 372 of 513 -> 12 of 16 passes);
//   * the banded mel projection runs CTA-wide, one thread per mel row over all 8 frames of the group (a weight is
//     loaded once per 8 frames; a per-warp projection iterates to the longest band of its lanes three times over);
//   * the 80 x 8 log-mel tile goes through shared memory so that every global store instruction writes runs of
//     8 consecutive frames (whole 32-byte sectors where the row is aligned) instead of 2.
// ------------------------------------------------------------------------------------------------
constexpr int kFe2Warps = 4;
constexpr int kFe2Threads = kFe2Warps * 32;
constexpr int kFe2Frames = 2 * kFe2Warps;
constexpr int kTr2 = 33;                   // float2 row pitch of the transpose tile: conflict-free for 64-bit accesses both ways
constexpr int kFe2Scratch = 32 * kTr2;     // float2 per warp: transpose tile, then the (|A|, |B|) magnitudes of its two frames
constexpr int kFe2OutOff = 520;            // float2 offset inside warp 0's scratch of the [n_mels][8] output tile (above its 513 magnitudes)
constexpr int kFe2MaxMels = (kFe2Scratch - kFe2OutOff) * 2 / kFe2Frames;  // 134

struct Frontend2Params {
  FrontendParams f;
  const float* window_half;  // 0.5 * window
  int64_t groups;            // ceil(F / 8) frame groups per utterance
  int64_t total;             // B * groups
  int nsmp;                  // staged samples per group: 7 * hop + 1024
  int nsmp_pad;              // rounded up to a multiple of 4 floats
  int nyquist;               // bin 512 carries mel weight
  const int4* items;         // mel projection work items: (mel row, first tap, taps, slices of the row | slice index << 8)
  int nitems;                // padded so that the slices of a row never straddle a warp
  const int* lens;           // ragged batch: samples per utterance (<= T; reflect padding and frame count follow it), or null
  int out_cl_pitch;          // > 0: store the log-mel CHANNELS-LAST, out [B, F, pitch] with channels n_mels .. pitch - 1 zero -- the
                             // layout conv_pre's tensor-core launch stages from (the fused wav -> wav call), no transpose pass
};

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// samples of group g -> ssmp, asynchronously: 16-byte copies where the group lies inside the utterance and is aligned,
// 4-byte copies through the reflect index otherwise
__device__ __forceinline__ void fe2_stage(const Frontend2Params& q, int64_t g, float* ssmp, int tid) {
  const FrontendParams& p = q.f;
  const int64_t b = g / q.groups;
  const int64_t f0 = (g - b * q.groups) * kFe2Frames;
  const int64_t Tb = q.lens ? (q.lens[b] < p.T ? q.lens[b] : p.T) : p.T, Fb = 1 + Tb / p.hop;  // this utterance's samples / frames
  const int nf = (int)((Fb - f0) < kFe2Frames ? (Fb - f0) : kFe2Frames);
  const float* __restrict__ yrow = p.y + b * p.y_stride;
  const int64_t s0 = f0 * p.hop - kNfft / 2;  // un-padded sample index of staged sample 0
  if (s0 >= 0 && s0 + q.nsmp <= Tb && ((reinterpret_cast<uintptr_t>(yrow + s0) & 15) == 0)) {
    for (int i = tid * 4; i + 3 < q.nsmp; i += kFe2Threads * 4) cp_async16(ssmp + i, yrow + s0 + i);
    for (int i = (q.nsmp & ~3) + tid; i < q.nsmp; i += kFe2Threads) ssmp[i] = __ldg(yrow + s0 + i);
  } else {
    const int need = (nf - 1) * p.hop + kNfft;  // samples the existing frames cover; the rest is never stored
    for (int i = tid; i < q.nsmp; i += kFe2Threads) {  // element-wise, still asynchronous
      if (i < need) cp_async4(ssmp + i, yrow + reflect_index(s0 + i, Tb));
      else ssmp[i] = 0.0f;
    }
  }
}

// K2N: passes of 32 bins that are separated (compile-time: the loop holds warp shuffles and register-indexed arrays)
template <int K2N>
__global__ void __launch_bounds__(kFe2Threads, 4) mel_frontend2_kernel(const Frontend2Params q) {
  extern __shared__ __align__(16) float smem[];
  const FrontendParams& p = q.f;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* ssmp = smem;
  float2* scr = reinterpret_cast<float2*>(smem + q.nsmp_pad);  // [4 warps][kFe2Scratch]
  float2* sz = scr + warp * kFe2Scratch;
  float* sout = reinterpret_cast<float*>(scr + kFe2OutOff);   // [n_mels][8], inside warp 0's scratch

  int64_t g = blockIdx.x;
  if (g < q.total) fe2_stage(q, g, ssmp, tid);
  for (; g < q.total; g += gridDim.x) {
    const int64_t b = g / q.groups;
    const int64_t f0 = (g - b * q.groups) * kFe2Frames;
    const int64_t Fb = q.lens ? 1 + (q.lens[b] < p.T ? q.lens[b] : p.T) / p.hop : p.F;
    const int nf = (int)((Fb - f0) < kFe2Frames ? (Fb - f0) : kFe2Frames);  // frames of this group that exist (<= 0: none stored)
    cp_async_wait_all();
    __syncthreads();  // samples of this group are in place; the previous group's output tile has been stored

    // Frame b of the last pair of an odd-length utterance does not exist: its samples are staged as zeros (or are
    // real samples further on), it is transformed like any other and never stored.  Pairs beyond the utterance
    // (fa >= nf) are transformed as well -- the warp would otherwise idle at the barriers below.
    const int fa = 2 * warp;
    float re[32], im[32];
    {
      const float* sa = ssmp + fa * p.hop + lane;
      const float* sb = sa + p.hop;
      const float* wh = q.window_half + lane;
#pragma unroll
      for (int m = 0; m < 32; ++m) {
        const float w = __ldg(wh + 32 * m);
        re[m] = sa[32 * m] * w;
        im[m] = sb[32 * m] * w;
      }
    }
    __syncthreads();  // every warp holds its frames: the sample buffer is free for the next group
    if (g + gridDim.x < q.total) fe2_stage(q, g + gridDim.x, ssmp, tid);

    fft32_dif(re, im);
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) {
      const int r = brev5(k1);
      const float2 t = __ldg(p.twiddle + k1 * 32 + lane);
      sz[k1 * kTr2 + lane] = make_float2(re[r] * t.x - im[r] * t.y, re[r] * t.y + im[r] * t.x);
    }
    __syncwarp();
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const float2 v = sz[lane * kTr2 + l];
      re[l] = v.x;
      im[l] = v.y;
    }
    __syncwarp();
    fft32_dif(re, im);  // Z[lane + 32*k2] (already halved through the window) is at register brev5(k2)

    // |A[k]|, |B[k]| for the bins that carry mel weight, as (a, b) pairs
    const int src_lane = (32 - lane) & 31;
#pragma unroll
    for (int k2 = 0; k2 < K2N; ++k2) {
      const float zr = re[brev5(k2)], zi = im[brev5(k2)];
      float pr = __shfl_sync(0xffffffffu, re[brev5(31 - k2)], src_lane);
      float pi = __shfl_sync(0xffffffffu, im[brev5(31 - k2)], src_lane);
      if (lane == 0) {
        pr = re[brev5((32 - k2) & 31)];
        pi = im[brev5((32 - k2) & 31)];
      }
      const float ar = zr + pr, ai = zi - pi;
      const float br = zr - pr, bi = zi + pi;
      sz[lane + 32 * k2] = make_float2(sqrt_approx(fmaf(ar, ar, ai * ai)), sqrt_approx(fmaf(br, br, bi * bi)));
    }
    if (q.nyquist && lane == 0) sz[512] = make_float2(2.0f * fabsf(re[brev5(16)]), 2.0f * fabsf(im[brev5(16)]));
    __syncthreads();  // the magnitudes of all 8 frames are in the four scratch tiles

    // Banded mel projection + log-clamp, all 8 frames at once.  The bands are 3 .. 27 bins long, so a row is cut into
    // 1, 2 or 4 slices of about equal length on adjacent lanes (work items, built at create time so that every
    // thread gets at most ~10 taps); the slices of a row are summed with two xor-shuffle rounds.
    for (int it = tid; it < q.nitems; it += kFe2Threads) {
      const int4 item = __ldg(q.items + it);
      const int m = item.x, cnt = item.z, nsl = item.w & 0xff, sl = item.w >> 8;
      float acc[kFe2Frames];
#pragma unroll
      for (int j = 0; j < kFe2Frames; ++j) acc[j] = 0.0f;
      if (m >= 0) {
        const float2* mg = scr + __ldg(p.band_lo + m) + item.y;
        const float* wp = p.wpack + (int64_t)item.y * p.n_mels + m;
        for (int t = 0; t < cnt; ++t) {
          const float w = __ldg(wp + (int64_t)t * p.n_mels);
#pragma unroll
          for (int wq = 0; wq < kFe2Warps; ++wq) {
            const float2 v = mg[wq * kFe2Scratch + t];
            acc[2 * wq] = fmaf(w, v.x, acc[2 * wq]);
            acc[2 * wq + 1] = fmaf(w, v.y, acc[2 * wq + 1]);
          }
        }
      }
#pragma unroll
      for (int off = 1; off <= 2; off <<= 1) {
#pragma unroll
        for (int j = 0; j < kFe2Frames; ++j) {
          const float o = __shfl_xor_sync(0xffffffffu, acc[j], off);
          if (nsl > off) acc[j] += o;
        }
      }
      if (m >= 0 && sl == 0) {
        float4* o = reinterpret_cast<float4*>(sout + m * kFe2Frames);
        o[0] = make_float4(__logf(fmaxf(acc[0], 1e-5f)), __logf(fmaxf(acc[1], 1e-5f)), __logf(fmaxf(acc[2], 1e-5f)), __logf(fmaxf(acc[3], 1e-5f)));
        o[1] = make_float4(__logf(fmaxf(acc[4], 1e-5f)), __logf(fmaxf(acc[5], 1e-5f)), __logf(fmaxf(acc[6], 1e-5f)), __logf(fmaxf(acc[7], 1e-5f)));
      }
    }
    __syncthreads();  // output tile complete

    // coalesced store of the [n_mels][8] tile: 8 consecutive lanes write 8 consecutive frames of one mel row
    if (q.out_cl_pitch > 0) {  // [frame][channel]: a frame's n_mels values (+ zero padding) are one contiguous run
      float* ob = p.out + (b * p.F + f0) * q.out_cl_pitch;
      for (int e = tid; e < kFe2Frames * q.out_cl_pitch; e += kFe2Threads) {
        const int j = e / q.out_cl_pitch, m = e - j * q.out_cl_pitch;
        if (j < nf) ob[e] = m < p.n_mels ? sout[m * kFe2Frames + j] : 0.0f;
      }
      continue;
    }
    float* obase = p.out + b * p.n_mels * p.F + f0;
    for (int e = tid; e < p.n_mels * kFe2Frames; e += kFe2Threads) {
      const int m = e >> 3, j = e & 7;
      if (j < nf) obase[(int64_t)m * p.F + j] = sout[e];
    }
  }
  cp_async_wait_all();
}

// ------------------------------------------------------------------------------------------------
// Backward of the front-end (the mel-L1 term of the generator loss back-propagates through
// mel_spectrogram(y_g_hat), train_time_wi_inv.py:173-179,231-235).  Per frame pair, one warp:
//   recompute the two spectra A, B (same packed FFT as the forward) -> |.| -> mel sums
//   d(mel sum) = dmel / sum  where sum >= 1e-5 (log + clamp, dataset.py:27-28), else 0
//   d|X[k]|   = sum_m basis[m][k] d(mel sum)[m];   G[k] = d|X[k]| * X[k] / |X[k]|   (0 where |X| = 0)
//   d frame[n] = w[n] * Re sum_{k=0}^{N/2} G[k] e^{+2 pi i k n / N}
// The last line for both frames is ONE more packed 1024-point FFT: with H the Hermitian extension of G
// (H[k] = G[k]/2, H[N-k] = conj(G[k])/2, H[0] = Re G[0], H[N/2] = Re G[N/2]),  ya + i yb = conj(FFT(conj(Ha + i Hb))).
// Frame gradients go to a scratch [B, F, 1024]; mel_overlap_add_kernel sums the (up to 4) overlapping
// frames and the reflected padding into dy in a fixed order (bit-reproducible, no atomics).
// ------------------------------------------------------------------------------------------------
struct FrontendBwdParams {
  FrontendParams f;
  const float* dmel;    // [B, n_mels, F]
  const int* bin_mlo;   // [513] first / last mel filter with a non-zero weight at the bin
  const int* bin_mhi;
  float* frames;        // scratch [B, F, 1024]
};

__device__ __forceinline__ void fft1024_warp(float (&re)[32], float (&im)[32], float* sre, float* sim,
                                             const float2* __restrict__ twiddle, int lane) {
  fft32_dif(re, im);
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) {
    const int r = brev5(k1);
    const float2 t = __ldg(twiddle + k1 * 32 + lane);
    sre[k1 * kTransposeStride + lane] = re[r] * t.x - im[r] * t.y;
    sim[k1 * kTransposeStride + lane] = re[r] * t.y + im[r] * t.x;
  }
  __syncwarp();
#pragma unroll
  for (int l = 0; l < 32; ++l) {
    re[l] = sre[lane * kTransposeStride + l];
    im[l] = sim[lane * kTransposeStride + l];
  }
  __syncwarp();
  fft32_dif(re, im);  // X[lane + 32*k2] is now at register brev5(k2)
}

constexpr int kBwdWarpSmemFloats = kWarpSmemFloats + 2 * 256;  // + d(mel sum) of both frames

__global__ void __launch_bounds__(kWarpsPerCta * 32) mel_frontend_bwd_kernel(const FrontendBwdParams q) {
  extern __shared__ float smem[];
  const FrontendParams& p = q.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t task = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  if (task >= p.B * p.pairs) return;
  float* sre = smem + warp * kBwdWarpSmemFloats;
  float* sim = sre + 32 * kTransposeStride;
  float* dm_a = sre + kWarpSmemFloats;
  float* dm_b = dm_a + 256;

  const int64_t b = task / p.pairs;
  const int64_t f0 = 2 * (task % p.pairs);
  const bool has_b = (f0 + 1) < p.F;
  const float* __restrict__ yrow = p.y + b * p.y_stride;
  const int64_t sa = f0 * p.hop - kNfft / 2;
  const int64_t sb = sa + p.hop;

  float re[32], im[32];
#pragma unroll
  for (int m = 0; m < 32; ++m) {
    const int n = lane + 32 * m;
    const float w = __ldg(p.window + n);
    re[m] = __ldg(yrow + reflect_index(sa + n, p.T)) * w;
    im[m] = has_b ? __ldg(yrow + reflect_index(sb + n, p.T)) * w : 0.0f;
  }
  fft1024_warp(re, im, sre, sim, p.twiddle, lane);

  // the two spectra (bins lane + 32*k2, k2 < 16; the Nyquist bin on lane 0) and their magnitudes
  float a_re[16], a_im[16], b_re[16], b_im[16];
  float* mag_a = sre;
  float* mag_b = sim;
  const int src_lane = (32 - lane) & 31;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const float zr = re[brev5(k2)], zi = im[brev5(k2)];
    float pr = __shfl_sync(0xffffffffu, re[brev5(31 - k2)], src_lane);
    float pi = __shfl_sync(0xffffffffu, im[brev5(31 - k2)], src_lane);
    if (lane == 0) {
      pr = re[brev5((32 - k2) & 31)];
      pi = im[brev5((32 - k2) & 31)];
    }
    a_re[k2] = 0.5f * (zr + pr); a_im[k2] = 0.5f * (zi - pi);   // A = (Z[k] + conj Z[N-k]) / 2
    b_re[k2] = 0.5f * (zi + pi); b_im[k2] = -0.5f * (zr - pr);  // B = (Z[k] - conj Z[N-k]) / 2i
    mag_a[lane + 32 * k2] = sqrtf(a_re[k2] * a_re[k2] + a_im[k2] * a_im[k2]);
    mag_b[lane + 32 * k2] = sqrtf(b_re[k2] * b_re[k2] + b_im[k2] * b_im[k2]);
  }
  const float a_ny = re[brev5(16)], b_ny = im[brev5(16)];  // meaningful on lane 0 only
  if (lane == 0) {
    mag_a[512] = fabsf(a_ny);
    mag_b[512] = fabsf(b_ny);
  }
  __syncwarp();

  // mel sums and their gradients
  for (int m = lane; m < p.n_mels; m += 32) {
    const int lo = __ldg(p.band_lo + m), len = __ldg(p.band_len + m);
    float acc_a = 0.0f, acc_b = 0.0f;
    for (int t = 0; t < len; ++t) {
      const float w = __ldg(p.wpack + (int64_t)t * p.n_mels + m);
      acc_a = fmaf(w, mag_a[lo + t], acc_a);
      acc_b = fmaf(w, mag_b[lo + t], acc_b);
    }
    const float* g = q.dmel + (b * p.n_mels + m) * p.F + f0;
    dm_a[m] = acc_a >= 1e-5f ? g[0] / acc_a : 0.0f;
    dm_b[m] = (has_b && acc_b >= 1e-5f) ? g[1] / acc_b : 0.0f;
  }
  __syncwarp();

  // d|X[k]| / |X[k]|, in place over the magnitudes
  for (int k = lane; k < kBins; k += 32) {
    float da = 0.0f, db = 0.0f;
    const int mlo = __ldg(q.bin_mlo + k), mhi = __ldg(q.bin_mhi + k);
    for (int m = mlo; m <= mhi; ++m) {
      const int t = k - __ldg(p.band_lo + m);
      if (t >= 0 && t < __ldg(p.band_len + m)) {
        const float w = __ldg(p.wpack + (int64_t)t * p.n_mels + m);
        da = fmaf(w, dm_a[m], da);
        db = fmaf(w, dm_b[m], db);
      }
    }
    const float ma = mag_a[k], mb = mag_b[k];
    mag_a[k] = ma > 0.0f ? da / ma : 0.0f;
    mag_b[k] = mb > 0.0f ? db / mb : 0.0f;
  }
  __syncwarp();

  // G = ratio * X;  scale by 1/2 for the Hermitian extension (bins 0 and N/2 keep their full real part)
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    const float ra = mag_a[lane + 32 * k2], rb = mag_b[lane + 32 * k2];
    const float h = (k2 == 0 && lane == 0) ? 1.0f : 0.5f;
    a_re[k2] *= ra * h; a_im[k2] *= ra * h;
    b_re[k2] *= rb * h; b_im[k2] *= rb * h;
  }
  const float ga_ny = lane == 0 ? a_ny * mag_a[512] : 0.0f, gb_ny = lane == 0 ? b_ny * mag_b[512] : 0.0f;
  __syncwarp();  // the magnitude planes are reused by the transpose below

  // input of the second FFT: conj(Ha + i Hb) at index lane + 32*m in register m
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    if (m == 0 && lane == 0) {  // DC: Re Ga + i Re Gb  (imaginary parts are exactly zero)
      re[m] = a_re[0];
      im[m] = -b_re[0];
    } else {
      re[m] = a_re[m] - b_im[m];
      im[m] = -(a_im[m] + b_re[m]);
    }
  }
#pragma unroll
  for (int m = 16; m < 32; ++m) {  // index N - k: conj(Ga[k])/2 + i conj(Gb[k])/2, G[k] owned by lane 32 - L at k2 = 31 - m
    float gar = __shfl_sync(0xffffffffu, a_re[31 - m], src_lane);
    float gai = __shfl_sync(0xffffffffu, a_im[31 - m], src_lane);
    float gbr = __shfl_sync(0xffffffffu, b_re[31 - m], src_lane);
    float gbi = __shfl_sync(0xffffffffu, b_im[31 - m], src_lane);
    if (lane == 0) {  // lane 0 pairs with itself: k2 = 32 - m (m = 16: the Nyquist bin, real)
      if (m == 16) {
        gar = ga_ny; gai = 0.0f; gbr = gb_ny; gbi = 0.0f;
      } else {
        gar = a_re[(32 - m) & 15]; gai = a_im[(32 - m) & 15]; gbr = b_re[(32 - m) & 15]; gbi = b_im[(32 - m) & 15];
      }
    }
    re[m] = gar + gbi;
    im[m] = -(gbr - gai);
  }
  fft1024_warp(re, im, sre, sim, p.twiddle, lane);

  float* fa = q.frames + (b * p.F + f0) * kNfft;
#pragma unroll
  for (int k2 = 0; k2 < 32; ++k2) {
    const int n = lane + 32 * k2;
    const float w = __ldg(p.window + n);
    fa[n] = w * re[brev5(k2)];
    if (has_b) fa[kNfft + n] = -w * im[brev5(k2)];
  }
}

// ------------------------------------------------------------------------------------------------
// dataset.amp_pha_specturm (reference dataset.py:124-139): the same packed STFT, with the complex spectrum
// written out as log-amplitude, phase, real and imaginary planes [B, 513, F] instead of being reduced to mels
// (the T-F vocoders' analysis front-end and STFT-consistency loss).
// ------------------------------------------------------------------------------------------------
struct StftParams {
  FrontendParams f;
  float* log_amp;  // each [B, 513, F] or null
  float* phase;
  float* real;
  float* imag;
};

// One CTA = kWarpsPerCta frame pairs = 8 CONSECUTIVE frames of one utterance: the planes of the 8 frames are collected
// in a shared tile [plane][bin][frame] and written out with 8 consecutive frames (32 B) per (plane, bin), so every store
// fills whole sectors -- a lane writing its own bins directly would touch one float per 32-byte sector.  Two passes of
// two planes each (log-amplitude + phase, then real + imaginary) over a tile that aliases the FFT scratch keep the CTA at
// ~37 KB of shared memory (5-6 CTAs per SM): the kernel is latency-bound like the mel front-end.
constexpr int kStftFrames = 2 * kWarpsPerCta;
constexpr int kStftPitch = kStftFrames + 1;                 // conflict-free: lanes write consecutive bins
constexpr int kStftTileFloats = 2 * kBins * kStftPitch;
constexpr int kStftSmemFloats = kStftTileFloats > kWarpSmemFloats * kWarpsPerCta ? kStftTileFloats : kWarpSmemFloats * kWarpsPerCta;

__global__ void __launch_bounds__(kWarpsPerCta * 32, 4) stft_amp_pha_kernel(const StftParams q, int groups_per_utt) {
  extern __shared__ float smem[];
  const FrontendParams& p = q.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile = smem;                                   // [2][kBins][kStftPitch], aliases the FFT scratch
  float* sre = smem + warp * kWarpSmemFloats;
  float* sim = sre + 32 * kTransposeStride;
  const int64_t b = blockIdx.x / groups_per_utt;
  const int64_t f_base = (int64_t)(blockIdx.x % groups_per_utt) * kStftFrames;
  const int64_t f0 = f_base + 2 * warp;
  const bool active = f0 < p.F;  // warp-uniform
  float re[32], im[32];
  if (active) {
    const bool has_b = (f0 + 1) < p.F;
    const float* __restrict__ yrow = p.y + b * p.y_stride;
    const int64_t sa = f0 * p.hop - kNfft / 2;
    const int64_t sb = sa + p.hop;
#pragma unroll
    for (int m = 0; m < 32; ++m) {
      const int n = lane + 32 * m;
      const float w = __ldg(p.window + n);
      re[m] = __ldg(yrow + reflect_index(sa + n, p.T)) * w;
      im[m] = has_b ? __ldg(yrow + reflect_index(sb + n, p.T)) * w : 0.0f;
    }
    fft1024_warp(re, im, sre, sim, p.twiddle, lane);
  }
  const int src_lane = (32 - lane) & 31;
  const int nf = (int)((p.F - f_base) < kStftFrames ? (p.F - f_base) : kStftFrames);
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    float* out0 = pass == 0 ? q.log_amp : q.real;
    float* out1 = pass == 0 ? q.phase : q.imag;
    if (!out0 && !out1) continue;  // uniform over the CTA
    __syncthreads();  // the scratch (pass 0) / the previous pass's tile is no longer read
    if (active) {
      auto emit = [&](int k, float xr, float xi, int frame) {
        float* t = tile + k * kStftPitch + 2 * warp + frame;
        if (pass == 0) {
          t[0] = logf(sqrtf(xr * xr + xi * xi) + 1e-7f);
          t[kBins * kStftPitch] = atan2f(xi, xr);
        } else {
          t[0] = xr;
          t[kBins * kStftPitch] = xi;
        }
      };
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        const float zr = re[brev5(k2)], zi = im[brev5(k2)];
        float pr = __shfl_sync(0xffffffffu, re[brev5(31 - k2)], src_lane);
        float pi = __shfl_sync(0xffffffffu, im[brev5(31 - k2)], src_lane);
        if (lane == 0) {
          pr = re[brev5((32 - k2) & 31)];
          pi = im[brev5((32 - k2) & 31)];
        }
        const int k = lane + 32 * k2;
        emit(k, 0.5f * (zr + pr), 0.5f * (zi - pi), 0);     // A = (Z[k] + conj Z[N-k]) / 2
        emit(k, 0.5f * (zi + pi), -0.5f * (zr - pr), 1);    // B = (Z[k] - conj Z[N-k]) / 2i
      }
      if (lane == 0) {  // Nyquist bin: Z[512] = A[512] + i B[512], both real
        emit(512, re[brev5(16)], 0.0f, 0);
        emit(512, im[brev5(16)], 0.0f, 1);
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kBins * kStftFrames; e += kWarpsPerCta * 32) {
      const int k = e / kStftFrames, fr = e % kStftFrames;
      if (fr < nf) {
        const int64_t o = (b * kBins + k) * p.F + f_base + fr;
        if (out0) out0[o] = tile[k * kStftPitch + fr];
        if (out1) out1[o] = tile[kBins * kStftPitch + k * kStftPitch + fr];
      }
    }
  }
}

// dataset.inverse_mel (reference dataset.py:94-121): out[b, k, f] = sum_m inv_basis[k, m] * exp(mel[b, m, f]).
// A small GEMM: one CTA computes 128 bins x 128 frames, each thread 8 bins x (4 + 4) frames (four LDS.128 per 64 FMAs; a
// warp's stores cover 64 consecutive frames of two bins); exp(mel) is applied once while staging.
constexpr int kInvBins = 128, kInvFrames = 128, kInvMelChunk = 16;
static_assert(kInvBins == kInvFrames, "one fetch loop stages both operands");
__global__ void __launch_bounds__(256) inverse_mel_kernel(const float* __restrict__ inv_basis, const float* __restrict__ mel,
                                                          float* __restrict__ out, int n_bins, int n_mels, int64_t F) {
  __shared__ __align__(16) float Es[kInvMelChunk][kInvFrames];
  __shared__ __align__(16) float Ws[kInvMelChunk][kInvBins + 4];  // +4: the transposing stores spread over the banks
  const int64_t b = blockIdx.z, f0 = (int64_t)blockIdx.x * kInvFrames;
  const int k0 = blockIdx.y * kInvBins;
  const int fx = threadIdx.x & 15, ky = threadIdx.x >> 4;  // 16 x 16 threads
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  // software pipeline: the next chunk's global loads are in flight (in registers) while this chunk is multiplied
  constexpr int kPerThread = kInvMelChunk * kInvFrames / 256;  // = kInvMelChunk * kInvBins / 256
  float pe[kPerThread], pw[kPerThread];
  auto fetch = [&](int m0) {
#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
      const int e = threadIdx.x + u * 256;
      const int m = e / kInvFrames, f = e % kInvFrames;
      pe[u] = (m0 + m < n_mels && f0 + f < F) ? mel[(b * n_mels + m0 + m) * F + f0 + f] : -INFINITY;  // exp(-inf) = 0
      const int k = e / kInvMelChunk, mw = e % kInvMelChunk;  // consecutive threads read consecutive m of a basis row
      pw[u] = (m0 + mw < n_mels && k0 + k < n_bins) ? __ldg(inv_basis + (int64_t)(k0 + k) * n_mels + m0 + mw) : 0.0f;
    }
  };
  fetch(0);
  for (int m0 = 0; m0 < n_mels; m0 += kInvMelChunk) {
#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
      const int e = threadIdx.x + u * 256;
      Es[e / kInvFrames][e % kInvFrames] = expf(pe[u]);
      Ws[e % kInvMelChunk][e / kInvMelChunk] = pw[u];
    }
    __syncthreads();
    if (m0 + kInvMelChunk < n_mels) fetch(m0 + kInvMelChunk);
#pragma unroll
    for (int m = 0; m < kInvMelChunk; ++m) {
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[m][ky * 8]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[m][ky * 8 + 4]);
      const float4 e0 = *reinterpret_cast<const float4*>(&Es[m][fx * 4]);
      const float4 e1 = *reinterpret_cast<const float4*>(&Es[m][64 + fx * 4]);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wv[i], ev[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + ky * 8 + i;
    if (k >= n_bins) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t f = f0 + (j < 4 ? fx * 4 + j : 64 + fx * 4 + (j - 4));
      if (f < F) out[(b * n_bins + k) * F + f] = acc[i][j];
    }
  }
}

// dy[b][t] = sum over the padded positions that read sample t (itself and its reflections) of the frames covering them
__global__ void __launch_bounds__(256) mel_overlap_add_kernel(const float* __restrict__ frames, float* __restrict__ dy, int64_t B,
                                                              int64_t T, int64_t F, int hop) {
  const int64_t n = B * T;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / T, t = e - b * T;
    const float* fb = frames + b * F * kNfft;
    int64_t qs[3];
    int nq = 0;
    qs[nq++] = t + kNfft / 2;
    if (t >= 1 && t <= kNfft / 2) qs[nq++] = kNfft / 2 - t;                              // left reflection: i = -t
    if (t <= T - 2 && 2 * (T - 1) - t <= T - 1 + kNfft / 2) qs[nq++] = 2 * (T - 1) - t + kNfft / 2;  // right reflection
    float acc = 0.0f;
    for (int i = 0; i < nq; ++i) {
      const int64_t qq = qs[i];
      int64_t f_lo = (qq - (kNfft - 1) + hop - 1) / hop;
      if (qq - (kNfft - 1) < 0) f_lo = 0;
      int64_t f_hi = qq / hop;
      if (f_hi > F - 1) f_hi = F - 1;
      for (int64_t f = f_lo; f <= f_hi; ++f) acc += fb[f * kNfft + (qq - f * hop)];
    }
    dy[e] = acc;
  }
}


// ------------------------------------------------------------------------------------------------
// Inverse STFT at n_fft = 1024, the head of the reference's T-F vocoders (Models/apnet.py:155, freeV.py:178,
// bsrnn.py:210: torch.istft(spec, n_fft, hop, win, window=hann, center=True)):
//   frame f:  s_f = irfft(spec[:, :, f]) (1/N scaling, imaginary parts of bins 0 and N/2 ignored),  * window
//   y[t]   =  sum_f s_f[t + N/2 - f hop] / sum_f window[t + N/2 - f hop]^2,   t in [0, hop (F - 1))
// Kernel 1: a CTA of 4 warps takes 8 consecutive frames; the [513 x 8] tile of the two planes is staged in shared
// memory (frames are the fastest axis of [B, 513, F], so 8 consecutive lanes read 32 contiguous bytes), each warp
// turns ONE frame pair into one packed 1024-point transform -- Z = A + i B with A, B the Hermitian extensions of the
// two spectra, s_a + i s_b = conj(FFT(conj Z)) / N -- and writes the two windowed frames to a scratch [B, F, 1024].
// Kernel 2 gathers the (at most ceil(N / hop)) overlapping frames of every output sample in frame order and divides
// by the window envelope: bit-reproducible, no atomics.
// ------------------------------------------------------------------------------------------------
struct IstftParams {
  const float* re;     // [B, 513, F]; with im == null: complex64 [B, 513, F] (interleaved re, im -- torch's layout)
  const float* im;
  int64_t B, F;
  int hop;
  const float* window;    // [1024]
  const float2* twiddle;  // forward twiddles of fft1024_warp
  float* frames;          // scratch [B, F, 1024]
  float* out;             // [B, hop * (F - 1)]
};

constexpr int kIsWarps = 4, kIsFrames = 2 * kIsWarps;

__global__ void __launch_bounds__(kIsWarps * 32) istft1024_frames_kernel(const IstftParams q) {
  extern __shared__ __align__(16) float smem[];
  float* tre = smem;                       // [513][8]
  float* tim = tre + kBins * kIsFrames;    // [513][8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* sre = tim + kBins * kIsFrames + warp * kWarpSmemFloats;
  float* sim = sre + 32 * kTransposeStride;
  const int64_t groups = (q.F + kIsFrames - 1) / kIsFrames;
  const int64_t b = blockIdx.x / groups;
  const int64_t f0 = (blockIdx.x - b * groups) * kIsFrames;
  const int nf = (int)((q.F - f0) < kIsFrames ? (q.F - f0) : kIsFrames);
  if (q.im) {
    const float* __restrict__ gre = q.re + b * kBins * q.F + f0;
    const float* __restrict__ gim = q.im + b * kBins * q.F + f0;
    for (int e = tid; e < kBins * kIsFrames; e += kIsWarps * 32) {
      const int k = e >> 3, j = e & 7;
      const bool in = j < nf;
      tre[e] = in ? __ldg(gre + (int64_t)k * q.F + j) : 0.0f;
      tim[e] = in ? __ldg(gim + (int64_t)k * q.F + j) : 0.0f;
    }
  } else {  // complex64 input: one 8-byte load per bin and frame, no de-interleaving pass before the kernel
    const float2* __restrict__ gc = reinterpret_cast<const float2*>(q.re) + b * kBins * q.F + f0;
    for (int e = tid; e < kBins * kIsFrames; e += kIsWarps * 32) {
      const int k = e >> 3, j = e & 7;
      const float2 v = j < nf ? __ldg(gc + (int64_t)k * q.F + j) : make_float2(0.0f, 0.0f);
      tre[e] = v.x;
      tim[e] = v.y;
    }
  }
  __syncthreads();
  const int fa = 2 * warp;
  if (fa >= nf) return;
  const bool has_b = fa + 1 < nf;

  // input of the forward FFT: conj(Z[k]) at index k = lane + 32 m in register m
  float re[32], im[32];
#pragma unroll
  for (int m = 0; m < 16; ++m) {  // k <= 511:  Z = A[k] + i B[k] = (Ar - Bi) + i (Ai + Br)
    const int k = lane + 32 * m;
    const float2 r = *reinterpret_cast<const float2*>(tre + k * kIsFrames + fa);  // (Ar, Br)
    const float2 i = *reinterpret_cast<const float2*>(tim + k * kIsFrames + fa);  // (Ai, Bi)
    if (m == 0 && lane == 0) {  // DC: the imaginary parts are ignored (C2R)
      re[m] = r.x;
      im[m] = -r.y;
    } else {
      re[m] = r.x - i.y;
      im[m] = -(i.x + r.y);
    }
  }
#pragma unroll
  for (int m = 16; m < 32; ++m) {  // k >= 512:  Z = conj A[N-k] + i conj B[N-k] = (Ar + Bi) + i (Br - Ai)
    const int kk = kNfft - (lane + 32 * m);  // 1 .. 512
    const float2 r = *reinterpret_cast<const float2*>(tre + kk * kIsFrames + fa);
    const float2 i = *reinterpret_cast<const float2*>(tim + kk * kIsFrames + fa);
    if (m == 16 && lane == 0) {  // Nyquist: real
      re[m] = r.x;
      im[m] = -r.y;
    } else {
      re[m] = r.x + i.y;
      im[m] = -(r.y - i.x);
    }
  }
  fft1024_warp(re, im, sre, sim, q.twiddle, lane);  // X[n], n = lane + 32 k2, at register brev5(k2);  s_a + i s_b = conj(X) / N

  float* fo = q.frames + (b * q.F + f0 + fa) * kNfft;
  constexpr float inv_n = 1.0f / (float)kNfft;
#pragma unroll
  for (int k2 = 0; k2 < 32; ++k2) {
    const int n = lane + 32 * k2;
    const float w = __ldg(q.window + n) * inv_n;
    fo[n] = w * re[brev5(k2)];
    if (has_b) fo[kNfft + n] = -w * im[brev5(k2)];
  }
}

__global__ void __launch_bounds__(256) istft1024_ola_kernel(const IstftParams q, int64_t Tout) {
  const int64_t n = q.B * Tout;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / Tout, t = e - b * Tout;
    const int64_t pos = t + kNfft / 2;  // position in the un-trimmed overlap-add buffer
    int64_t f_lo = (pos - (kNfft - 1) + q.hop - 1) / q.hop;
    if (pos - (kNfft - 1) < 0) f_lo = 0;
    int64_t f_hi = pos / q.hop;
    if (f_hi > q.F - 1) f_hi = q.F - 1;
    const float* fb = q.frames + b * q.F * kNfft;
    float acc = 0.0f, env = 0.0f;
    for (int64_t f = f_lo; f <= f_hi; ++f) {
      const int64_t i = pos - f * q.hop;
      const float w = __ldg(q.window + i);
      acc += fb[f * kNfft + i];
      env = fmaf(w, w, env);
    }
    q.out[e] = acc / env;
  }
}

}  // namespace

}  // namespace nvse

struct nvse_frontend {
  int n_fft, hop, n_mels, max_band;
  float* window = nullptr;
  float* window_half = nullptr;  // 0.5 * window: the staged forward kernel folds the FFT-separation factor into it
  int k2n = 16, nyquist = 1;     // which bins carry mel weight (staged forward kernel)
  int4* items = nullptr;         // mel projection work items of the staged forward kernel
  int nitems = 0;
  float2* twiddle = nullptr;
  float* wpack = nullptr;
  int* band_lo = nullptr;
  int* band_len = nullptr;
  int* bin_mlo = nullptr;  // backward: range of mel filters touching each bin
  int* bin_mhi = nullptr;
};

extern "C" int nvse_frontend_create(int n_fft, int hop, int n_mels, const float* window_host,
                                    const float* mel_basis_host, nvse_frontend** out) {
  using namespace nvse;
  NVSE_REQUIRE(out && window_host && mel_basis_host, NVSE_ERR_INVALID, "nvse_frontend_create: null argument");
  NVSE_REQUIRE(n_fft == kNfft, NVSE_ERR_UNSUPPORTED,
               "nvse_frontend_create: n_fft=%d not supported (this build implements n_fft=1024)", n_fft);
  NVSE_REQUIRE(hop >= 1 && n_mels >= 1 && n_mels <= 256, NVSE_ERR_INVALID,
               "nvse_frontend_create: bad hop=%d / n_mels=%d", hop, n_mels);
  int ndev = 0;
  NVSE_CUDA_CHECK(cudaGetDeviceCount(&ndev));
  NVSE_REQUIRE(ndev > 0, NVSE_ERR_CUDA, "nvse_frontend_create: no CUDA device (there is no CPU fallback)");

  // band structure of the mel basis: [lo, lo+len) covers every non-zero of the row
  std::vector<int> lo(n_mels, 0), len(n_mels, 0);
  int max_band = 1;
  for (int m = 0; m < n_mels; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < kBins; ++k)
      if (mel_basis_host[(size_t)m * kBins + k] != 0.0f) {
        if (first < 0) first = k;
        last = k;
      }
    if (first >= 0) {
      lo[m] = first;
      len[m] = last - first + 1;
      if (len[m] > max_band) max_band = len[m];
    }
  }
  std::vector<float> wpack((size_t)max_band * n_mels, 0.0f);
  for (int m = 0; m < n_mels; ++m)
    for (int t = 0; t < len[m]; ++t) wpack[(size_t)t * n_mels + m] = mel_basis_host[(size_t)m * kBins + lo[m] + t];
  std::vector<int> bin_mlo(kBins, 0), bin_mhi(kBins, -1);
  for (int k = 0; k < kBins; ++k) {
    int first = -1, last = -1;
    for (int m = 0; m < n_mels; ++m)
      if (mel_basis_host[(size_t)m * kBins + k] != 0.0f) {
        if (first < 0) first = m;
        last = m;
      }
    if (first >= 0) { bin_mlo[k] = first; bin_mhi[k] = last; }
  }
  std::vector<float2> tw(32 * 32);
  for (int k1 = 0; k1 < 32; ++k1)
    for (int l = 0; l < 32; ++l) {
      const double a = -2.0 * M_PI * (double)(l * k1) / (double)kNfft;
      tw[k1 * 32 + l] = make_float2((float)std::cos(a), (float)std::sin(a));
    }

  nvse_frontend* fe = new nvse_frontend();
  fe->n_fft = n_fft;
  fe->hop = hop;
  fe->n_mels = n_mels;
  fe->max_band = max_band;
  auto cleanup = [&]() { nvse_frontend_destroy(fe); };
#define FE_TRY(expr)                                                                           \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      cleanup();                                                                               \
      return fail(NVSE_ERR_CUDA, "nvse_frontend_create: %s: %s", #expr, cudaGetErrorString(_e)); \
    }                                                                                          \
  } while (0)
  FE_TRY(cudaMalloc(&fe->window, sizeof(float) * kNfft));
  {
    std::vector<float> wh(kNfft);
    for (int i = 0; i < kNfft; ++i) wh[i] = 0.5f * window_host[i];
    FE_TRY(cudaMalloc(&fe->window_half, sizeof(float) * kNfft));
    FE_TRY(cudaMemcpy(fe->window_half, wh.data(), sizeof(float) * kNfft, cudaMemcpyHostToDevice));
    int top = 0;  // one past the highest bin below the Nyquist bin with a non-zero weight in any filter
    bool nyq = false;
    for (int m = 0; m < n_mels; ++m) {
      if (len[m] > 0 && lo[m] + len[m] - 1 == kBins - 1) nyq = true;
      for (int k = 0; k < kBins - 1; ++k)
        if (mel_basis_host[(size_t)m * kBins + k] != 0.0f && k + 1 > top) top = k + 1;
    }
    fe->k2n = std::min(16, std::max(1, (top + 31) / 32));
    fe->nyquist = nyq ? 1 : 0;
    // work items of the mel projection: the smallest chunk size for which every row, cut into 1 / 2 / 4 slices of at most
    // that many taps on an aligned group of lanes, fits the 128 threads of a CTA in one round
    std::vector<int4> items;
    for (int chunk = 4; chunk <= kBins; ++chunk) {
      items.clear();
      bool ok = true;
      for (int m = 0; m < n_mels && ok; ++m) {
        const int need = std::max(1, (len[m] + chunk - 1) / chunk);
        const int nsl = need <= 1 ? 1 : (need <= 2 ? 2 : 4);
        if (need > 4) { ok = false; break; }
        while (items.size() % nsl) items.push_back(make_int4(-1, 0, 0, 1));  // aligned group of lanes
        const int per = (len[m] + nsl - 1) / nsl;
        for (int sidx = 0; sidx < nsl; ++sidx) {
          const int t0 = std::min(len[m], sidx * per), cnt = std::max(0, std::min(per, len[m] - t0));
          items.push_back(make_int4(m, t0, cnt, nsl | (sidx << 8)));
        }
      }
      if (ok && (int)items.size() <= 128) break;
      if (chunk == kBins) {  // does not fit one round even uncut (n_mels > 128): several rounds of whole rows
        items.clear();
        for (int m = 0; m < n_mels; ++m) items.push_back(make_int4(m, 0, len[m], 1));
      }
    }
    while (items.size() % 32) items.push_back(make_int4(-1, 0, 0, 1));  // whole warps: the shuffles need every lane
    fe->nitems = (int)items.size();
    FE_TRY(cudaMalloc(&fe->items, sizeof(int4) * items.size()));
    FE_TRY(cudaMemcpy(fe->items, items.data(), sizeof(int4) * items.size(), cudaMemcpyHostToDevice));
  }
  FE_TRY(cudaMalloc(&fe->twiddle, sizeof(float2) * tw.size()));
  FE_TRY(cudaMalloc(&fe->wpack, sizeof(float) * wpack.size()));
  FE_TRY(cudaMalloc(&fe->band_lo, sizeof(int) * n_mels));
  FE_TRY(cudaMalloc(&fe->band_len, sizeof(int) * n_mels));
  FE_TRY(cudaMemcpy(fe->window, window_host, sizeof(float) * kNfft, cudaMemcpyHostToDevice));
  FE_TRY(cudaMemcpy(fe->twiddle, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
  FE_TRY(cudaMemcpy(fe->wpack, wpack.data(), sizeof(float) * wpack.size(), cudaMemcpyHostToDevice));
  FE_TRY(cudaMemcpy(fe->band_lo, lo.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
  FE_TRY(cudaMemcpy(fe->band_len, len.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
  FE_TRY(cudaMalloc(&fe->bin_mlo, sizeof(int) * kBins));
  FE_TRY(cudaMalloc(&fe->bin_mhi, sizeof(int) * kBins));
  FE_TRY(cudaMemcpy(fe->bin_mlo, bin_mlo.data(), sizeof(int) * kBins, cudaMemcpyHostToDevice));
  FE_TRY(cudaMemcpy(fe->bin_mhi, bin_mhi.data(), sizeof(int) * kBins, cudaMemcpyHostToDevice));
#undef FE_TRY
  *out = fe;
  return NVSE_OK;
}

extern "C" int nvse_frontend_destroy(nvse_frontend* fe) {
  if (!fe) return NVSE_OK;
  cudaFree(fe->window);
  cudaFree(fe->window_half);
  cudaFree(fe->items);
  cudaFree(fe->twiddle);
  cudaFree(fe->wpack);
  cudaFree(fe->band_lo);
  cudaFree(fe->band_len);
  cudaFree(fe->bin_mlo);
  cudaFree(fe->bin_mhi);
  delete fe;
  return NVSE_OK;
}

extern "C" int64_t nvse_frontend_num_frames(const nvse_frontend* fe, int64_t T) {
  if (!fe || T < 0) return -1;
  return 1 + T / fe->hop;
}

static int frontend_mel_impl(const nvse_frontend* fe, const float* y, int64_t B, int64_t T, int64_t y_row_stride, const int* lens,
                             float* out, void* stream, int cl_pitch = 0);

namespace nvse {
// the log-mel written channels-last [B, F, pitch] (zero above n_mels): what conv_pre's tensor-core launch reads (generator.cu)
int frontend_mel_cl(const nvse_frontend* fe, const float* y, int64_t B, int64_t T, int64_t y_row_stride, const int* lens, int pitch,
                    float* out, cudaStream_t st) {
  NVSE_REQUIRE(fe && pitch >= fe->n_mels, NVSE_ERR_INVALID, "frontend_mel_cl: pitch %d below n_mels", pitch);
  return frontend_mel_impl(fe, y, B, T, y_row_stride, lens, out, st, pitch);
}
// frames_dev[b] = 1 + samples_dev[b] / hop
__global__ void frames_from_samples_kernel(const int* __restrict__ samples, int* __restrict__ frames, int64_t B, int hop) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) frames[i] = 1 + samples[i] / hop;
}
int launch_frames_from_samples(const nvse_frontend* fe, const int* samples_dev, int* frames_dev, int64_t B, cudaStream_t st) {
  frames_from_samples_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(samples_dev, frames_dev, B, fe->hop);
  NVSE_LAUNCH_CHECK("frames_from_samples_kernel");
  return NVSE_OK;
}
}  // namespace nvse

extern "C" int nvse_frontend_mel_f32(const nvse_frontend* fe, const float* y, int64_t B, int64_t T,
                                     int64_t y_row_stride, float* out, void* stream) {
  return frontend_mel_impl(fe, y, B, T, y_row_stride, nullptr, out, stream);
}

extern "C" int nvse_frontend_mel_ragged_f32(const nvse_frontend* fe, const float* y, int64_t B, int64_t T, int64_t y_row_stride,
                                            const int32_t* samples_dev, float* out, void* stream) {
  NVSE_REQUIRE(samples_dev, NVSE_ERR_INVALID, "nvse_frontend_mel_ragged_f32: null lengths");
  return frontend_mel_impl(fe, y, B, T, y_row_stride, samples_dev, out, stream);
}

static int frontend_mel_impl(const nvse_frontend* fe, const float* y, int64_t B, int64_t T, int64_t y_row_stride, const int* lens,
                             float* out, void* stream, int cl_pitch) {
  using namespace nvse;
  NVSE_REQUIRE(fe && y && out, NVSE_ERR_INVALID, "nvse_frontend_mel_f32: null argument");
  NVSE_REQUIRE(B >= 0 && y_row_stride >= T, NVSE_ERR_INVALID, "nvse_frontend_mel_f32: bad B/stride");
  // torch.stft(center=True, pad_mode='reflect') raises unless pad < T
  NVSE_REQUIRE(T > fe->n_fft / 2, NVSE_ERR_INVALID,
               "nvse_frontend_mel_f32: reflect padding needs T > n_fft/2 (T=%lld, n_fft=%d)", (long long)T, fe->n_fft);
  if (B == 0) return NVSE_OK;
  FrontendParams p;
  p.y = y;
  p.y_stride = y_row_stride;
  p.T = T;
  p.B = B;
  p.F = 1 + T / fe->hop;
  p.pairs = (p.F + 1) / 2;
  p.hop = fe->hop;
  p.n_mels = fe->n_mels;
  p.window = fe->window;
  p.twiddle = fe->twiddle;
  p.wpack = fe->wpack;
  p.band_lo = fe->band_lo;
  p.band_len = fe->band_len;
  p.out = out;
  // algorithmic bytes: waveform in + log-mel out
  ProfScope prof("mel_frontend", 1, fe->n_mels, 0.0, 4.0 * (double)B * (double)T + 4.0 * (double)B * fe->n_mels * (double)p.F,
                 as_stream(stream));
  static const bool legacy = [] { const char* e = std::getenv("NVSE_FE_LEGACY"); return e && e[0] == '1'; }();
  Frontend2Params q;
  q.f = p;
  q.window_half = fe->window_half;
  q.groups = (p.F + kFe2Frames - 1) / kFe2Frames;
  q.total = B * q.groups;
  q.nsmp = (kFe2Frames - 1) * fe->hop + kNfft;
  q.nsmp_pad = (q.nsmp + 3) & ~3;
  q.nyquist = fe->nyquist;
  q.items = fe->items;
  q.nitems = fe->nitems;
  q.lens = lens;
  q.out_cl_pitch = cl_pitch;
  const size_t smem2 = sizeof(float) * ((size_t)q.nsmp_pad + (size_t)kFe2Warps * 2 * kFe2Scratch);
  if (!legacy && smem2 <= 100 * 1024 && fe->n_mels <= kFe2MaxMels) {  // the staged kernel (persistent CTAs, 8 consecutive frames per step)
    static const int ctas_per_sm = [] { const char* e = std::getenv("NVSE_FE_CTAS"); const int v = e ? std::atoi(e) : 0; return v > 0 ? v : 4; }();
    const int64_t ctas = std::min<int64_t>(q.total, (int64_t)device_sm_count() * ctas_per_sm);
    if (fe->k2n <= 12) {  // fmax 8000 at 22.05 kHz: bins >= 372 carry no mel weight
      NVSE_CUDA_CHECK(cudaFuncSetAttribute(mel_frontend2_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      mel_frontend2_kernel<12><<<(unsigned)ctas, kFe2Threads, smem2, as_stream(stream)>>>(q);
    } else {
      NVSE_CUDA_CHECK(cudaFuncSetAttribute(mel_frontend2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      mel_frontend2_kernel<16><<<(unsigned)ctas, kFe2Threads, smem2, as_stream(stream)>>>(q);
    }
    NVSE_LAUNCH_CHECK("mel_frontend2_kernel");
    return NVSE_OK;
  }
  // very large hops: one warp per frame pair straight from global memory
  NVSE_REQUIRE(cl_pitch == 0, NVSE_ERR_UNSUPPORTED, "the fused wav -> wav call needs the staged front-end kernel (hop too large)");
  NVSE_REQUIRE(!lens, NVSE_ERR_UNSUPPORTED, "nvse_frontend_mel_ragged_f32: per-utterance lengths need the staged kernel (hop too large)");
  const int64_t tasks = B * p.pairs;
  const int64_t ctas = (tasks + kWarpsPerCta - 1) / kWarpsPerCta;
  NVSE_REQUIRE(ctas <= 0x7fffffff, NVSE_ERR_INVALID, "nvse_frontend_mel_f32: problem too large for one launch");
  const size_t smem = sizeof(float) * kWarpSmemFloats * kWarpsPerCta;
  NVSE_CUDA_CHECK(cudaFuncSetAttribute(mel_frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mel_frontend_kernel<<<(unsigned)ctas, kWarpsPerCta * 32, smem, as_stream(stream)>>>(p);
  NVSE_LAUNCH_CHECK("mel_frontend_kernel");
  return NVSE_OK;
}

extern "C" size_t nvse_frontend_backward_scratch_bytes(const nvse_frontend* fe, int64_t B, int64_t T) {
  if (!fe || B < 0 || T < 0) return 0;
  return (size_t)B * (size_t)(1 + T / fe->hop) * nvse::kNfft * sizeof(float) + 256;
}

extern "C" int nvse_frontend_mel_backward_f32(const nvse_frontend* fe, const float* y, int64_t B, int64_t T,
                                              int64_t y_row_stride, const float* dmel, float* dy, void* scratch,
                                              size_t scratch_bytes, void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(fe && y && dmel && dy && scratch, NVSE_ERR_INVALID, "nvse_frontend_mel_backward_f32: null argument");
  NVSE_REQUIRE(B >= 0 && y_row_stride >= T, NVSE_ERR_INVALID, "nvse_frontend_mel_backward_f32: bad B/stride");
  NVSE_REQUIRE(T > fe->n_fft / 2, NVSE_ERR_INVALID,
               "nvse_frontend_mel_backward_f32: reflect padding needs T > n_fft/2 (T=%lld, n_fft=%d)", (long long)T, fe->n_fft);
  NVSE_REQUIRE(scratch_bytes >= nvse_frontend_backward_scratch_bytes(fe, B, T), NVSE_ERR_INVALID, "scratch too small");
  if (B == 0) return NVSE_OK;
  FrontendBwdParams q;
  FrontendParams& p = q.f;
  p.y = y; p.y_stride = y_row_stride; p.T = T; p.B = B;
  p.F = 1 + T / fe->hop; p.pairs = (p.F + 1) / 2; p.hop = fe->hop; p.n_mels = fe->n_mels;
  p.window = fe->window; p.twiddle = fe->twiddle; p.wpack = fe->wpack; p.band_lo = fe->band_lo; p.band_len = fe->band_len;
  p.out = nullptr;
  q.dmel = dmel; q.bin_mlo = fe->bin_mlo; q.bin_mhi = fe->bin_mhi;
  q.frames = reinterpret_cast<float*>((reinterpret_cast<size_t>(scratch) + 255) / 256 * 256);
  const int64_t tasks = B * p.pairs;
  const int64_t ctas = (tasks + kWarpsPerCta - 1) / kWarpsPerCta;
  NVSE_REQUIRE(ctas <= 0x7fffffff, NVSE_ERR_INVALID, "nvse_frontend_mel_backward_f32: problem too large for one launch");
  const size_t smem = sizeof(float) * kBwdWarpSmemFloats * kWarpsPerCta;
  NVSE_CUDA_CHECK(cudaFuncSetAttribute(mel_frontend_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = as_stream(stream);
  mel_frontend_bwd_kernel<<<(unsigned)ctas, kWarpsPerCta * 32, smem, st>>>(q);
  NVSE_LAUNCH_CHECK("mel_frontend_bwd_kernel");
  const int64_t n = B * T;
  mel_overlap_add_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 1 << 20), 256, 0, st>>>(q.frames, dy, B, T, p.F, fe->hop);
  NVSE_LAUNCH_CHECK("mel_overlap_add_kernel");
  return NVSE_OK;
}

extern "C" int nvse_frontend_stft_f32(const nvse_frontend* fe, const float* y, int64_t B, int64_t T, int64_t y_row_stride,
                                      float* log_amp, float* phase, float* real, float* imag, void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(fe && y, NVSE_ERR_INVALID, "nvse_frontend_stft_f32: null argument");
  NVSE_REQUIRE(B >= 0 && y_row_stride >= T, NVSE_ERR_INVALID, "nvse_frontend_stft_f32: bad B/stride");
  NVSE_REQUIRE(T > fe->n_fft / 2, NVSE_ERR_INVALID, "nvse_frontend_stft_f32: reflect padding needs T > n_fft/2 (T=%lld, n_fft=%d)",
               (long long)T, fe->n_fft);
  if (B == 0) return NVSE_OK;
  StftParams q;
  FrontendParams& p = q.f;
  p.y = y; p.y_stride = y_row_stride; p.T = T; p.B = B;
  p.F = 1 + T / fe->hop; p.pairs = (p.F + 1) / 2; p.hop = fe->hop; p.n_mels = fe->n_mels;
  p.window = fe->window; p.twiddle = fe->twiddle; p.wpack = fe->wpack; p.band_lo = fe->band_lo; p.band_len = fe->band_len;
  p.out = nullptr;
  q.log_amp = log_amp; q.phase = phase; q.real = real; q.imag = imag;
  const int groups_per_utt = (int)((p.F + kStftFrames - 1) / kStftFrames);
  const int64_t ctas = B * groups_per_utt;
  NVSE_REQUIRE(ctas <= 0x7fffffff, NVSE_ERR_INVALID, "nvse_frontend_stft_f32: problem too large for one launch");
  const size_t smem = sizeof(float) * kStftSmemFloats;
  NVSE_CUDA_CHECK(cudaFuncSetAttribute(stft_amp_pha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int planes = (log_amp ? 1 : 0) + (phase ? 1 : 0) + (real ? 1 : 0) + (imag ? 1 : 0);
  ProfScope prof("stft_amp_pha", 1, planes, 0.0, 4.0 * (double)B * (double)T + 4.0 * planes * (double)B * kBins * (double)p.F, as_stream(stream));
  stft_amp_pha_kernel<<<(unsigned)ctas, kWarpsPerCta * 32, smem, as_stream(stream)>>>(q, groups_per_utt);
  NVSE_LAUNCH_CHECK("stft_amp_pha_kernel");
  return NVSE_OK;
}

extern "C" int nvse_inverse_mel_f32(const float* inv_basis, const float* mel, float* out, int64_t B, int n_bins, int n_mels,
                                    int64_t frames, void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(inv_basis && mel && out, NVSE_ERR_INVALID, "nvse_inverse_mel_f32: null argument");
  NVSE_REQUIRE(B >= 0 && B <= 65535 && n_bins >= 1 && n_mels >= 1 && n_mels <= 1024 && frames >= 0, NVSE_ERR_INVALID,
               "nvse_inverse_mel_f32: bad shape");
  if (B == 0 || frames == 0) return NVSE_OK;
  dim3 grid((unsigned)((frames + kInvFrames - 1) / kInvFrames), (unsigned)((n_bins + kInvBins - 1) / kInvBins), (unsigned)B);
  ProfScope prof("inverse_mel", n_mels, n_bins, 2.0 * (double)B * frames * n_bins * n_mels,
                 4.0 * (double)B * frames * (n_bins + n_mels), as_stream(stream));
  inverse_mel_kernel<<<grid, 256, 0, as_stream(stream)>>>(inv_basis, mel, out, n_bins, n_mels, frames);
  NVSE_LAUNCH_CHECK("inverse_mel_kernel");
  return NVSE_OK;
}


extern "C" size_t nvse_frontend_istft_scratch_bytes(const nvse_frontend* fe, int64_t B, int64_t frames) {
  if (!fe || B < 0 || frames < 0) return 0;
  return (size_t)B * (size_t)frames * nvse::kNfft * sizeof(float) + 256;
}

static int frontend_istft_impl(const nvse_frontend* fe, const float* real, const float* imag, int64_t B, int64_t frames, float* out,
                               void* scratch, size_t scratch_bytes, void* stream);

extern "C" int nvse_frontend_istft_f32(const nvse_frontend* fe, const float* real, const float* imag, int64_t B, int64_t frames,
                                       float* out, void* scratch, size_t scratch_bytes, void* stream) {
  NVSE_REQUIRE(imag, NVSE_ERR_INVALID, "nvse_frontend_istft_f32: null argument");
  return frontend_istft_impl(fe, real, imag, B, frames, out, scratch, scratch_bytes, stream);
}

extern "C" int nvse_frontend_istft_c64(const nvse_frontend* fe, const float* spec_c64, int64_t B, int64_t frames, float* out,
                                       void* scratch, size_t scratch_bytes, void* stream) {
  NVSE_REQUIRE((reinterpret_cast<size_t>(spec_c64) & 7) == 0, NVSE_ERR_INVALID, "nvse_frontend_istft_c64: the spectrum must be 8-byte aligned");
  return frontend_istft_impl(fe, spec_c64, nullptr, B, frames, out, scratch, scratch_bytes, stream);
}

static int frontend_istft_impl(const nvse_frontend* fe, const float* real, const float* imag, int64_t B, int64_t frames, float* out,
                               void* scratch, size_t scratch_bytes, void* stream) {
  using namespace nvse;
  NVSE_REQUIRE(fe && real && out && scratch, NVSE_ERR_INVALID, "nvse_frontend_istft_f32: null argument");
  NVSE_REQUIRE(B >= 0 && frames >= 1, NVSE_ERR_INVALID, "nvse_frontend_istft_f32: bad B=%lld / frames=%lld", (long long)B, (long long)frames);
  NVSE_REQUIRE(fe->hop <= kNfft, NVSE_ERR_UNSUPPORTED, "nvse_frontend_istft_f32: hop %d > n_fft leaves gaps (torch.istft rejects it: zero envelope)", fe->hop);
  NVSE_REQUIRE(scratch_bytes >= nvse_frontend_istft_scratch_bytes(fe, B, frames), NVSE_ERR_INVALID, "nvse_frontend_istft_f32: scratch too small");
  const int64_t Tout = (int64_t)fe->hop * (frames - 1);
  if (B == 0 || Tout == 0) return NVSE_OK;
  cudaStream_t st = as_stream(stream);
  IstftParams q;
  q.re = real; q.im = imag; q.B = B; q.F = frames; q.hop = fe->hop;
  q.window = fe->window; q.twiddle = fe->twiddle;
  q.frames = reinterpret_cast<float*>((reinterpret_cast<size_t>(scratch) + 255) / 256 * 256);
  q.out = out;
  const int64_t ctas = B * ((frames + kIsFrames - 1) / kIsFrames);
  NVSE_REQUIRE(ctas <= 0x7fffffff, NVSE_ERR_INVALID, "nvse_frontend_istft_f32: problem too large for one launch");
  const size_t smem = sizeof(float) * (2 * (size_t)kBins * kIsFrames + (size_t)kIsWarps * kWarpSmemFloats);
  NVSE_CUDA_CHECK(cudaFuncSetAttribute(istft1024_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    ProfScope prof("istft1024_frames", 513, 1, 0.0, 8.0 * (double)B * kBins * (double)frames + 4.0 * (double)B * (double)frames * kNfft, st);
    istft1024_frames_kernel<<<(unsigned)ctas, kIsWarps * 32, smem, st>>>(q);
    NVSE_LAUNCH_CHECK("istft1024_frames_kernel");
  }
  {
    ProfScope prof("istft1024_ola", 1, 1, 0.0, 4.0 * (double)B * (double)Tout * (1.0 + (double)kNfft / fe->hop), st);
    const unsigned grid = (unsigned)std::min<int64_t>((B * Tout + 255) / 256, 148 * 16);
    istft1024_ola_kernel<<<grid, 256, 0, st>>>(q, Tout);
    NVSE_LAUNCH_CHECK("istft1024_ola_kernel");
  }
  return NVSE_OK;
}
