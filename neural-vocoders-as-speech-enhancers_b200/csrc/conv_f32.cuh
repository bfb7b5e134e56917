// Launch interface of the fp32 tap-list convolution kernels (conv_f32.cu).
#pragma once

#include <algorithm>

#include "common.cuh"

namespace nvse {

struct ConvF32Args {
  const float* x;     // [B, Tin_actual, Cin] channels-last
  int64_t x_bstride;  // elements between batch items of x
  int Tin;            // rows addressable by taps (virtual length when reflect_left > 0)
  int Cin;
  const float* w;     // [slices][Cin][Cout]
  const float* bias;  // [Cout] or null
  const float* residual;  // same shape as y, or null
  float* y;           // [B, Tout, Cout]
  int64_t y_bstride;
  int Tout;
  int Cout;
  ConvTaps taps;
  int out_mul, out_add;  // output row = out_mul * t + out_add
  int Trows;             // t runs over [0, Trows)
  float in_slope;        // leaky_relu slope applied to x on load (1 = identity)
  float out_scale;       // y = [accumulate ? y : 0] + out_scale * (conv + bias + residual)
  int accumulate;
  int out_act;           // 0 none, 1 tanh
  int16_t* y_pcm16;      // when set (conv_post of the T32 path only): store round(y * 32767) clipped to int16 here INSTEAD of
                         // y -- the PCM_16 quantisation of sf.write (infers/inference_hifigan.py:93) fused into the last kernel
  int x_t32;             // x is in the T32 layout (thin kernel only: conv_post of the tensor-core path)
  int reflect_left;      // x is viewed through ReflectionPad1d((reflect_left, 0)) (istftnet.py:296,312)
  // backward (grad.cu): a strided input row map and the leaky_relu derivative fused into the epilogue
  int in_stride;         // input row of tap i = max(in_stride, 1) * t + off[i]  (dgrad of a ConvTranspose1d)
  const float* mask;     // same shape as y, or null:  conv *= (mask > 0 ? 1 : mask_slope)  before `residual` is added
  float mask_slope;
  RowLens in_lens;       // ragged batch (T32 conv_post kernel only): valid input rows per utterance; null lens: Tin
};

int launch_conv_f32(const ConvF32Args& a, int64_t B, cudaStream_t st);
int launch_repack_weight(const float* src, float* dst, int Cin, int Cout, int k, bool transposed, cudaStream_t st);
int launch_transpose(const float* x, float* y, int64_t B, int64_t R, int64_t C, cudaStream_t st);
// same with y rows padded to Rpad >= R entries (zero fill): x [B, R, C] -> y [B, C, Rpad]
int launch_transpose_pad(const float* x, float* y, int64_t B, int64_t R, int64_t C, int64_t Rpad, cudaStream_t st);
int launch_pcm16(const float* x, int16_t* y, int64_t n, cudaStream_t st);
int launch_add3(float* a, const float* b, const float* c, int64_t n, cudaStream_t st);  // a = (a + b) + c
// channels-last [B, T, C] <-> T32 layout (common.cuh)
int launch_relayout_t32(const float* x, float* y, int64_t B, int64_t T, int C, bool to_t32, cudaStream_t st);
void conv1d_taps(int k, int dilation, ConvTaps* taps);
// taps of output phase `phase` of a ConvTranspose1d; returns the tap count or -1 if > kMaxTaps
int conv_transpose_phase_taps(int k, int stride, int padding, int phase, ConvTaps* taps);

}  // namespace nvse
