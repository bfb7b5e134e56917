// Launch interface of the fused ResBlock1 chain kernel (resblock_tc.cu): up to three
// [lrelu -> dilated conv -> lrelu -> conv -> + x] pairs (Models/hifigan.py:43-50) in ONE
// launch, with the fp32 residual stream held in tensor memory between the pairs.
#pragma once

#include "common.cuh"

namespace nvse {

constexpr int kRbMaxPairs = 3;

struct RbPair {
  const __nv_bfloat16* w1;  // tensor-core weight image of convs1[m] (conv_tc.cuh layout), dilation `dil`
  const __nv_bfloat16* w2;  // image of convs2[m], dilation 1 (IEEE half bit patterns when h_fp16)
  const float* b1;          // [C]
  const float* b2;          // [C]
  int dil;
};

struct ResblockTcArgs {
  const float* x;    // [B, T, C] channels-last fp32 residual stream in
  float* y;          // [B, T, C]   y = [accumulate ? y : 0] + out_scale * resblock(x)
  int T, C, k;       // k odd
  int t32;           // x and y are in the T32 layout (common.cuh) instead of channels-last
  int64_t bstride;   // elements between utterances of x and of y (0: dense, T * C or t32_rows(T) * C)
  int npairs;
  RbPair pair[kRbMaxPairs];
  float slope;       // leaky_relu slope in front of every conv (LRELU_SLOPE, hifigan.py:7)
  float out_scale;
  int accumulate;
  int h_fp16;        // the c1 -> c2 intermediate and the w2 images are IEEE half instead of bf16 (same range of values,
                     // three more mantissa bits): where rounding of that intermediate costs the most SNR (tests/bf16_budget.py)
  RowLens lens;      // ragged batch: valid rows per utterance (rows beyond are zero, like rows beyond T); null lens: T
};

// true when the fused kernel handles this shape (otherwise the caller uses the per-layer kernels)
bool rb_supported(int C, int k, const int* dil, int npairs);
// modelled SM cycles per output row of one launch covering `npairs` pairs (< 0 when unsupported): lets the
// caller choose between one launch for the whole ResBlock and one launch per pair
double rb_cost_per_row(int C, int k, const int* dil, int npairs);
int launch_resblock_tc(const ResblockTcArgs& a, int64_t B, cudaStream_t st);
int rb_abort_status(bool reset, unsigned int* flag);
int rb_abort_bind(unsigned int* host_word_dev);
int rb_abort_clear(cudaStream_t st);
long long* rb_trace_buffer();  // debug stamps (NVSE_RB_TRACE), null when off

// pair_tc.cu: persistent, software-pipelined single pair (C = 128), T32 layout only
bool pair_supported(int C, int k, int dil);
int launch_pair_tc(const ResblockTcArgs& a, int64_t B, cudaStream_t st);
int pair_abort_status(bool reset, unsigned int* flag);
int pair_abort_bind(unsigned int* host_word_dev);
int pair_abort_clear(cudaStream_t st);

}  // namespace nvse
