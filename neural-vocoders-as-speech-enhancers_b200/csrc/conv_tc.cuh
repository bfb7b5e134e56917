// Launch interface of the tcgen05 (5th-gen tensor core) tap-list convolution (conv_tc.cu).
#pragma once

#include "common.cuh"

namespace nvse {

// Weight image for the tensor-core kernel: bf16, laid out as the exact shared-memory image of
// every (tap slice, K-chunk) stage so that one cp.async.bulk (TMA 1-D) moves a stage:
//   [slice j][ci / KC][(ci % KC) / 8][co][ci % 8],   KC = min(Cin, 64)
// which is the canonical no-swizzle K-major UMMA layout (8 x 16-byte core matrices).
inline int tc_kchunk(int Cin) { return Cin < 64 ? Cin : 64; }
inline size_t tc_weight_image_elems(int Cin, int Cout, int k) { return (size_t)Cin * Cout * k; }
inline bool tc_supported(int Cin, int Cout) {
  return (Cin == 32 || Cin == 64 || Cin == 128 || Cin == 256 || Cin == 512) &&
         (Cout == 32 || Cout == 64 || Cout == 128 || Cout == 256);
}

struct ConvTcArgs {
  const void* x;       // [B, Tin, Cin] channels-last: fp32 (leaky_relu fused on load) or bf16 (used as is)
  int64_t x_bstride;   // elements
  int Tin;
  int Cin, Cout;
  int in_bf16;
  int x_t32;           // fp32 x is in the T32 layout (common.cuh); x_bstride then counts padded rows
  const __nv_bfloat16* wimg;
  const float* bias;      // [Cout] or null
  const float* residual;  // fp32, same shape as y (fp32 output only; added after the out_slope activation)
  void* y;                // [B, Tout, Cout] fp32, or bf16 when out_bf16
  int64_t y_bstride;
  int Tout;
  int out_bf16;           // y = bf16(lrelu(conv + bias, out_slope))
  int y_t32;              // fp32 y (and residual) are in the T32 layout
  int y_ld;               // fp32 channels-last y only: row pitch in elements (0 = Cout); y then points at this launch's first column
  ConvTaps taps;
  int out_mul, out_add, Trows;
  float in_slope, out_slope, out_scale;
  int accumulate;
  int split_act;          // stage activations as hi + lo bf16 planes (fp32 input only): 2 MMAs per K step
  const float* mask;      // backward (dgrad): same shape as y, or null -- conv *= (mask > 0 ? 1 : mask_slope) before `residual`
  float mask_slope;       //   is added (the leaky_relu derivative of the layer input); fp32 channels-last output only
  int ops_f16;            // stage the (fp32) activations as IEEE half; `wimg` must then be a half image.  Used by the
                          // upsamplers, where bf16 rounding of the WEIGHTS is the largest error of the whole path
  RowLens in_lens;        // ragged batch: valid INPUT rows per utterance (rows beyond read as zero); lens == null: Tin
  int ntile_hint;         // 128-row tiles per CTA (0: chosen by the launcher).  Memory-bound thin layers want 1: several small CTAs per SM
};

// shared memory one CTA of the kernel needs for this shape (used to decide whether split fits)
size_t tc_smem_bytes(int Cin, int Cout, int tap_span, bool split, int stages);
bool tc_split_fits(int Cin, int Cout, int tap_span);

int launch_conv_tc(const ConvTcArgs& a, int64_t B, cudaStream_t st);
// Several output phases (ConvTranspose1d polyphase branches) computed from ONE staged activation tile;
// a.taps / a.out_add are ignored, phase p uses phase_taps[p] and output rows out_mul * t + phase_out_add[p].
constexpr int kTcMaxPhases = 8;
int launch_conv_tc_phases(const ConvTcArgs& a, const ConvTaps* phase_taps, const int* phase_out_add, int nphase, int64_t B,
                          cudaStream_t st);
// fp32 [k][Cin][Cout] (Layer::w layout) -> bf16 tensor-core image
// as_fp16: IEEE half instead of bf16 in the same image layout (operands of the C = 32 c2 convs, resblock_tc.cu)
int launch_pack_weight_tc(const float* w_kio, __nv_bfloat16* img, int Cin, int Cout, int k, cudaStream_t st, bool as_fp16 = false);
// image of the [Cin x Cout] slice starting at output channel co0 of a [k][cin_src][cout_src] weight, input channels zero-padded to Cin
int launch_pack_weight_tc_slice(const float* w_kio, int cin_src, int cout_src, int co0, __nv_bfloat16* img, int Cin, int Cout, int k,
                                cudaStream_t st, bool as_fp16);

}  // namespace nvse
