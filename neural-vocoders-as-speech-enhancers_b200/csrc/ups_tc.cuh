// Launch interface of the persistent tcgen05 ConvTranspose1d kernel (ups_tc.cu): the upsamplers of the
// HiFi-GAN family (reference Models/hifigan.py:93-96, 111-112; k = 2 * stride, padding = stride / 2) between two
// activations in the T32 layout.
#pragma once

#include "common.cuh"

namespace nvse {

struct UpsTcArgs {
  const float* x;      // [B][t32_rows(Tin)][Cin] fp32, T32 layout (or, x_cl, channels-last [B][Tin][Cin]); leaky_relu(in_slope) on load
  int64_t x_bstride;   // elements per utterance
  int Tin, Cin, Cout, stride;
  const void* wimg;    // IEEE-half image built by launch_pack_weight_ups
  const float* bias;   // [Cout]
  float* y;            // [B][t32_rows(stride * Tin)][Cout] fp32, T32 layout
  int64_t y_bstride;
  float in_slope;
  RowLens in_lens;     // ragged batch: valid input rows per utterance (null lens: Tin)
  int x_cl;            // x is channels-last [B][Tin][Cin] instead of T32 (stride 8 only: the first upsampler reads conv_pre's output)
};

// k = 2 * stride, padding = stride / 2, stride 2 or 8, Cin a multiple of 32, channel slices of 16-channel chunks
bool ups_tc_supported(int Cin, int Cout, int k, int stride, int padding);
inline size_t ups_tc_image_elems(int Cin, int Cout, int k) { return (size_t)Cin * Cout * k; }
// fp32 [k][Cin][Cout] (Layer::w layout of a ConvTranspose1d) -> the kernel's stage-ordered half image
int launch_pack_weight_ups(const float* w_kio, void* img, int Cin, int Cout, int stride, cudaStream_t st);
int launch_ups_tc(const UpsTcArgs& a, int64_t B, cudaStream_t st);
int ups_abort_status(bool reset, unsigned int* flag);
int ups_abort_bind(unsigned int* host_word_dev);
int ups_abort_clear(cudaStream_t st);

}  // namespace nvse
