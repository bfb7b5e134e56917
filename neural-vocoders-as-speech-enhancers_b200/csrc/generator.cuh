// Generator handle internals shared by generator.cu (fp32 path) and generator_bf16.cu.
#pragma once

#include <string>
#include <unordered_map>
#include <vector>

#include "conv_f32.cuh"

namespace nvse {

struct Layer {
  std::string name;   // reference state-dict prefix, e.g. "resblocks.4.convs1.2"
  bool transposed = false;  // ConvTranspose1d
  int Cin = 0, Cout = 0, k = 0, dilation = 1, stride = 1, padding = 0;
  float* w = nullptr;     // fp32 [k][Cin][Cout]
  float* bias = nullptr;  // fp32 [Cout]
  void* w_bf16 = nullptr;  // tensor-core weight image (see conv_tc.cuh), null for fp32-only layers
  void* w_f16 = nullptr;   // the same image in IEEE half: second conv of a ResBlock1 pair at C <= 32 (resblock_tc.cuh h_fp16)
  bool tc_split = false;   // activations as hi+lo bf16 pairs on the tensor-core path (accuracy)
  bool tc_f16 = false;     // IEEE-half operands (w_f16) on the tensor-core path: the upsamplers
  // conv_pre on the tensor cores (16-bit path): IEEE-half images of its output-channel slices, input channels zero-padded
  // to pre_cin (80 -> 128); pre_cout channels per slice (N <= 256 per MMA)
  void* w_ups = nullptr;   // ConvTranspose1d with k = 2 * stride: stage-ordered half image of the persistent all-phase kernel (ups_tc.cuh)
  void* w_pre[4] = {nullptr, nullptr, nullptr, nullptr};
  int pre_cin = 0, pre_cout = 0, pre_n = 0;
  // iSTFTNet's conv_post on the tensor cores (16-bit path): IEEE-half image with the output channels zero-padded to
  // post_cout (n_fft + 2 = 18 -> 32), and the bias padded likewise
  void* w_post = nullptr;
  float* bias_post = nullptr;
  int post_cout = 0;
  bool have_w = false, have_bias = false;
  float* wT = nullptr;     // training path: per-tap transposed weights [k][Cout][Cin] (the dgrad operand), built lazily
  void* wT_bf16 = nullptr; // training path, tensor-core dgrad: the bf16 image of wT (as a layer with Cin <-> Cout)
  int64_t grad_off = 0;    // offset of this layer's (dW, dbias) in the flat gradient buffer of nvse_generator_backward
};

}  // namespace nvse

namespace nvse { struct WeightLoader; void destroy_weight_loader(WeightLoader*); }

struct nvse_generator {
  nvse_generator_config cfg;
  std::vector<nvse::Layer> layers;
  std::unordered_map<std::string, int> index;
  bool finalized = false;
  bool train_ready = false;  // wT of every layer matches w
  nvse::WeightLoader* loader = nullptr;  // tables of the batched weight load (weights.cu)
  // small-batch mode: the ResBlocks of a stage run concurrently on the caller's stream + two side streams
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  const nvse::Layer& layer(const std::string& name) const { return layers[index.at(name)]; }
};

namespace nvse {

// lens: ragged batch, valid conv_post rows per utterance EXCLUDING the reflected one (null: all Tp - 1)
// zpitch: floats per row of z (0: n_fft + 2, dense)
int launch_istft_head(const float* z, float* out, int64_t B, int64_t Tp, int n_fft, int hop, cudaStream_t st, RowLens lens = RowLens{nullptr, 1, 0},
                      int zpitch = 0);
// dz rows have `pitch` >= n_fft + 2 floats (extra columns zeroed)
int launch_istft_head_bwd(const float* z, const float* gout, float* dz, int64_t B, int64_t Tp, int n_fft, int hop, int pitch,
                          cudaStream_t st);
// frontend.cu: the log-mel channels-last [B, F, pitch] for the fused wav -> wav call; frames per utterance of a ragged batch
int frontend_mel_cl(const nvse_frontend* fe, const float* y, int64_t B, int64_t T, int64_t y_row_stride, const int* lens, int pitch,
                    float* out, cudaStream_t st);
int launch_frames_from_samples(const nvse_frontend* fe, const int* samples_dev, int* frames_dev, int64_t B, cudaStream_t st);
int launch_pad_reflect_left(const float* x, float* y, int64_t B, int64_t T, int C, cudaStream_t st);      // [B,T,C] -> [B,T+1,C]
int launch_unpad_reflect_left(const float* dy, float* dx, int64_t B, int64_t T, int C, cudaStream_t st);  // the adjoint

// ConvTranspose1d layer (polyphase): tensor cores (IEEE-half operands) when tc and the layer has an image, else fp32
int run_conv_transpose(const Layer& L, bool tc, const float* x, int64_t B, int64_t Tin, float* y, float in_slope,
                       cudaStream_t st, bool x_t32 = false, bool y_t32 = false, RowLens in_lens = RowLens{nullptr, 1, 0});
int ensure_side_streams(nvse_generator* g);  // g->side[], ev_fork, ev_join[]: the ResBlocks of an MRF run concurrently
int finalize_plan(nvse_generator* g);                    // allocates the tensor-core image buffers, sets the per-layer precision flags
int finalize_bf16(nvse_generator* g, cudaStream_t st);  // finalize_plan + builds the tensor-core weight images
int build_extra_images(nvse_generator* g, cudaStream_t st);  // images derived from Layer::w that the batched loader does not write (conv_pre)
int tc_abort_status(bool reset, unsigned int* flag);
int tc_abort_bind(unsigned int* host_word_dev);
int tc_abort_clear(cudaStream_t st);
// tc_abort.cu: bind this device to the shared host word; poll it on entry of a product call (NVSE_ERR_STATE if a
// kernel of an earlier call timed out; the flags are cleared so later calls run)
int tc_abort_bind_device();
int tc_abort_poll(cudaStream_t st);
void tc_abort_host_clear();

}  // namespace nvse
