/* nvse_b200.h -- C ABI of the B200-native (sm_100a) time-domain vocoding hot path.
 *
 * The reference (Andong-Li-speech/Neural-Vocoders-as-Speech-Enhancers) is pure Python /
 * PyTorch and has no FFI of its own; its boundary for this path is a Python API
 * (SURVEY.md §8b).  Each entry point below names the reference interface it replaces
 * (paths relative to the reference checkout).  The Python drop-ins in
 * neural-vocoders-as-speech-enhancers_b200/{dataset.py,Models/} bind these with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All tensor pointers are DEVICE
 *     pointers unless the name ends in _host.  `stream` is a cudaStream_t / CUstream.
 *   - every function returns 0 on success, non-zero on failure; nvse_last_error()
 *     returns a thread-local description of the last failure.
 *   - the caller owns every buffer, including workspaces; the library owns only what a
 *     handle holds (packed weights, constant tables).  One handle per device, not
 *     re-entrant per handle.  Work is enqueued on `stream`; nothing synchronises.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef NVSE_B200_H_
#define NVSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NVSE_API __attribute__((visibility("default")))
#else
#define NVSE_API
#endif

#define NVSE_ABI_VERSION 1

enum {
  NVSE_OK = 0,
  NVSE_ERR_INVALID = 1,     /* bad argument / shape */
  NVSE_ERR_CUDA = 2,        /* CUDA runtime error */
  NVSE_ERR_UNSUPPORTED = 3, /* valid for the reference, not implemented by this library */
  NVSE_ERR_STATE = 4        /* handle not ready (e.g. weights missing) */
};

enum { NVSE_PRECISION_F32 = 0, NVSE_PRECISION_BF16 = 1 };
enum { NVSE_GEN_HIFIGAN = 0, NVSE_GEN_ISTFTNET = 1 };

NVSE_API int nvse_abi_version(void);
NVSE_API const char* nvse_last_error(void);
/* Number of kernel launches issued by this library in the calling process so far
 * (bench.py reports the delta over the timed region as `gpu_launches`). */
NVSE_API uint64_t nvse_launch_count(void);

/* Per-launch timing for bench.py's roofline numbers.  Between begin and end every kernel this
 * library launches is bracketed by CUDA events on its stream; end() waits for them and writes a
 * JSON array [{"kernel", "launches", "ms", "flops", "bytes"}...] aggregated per kernel and channel
 * class (flops/bytes are ALGORITHMIC, as defined in DESIGN.md), then disables timing. */
NVSE_API int nvse_profile_begin(void);
NVSE_API int nvse_profile_end(char* json_out, size_t capacity);

/* ------------------------------------------------------------------------------------
 * Front-end: dataset.mel_spectrogram  (dataset.py:53-91)
 *   reflect-pad n_fft/2 | frame (n_fft, hop) | x window | R2C FFT | magnitude |
 *   mel_basis @ . | log(clamp(., 1e-5))        -- one fused kernel, fp32.
 * The handle replaces the reference's per-parameter cache `mel_window[param_string]`
 * (dataset.py:45-50,68-76): it holds the window, the mel basis and the FFT twiddles.
 * ---------------------------------------------------------------------------------- */
typedef struct nvse_frontend nvse_frontend;

/* window_host: [n_fft] (a win_size < n_fft window must be centre-padded by the caller,
 * as torch.stft does).  mel_basis_host: [n_mels, n_fft/2+1] row-major (librosa layout).
 * Supported: n_fft == 1024, hop >= 1, 1 <= n_mels <= 256. */
NVSE_API int nvse_frontend_create(int n_fft, int hop, int n_mels, const float* window_host,
                         const float* mel_basis_host, nvse_frontend** out);
NVSE_API int nvse_frontend_destroy(nvse_frontend* fe);
/* frames = 1 + T / hop  (torch.stft center=True, dataset.py:78-86) */
NVSE_API int64_t nvse_frontend_num_frames(const nvse_frontend* fe, int64_t T);
/* y: [B, T] fp32 with row stride y_row_stride (elements); out: [B, n_mels, F] fp32.
 * Requires T > n_fft/2 (reflect padding), like torch.stft. */
NVSE_API int nvse_frontend_mel_f32(const nvse_frontend* fe, const float* y, int64_t B, int64_t T,
                          int64_t y_row_stride, float* out, void* stream);
/* Ragged batch: y [B, T] padded to the longest utterance, samples_dev [B] int32 ON THE DEVICE = samples of each utterance
 * (n_fft/2 < samples_dev[b] <= T).  Utterance b gets the log-mel of its own samples (reflect padding at ITS end, bit-identical
 * to a single-utterance call): frames [0, 1 + samples_dev[b] / hop) of out [B, n_mels, 1 + T / hop]; the rest is not written. */
NVSE_API int nvse_frontend_mel_ragged_f32(const nvse_frontend* fe, const float* y, int64_t B, int64_t T, int64_t y_row_stride,
                                 const int32_t* samples_dev, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Generator: Models.HiFiGAN / Models.iSTFTNet forward
 *   (Models/hifigan.py:83-133, Models/istftnet.py:271-328), hyper-parameters as in
 *   cfgs/hifigan_v1_config.json:28-33 and cfgs/istftnet_config.json:28-35.
 * ---------------------------------------------------------------------------------- */
#define NVSE_MAX_UPS 8
#define NVSE_MAX_KERNELS 8
#define NVSE_MAX_DILATIONS 8

typedef struct nvse_generator_config {
  int32_t kind;                 /* NVSE_GEN_HIFIGAN | NVSE_GEN_ISTFTNET  (h.model_name) */
  int32_t in_channels;          /* 80, hard-coded at hifigan.py:89 */
  int32_t initial_channel;      /* h.upsample_initial_channel */
  int32_t num_upsamples;        /* len(h.upsample_rates) */
  int32_t upsample_rates[NVSE_MAX_UPS];
  int32_t upsample_kernel_sizes[NVSE_MAX_UPS];
  int32_t resblock_type;        /* 1 | 2  (h.resblock) */
  int32_t num_kernels;          /* len(h.resblock_kernel_sizes) */
  int32_t resblock_kernel_sizes[NVSE_MAX_KERNELS];
  int32_t num_dilations[NVSE_MAX_KERNELS];
  int32_t resblock_dilations[NVSE_MAX_KERNELS][NVSE_MAX_DILATIONS];
  int32_t istft_n_fft;          /* h.gen_istft_n_fft  (iSTFTNet only) */
  int32_t istft_hop;            /* h.gen_istft_hop_size */
} nvse_generator_config;

typedef struct nvse_generator nvse_generator;

NVSE_API int nvse_generator_create(const nvse_generator_config* cfg, nvse_generator** out);
NVSE_API int nvse_generator_destroy(nvse_generator* g);

/* Load one FOLDED fp32 tensor (weight-norm already removed, see nvse_weight_norm_fold_f32)
 * by its reference state-dict name after remove_weight_norm(): "conv_pre.weight",
 * "conv_pre.bias", "ups.<i>.weight", "resblocks.<n>.convs1.<m>.weight", "conv_post.bias"...
 * PyTorch layouts: Conv1d [Cout,Cin,k], ConvTranspose1d [Cin,Cout,k], bias [Cout].
 * The data is copied/re-packed into the handle on `stream`. */
NVSE_API int nvse_generator_set_weight(nvse_generator* g, const char* name, const float* data,
                              const int64_t* shape, int ndim, void* stream);
/* Verifies every tensor was set and builds the bf16 tensor-core weight images. */
NVSE_API int nvse_generator_finalize(nvse_generator* g, void* stream);

/* Batched alternative to set_weight x N + finalize: load EVERY layer in two launches, straight from the module's
 * parameters (what a training step does once per optimiser step).  Arrays of nvse_generator_num_layers() device
 * pointers in the order of nvse_generator_layer_name(): weight[i] = weight_v (weight_g[i] = weight_g, folded on the
 * fly exactly like nvse_weight_norm_fold_f32) or the plain weight (weight_g[i] = NULL), bias[i]; fp32, contiguous,
 * PyTorch layouts.  with_train != 0 also builds the transposed copies the backward reads.  Leaves the handle finalized. */
NVSE_API int nvse_generator_num_layers(const nvse_generator* g);
NVSE_API int nvse_generator_layer_name(const nvse_generator* g, int index, char* out, size_t capacity);
NVSE_API int nvse_generator_load_weights(nvse_generator* g, const float* const* weight, const float* const* weight_g,
                                const float* const* bias, int n_layers, int with_train, void* stream);

/* samples out per mel frame in: prod(upsample_rates) (* istft_hop for iSTFTNet) */
NVSE_API int64_t nvse_generator_out_samples(const nvse_generator* g, int64_t frames);
NVSE_API size_t nvse_generator_workspace_bytes(const nvse_generator* g, int64_t B, int64_t frames, int precision);
/* mel: [B, in_channels, frames] fp32 (the reference's layout); out: [B, out_samples] fp32.
 * precision: NVSE_PRECISION_F32 (CUDA-core fp32 everywhere) or NVSE_PRECISION_BF16
 * (tcgen05 bf16 operands / fp32 accumulate for the upsampling and MRF convolutions,
 * fp32 residual stream, fp32 conv_pre / conv_post). */
NVSE_API int nvse_generator_forward(nvse_generator* g, const float* mel, int64_t B, int64_t frames, float* out,
                           void* workspace, size_t workspace_bytes, int precision, void* stream);
/* Ragged batch (the reference's inference loop, infers/inference_hifigan.py:67-95, runs utterances of different lengths one
 * at a time): mel [B, in_channels, frames] padded to the longest utterance, frames_dev [B] int32 ON THE DEVICE = the mel
 * frames of each utterance (1 <= frames_dev[b] <= frames).  Utterance b is computed exactly -- bit for bit -- as if it were
 * passed alone with frames_dev[b] frames: every kernel treats its rows beyond that length like rows beyond the end of the
 * sequence.  out [B, out_samples(frames)]; the samples of utterance b beyond out_samples(frames_dev[b]) are undefined.
 * NVSE_PRECISION_BF16 (the fused tensor-core plan) only; HiFiGAN and iSTFTNet (whose reflection-padded conv_post and iSTFT
 * overlap-add stop at every utterance's own last frame). */
NVSE_API int nvse_generator_forward_ragged(nvse_generator* g, const float* mel, int64_t B, int64_t frames, const int32_t* frames_dev,
                                  float* out, void* workspace, size_t workspace_bytes, int precision, void* stream);
/* The same forward with the waveform delivered as PCM_16 (int16 [B, out_samples]) instead of float: the quantisation of
 * sf.write(..., 'PCM_16') (infers/inference_hifigan.py:93-95: round(x * 32767), clipped) fused into the last kernel
 * where that kernel is conv_post of the tensor-core plan, one extra pass otherwise.  Bit-identical to
 * nvse_pcm16_from_f32 applied to the output of nvse_generator_forward. */
NVSE_API int nvse_generator_forward_pcm16(nvse_generator* g, const float* mel, int64_t B, int64_t frames, int16_t* out,
                                 void* workspace, size_t workspace_bytes, int precision, void* stream);

/* ------------------------------------------------------------------------------------
 * Layer-level entry points: the same kernels the generator launches, exposed so the
 * parity tests can check each shape class in isolation.
 * Activations here are CHANNELS-LAST fp32 [B, T, C] (the library's internal layout).
 * ---------------------------------------------------------------------------------- */

/* torch.nn.utils.weight_norm fold (remove_weight_norm, hifigan.py:52-56,126-133):
 * w[r,:] = g[r] * v[r,:] / ||v[r,:]||_2,  v: [rows, cols]. */
NVSE_API int nvse_weight_norm_fold_f32(const float* v, const float* g, float* w, int64_t rows, int64_t cols, void* stream);

/* Float waveform -> PCM_16 samples exactly as the reference's output step writes them
 * (sf.write(path, audio, sr, 'PCM_16'), infers/inference_hifigan.py:93-95: libsndfile's lrint(x * 0x7FFF)),
 * so a host pipeline copies half the bytes back.  x, y: device pointers, n samples. */
NVSE_API int nvse_pcm16_from_f32(const float* x, int16_t* y, int64_t n, void* stream);

/* [B, C, T] -> [B, T, C] and back (module-boundary layout changes). */
NVSE_API int nvse_transpose_bct_to_btc_f32(const float* x, float* y, int64_t B, int64_t C, int64_t T, void* stream);
NVSE_API int nvse_transpose_btc_to_bct_f32(const float* x, float* y, int64_t B, int64_t T, int64_t C, void* stream);

/* y = [accumulate ? y : 0] + out_scale * ( conv1d(lrelu(x, in_slope), w, dilation, "same") + bias [+ residual] )
 * w: [Cout, Cin, k] (PyTorch Conv1d layout), k odd.  in_slope = 1 disables the activation.
 * Covers Conv1d + the fused leaky_relu / residual / MRF-average of hifigan.py:45-49,113-119. */
NVSE_API int nvse_conv1d_f32(const float* x, const float* w, const float* bias, const float* residual, float* y,
                    int64_t B, int64_t T, int Cin, int Cout, int k, int dilation,
                    float in_slope, float out_scale, int accumulate, void* stream);
/* Same contract on the tcgen05 path (bf16 operands, fp32 accumulate).  Cin, Cout multiples of 16. */
NVSE_API int nvse_conv1d_bf16(const float* x, const float* w, const float* bias, const float* residual, float* y,
                     int64_t B, int64_t T, int Cin, int Cout, int k, int dilation,
                     float in_slope, float out_scale, int accumulate, void* stream);

/* One fused launch of a ResBlock1 chain (hifigan.py:43-50) on the tcgen05 path:
 *   for m in 0..npairs-1:  xt = conv1d(lrelu(x, 0.1), w1[m], dilation[m]) + b1[m]
 *                          xt = conv1d(lrelu(xt, 0.1), w2[m], 1) + b2[m];   x = xt + x
 *   y = [accumulate ? y : 0] + out_scale * x
 * w1, b1, w2, b2: HOST arrays of npairs DEVICE pointers, each weight [C, C, k] (PyTorch Conv1d
 * layout), each bias [C].  bf16 operands, fp32 accumulate, fp32 residual stream kept in tensor
 * memory.  Returns NVSE_ERR_UNSUPPORTED for shapes the fused kernel does not take (the generator
 * then falls back to per-layer launches): C in {32, 64, 128, 256}, k odd <= 15, npairs <= 3. */
NVSE_API int nvse_resblock1_bf16(const float* x, const float* const* w1, const float* const* b1,
                        const float* const* w2, const float* const* b2, const int* dilations, int npairs,
                        float* y, int64_t B, int64_t T, int C, int k, float out_scale, int accumulate, void* stream);

/* y = conv_transpose1d(lrelu(x, in_slope), w, stride, padding) + bias;  w: [Cin, Cout, k]
 * (hifigan.py:93-96,111-112).  T_out = (T-1)*stride - 2*padding + k. */
NVSE_API int nvse_conv_transpose1d_f32(const float* x, const float* w, const float* bias, float* y,
                              int64_t B, int64_t T, int Cin, int Cout, int k, int stride, int padding,
                              float in_slope, void* stream);
NVSE_API int nvse_conv_transpose1d_bf16(const float* x, const float* w, const float* bias, float* y,
                               int64_t B, int64_t T, int Cin, int Cout, int k, int stride, int padding,
                               float in_slope, void* stream);

/* iSTFT head, istftnet.py:314-316,183-188:  z: [B, Tp, n_fft+2] channels-last conv_post output;
 * mag = exp(z[..., :n_fft/2+1]), phase = sin(z[..., n_fft/2+1:]); out: [B, hop*(Tp-1)].
 * Supported: n_fft in {4..64} even, hop dividing n_fft. */
NVSE_API int nvse_istft_head_f32(const float* z, float* out, int64_t B, int64_t Tp, int n_fft, int hop, void* stream);

/* ---- training path: backward of HiFiGAN.forward / iSTFTNet.forward -------
 * Replaces what loss_gen_all.backward() runs through the generator in train_time_wi_inv.py:222-236
 * (autograd over Models/hifigan.py:108-124 / Models/istftnet.py:299-318).  Gradients are taken w.r.t. the FOLDED tensors loaded
 * with nvse_generator_set_weight; weight_norm's own backward is nvse_weight_norm_backward_f32.
 *
 * forward_train = the fp32 forward, keeping the input of every convolution on the caller-owned `tape`
 * (nvse_generator_tape_bytes).  backward consumes the tape, the forward's output and dL/dout [B, samples]
 * and writes every layer's (dW, dbias) into the flat buffer `grads` (nvse_generator_grad_elems floats;
 * nvse_generator_grad_offset maps "<layer>.weight" / "<layer>.bias" to its slice, PyTorch layouts), and
 * dL/dmel [B, 80, frames] when dmel is not null.  Bit-reproducible (no atomics).
 * precision: NVSE_PRECISION_F32 = every gradient on the fp32 CUDA cores (parity gate: the fp32 reference);
 * NVSE_PRECISION_BF16 = the MRF convolutions (96 % of the FLOPs) on the tcgen05 tensor cores in the forward, the data
 * gradients and the weight gradients, with bf16 operands and fp32 accumulation (mixed-precision training arithmetic;
 * activations, gradients and the tape stay fp32); conv_pre, the upsamplers and conv_post stay on the fp32 path. */
NVSE_API size_t nvse_generator_tape_bytes(const nvse_generator* g, int64_t B, int64_t frames);
NVSE_API size_t nvse_generator_backward_workspace_bytes(const nvse_generator* g, int64_t B, int64_t frames);
NVSE_API int64_t nvse_generator_grad_elems(const nvse_generator* g);
NVSE_API int nvse_generator_grad_offset(const nvse_generator* g, const char* name, int64_t* offset, int64_t* numel);
NVSE_API int nvse_generator_forward_train(nvse_generator* g, const float* mel, int64_t B, int64_t frames, float* out,
                                 void* tape, size_t tape_bytes, int precision, void* stream);
NVSE_API int nvse_generator_backward(nvse_generator* g, int64_t B, int64_t frames, const float* out, const float* dout,
                            const void* tape, size_t tape_bytes, float* grads, float* dmel, void* workspace,
                            size_t workspace_bytes, int precision, void* stream);

/* dataset.amp_pha_specturm (dataset.py:124-139): the STFT of the front-end (torch.stft(y, n_fft, hop, win, hann,
 * center=True)) written out as log(|X| + 1e-7), atan2(Im, Re), Re, Im -- each [B, n_fft/2+1, frames]; any plane may be
 * NULL.  The handle's mel basis is not used. */
NVSE_API int nvse_frontend_stft_f32(const nvse_frontend* fe, const float* y, int64_t B, int64_t T, int64_t y_row_stride,
                           float* log_amp, float* phase, float* real, float* imag, void* stream);
/* dataset.inverse_mel (dataset.py:94-121): out[b, k, f] = sum_m inv_basis[k, m] * exp(mel[b, m, f]);
 * inv_basis [n_bins, n_mels] = pinverse of the mel basis (device), mel [B, n_mels, frames], out [B, n_bins, frames]. */
NVSE_API int nvse_inverse_mel_f32(const float* inv_basis, const float* mel, float* out, int64_t B, int n_bins, int n_mels,
                         int64_t frames, void* stream);

/* Inverse STFT at n_fft = 1024, the head of the reference's T-F vocoders (Models/apnet.py:155, Models/freeV.py:178,
 * Models/bsrnn.py:210): torch.istft(complex(real, imag), n_fft, hop, win, window, center=True) with the handle's window
 * and hop.  real / imag [B, n_fft/2+1, frames] -> out [B, hop * (frames - 1)].  scratch: caller-owned, at least
 * nvse_frontend_istft_scratch_bytes (the windowed frames before the overlap-add).  Bit-reproducible. */
NVSE_API size_t nvse_frontend_istft_scratch_bytes(const nvse_frontend* fe, int64_t B, int64_t frames);
NVSE_API int nvse_frontend_istft_f32(const nvse_frontend* fe, const float* real, const float* imag, int64_t B, int64_t frames,
                            float* out, void* scratch, size_t scratch_bytes, void* stream);
/* The same from a complex64 spectrum [B, n_fft/2+1, frames] (interleaved re, im: the tensor torch.istft takes, 8-byte aligned):
 * no de-interleaving pass in front of the kernel. */
NVSE_API int nvse_frontend_istft_c64(const nvse_frontend* fe, const float* spec_c64, int64_t B, int64_t frames, float* out,
                            void* scratch, size_t scratch_bytes, void* stream);

/* Backward of nvse_frontend_mel_f32 (the mel-L1 term of the generator loss differentiates
 * mel_spectrogram(y_g_hat), train_time_wi_inv.py:173-179,231-235): dmel [B, n_mels, frames] -> dy [B, T] (dense).
 * The spectra are recomputed from y (nothing is saved by the forward).  Bit-reproducible. */
NVSE_API size_t nvse_frontend_backward_scratch_bytes(const nvse_frontend* fe, int64_t B, int64_t T);
NVSE_API int nvse_frontend_mel_backward_f32(const nvse_frontend* fe, const float* y, int64_t B, int64_t T,
                                   int64_t y_row_stride, const float* dmel, float* dy, void* scratch,
                                   size_t scratch_bytes, void* stream);

/* Layer-level backward (channels-last fp32, the layouts of nvse_conv1d_f32 / nvse_conv_transpose1d_f32):
 *   y = conv1d(lrelu(x, in_slope), w, dilation, "same") + bias [+ residual]
 *   dx = lrelu'(x) * conv^T(dy) [+ dresidual_in],  dw [Cout,Cin,k],  dbias [Cout];  any of dx/dw/dbias may be null.
 *   precision = NVSE_PRECISION_BF16: dw on the tensor cores (bf16 operands) where the shape allows, else fp32. */
NVSE_API int nvse_conv1d_backward_f32(const float* x, const float* w, const float* dy, const float* dresidual_in,
                             float* dx, float* dw, float* dbias, int64_t B, int64_t T, int Cin, int Cout, int k,
                             int dilation, float in_slope, int precision, void* stream);
/*   y = conv_transpose1d(lrelu(x, in_slope), w, stride, padding) + bias;  dy: [B, T_out, Cout];  dw [Cin,Cout,k] */
NVSE_API int nvse_conv_transpose1d_backward_f32(const float* x, const float* w, const float* dy, float* dx, float* dw,
                                       float* dbias, int64_t B, int64_t T, int Cin, int Cout, int k, int stride,
                                       int padding, float in_slope, void* stream);
/* weight_norm backward of EVERY layer in one launch (after nvse_generator_load_weights; arrays as there).  For layer i
 * with weight_g[i] != NULL:  dv is written at dv_flat + (offset of "<layer>.weight" in `grads`), dg at dg_flat + (rows of
 * the layers before it; nvse_generator_total_rows floats in all).  Layers without weight norm are skipped. */
NVSE_API int64_t nvse_generator_total_rows(const nvse_generator* g);
NVSE_API int nvse_generator_weight_norm_backward(nvse_generator* g, const float* const* weight_v, const float* const* weight_g,
                                        const float* grads, float* dv_flat, float* dg_flat, int n_layers, void* stream);
/* Backward of nvse_weight_norm_fold_f32:  dg[r] = <dw[r], v[r]> / ||v[r]||,
 * dv[r] = g[r] / ||v[r]|| * (dw[r] - v[r] * <dw[r], v[r]> / ||v[r]||^2). */
NVSE_API int nvse_weight_norm_backward_f32(const float* v, const float* g, const float* dw, float* dv, float* dg,
                                  int64_t rows, int64_t cols, void* stream);

/* Backward of nvse_istft_head_f32 (autograd through exp / sin / torch.istft, istftnet.py:314-316,183-188):
 * dout [B, hop*(Tp-1)] -> dz [B, Tp, n_fft+2].  n_fft in {4, 8, 16, 32}. */
NVSE_API int nvse_istft_head_backward_f32(const float* z, const float* dout, float* dz, int64_t B, int64_t Tp, int n_fft,
                                 int hop, void* stream);

/* ---- fused wav -> wav call: the loop body of infers/inference_hifigan.py:82-95 for a batch ------------
 *   x = mel_spectrogram(wav, ...);  y = generator(x);  [sf.write(..., 'PCM_16')]
 * in one library call on the 16-bit tensor-core path: the front-end writes its log-mel straight into the channels-last,
 * zero-padded layout conv_pre's tensor-core launch stages from (no [B, 80, F] tensor, no transpose pass), the generator runs
 * as nvse_generator_forward / _pcm16 / _ragged do, bit-identical to calling the two stages separately.
 * wav [B, T] (row stride wav_row_stride); samples_dev: NULL, or [B] int32 on the device = samples per utterance of a padded
 * batch (float output only); exactly one of out [B, out_samples(frames)] / out_pcm16 (same shape, int16) is non-NULL;
 * frames = nvse_frontend_num_frames(fe, T).  workspace: nvse_vocoder_workspace_bytes.  nvse_vocoder_mel_pitch: channels per
 * frame of that internal layout (0 when this generator's conv_pre cannot run on the tensor cores: use the two-call path). */
NVSE_API int nvse_vocoder_mel_pitch(const nvse_generator* g);
NVSE_API size_t nvse_vocoder_workspace_bytes(const nvse_frontend* fe, const nvse_generator* g, int64_t B, int64_t T);
NVSE_API int nvse_vocoder_forward(const nvse_frontend* fe, nvse_generator* g, const float* wav, int64_t B, int64_t T,
                         int64_t wav_row_stride, const int32_t* samples_dev, float* out, int16_t* out_pcm16, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- discriminators: MultiPeriodDiscriminator / MultiScaleDiscriminator -------------------------
 * Replaces the convolution stacks of DiscriminatorP (Models/models.py:15-87: Conv2d (k,1) / (stride,1) over
 * [B, C, T/period, period]) and DiscriminatorS (Models/models.py:187-214: grouped strided Conv1d over [B, C, T]) and what
 * autograd runs through them in train_time_wi_inv.py:188-236.  One operation covers both: a strided grouped convolution
 * along axis 2 of a CHANNELS-FIRST tensor x [B, Cin, L, W] (W = period for DiscriminatorP, 1 for DiscriminatorS) with the
 * following leaky_relu fused:   y = lrelu(conv(x, w, stride, pad, groups) + bias, out_slope),   y [B, Cout, Lo, W],
 * Lo = (L + 2 pad - k) / stride + 1 (nvse_disc_conv_out_len), w [Cout, Cin/groups, k] (the PyTorch Conv1d / Conv2d(k,1)
 * layout), bias [Cout] or NULL, out_slope = 1 for no activation (conv_post).  These layouts ARE the feature maps the
 * reference returns.  fp32, bit-reproducible. */
NVSE_API int64_t nvse_disc_conv_out_len(int64_t L, int k, int stride, int pad);
NVSE_API int nvse_disc_conv_forward_f32(const float* x, const float* w, const float* bias, float* y, int64_t B, int Cin, int Cout,
                               int64_t L, int W, int k, int stride, int pad, int groups, float out_slope, void* stream);
/* Backward of the above: dy [B, Cout, Lo, W] is the gradient w.r.t. the ACTIVATED output y (the leaky_relu derivative is taken
 * from the sign of y); dx [B, Cin, L, W], dw [Cout, Cin/groups, k], dbias [Cout] -- any of the three may be NULL.  scratch:
 * caller-owned, nvse_disc_conv_backward_scratch_bytes (masked gradient, per-phase transposed sub-filters of the data gradient,
 * partial sums of the split weight-gradient reduction, added in a fixed order). */
NVSE_API size_t nvse_disc_conv_backward_scratch_bytes(int64_t B, int Cin, int Cout, int64_t L, int W, int k, int stride, int pad,
                                             int groups);
NVSE_API int nvse_disc_conv_backward_f32(const float* x, const float* w, const float* y, const float* dy, float* dx, float* dw,
                                float* dbias, int64_t B, int Cin, int Cout, int64_t L, int W, int k, int stride, int pad,
                                int groups, float out_slope, void* scratch, size_t scratch_bytes, void* stream);
/* AvgPool1d(k, stride, padding=pad) with count_include_pad (MultiScaleDiscriminator.meanpools, Models/models.py:225-228):
 * x [rows, T] -> y [rows, (T + 2 pad - k) / stride + 1], and its backward dy -> dx [rows, T]. */
NVSE_API int nvse_avgpool1d_f32(const float* x, float* y, int64_t rows, int64_t T, int k, int stride, int pad, void* stream);
NVSE_API int nvse_avgpool1d_backward_f32(const float* dy, float* dx, int64_t rows, int64_t T, int k, int stride, int pad, void* stream);

/* The tensor-core kernels bound every mbarrier wait (a protocol bug must never hang the GPU); a
 * tripped timeout sets a device flag and later tensor-core launches return without computing.
 * *flag = 1 if it is set; reset != 0 clears it.  Synchronises the device. */
NVSE_API int nvse_tc_abort_status(int reset, int* flag);

/* Debug hook of tools/rb_trace.py: with NVSE_RB_TRACE set in the environment the fused ResBlock kernel
 * records clock64() stamps of one CTA's phase boundaries (slots 0..62: MMA warp, 63: cycles spent
 * waiting for weight stages, 64..127: epilogue warp 0); this copies them to the host.  Returns -1
 * when tracing is off. */
NVSE_API int nvse_debug_rb_trace(long long* out128);

#ifdef __cplusplus
}
#endif
#endif /* NVSE_B200_H_ */
