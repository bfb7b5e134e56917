// Launch interface of the backward-pass kernels (grad.cu): weight gradients of the tap-list
// convolutions, bias gradients, and the small element-wise derivatives.  The data gradients
// reuse the forward kernels of conv_f32.cu (a dgrad is a convolution with the per-tap
// transposed weights, a mirrored tap list and the leaky_relu derivative in the epilogue).
#pragma once

#include "conv_f32.cuh"

namespace nvse {

// G[j][ca][cb] = sum_{b,t}  act_u(U[b, u_stride*t + off[j], ca]) * act_v(V[b, t, cb]),   t in [0, Tv), U rows outside
// [0, Tu) contribute zero.  Written as  dst[(cb*Ca + ca)*ntaps + j] = scale * G[j][ca][cb]  -- which is the PyTorch
// layout of both weight gradients:
//   Conv1d           dW[co][ci][j]:  U = layer input (ca = ci, u_slope = the fused leaky_relu), V = dy (cb = co),
//                                    off[j] = j*dilation - pad, u_stride = 1
//   ConvTranspose1d  dW[ci][co][j]:  U = dy (ca = co), V = layer input (cb = ci, v_slope), off[j] = j - pad, u_stride = stride
struct WgradArgs {
  const float* U; int64_t u_bstride; int Tu; int Ca; float u_slope;
  const float* V; int64_t v_bstride; int Tv; int Cb; float v_slope;
  int u_stride;
  int ntaps; int off[kMaxTaps];
  float* dst;
  float scale;
  int tc;   // use the tcgen05 kernel (wgrad_tc.cu: bf16 operands, fp32 accumulate) when it supports the shape
};
// scratch floats needed by launch_wgrad for this shape (partial sums of the split reduction, deterministic order)
size_t wgrad_scratch_elems(int Ca, int Cb, int ntaps, int64_t B, int Tv);
int launch_wgrad(const WgradArgs& a, int64_t B, float* scratch, cudaStream_t st);
// dst[(cb*Ca + ca)*ntaps + j] = scale * sum_s partial[s][j][ca][cb], s ascending
int launch_wgrad_reduce(const float* partial, int nsplit, int ntaps, int Ca, int Cb, float* dst, float scale, cudaStream_t st);

// wgrad_tc.cu: the same correlation on the tensor cores (stride-1 U, Cb in {32, 64, 128, 256}, Ca a multiple of 8)
bool wgrad_tc_supported(int Ca, int Cb, int ntaps, const int* off, int u_stride, int64_t B, int Tv);
size_t wgrad_tc_scratch_elems(int Ca, int Cb, int ntaps, int64_t B, int Tv);
int launch_wgrad_tc(const WgradArgs& a, int64_t B, float* scratch, cudaStream_t st);
int wgrad_abort_status(bool reset, unsigned int* flag);
int wgrad_abort_bind(unsigned int* host_word_dev);
int wgrad_abort_clear(cudaStream_t st);

// dst[c] = scale * sum over rows of V[row][c]   (bias gradient; V dense [rows, C]);  scratch: colsum_scratch_elems floats
size_t colsum_scratch_elems(int C, int64_t rows);
int launch_colsum(const float* V, int64_t rows, int C, float* dst, float scale, float* scratch, cudaStream_t st);

// dz = dout * (1 - out^2)   (tanh, hifigan.py:122)
int launch_tanh_bwd(const float* out, const float* dout, float* dz, int64_t n, cudaStream_t st);

// [k][Cin][Cout] -> [k][Cout][Cin]
int launch_transpose_taps(const float* src, float* dst, int k, int Cin, int Cout, cudaStream_t st);

// weight_norm (dim 0) backward:  w = g * v / ||v||  per row r
//   dg[r] = <dw[r], v[r]> / ||v[r]||;   dv[r] = g[r] / ||v[r]|| * (dw[r] - v[r] * <dw[r], v[r]> / ||v[r]||^2)
int launch_weight_norm_bwd(const float* v, const float* g, const float* dw, float* dv, float* dg, int64_t rows, int64_t cols,
                           cudaStream_t st);

}  // namespace nvse
