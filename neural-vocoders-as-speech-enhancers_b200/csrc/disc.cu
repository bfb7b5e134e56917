// Convolutions of the MPD / MSD discriminators (reference Models/models.py:15-113,187-246), forward and backward.
//
// Both discriminator families are stacks of ONE operation: a strided, possibly grouped convolution along one axis of a
// channels-first tensor [B, C, L, W], followed by leaky_relu(0.1):
//   DiscriminatorS  Conv1d(Cin, Cout, k, stride, groups, padding)            over [B, C, T]         (W = 1)
//   DiscriminatorP  Conv2d(Cin, Cout, (k, 1), (stride, 1), padding=(p, 0))   over [B, C, T/period, period]  (W = period)
// The feature maps the reference returns (and feature_loss reads, models.py:604-610) are the layer outputs in exactly that
// layout, so the kernels work in it: no transposes at the module boundary.
//
//   dconv_kernel   "tap list" implicit GEMM: out[b, oc, out_mul*m + out_add, w] = sum_{ic in group, t} Wt[oc][ic][t] *
//                  in[b, ic, s_in*m + off0 + dt*t, w].  Forward conv: s_in = stride, off = t - pad.  Data gradient: one launch
//                  per residue r of the input row modulo the stride (polyphase: taps j = r + stride*t, s_in = 1, off = m_min - t,
//                  output rows stride*m + const), on per-phase transposed sub-filters.
//                  A CTA computes TCO output channels x TP flattened (m, w) positions; the input rows the tile touches are
//                  staged in shared memory DE-INTERLEAVED by (row mod s_in), so that for every tap the 32 lanes of a warp read
//                  32 consecutive floats whatever the stride (conflict-free), weights are staged [ci, t][oc] and read as
//                  warp-broadcast 128-bit loads; each thread holds NCO x NPI accumulators.
//   dwgrad_kernel  dW[oc][ic][j] = sum_{b, m, w} dz[b, oc, m, w] * x[b, ic, stride*m + j - pad, w]: the same tiling with the
//                  roles swapped -- lanes over (ic, j) columns, the reduction runs over staged position chunks; the position
//                  range is split over CTAs and the partial sums are added in a fixed order (bit-reproducible, no atomics).
//   small: leaky_relu derivative mask, bias gradient, AvgPool1d(4, 2, 2) of MultiScaleDiscriminator (models.py:225-228) fwd/bwd.
// fp32 throughout (the reference trains the discriminators in fp32).
#include "common.cuh"

#include <algorithm>

namespace nvse {

namespace {

constexpr int kDThreads = 256;

struct DConvArgs {
  const float* in;
  const float* w;     // [groups * Cout_g][Cin_g][nt]
  const float* bias;  // [groups * Cout_g] or null
  float* out;
  int64_t in_bstride, out_bstride;  // elements per batch item
  int Cin_g, Cout_g, groups;
  int Lin, Lout, W;
  int M;                 // tile rows m in [0, M)
  int s_in, nt, off0, dt;  // input row of tap t at tile row m: s_in * m + off0 + dt * t
  int out_mul, out_add;  // output row of tile row m
  float slope;           // leaky_relu on the output (1 = none)
  int ci_chunk, pitch;   // input channels per shared-memory stage; floats per (channel, phase) of the staged slab
};

template <int NCO>
struct WVec;
template <>
struct WVec<16> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = *reinterpret_cast<const float4*>(p + 4 * q);
      v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
    }
  }
};
template <>
struct WVec<8> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <>
struct WVec<4> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
};
template <>
struct WVec<2> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[2]) {
    const float2 a = *reinterpret_cast<const float2*>(p);
    v[0] = a.x; v[1] = a.y;
  }
};

// n / d for 0 <= n < 2^14 with a precomputed reciprocal m = ceil(2^20 / d), d <= 64: exact (n * (m*d - 2^20) < 2^20), no division
__host__ __device__ __forceinline__ uint32_t fdiv_magic(int d) { return ((1u << 20) + d - 1) / d; }
__device__ __forceinline__ int fdiv(int n, uint32_t magic) { return (int)(((uint64_t)(uint32_t)n * magic) >> 20); }

// 4-byte asynchronous global -> shared copy; `valid` = false writes a zero instead (src-size 0)
__device__ __forceinline__ void cp_async4_zfill(float* dst_smem, const float* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src),
               "r"(valid ? 4 : 0)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NCO, int NPI>
__global__ void __launch_bounds__(kDThreads, 2) dconv_kernel(const DConvArgs a) {
  constexpr int TCO = 8 * NCO, TP = 32 * NPI, TCOP = TCO + 4;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int W = a.W, s = a.s_in, nt = a.nt;
  const int ntiles_co = (a.Cout_g + TCO - 1) / TCO;
  const int g = blockIdx.y / ntiles_co, co0 = (blockIdx.y - g * ntiles_co) * TCO;
  const int64_t b = blockIdx.z;
  const int p0 = blockIdx.x * TP;
  const int m0 = p0 / W, pl0 = p0 - m0 * W;
  const int mlast = min((p0 + TP - 1) / W, a.M - 1);
  const int off_min = a.dt > 0 ? a.off0 : a.off0 - (nt - 1);
  const int nrel = s * (mlast - m0) + nt;  // input rows this tile touches
  const int pitch = a.pitch, cpitch = s * pitch;
  const int xs_floats = (a.ci_chunk * cpitch + 3) & ~3;
  const int buf_floats = xs_floats + a.ci_chunk * nt * TCOP;  // one stage: x slab [ci_chunk][s][pitch] + weights [ci_chunk * nt][TCOP]
  float acc[NCO][NPI];
#pragma unroll
  for (int c = 0; c < NCO; ++c)
#pragma unroll
    for (int i = 0; i < NPI; ++i) acc[c][i] = 0.0f;

  const float* inb = a.in + b * a.in_bstride + (int64_t)g * a.Cin_g * a.Lin * W;
  const float* wg = a.w + (int64_t)(g * a.Cout_g + co0) * a.Cin_g * nt;
  const int row0 = s * m0 + off_min;
  const int nrw = nrel * W;
  const uint32_t magic_w = fdiv_magic(W), magic_s = fdiv_magic(s);
  // stage input channels [ci0, ci0 + cic) into buffer `buf`, asynchronously (out-of-range rows / channels as zeros)
  auto stage = [&](int ci0, int buf) {
    float* xs = smem + buf * buf_floats;
    float* ws = xs + xs_floats;
    const int cic = min(a.ci_chunk, a.Cin_g - ci0);
    // one warp per input channel: (row, w) of the slab is a contiguous run of global memory
    for (int ci = ty; ci < cic; ci += 8) {
      const float* src = inb + (int64_t)(ci0 + ci) * a.Lin * W + (int64_t)row0 * W;
      float* dst = xs + ci * cpitch;
      for (int r = tx; r < nrw; r += 32) {
        const int rel = fdiv(r, magic_w), w = r - rel * W;
        const int q = fdiv(rel, magic_s), ph = rel - q * s;
        const int gr = row0 + rel;
        const bool ok = gr >= 0 && gr < a.Lin;
        cp_async4_zfill(dst + ph * pitch + q * W + w, ok ? src + r : inb, ok);
      }
    }
    const int nk = cic * nt;
    for (int oc = ty; oc < TCO; oc += 8) {  // one warp per output channel: its [cic][nt] weights are contiguous
      const bool ok = co0 + oc < a.Cout_g;
      const float* src = wg + ((int64_t)(ok ? oc : 0) * a.Cin_g + ci0) * nt;
      for (int e = tx; e < nk; e += 32) cp_async4_zfill(ws + e * TCOP + oc, src + e, ok);
    }
    cp_async_commit();
  };
  const int nchunks = (a.Cin_g + a.ci_chunk - 1) / a.ci_chunk;
  stage(0, 0);
  for (int ch = 0; ch < nchunks; ++ch) {
    if (ch + 1 < nchunks) {
      stage((ch + 1) * a.ci_chunk, (ch + 1) & 1);  // the next stage loads while this one is multiplied
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int cic = min(a.ci_chunk, a.Cin_g - ch * a.ci_chunk);
    const float* xs = smem + (ch & 1) * buf_floats;
    const float* ws = xs + xs_floats;
    // taps are walked in slab order (e = row offset above the lowest tap: phase e % s, row e / s of the de-interleaved
    // slab), so the slab pointer advances by `pitch` within a row group and wraps to the next row after s taps
    const int wstep = a.dt > 0 ? TCOP : -TCOP;
    const int wrap = W - (s - 1) * pitch;
    for (int ci = 0; ci < cic; ++ci) {
      const float* xp = xs + ci * cpitch + pl0 + tx;
      const float* wp = ws + (ci * nt + (a.dt > 0 ? 0 : nt - 1)) * TCOP + ty * NCO;
      int ph = 0;
#pragma unroll 2
      for (int e = 0; e < nt; ++e, wp += wstep) {
        float wv[NCO], xv[NPI];
        WVec<NCO>::load(wp, wv);
#pragma unroll
        for (int i = 0; i < NPI; ++i) xv[i] = xp[32 * i];
#pragma unroll
        for (int c = 0; c < NCO; ++c)
#pragma unroll
          for (int i = 0; i < NPI; ++i) acc[c][i] = fmaf(wv[c], xv[i], acc[c][i]);
        if (++ph == s) {
          ph = 0;
          xp += wrap;
        } else {
          xp += pitch;
        }
      }
    }
    __syncthreads();  // this buffer is refilled by the stage issued in the next iteration
  }

  float* outb = a.out + b * a.out_bstride;
#pragma unroll
  for (int i = 0; i < NPI; ++i) {
    const int p = p0 + tx + 32 * i;
    const int m = p / W, w = p - m * W;
    if (m >= a.M) continue;
    const int orow = a.out_mul * m + a.out_add;
    if (orow < 0 || orow >= a.Lout) continue;
#pragma unroll
    for (int c = 0; c < NCO; ++c) {
      const int oc = co0 + ty * NCO + c;
      if (oc >= a.Cout_g) continue;
      const int och = g * a.Cout_g + oc;
      float v = acc[c][i] + (a.bias ? __ldg(a.bias + och) : 0.0f);
      v = v > 0.0f ? v : v * a.slope;
      outb[((int64_t)och * a.Lout + orow) * W + w] = v;
    }
  }
}

// Cout = 1 (conv_post of both discriminators): one warp-lane per output position, the input channels split over the 8 warps
// of a CTA and summed through shared memory in a fixed order.  out[b, 0, m, w] = bias + sum_{ci, t} w[ci][t] x[b, ci, s m + t - pad, w]
__global__ void __launch_bounds__(kDThreads) dconv_cout1_kernel(const DConvArgs a) {
  __shared__ float red[8][33];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int W = a.W, nt = a.nt;
  const int64_t b = blockIdx.y;
  const int p = blockIdx.x * 32 + tx;
  const int m = p / W, w = p - m * W;
  const bool live = m < a.M;
  const float* inb = a.in + b * a.in_bstride;
  float acc = 0.0f;
  if (live) {
    const int r0 = a.s_in * m + a.off0;
    for (int ci = ty; ci < a.Cin_g; ci += 8) {
      const float* xc = inb + (int64_t)ci * a.Lin * W + w;
      const float* wc = a.w + ci * nt;
      for (int t = 0; t < nt; ++t) {
        const int gr = r0 + t;
        if (gr >= 0 && gr < a.Lin) acc = fmaf(__ldg(wc + t), __ldg(xc + (int64_t)gr * W), acc);
      }
    }
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && live) {
    float v = red[0][tx];
#pragma unroll
    for (int q = 1; q < 8; ++q) v += red[q][tx];
    v += a.bias ? __ldg(a.bias) : 0.0f;
    v = v > 0.0f ? v : v * a.slope;
    a.out[b * a.out_bstride + (int64_t)m * W + w] = v;
  }
}

int dconv_pitch(int TP, int W, int nt, int s) { return (TP + W * (2 + (nt - 1) / s)) | 1; }

int launch_dconv(DConvArgs a, int64_t B, cudaStream_t st) {
  NVSE_REQUIRE(a.nt >= 1 && a.nt <= 64, NVSE_ERR_UNSUPPORTED, "discriminator conv: %d taps (supported: 1..64)", a.nt);
  constexpr int TP = 128;
  a.ci_chunk = std::max(1, std::min(std::min(16, a.Cin_g), 96 / a.nt));
  a.pitch = dconv_pitch(TP, a.W, a.nt, a.s_in);
  const int64_t positions = (int64_t)a.M * a.W;
  if (positions <= 0 || B <= 0) return NVSE_OK;
  if (a.Cout_g == 1 && a.groups == 1 && a.dt > 0 && a.out_mul == 1 && a.out_add == 0) {
    dim3 grid1((unsigned)((positions + 31) / 32), (unsigned)B);
    NVSE_REQUIRE(B <= 65535, NVSE_ERR_UNSUPPORTED, "discriminator conv: batch too large");
    ProfScope prof("dconv", a.Cin_g, 1, 2.0 * B * positions * (double)a.Cin_g * a.nt, 4.0 * B * ((double)positions + (double)a.Lin * a.W * a.Cin_g), st);
    dconv_cout1_kernel<<<grid1, kDThreads, 0, st>>>(a);
    NVSE_LAUNCH_CHECK("dconv_cout1_kernel");
    return NVSE_OK;
  }
  const int nco = a.Cout_g >= 128 ? 16 : (a.Cout_g >= 64 ? 8 : (a.Cout_g >= 32 ? 4 : 2));
  const int TCO = 8 * nco;
  if (nco == 16) a.ci_chunk = std::max(1, std::min(a.ci_chunk, 64 / a.nt));
  const size_t smem = 2 * sizeof(float) * (size_t)(((a.ci_chunk * a.s_in * a.pitch + 3) & ~3) + a.ci_chunk * a.nt * (TCO + 4));
  dim3 grid((unsigned)((positions + TP - 1) / TP), (unsigned)(a.groups * ((a.Cout_g + TCO - 1) / TCO)), (unsigned)B);
  NVSE_REQUIRE(B <= 65535 && grid.y <= 65535, NVSE_ERR_UNSUPPORTED, "discriminator conv: grid too large");
  auto go = [&](auto kern) -> int {
    if (smem > 48 * 1024) NVSE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kDThreads, smem, st>>>(a);
    NVSE_LAUNCH_CHECK("dconv_kernel");
    return NVSE_OK;
  };
  ProfScope prof("dconv", a.Cin_g * a.groups, a.Cout_g * a.groups,
                 2.0 * B * positions * a.Cout_g * a.groups * (double)a.Cin_g * a.nt,
                 4.0 * B * ((double)positions * a.Cout_g * a.groups + (double)a.Lin * a.W * a.Cin_g * a.groups), st);
  if (nco == 16) return go(dconv_kernel<16, 4>);
  if (nco == 8) return go(dconv_kernel<8, 4>);
  if (nco == 4) return go(dconv_kernel<4, 4>);
  return go(dconv_kernel<2, 4>);
}

// ---- weight gradient ----------------------------------------------------------------------------
struct DWgradArgs {
  const float* x;   // [B, groups * Cin_g, L, W]
  const float* dz;  // [B, groups * Cout_g, Lo, W]
  float* dst;       // [nsplit][groups * Cout_g][Cin_g * k]
  int64_t x_bstride, dz_bstride;
  int Cin_g, Cout_g, groups, L, Lo, W, k, stride, pad;
  int chunks_per_b, total_chunks, chunks_per_split;
  int pitch, nic_max;
};

constexpr int kWgTPK = 64;   // positions per staged chunk
constexpr int kWgNCJ = 4;    // (ic, j) columns per thread -> 128 columns per CTA

template <int NCO>
__global__ void __launch_bounds__(kDThreads, 2) dwgrad_kernel(const DWgradArgs a) {
  constexpr int TCO = 8 * NCO, TC = 32 * kWgNCJ, OTP = TCO + 4;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int W = a.W, s = a.stride, k = a.k;
  const int ncols = a.Cin_g * k;
  const int ntiles_co = (a.Cout_g + TCO - 1) / TCO;
  const int g = blockIdx.y / ntiles_co, co0 = (blockIdx.y - g * ntiles_co) * TCO;
  const int col0 = blockIdx.x * TC;
  const int ic_first = col0 / k;
  const int nic = min(a.Cin_g - 1, (col0 + TC - 1) / k) - ic_first + 1;
  const int pitch = a.pitch, cpitch = s * pitch;
  const int xs_floats = (a.nic_max * cpitch + 3) & ~3;
  const int buf_floats = xs_floats + kWgTPK * OTP;          // one stage: x slab [nic_max][s][pitch] + dz tile [kWgTPK][OTP]

  int colbase[kWgNCJ];
#pragma unroll
  for (int i = 0; i < kWgNCJ; ++i) {
    const int col = col0 + tx + 32 * i;
    if (col < ncols) {
      const int ic = col / k, j = col - ic * k;
      colbase[i] = (ic - ic_first) * cpitch + (j % s) * pitch + (j / s) * W;
    } else {
      colbase[i] = 0;
    }
  }
  float acc[NCO][kWgNCJ];
#pragma unroll
  for (int c = 0; c < NCO; ++c)
#pragma unroll
    for (int i = 0; i < kWgNCJ; ++i) acc[c][i] = 0.0f;

  const int positions = a.Lo * W;
  const int qn = pitch / W;  // staged rows per phase (every entry a chunk can read is initialised)
  const int nrw = qn * s * W;
  const uint32_t magic_w = fdiv_magic(W), magic_s = fdiv_magic(s);
  const int c_begin = blockIdx.z * a.chunks_per_split, c_end = min(a.total_chunks, c_begin + a.chunks_per_split);
  auto stage = [&](int ch, int buf) {
    float* xs = smem + buf * buf_floats;
    float* dzs = xs + xs_floats;
    const int64_t b = ch / a.chunks_per_b;
    const int pc0 = (ch - (int)b * a.chunks_per_b) * kWgTPK;
    const int row0 = s * (pc0 / W) - a.pad;
    const float* xb = a.x + b * a.x_bstride + (int64_t)(g * a.Cin_g + ic_first) * a.L * W;
    const float* dzb = a.dz + b * a.dz_bstride + (int64_t)(g * a.Cout_g + co0) * positions;
    for (int ic = ty; ic < nic; ic += 8) {
      const float* src = xb + (int64_t)ic * a.L * W + (int64_t)row0 * W;
      float* dst = xs + ic * cpitch;
      for (int r = tx; r < nrw; r += 32) {
        const int rel = fdiv(r, magic_w), w = r - rel * W;
        const int q = fdiv(rel, magic_s), ph = rel - q * s;
        const int gr = row0 + rel;
        const bool ok = gr >= 0 && gr < a.L;
        cp_async4_zfill(dst + ph * pitch + q * W + w, ok ? src + r : xb, ok);
      }
    }
    for (int idx = tid; idx < TCO * kWgTPK; idx += kDThreads) {
      const int oc = idx / kWgTPK, pp = idx - oc * kWgTPK;
      const int p = pc0 + pp;
      const bool ok = p < positions && co0 + oc < a.Cout_g;
      cp_async4_zfill(dzs + pp * OTP + oc, dzb + (ok ? (int64_t)oc * positions + p : 0), ok);
    }
    cp_async_commit();
  };
  if (c_begin < c_end) stage(c_begin, 0);
  for (int ch = c_begin; ch < c_end; ++ch) {
    const int buf = (ch - c_begin) & 1;
    if (ch + 1 < c_end) {
      stage(ch + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int pc0 = (ch % a.chunks_per_b) * kWgTPK;
    const int pl0 = pc0 - (pc0 / W) * W;
    const float* xp = smem + buf * buf_floats + pl0;
    const float* dp = smem + buf * buf_floats + xs_floats + ty * NCO;
#pragma unroll 4
    for (int pp = 0; pp < kWgTPK; ++pp) {
      float dv[NCO], xv[kWgNCJ];
      WVec<NCO>::load(dp + pp * OTP, dv);
#pragma unroll
      for (int i = 0; i < kWgNCJ; ++i) xv[i] = xp[colbase[i] + pp];
#pragma unroll
      for (int c = 0; c < NCO; ++c)
#pragma unroll
        for (int i = 0; i < kWgNCJ; ++i) acc[c][i] = fmaf(dv[c], xv[i], acc[c][i]);
    }
    __syncthreads();
  }
  float* dst = a.dst + (int64_t)blockIdx.z * a.groups * a.Cout_g * ncols;
#pragma unroll
  for (int c = 0; c < NCO; ++c) {
    const int oc = co0 + ty * NCO + c;
    if (oc >= a.Cout_g) continue;
#pragma unroll
    for (int i = 0; i < kWgNCJ; ++i) {
      const int col = col0 + tx + 32 * i;
      if (col < ncols) dst[(int64_t)(g * a.Cout_g + oc) * ncols + col] = acc[c][i];
    }
  }
}

__global__ void dsplit_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dst, int64_t n, int nsplit) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = partial[i];
  for (int z = 1; z < nsplit; ++z) s += partial[(int64_t)z * n + i];
  dst[i] = s;
}

struct WgradPlan {
  int nco, nsplit, chunks_per_b, total_chunks, chunks_per_split, pitch, nic_max;
  size_t smem;
};

WgradPlan wgrad_plan(int64_t B, int Cin_g, int Cout_g, int groups, int Lo, int W, int k, int stride) {
  WgradPlan p;
  p.nco = Cout_g >= 128 ? 16 : (Cout_g >= 64 ? 8 : (Cout_g >= 32 ? 4 : 2));
  const int TCO = 8 * p.nco, TC = 32 * kWgNCJ;
  const int nm_max = kWgTPK / W + 2;
  const int nrel = stride * (nm_max - 1) + k;
  int qn = (nrel + stride - 1) / stride;
  if ((qn * W) % 2 == 0 && W % 2 == 1) ++qn;  // odd pitch where that is possible: fewer bank conflicts between phases
  p.pitch = qn * W;
  p.nic_max = std::min(Cin_g, TC / k + 2);
  p.smem = 2 * sizeof(float) * (size_t)(((p.nic_max * stride * p.pitch + 3) & ~3) + kWgTPK * (TCO + 4));
  p.chunks_per_b = (Lo * W + kWgTPK - 1) / kWgTPK;
  p.total_chunks = (int)(B * p.chunks_per_b);
  const int tiles = ((Cin_g * k + TC - 1) / TC) * groups * ((Cout_g + TCO - 1) / TCO);
  int want = std::max(1, (4 * device_sm_count() + tiles - 1) / tiles);
  want = std::min(want, std::max(1, p.total_chunks));
  p.chunks_per_split = (p.total_chunks + want - 1) / want;
  p.nsplit = p.chunks_per_split > 0 ? (p.total_chunks + p.chunks_per_split - 1) / p.chunks_per_split : 1;
  return p;
}

// ---- small kernels ------------------------------------------------------------------------------
__global__ void dmask_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dz, int64_t n, float slope) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dz[i] = __ldg(y + i) > 0.0f ? __ldg(dy + i) : __ldg(dy + i) * slope;
}

// db[oc] = sum_{b, p} dz[b, oc, p], one CTA per channel, fixed summation order
__global__ void __launch_bounds__(256) dbias_kernel(const float* __restrict__ dz, float* __restrict__ db, int64_t B, int C, int64_t P) {
  __shared__ float red[256];
  const int oc = blockIdx.x;
  float s = 0.0f;
  for (int64_t b = 0; b < B; ++b) {
    const float* row = dz + (b * C + oc) * P;
    for (int64_t p = threadIdx.x; p < P; p += 256) s += __ldg(row + p);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) db[oc] = red[0];
}

// wt[phase r][g * Cin_g + ci][co][t] = w[g * Cout_g + co][ci][r + stride * t]: the sub-filters of the data gradient
__global__ void dphase_weights_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cin_g, int Cout_g, int groups, int k,
                                      int stride) {
  const int64_t n = (int64_t)groups * Cout_g * Cin_g * k;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = (int)(i % k);
  const int ci = (int)((i / k) % Cin_g);
  const int och = (int)(i / ((int64_t)k * Cin_g));
  const int g = och / Cout_g, co = och - g * Cout_g;
  const int r = j % stride, t = j / stride;
  int64_t base = 0;  // elements of the phases before r
  for (int q = 0; q < r; ++q) base += (int64_t)groups * Cin_g * Cout_g * ((k - q + stride - 1) / stride);
  const int ntr = (k - r + stride - 1) / stride;
  wt[base + ((int64_t)(g * Cin_g + ci) * Cout_g + co) * ntr + t] = w[i];
}

__global__ void avgpool1d_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int T, int To, int k, int stride, int pad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * To) return;
  const int64_t r = i / To;
  const int t = (int)(i - r * To);
  float s = 0.0f;
  for (int q = 0; q < k; ++q) {
    const int n = t * stride - pad + q;
    if (n >= 0 && n < T) s += __ldg(x + r * T + n);
  }
  y[i] = s / (float)k;  // count_include_pad = True (the AvgPool1d default)
}

__global__ void avgpool1d_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t rows, int T, int To, int k, int stride, int pad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * T) return;
  const int64_t r = i / T;
  const int n = (int)(i - r * T);
  float s = 0.0f;  // windows t with t*stride - pad <= n <= t*stride - pad + k - 1, ascending t
  int t_lo = n + pad - k + 1;
  t_lo = t_lo <= 0 ? 0 : (t_lo + stride - 1) / stride;
  const int t_hi = min(To - 1, (n + pad) / stride);
  for (int t = t_lo; t <= t_hi; ++t) s += __ldg(dy + r * To + t);
  dx[i] = s / (float)k;
}

int conv_out_len(int64_t L, int k, int stride, int pad) { return (int)((L + 2 * pad - k) / stride + 1); }

int check_conv_shape(int64_t B, int Cin, int Cout, int64_t L, int W, int k, int stride, int pad, int groups) {
  NVSE_REQUIRE(B >= 1 && Cin >= 1 && Cout >= 1 && L >= 1 && W >= 1 && k >= 1 && stride >= 1 && pad >= 0 && groups >= 1,
               NVSE_ERR_INVALID, "discriminator conv: bad shape");
  NVSE_REQUIRE(Cin % groups == 0 && Cout % groups == 0, NVSE_ERR_INVALID, "discriminator conv: channels not divisible by groups");
  NVSE_REQUIRE(k <= 64 && stride <= 8 && W <= 64, NVSE_ERR_UNSUPPORTED, "discriminator conv: k <= 64, stride <= 8, W <= 64 supported");
  NVSE_REQUIRE(L + 2 * pad >= k, NVSE_ERR_INVALID, "discriminator conv: input shorter than the kernel");
  NVSE_REQUIRE((int64_t)std::max(Cin, Cout) * (L + 2 * pad) * W < (int64_t)1 << 31, NVSE_ERR_UNSUPPORTED,
               "discriminator conv: one batch item exceeds 2^31 elements");
  return NVSE_OK;
}

}  // namespace
}  // namespace nvse

using namespace nvse;

extern "C" int64_t nvse_disc_conv_out_len(int64_t L, int k, int stride, int pad) {
  if (k < 1 || stride < 1 || pad < 0 || L + 2 * pad < k) return -1;
  return conv_out_len(L, k, stride, pad);
}

extern "C" int nvse_disc_conv_forward_f32(const float* x, const float* w, const float* bias, float* y, int64_t B, int Cin, int Cout,
                                          int64_t L, int W, int k, int stride, int pad, int groups, float out_slope, void* stream) {
  if (int rc = check_conv_shape(B, Cin, Cout, L, W, k, stride, pad, groups)) return rc;
  NVSE_REQUIRE(x && w && y, NVSE_ERR_INVALID, "nvse_disc_conv_forward_f32: null pointer");
  const int Lo = conv_out_len(L, k, stride, pad);
  DConvArgs a{};
  a.in = x; a.w = w; a.bias = bias; a.out = y;
  a.in_bstride = (int64_t)Cin * L * W;
  a.out_bstride = (int64_t)Cout * Lo * W;
  a.Cin_g = Cin / groups; a.Cout_g = Cout / groups; a.groups = groups;
  a.Lin = (int)L; a.Lout = Lo; a.W = W; a.M = Lo;
  a.s_in = stride; a.nt = k; a.off0 = -pad; a.dt = 1;
  a.out_mul = 1; a.out_add = 0;
  a.slope = out_slope;
  return launch_dconv(a, B, as_stream(stream));
}

extern "C" size_t nvse_disc_conv_backward_scratch_bytes(int64_t B, int Cin, int Cout, int64_t L, int W, int k, int stride, int pad,
                                                        int groups) {
  if (B < 1 || Cin < 1 || Cout < 1 || groups < 1 || Cin % groups || Cout % groups || k < 1 || stride < 1 || L + 2 * pad < k) return 0;
  const int Lo = conv_out_len(L, k, stride, pad);
  const WgradPlan p = wgrad_plan(B, Cin / groups, Cout / groups, groups, Lo, W, k, stride);
  const size_t wn = (size_t)Cout * (Cin / groups) * k;
  return sizeof(float) * ((size_t)B * Cout * Lo * W + wn + (size_t)p.nsplit * wn) + 256;
}

extern "C" int nvse_disc_conv_backward_f32(const float* x, const float* w, const float* y, const float* dy, float* dx, float* dw,
                                           float* dbias, int64_t B, int Cin, int Cout, int64_t L, int W, int k, int stride, int pad,
                                           int groups, float out_slope, void* scratch, size_t scratch_bytes, void* stream) {
  if (int rc = check_conv_shape(B, Cin, Cout, L, W, k, stride, pad, groups)) return rc;
  NVSE_REQUIRE(x && w && y && dy, NVSE_ERR_INVALID, "nvse_disc_conv_backward_f32: null pointer");
  NVSE_REQUIRE(scratch && scratch_bytes >= nvse_disc_conv_backward_scratch_bytes(B, Cin, Cout, L, W, k, stride, pad, groups),
               NVSE_ERR_INVALID, "nvse_disc_conv_backward_f32: scratch too small");
  cudaStream_t st = as_stream(stream);
  const int Lo = conv_out_len(L, k, stride, pad);
  const int Cin_g = Cin / groups, Cout_g = Cout / groups;
  const int64_t ny = (int64_t)B * Cout * Lo * W;
  const size_t wn = (size_t)Cout * Cin_g * k;
  float* dzbuf = reinterpret_cast<float*>(scratch);
  float* wt = dzbuf + ny;
  float* partial = wt + wn;
  const float* dz = dy;
  if (out_slope != 1.0f) {  // dz = dy * leaky_relu'(pre-activation); the sign of y is the sign of the pre-activation (slope > 0)
    NVSE_REQUIRE(out_slope > 0.0f, NVSE_ERR_UNSUPPORTED, "discriminator conv backward: out_slope must be positive");
    dmask_kernel<<<(unsigned)((ny + 255) / 256), 256, 0, st>>>(y, dy, dzbuf, ny, out_slope);
    NVSE_LAUNCH_CHECK("dmask_kernel");
    dz = dzbuf;
  }
  if (dbias) {
    dbias_kernel<<<Cout, 256, 0, st>>>(dz, dbias, B, Cout, (int64_t)Lo * W);
    NVSE_LAUNCH_CHECK("dbias_kernel");
  }
  if (dw) {
    const WgradPlan p = wgrad_plan(B, Cin_g, Cout_g, groups, Lo, W, k, stride);
    DWgradArgs a{};
    a.x = x; a.dz = dz;
    a.dst = p.nsplit > 1 ? partial : dw;
    a.x_bstride = (int64_t)Cin * L * W; a.dz_bstride = (int64_t)Cout * Lo * W;
    a.Cin_g = Cin_g; a.Cout_g = Cout_g; a.groups = groups; a.L = (int)L; a.Lo = Lo; a.W = W; a.k = k; a.stride = stride; a.pad = pad;
    a.chunks_per_b = p.chunks_per_b; a.total_chunks = p.total_chunks; a.chunks_per_split = p.chunks_per_split;
    a.pitch = p.pitch; a.nic_max = p.nic_max;
    const int TCO = 8 * p.nco, TC = 32 * kWgNCJ;
    dim3 grid((unsigned)((Cin_g * k + TC - 1) / TC), (unsigned)(groups * ((Cout_g + TCO - 1) / TCO)), (unsigned)p.nsplit);
    NVSE_REQUIRE(grid.y <= 65535 && grid.z <= 65535, NVSE_ERR_UNSUPPORTED, "discriminator wgrad: grid too large");
    auto go = [&](auto kern) -> int {
      if (p.smem > 48 * 1024) NVSE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
      kern<<<grid, kDThreads, p.smem, st>>>(a);
      NVSE_LAUNCH_CHECK("dwgrad_kernel");
      return NVSE_OK;
    };
    {
      ProfScope prof("dwgrad", Cin, Cout, 2.0 * B * Lo * W * Cout * (double)Cin_g * k,
                     4.0 * B * ((double)Lo * W * Cout + (double)L * W * Cin), st);
      int rc = p.nco == 16 ? go(dwgrad_kernel<16>) : p.nco == 8 ? go(dwgrad_kernel<8>) : (p.nco == 4 ? go(dwgrad_kernel<4>) : go(dwgrad_kernel<2>));
      if (rc) return rc;
    }
    if (p.nsplit > 1) {
      dsplit_reduce_kernel<<<(unsigned)((wn + 255) / 256), 256, 0, st>>>(partial, dw, (int64_t)wn, p.nsplit);
      NVSE_LAUNCH_CHECK("dsplit_reduce_kernel");
    }
  }
  if (dx) {
    NVSE_REQUIRE(k >= stride, NVSE_ERR_UNSUPPORTED, "discriminator conv backward: kernel shorter than the stride");
    dphase_weights_kernel<<<(unsigned)((wn + 255) / 256), 256, 0, st>>>(w, wt, Cin_g, Cout_g, groups, k, stride);
    NVSE_LAUNCH_CHECK("dphase_weights_kernel");
    int64_t base = 0;
    for (int r = 0; r < stride; ++r) {
      const int ntr = (k - r + stride - 1) / stride;
      // input rows li = stride * m + r - pad in [0, L)
      const int m_min = r >= pad ? 0 : (pad - r + stride - 1) / stride;
      const int64_t li_max = L - 1 + pad - r;
      const int64_t m_max = li_max < 0 ? -1 : li_max / stride;
      const int M = (int)(m_max - m_min + 1);
      if (M > 0) {
        DConvArgs a{};
        a.in = dz; a.w = wt + base; a.bias = nullptr; a.out = dx;
        a.in_bstride = (int64_t)Cout * Lo * W; a.out_bstride = (int64_t)Cin * L * W;
        a.Cin_g = Cout_g; a.Cout_g = Cin_g; a.groups = groups;
        a.Lin = Lo; a.Lout = (int)L; a.W = W; a.M = M;
        a.s_in = 1; a.nt = ntr; a.off0 = m_min; a.dt = -1;
        a.out_mul = stride; a.out_add = stride * m_min + r - pad;
        a.slope = 1.0f;
        if (int rc = launch_dconv(a, B, st)) return rc;
      }
      base += (int64_t)groups * Cin_g * Cout_g * ntr;
    }
  }
  return NVSE_OK;
}

extern "C" int nvse_avgpool1d_f32(const float* x, float* y, int64_t rows, int64_t T, int k, int stride, int pad, void* stream) {
  NVSE_REQUIRE(x && y && rows >= 1 && T >= 1 && k >= 1 && stride >= 1 && pad >= 0 && 2 * pad <= k && T + 2 * pad >= k, NVSE_ERR_INVALID,
               "nvse_avgpool1d_f32: bad arguments");
  const int To = conv_out_len(T, k, stride, pad);
  const int64_t n = rows * To;
  avgpool1d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(x, y, rows, (int)T, To, k, stride, pad);
  NVSE_LAUNCH_CHECK("avgpool1d_kernel");
  return NVSE_OK;
}

extern "C" int nvse_avgpool1d_backward_f32(const float* dy, float* dx, int64_t rows, int64_t T, int k, int stride, int pad, void* stream) {
  NVSE_REQUIRE(dy && dx && rows >= 1 && T >= 1 && k >= 1 && stride >= 1 && pad >= 0 && 2 * pad <= k && T + 2 * pad >= k, NVSE_ERR_INVALID,
               "nvse_avgpool1d_backward_f32: bad arguments");
  const int To = conv_out_len(T, k, stride, pad);
  const int64_t n = rows * T;
  avgpool1d_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(dy, dx, rows, (int)T, To, k, stride, pad);
  NVSE_LAUNCH_CHECK("avgpool1d_bwd_kernel");
  return NVSE_OK;
}
