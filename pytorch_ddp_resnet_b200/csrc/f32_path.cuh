// fp32 activation kernels of the fp32 / TF32 precision mode: the reference evaluates WITHOUT autocast
// (resnet/algos/evaluation.py:32-39), i.e. fp32 tensors with TF32 convolutions. These are the kernels around
// the kind::tf32 convolutions (conv_tc.cuh, TF32 = true) for that forward pass: layout change, few-input-channel
// im2col, eval-mode batch norm + ReLU + skip, subsample, pooling, the classifier head and the loss / top-k
// metrics, all on fp32 NHWC activations with NO intermediate rounding. They are plain grid-stride kernels with
// 16-byte accesses: this mode exists for numerical parity (1e-3), the bf16 kernels are the tuned training path.
#pragma once
#include "common.cuh"
#include "conv_direct.cuh"
#include "head_sgd.cuh"

namespace b200 {

// fp32 NCHW -> fp32 NHWC
__global__ void nchw_to_nhwc_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int C, int H,
                                        int W) {
  const size_t total = (size_t)N * H * W * C;
  const size_t hw = (size_t)H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t p = i / C;
    const size_t n = p / hw, r = p % hw;
    y[i] = x[(n * C + c) * hw + r];
  }
}

// col[pix][kk], kk = (r*S + s)*C + c, zero padded to Kpad columns (see im2col_kernel)
__global__ void im2col_f32_kernel(const float* __restrict__ x, float* __restrict__ col, int N, int H, int W, int C,
                                  int R, int S, int stride, int pad, int P, int Q, int Kpad) {
  const size_t total = (size_t)N * P * Q * Kpad;
  const int rsc = R * S * C;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % Kpad);
    const size_t pix = idx / Kpad;
    float v = 0.f;
    if (kk < rsc) {
      const int c = kk % C, s = (kk / C) % S, r = kk / (C * S);
      const int q = (int)(pix % Q);
      const int p = (int)((pix / Q) % P);
      const int n = (int)(pix / ((size_t)Q * P));
      const int ih = p * stride + r - pad, iw = q * stride + s - pad;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = x[(((size_t)n * H + ih) * W + iw) * C + c];
    }
    col[idx] = v;
  }
}

// exact fp32 direct convolution for the shapes the TF32 tensor path does not take (few channels, odd tiles)
__global__ void conv_fprop_direct_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                             const float* __restrict__ bias, const float* __restrict__ residual,
                                             float* __restrict__ y, ConvDims d, int relu) {
  const size_t total = (size_t)d.N * d.P * d.Q * d.K;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % d.K);
    size_t pix = idx / d.K;
    const int q = (int)(pix % d.Q);
    const int p = (int)((pix / d.Q) % d.P);
    const int n = (int)(pix / ((size_t)d.Q * d.P));
    float acc = 0.f;
    for (int r = 0; r < d.R; ++r) {
      const int ih = p * d.stride + r - d.pad;
      if (ih < 0 || ih >= d.H) continue;
      for (int s = 0; s < d.S; ++s) {
        const int iw = q * d.stride + s - d.pad;
        if (iw < 0 || iw >= d.W) continue;
        const float* xp = x + (((size_t)n * d.H + ih) * d.W + iw) * d.C;
        const float* wp = w + (((size_t)k * d.R + r) * d.S + s) * d.C;
        for (int c = 0; c < d.C; ++c) acc = fmaf(__ldg(xp + c), __ldg(wp + c), acc);
      }
    }
    if (bias) acc += bias[k];
    if (residual) acc += residual[idx];
    y[idx] = relu ? fmaxf(acc, 0.f) : acc;
  }
}

struct BnActF32Args {
  const float* x;
  float* y;
  const float* skip;
  const float* mean;
  const float* stat;     // invstd, or the variance when stat_is_var
  const float* gamma;
  const float* beta;
  int N, H, W, C;
  int skip_mode, skip_C;
  int stat_is_var, relu, affine;
  float eps;
};

// y = act( (x - mean) * invstd * gamma + beta [+ skip] ), four channels per thread
__global__ void bn_act_fwd_f32_kernel(const BnActF32Args a) {
  const int C4 = a.C / 4;
  const size_t nvec = (size_t)a.N * a.H * a.W * C4;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(v % C4) * 4;
    const size_t r = v / C4;
    const float4 xv = *reinterpret_cast<const float4*>(a.x + v * 4);
    float f[4] = {xv.x, xv.y, xv.z, xv.w};
    if (a.affine) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        const float is = a.stat_is_var ? rsqrtf(a.stat[c] + a.eps) : a.stat[c];
        const float sc = a.gamma[c] * is;
        f[j] = fmaf(f[j], sc, a.beta[c] - a.mean[c] * sc);
      }
    }
    if (a.skip_mode == 1) {
      const float4 s = *reinterpret_cast<const float4*>(a.skip + v * 4);
      f[0] += s.x; f[1] += s.y; f[2] += s.z; f[3] += s.w;
    } else if (a.skip_mode == 2 && c0 < a.skip_C) {
      const int w = (int)(r % a.W);
      const int h = (int)((r / a.W) % a.H);
      const int n = (int)(r / ((size_t)a.W * a.H));
      const float4 s = *reinterpret_cast<const float4*>(
          a.skip + (((size_t)n * (2 * a.H) + 2 * h) * (2 * a.W) + 2 * w) * a.skip_C + c0);
      f[0] += s.x; f[1] += s.y; f[2] += s.z; f[3] += s.w;
    }
    if (a.relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    *reinterpret_cast<float4*>(a.y + v * 4) = make_float4(f[0], f[1], f[2], f[3]);
  }
}

// y[n,h,w,:] = x[n,2h,2w,:]
__global__ void subsample2_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W,
                                      int C) {
  const int C4 = C / 4;
  const size_t nvec = (size_t)N * H * W * C4;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
    const int cg = (int)(v % C4);
    const size_t pix = v / C4;
    const int w = (int)(pix % W);
    const int h = (int)((pix / W) % H);
    const int n = (int)(pix / ((size_t)W * H));
    const size_t spix = ((size_t)n * (2 * H) + 2 * h) * (2 * W) + 2 * w;
    *reinterpret_cast<float4*>(y + v * 4) = *reinterpret_cast<const float4*>(x + spix * C + (size_t)cg * 4);
  }
}

// average (count_include_pad) or max pooling, four channels per thread
template <bool MAX>
__global__ void pool_fwd_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C,
                                    int k, int stride, int pad, int P, int Q) {
  const int C4 = C / 4;
  const size_t nvec = (size_t)N * P * Q * C4;
  const float inv = 1.f / (float)(k * k);
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
    const int cg = (int)(v % C4);
    const size_t pix = v / C4;
    const int q = (int)(pix % Q);
    const int p = (int)((pix / Q) % P);
    const int n = (int)(pix / ((size_t)Q * P));
    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = MAX ? -INFINITY : 0.f;
    for (int r = 0; r < k; ++r) {
      const int h = p * stride + r - pad;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int w = q * stride + s - pad;
        if (w < 0 || w >= W) continue;
        const float4 t = *reinterpret_cast<const float4*>(x + (((size_t)n * H + h) * W + w) * C + (size_t)cg * 4);
        const float f[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = MAX ? fmaxf(acc[j], f[j]) : acc[j] + f[j];
      }
    }
    if (!MAX) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] *= inv;
    }
    *reinterpret_cast<float4*>(y + v * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// logits[b][o] = sum_i x[b][i] * w[o][i] + bias[o], all fp32; one warp per (b, o)
__global__ void linear_fwd_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ bias, float* __restrict__ y, int B, int I, int O) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * O) return;
  const int b = warp / O, o = warp % O;
  float acc = 0.f;
  for (int i = lane; i < I; i += 32) acc = fmaf(x[(size_t)b * I + i], w[(size_t)o * I + i], acc);
  acc = warp_sum(acc);
  if (lane == 0) y[(size_t)b * O + o] = acc + (bias ? bias[o] : 0.f);
}

// mean cross entropy, top-1 / top-5 error of fp32 logits; one block, fixed-order sums (see ce_topk_kernel)
__global__ void __launch_bounds__(CE_WARPS * 32)
ce_topk_f32_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                   float* __restrict__ out, int B, int O) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  float loss = 0.f, e1 = 0.f, e5 = 0.f;
  for (int b = warp; b < B; b += CE_WARPS) {
    const float* row = logits + (size_t)b * O;
    const int label = (int)labels[b];
    float mx = -INFINITY;
    for (int o = lane; o < O; o += 32) mx = fmaxf(mx, row[o]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    float se = 0.f;
    for (int o = lane; o < O; o += 32) se += expf(row[o] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    const float zl = row[label];
    float rank = 0.f;
    for (int o = lane; o < O; o += 32) {
      const float z = row[o];
      rank += (z > zl || (z == zl && o < label)) ? 1.f : 0.f;
    }
    rank = warp_sum(rank);
    loss += lse - zl;
    e1 += rank >= 1.f ? 1.f : 0.f;
    e5 += rank >= 5.f ? 1.f : 0.f;
  }
  ce_block_sum3(loss, e1, e5, out, B);
}

}  // namespace b200
