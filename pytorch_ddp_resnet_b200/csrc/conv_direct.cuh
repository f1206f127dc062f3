// Direct (CUDA-core, fp32 accumulate) NHWC convolution kernels. They cover what the tcgen05 path
// does not: the 3-channel stem (K = 27, memory bound), channel counts that are not multiples of 16,
// filters larger than 3x3, and they double as an independent cross-check of the tensor-core kernels.
#pragma once
#include "common.cuh"

namespace b200 {

struct ConvDims {
  int N, H, W, C, K, R, S, stride, pad, P, Q;
};

// dot product over `C` contiguous bf16 channels
__device__ __forceinline__ float dot_bf16(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                          int C, bool vec) {
  float acc = 0.f;
  if (vec) {
    for (int c = 0; c < C; c += 8) {
      Vec8 va, vb;
      va.raw = __ldg(reinterpret_cast<const uint4*>(a + c));
      vb.raw = __ldg(reinterpret_cast<const uint4*>(b + c));
      float fa[8], fb[8];
      va.to_float(fa);
      vb.to_float(fb);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(fa[j], fb[j], acc);
    }
  } else {
    for (int c = 0; c < C; ++c) acc = fmaf(__bfloat162float(a[c]), __bfloat162float(b[c]), acc);
  }
  return acc;
}

// one thread per output element (pixel, k); consecutive threads -> consecutive k
__global__ void conv_fprop_direct_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                         const float* __restrict__ bias,
                                         const bf16* __restrict__ residual, bf16* __restrict__ y,
                                         ConvDims d, int relu) {
  const size_t total = (size_t)d.N * d.P * d.Q * d.K;
  const bool vec = (d.C % 8) == 0;
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
        const bf16* xp = x + (((size_t)n * d.H + ih) * d.W + iw) * d.C;
        const bf16* wp = w + (((size_t)k * d.R + r) * d.S + s) * d.C;
        acc += dot_bf16(xp, wp, d.C, vec);
      }
    }
    if (bias) acc += round_bf16(bias[k]);
    float o = round_bf16(acc);
    if (residual) o = round_bf16(o + __bfloat162float(residual[idx]));
    if (relu) o = fmaxf(o, 0.f);
    y[idx] = __float2bfloat16_rn(o);
  }
}

// one thread per input-gradient element (pixel, c); filter in CRSK order (k contiguous)
__global__ void conv_dgrad_direct_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ wt,
                                         const bf16* __restrict__ addend, bf16* __restrict__ dx,
                                         ConvDims d) {
  const size_t total = (size_t)d.N * d.H * d.W * d.C;
  const bool vec = (d.K % 8) == 0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % d.C);
    size_t pix = idx / d.C;
    const int iw = (int)(pix % d.W);
    const int ih = (int)((pix / d.W) % d.H);
    const int n = (int)(pix / ((size_t)d.W * d.H));
    float acc = 0.f;
    for (int r = 0; r < d.R; ++r) {
      const int hp = ih + d.pad - r;
      if (hp < 0 || (hp % d.stride) != 0) continue;
      const int p = hp / d.stride;
      if (p >= d.P) continue;
      for (int s = 0; s < d.S; ++s) {
        const int wp_ = iw + d.pad - s;
        if (wp_ < 0 || (wp_ % d.stride) != 0) continue;
        const int q = wp_ / d.stride;
        if (q >= d.Q) continue;
        const bf16* dyp = dy + (((size_t)n * d.P + p) * d.Q + q) * d.K;
        const bf16* wp = wt + (((size_t)c * d.R + r) * d.S + s) * d.K;
        acc += dot_bf16(dyp, wp, d.K, vec);
      }
    }
    float o = round_bf16(acc);
    if (addend) o = round_bf16(o + __bfloat162float(addend[idx]));
    dx[idx] = __float2bfloat16_rn(o);
  }
}

// thread per filter element (k, r, s, c) and pixel chunk (blockIdx.y); fp32 atomics across chunks, or (part !=
// nullptr, deterministic mode) one partial per chunk at part + chunk * total for wgrad_reduce_splits_kernel
__global__ void conv_wgrad_direct_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                         float* __restrict__ dw, ConvDims d, int pix_per_chunk,
                                         float* __restrict__ part) {
  const int total = d.K * d.R * d.S * d.C;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % d.C;
  const int s = (idx / d.C) % d.S;
  const int r = (idx / (d.C * d.S)) % d.R;
  const int k = idx / (d.C * d.S * d.R);
  const size_t npix = (size_t)d.N * d.P * d.Q;
  const size_t p0 = (size_t)blockIdx.y * pix_per_chunk;
  const size_t p1 = min(npix, p0 + (size_t)pix_per_chunk);
  float acc = 0.f;
  for (size_t pix = p0; pix < p1; ++pix) {
    const int q = (int)(pix % d.Q);
    const int p = (int)((pix / d.Q) % d.P);
    const int n = (int)(pix / ((size_t)d.Q * d.P));
    const int ih = p * d.stride + r - d.pad;
    const int iw = q * d.stride + s - d.pad;
    if (ih < 0 || ih >= d.H || iw < 0 || iw >= d.W) continue;
    acc = fmaf(__bfloat162float(dy[pix * d.K + k]),
               __bfloat162float(x[(((size_t)n * d.H + ih) * d.W + iw) * d.C + c]), acc);
  }
  if (gridDim.y == 1) dw[idx] = acc;
  else if (part) part[(size_t)blockIdx.y * total + idx] = acc;
  else atomicAdd(dw + idx, acc);
}

// dbias[k] = sum over pixels of dy[pix][k]: block = (8-channel group lanes) x (pixel lanes) over a chunk
// of pixels, 16-byte loads, shared-memory reduction over the pixel lanes, one atomic per channel per block (or,
// part != nullptr: the block's sums stored at part + blockIdx.x * K for wgrad_reduce_splits_kernel)
__global__ void conv_dbias_kernel(const bf16* __restrict__ dy, float* __restrict__ dbias,
                                  size_t npix, int K, int pix_per_chunk, float* __restrict__ part) {
  __shared__ float red[256 * 8];
  const int CG = K / 8;
  const int CGb = min(256, CG);
  const int RP = 256 / CGb;
  const int cg = threadIdx.x % CGb, rl = threadIdx.x / CGb;
  const size_t p0 = (size_t)blockIdx.x * pix_per_chunk;
  const size_t p1 = min(npix, p0 + (size_t)pix_per_chunk);
  for (int cgb = blockIdx.y * CGb; cgb < CG; cgb += gridDim.y * CGb) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int cgi = cgb + cg;
    if (rl < RP && cgi < CG) {
      // 8 independent 16-byte loads in flight per thread (one per iteration ran at ~1 TB/s: 45.9 us for 42 MB)
      for (size_t pix = p0 + rl; pix < p1; pix += 8 * (size_t)RP) {
        Vec8 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const size_t pu = pix + (size_t)u * RP;
          v[u].raw = make_uint4(0u, 0u, 0u, 0u);
          if (pu < p1) v[u].raw = ldg_stream(dy + pu * K + (size_t)cgi * 8);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float f[8];
          v[u].to_float(f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = acc[j];
    __syncthreads();
    if (rl == 0 && cgi < CG) {
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
        for (int r = 0; r < RP; ++r) s += red[(r * CGb + cg) * 8 + j];
        if (part) part[(size_t)blockIdx.x * K + cgi * 8 + j] = s;
        else atomicAdd(dbias + cgi * 8 + j, s);
      }
    }
    __syncthreads();
  }
}

}  // namespace b200
