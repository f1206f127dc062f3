// On-device input pipeline: the per-sample CPU transform chain of the reference
// (resnet/utils/transform_util.py:32-205, composed in resnet/utils/data_util.py:66-113 and applied one
// sample at a time by the DataLoader, data_util.py:218-227) as ONE kernel over a device-resident uint8 dataset:
//
//   ToTensorTransform            u8 HWC -> float CHW, x / 255                       (transform_util.py:32-44)
//   Zero-mean / Standardize      (x - image_mean) [/ image_stddev], per pixel+channel (:47-105)
//   FlipTransform                horizontal flip of the whole image, drawn per sample (:156-166)
//   PaddingTransform             zero or mirror ('reflect') padding by pad_size        (:169-187)
//   RandomCropTransform          crop_size window at a per-sample (top, left)          (:190-205)
//
// in exactly that order, so whitening uses the statistics of the SOURCE pixel, padding pads the whitened
// image and the crop offsets index the padded image. The random draws (flip, top, left) are inputs: the host
// side draws them on the device (no host sync); tests inject them to compare bit for bit with the reference.
// Arithmetic is fp32 with IEEE division, the operation order of the reference => the fp32 output is bit-exact.
#pragma once
#include "common.cuh"

namespace b200 {

struct AugmentArgs {
  const uint8_t* data;     // [M][H][W][C] uint8 (the dataset, device resident)
  const int64_t* index;    // [B] rows of `data` that make up the batch
  const uint8_t* flip;     // [B] 0/1 or nullptr (no FlipTransform)
  const int32_t* top;      // [B] crop offsets into the padded image, or nullptr (offset 0)
  const int32_t* left;
  const float* mean;       // [C][H][W] or nullptr (no whitening)
  const float* stddev;     // [C][H][W] or nullptr (zero-mean whitening only)
  int B, H, W, C;
  int pad, mirror;         // PaddingTransform: pad_size, pad_type == 'mirror'
  int OH, OW;              // output extent (crop_size, or H + 2 pad without a crop)
  int to_tensor;           // 1: divide by 255 (ToTensorTransform); 0: keep 0..255
  float* out_f32;          // [B][C][OH][OW] fp32 (the reference's tensor), or nullptr
  bf16* out_bf16;          // [B][OH][OW][C] bf16 (what the stem kernel consumes), or nullptr
};

__global__ void augment_batch_kernel(const AugmentArgs a) {
  const size_t npix = (size_t)a.B * a.OH * a.OW;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix;
       p += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(p % a.OW);
    const int i = (int)((p / a.OW) % a.OH);
    const int b = (int)(p / ((size_t)a.OW * a.OH));
    // position in the padded image -> position in the (flipped, whitened) image
    int ii = i + (a.top ? a.top[b] : 0) - a.pad;
    int jj = j + (a.left ? a.left[b] : 0) - a.pad;
    bool inside = true;
    if (a.mirror) {   // torch 'reflect': -k -> k, (H-1)+k -> (H-1)-k
      if (ii < 0) ii = -ii;
      if (ii >= a.H) ii = 2 * (a.H - 1) - ii;
      if (jj < 0) jj = -jj;
      if (jj >= a.W) jj = 2 * (a.W - 1) - jj;
    } else {
      inside = ii >= 0 && ii < a.H && jj >= 0 && jj < a.W;
    }
    // the flip acts on the whitened image: column jj of the flipped image is column W-1-jj of the source
    const int js = (a.flip && a.flip[b]) ? (a.W - 1 - jj) : jj;
    const uint8_t* src = a.data + (((size_t)a.index[b] * a.H + ii) * a.W + js) * a.C;
    for (int c = 0; c < a.C; ++c) {
      float v = 0.f;
      if (inside) {
        v = (float)src[c];
        if (a.to_tensor) v = __fdiv_rn(v, 255.f);
        if (a.mean) {
          const size_t s = ((size_t)c * a.H + ii) * a.W + js;
          v = __fsub_rn(v, a.mean[s]);
          if (a.stddev) v = __fdiv_rn(v, a.stddev[s]);
        }
      }
      if (a.out_f32) a.out_f32[(((size_t)b * a.C + c) * a.OH + i) * a.OW + j] = v;
      if (a.out_bf16) a.out_bf16[p * a.C + c] = __float2bfloat16_rn(v);
    }
  }
}

}  // namespace b200
