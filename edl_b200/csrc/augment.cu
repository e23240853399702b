// Device-side image augmentation: random-resized crop (bilinear) + horizontal flip + normalisation, from decoded
// full-size uint8 RGB images of DIFFERENT sizes to one bf16 NHWC training batch.
//
// Reference: the DALI pipeline of the ImageNet examples -- ops.ImageDecoderRandomCrop(device="mixed") -> ops.Resize ->
// ops.CropMirrorNormalize (example/distill/resnet/dali.py:60-106) -- or its CPU twin (utils/img_tool.py:60-157:
// random_crop -> cv2.resize(INTER_LINEAR) -> flip -> (x/255 - mean)/std).  Here the three steps are one kernel behind the
// nvJPEG batched decode (jpeg_decode.cpp): the decoded images never go back to the host and the cropped / resized
// uint8 intermediate never exists.  Sampling convention = cv2.INTER_LINEAR (half-pixel centres, edge clamp), so the
// CPU loader (utils/image_pipeline.py decode_train) and this path produce the same pixels up to rounding.
//
// One thread = one output pixel (3 channels, 12 source bytes).  The batch is ~77 MB of source at 500x375 and 9.6 MB of
// output at batch 32: a few microseconds of HBM time; no shared memory, no tensor cores.
#include "common.cuh"
#include "kernels.h"

namespace edl {
namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
crop_resize_normalize_kernel(const uint8_t* __restrict__ pool, const AugmentItem* __restrict__ items,
                             __nv_bfloat16* __restrict__ y, int S, float m0, float m1, float m2, float r0, float r1,
                             float r2) {
  const int n = blockIdx.y;
  const AugmentItem it = items[n];
  const uint8_t* __restrict__ img = pool + it.offset;
  const float sx = (float)it.cw / (float)S, sy = (float)it.ch / (float)S;
  for (int p = blockIdx.x * kThreads + threadIdx.x; p < S * S; p += gridDim.x * kThreads) {
    const int oy = p / S, ox0 = p % S;
    const int ox = it.flip ? S - 1 - ox0 : ox0;                 // mirror AFTER the resize = sample the mirrored column
    // source coordinates inside the crop box (cv2.INTER_LINEAR: centre alignment, negative coordinates clamp to 0)
    float fx = ((float)ox + 0.5f) * sx - 0.5f;
    float fy = ((float)oy + 0.5f) * sy - 0.5f;
    int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
    float ax = fx - (float)x0, ay = fy - (float)y0;
    if (x0 < 0) { x0 = 0; ax = 0.f; }
    if (y0 < 0) { y0 = 0; ay = 0.f; }
    int x1 = x0 + 1, y1 = y0 + 1;
    if (x1 >= it.cw) { x1 = it.cw - 1; if (x0 >= it.cw) x0 = it.cw - 1; }
    if (y1 >= it.ch) { y1 = it.ch - 1; if (y0 >= it.ch) y0 = it.ch - 1; }
    const uint8_t* r0p = img + (int64_t)(it.y + y0) * it.pitch;
    const uint8_t* r1p = img + (int64_t)(it.y + y1) * it.pitch;
    const int c0 = (it.x + x0) * 3, c1 = (it.x + x1) * 3;
    const float w00 = (1.f - ax) * (1.f - ay), w01 = ax * (1.f - ay), w10 = (1.f - ax) * ay, w11 = ax * ay;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      v[c] = w00 * r0p[c0 + c] + w01 * r0p[c1 + c] + w10 * r1p[c0 + c] + w11 * r1p[c1 + c];
    __nv_bfloat16* d = y + ((int64_t)n * S * S + p) * 3;
    d[0] = __float2bfloat16((v[0] * (1.f / 255.f) - m0) * r0);
    d[1] = __float2bfloat16((v[1] * (1.f / 255.f) - m1) * r1);
    d[2] = __float2bfloat16((v[2] * (1.f / 255.f) - m2) * r2);
  }
}

}  // namespace

void crop_resize_normalize(const uint8_t* pool, const AugmentItem* items, void* y, int N, int S, const float* mean,
                           const float* stdv, cudaStream_t s) {
  if (N <= 0) return;
  int bx = (S * S + kThreads - 1) / kThreads;
  const int cap = (kNumSMs * 8 + N - 1) / N;                     // ~8 CTAs per SM over the whole batch
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  crop_resize_normalize_kernel<<<dim3(bx, N), kThreads, 0, s>>>(
      pool, items, reinterpret_cast<__nv_bfloat16*>(y), S, mean[0], mean[1], mean[2], 1.f / stdv[0], 1.f / stdv[1],
      1.f / stdv[2]);
}

}  // namespace edl
