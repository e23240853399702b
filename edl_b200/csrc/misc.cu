// Small fused ops for sm_100a: rotary position embedding (SURVEY K18 -- no reference counterpart, a
// north-star deliverable), embedding-bag forward / backward for the CTR-DNN model (K16; reference
// graph ops: embedding(is_sparse=True) + sequence_pool(avg), example/ctr/ctr/save_program.py:75-144)
// and the uint8 -> bf16 normalise input kernel (K15; reference: cv2 / DALI CropMirrorNormalize,
// example/distill/resnet/utils/img_tool.py:106-157).
#include "common.cuh"
#include "kernels.h"

namespace edl {
namespace {

constexpr int kThreads = 256;

// x: [T, H, D] bf16 (T tokens, H heads, D head dim, D even), cos/sin: [T, D/2] fp32.
// Rotate-half convention: (x1, x2) -> (x1*cos - x2*sin, x2*cos + x1*sin), x1 = x[..., :D/2].
// `inverse` applies the transposed rotation (backward pass).
__global__ void __launch_bounds__(kThreads)
rope_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ cosv, const float* __restrict__ sinv,
            __nv_bfloat16* __restrict__ y, int64_t T, int H, int D, int inverse) {
  const int half = D / 2;
  const int64_t total = T * H * (int64_t)half;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % half);
    const int64_t th = i / half;
    const int64_t t = th / H;
    const float c = cosv[t * half + j];
    float s = sinv[t * half + j];
    if (inverse) s = -s;
    const int64_t base = th * D;
    const float x1 = __bfloat162float(x[base + j]);
    const float x2 = __bfloat162float(x[base + half + j]);
    y[base + j] = __float2bfloat16(x1 * c - x2 * s);
    y[base + half + j] = __float2bfloat16(x2 * c + x1 * s);
  }
}

// Embedding bag (mean): out[b, :] = mean_{l < L} table[ids[b, l], :]   (fixed bag length L)
// One warp per bag row; D is small (10 in the reference) so lanes cover D with a stride loop.
template <typename TT>
__global__ void __launch_bounds__(kThreads)
embedding_bag_fwd_kernel(const TT* __restrict__ table, const int64_t* __restrict__ ids, TT* __restrict__ out,
                         int64_t B, int L, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (kThreads / 32) + warp;
  if (b >= B) return;
  const float inv = 1.f / (float)L;
  for (int d = lane; d < D; d += 32) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) {
      const int64_t id = ids[b * L + l];
      if constexpr (sizeof(TT) == 2) acc += __bfloat162float(table[id * D + d]);
      else acc += table[id * D + d];
    }
    if constexpr (sizeof(TT) == 2) out[b * D + d] = __float2bfloat16(acc * inv);
    else out[b * D + d] = acc * inv;
  }
}

// dtable[ids[b, l], :] += dout[b, :] / L   (fp32 atomics into a dense gradient table)
template <typename TT>
__global__ void __launch_bounds__(kThreads)
embedding_bag_bwd_kernel(const TT* __restrict__ dout, const int64_t* __restrict__ ids, float* __restrict__ dtable,
                         int64_t B, int L, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (kThreads / 32) + warp;
  if (b >= B) return;
  const float inv = 1.f / (float)L;
  for (int d = lane; d < D; d += 32) {
    float g;
    if constexpr (sizeof(TT) == 2) g = __bfloat162float(dout[b * D + d]) * inv;
    else g = dout[b * D + d] * inv;
    for (int l = 0; l < L; ++l) atomicAdd(&dtable[ids[b * L + l] * D + d], g);
  }
}

// uint8 NHWC [N,H,W,3] -> bf16 NHWC, (x/255 - mean[c]) / std[c], optional horizontal flip per image
__global__ void __launch_bounds__(kThreads)
normalize_u8_kernel(const uint8_t* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t npix, int W, float m0,
                    float m1, float m2, float s0, float s1, float s2, const uint8_t* __restrict__ flip,
                    int64_t pix_per_img) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    int64_t src = p;
    if (flip != nullptr && flip[p / pix_per_img]) {
      const int64_t w = p % W;
      src = p - w + (W - 1 - w);
    }
    const uint8_t* s = x + src * 3;
    __nv_bfloat16* d = y + p * 3;
    d[0] = __float2bfloat16((s[0] * (1.f / 255.f) - m0) / s0);
    d[1] = __float2bfloat16((s[1] * (1.f / 255.f) - m1) / s1);
    d[2] = __float2bfloat16((s[2] * (1.f / 255.f) - m2) / s2);
  }
}

inline int grid_for(int64_t n) {
  int64_t b = (n + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

void rope(const void* x, const float* cosv, const float* sinv, void* y, int64_t T, int H, int D, bool inverse,
          cudaStream_t s) {
  rope_kernel<<<grid_for(T * H * (D / 2)), kThreads, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), cosv, sinv, reinterpret_cast<__nv_bfloat16*>(y), T, H, D,
      inverse ? 1 : 0);
}

void embedding_bag_fwd(const void* table, bool bf16, const int64_t* ids, void* out, int64_t B, int L, int D,
                       cudaStream_t s) {
  const int grid = (int)((B + kThreads / 32 - 1) / (kThreads / 32));
  if (bf16)
    embedding_bag_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(table), ids, reinterpret_cast<__nv_bfloat16*>(out), B, L, D);
  else
    embedding_bag_fwd_kernel<float><<<grid, kThreads, 0, s>>>(reinterpret_cast<const float*>(table), ids,
                                                             reinterpret_cast<float*>(out), B, L, D);
}

void embedding_bag_bwd(const void* dout, bool bf16, const int64_t* ids, float* dtable, int64_t B, int L, int D,
                       cudaStream_t s) {
  const int grid = (int)((B + kThreads / 32 - 1) / (kThreads / 32));
  if (bf16)
    embedding_bag_bwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(dout), ids, dtable, B, L, D);
  else
    embedding_bag_bwd_kernel<float><<<grid, kThreads, 0, s>>>(reinterpret_cast<const float*>(dout), ids, dtable,
                                                             B, L, D);
}

void normalize_u8(const uint8_t* x, void* y, int64_t N, int H, int W, const float* mean, const float* stdv,
                  const uint8_t* flip, cudaStream_t s) {
  const int64_t npix = N * H * W;
  normalize_u8_kernel<<<grid_for(npix), kThreads, 0, s>>>(x, reinterpret_cast<__nv_bfloat16*>(y), npix, W, mean[0],
                                                          mean[1], mean[2], stdv[0], stdv[1], stdv[2], flip,
                                                          (int64_t)H * W);
}

}  // namespace edl
