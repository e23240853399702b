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

// ---------------------------------------------------------------------------------------------------------------
// "Pixel-pair" form of a 3x3 / stride 1 / pad 1 convolution with 32 input channels (the stem convolutions conv1_2 and
// conv1_3 of ResNet_vd): an NHWC tensor [N, H, W, 32] IS an NHWC tensor [N, H, W/2, 64] (two neighbouring pixels =
// one 64-channel "pair pixel"), and the convolution becomes a 3x3 convolution over pair pixels with 64 input and
// 2 * Cout output channels whose weight is a re-arrangement of the original one:
//     W2[(p, co), r, t, (q, c)] = W[co, r, s, c]   with  s = 2 (t - 1) + q - p + 1   if 0 <= s <= 2, else 0
// (p = parity of the output pixel inside its pair, q = parity of the input pixel inside its pair, t = tap over pairs).
// Half of W2 is zero, but the layers are memory-bound and this puts them on the 64-channel k-blocks of the tcgen05
// 3x3 kernels (fprop, dgrad, wgrad) instead of the vendor library.  Reference call site:
// example/distill/resnet/models/resnet_vd.py:49-60.
__global__ void __launch_bounds__(kThreads)
pair_weight_expand_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ w2, int cout) {
  // one thread per element of W2 [2*cout][3][3][64]
  const int total = 2 * cout * 9 * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int qc = i & 63, q = qc >> 5, c = qc & 31;
    int rest = i >> 6;
    const int t = rest % 3;
    rest /= 3;
    const int r = rest % 3;
    const int pco = rest / 3;
    const int p = pco / cout, co = pco - p * cout;
    const int s = 2 * (t - 1) + q - p + 1;
    w2[i] = (s >= 0 && s <= 2) ? w[((co * 3 + r) * 3 + s) * 32 + c] : __float2bfloat16(0.f);
  }
}

// dW[co, r, s, c] (+)= sum over p of dW2[(p, co), r, t(s, p), (q(s, p), c)]: every element of dW has exactly two sources
__global__ void __launch_bounds__(kThreads)
pair_weight_fold_kernel(const __nv_bfloat16* __restrict__ dw2, __nv_bfloat16* __restrict__ dw, int cout, int accumulate) {
  const int total = cout * 9 * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i & 31;
    int rest = i >> 5;
    const int s = rest % 3;
    rest /= 3;
    const int r = rest % 3;
    const int co = rest / 3;
    float acc = accumulate ? __bfloat162float(dw[i]) : 0.f;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      // s = 2 (t - 1) + q - p + 1  =>  u = s + p - 1 = 2 (t - 1) + q  with q in {0, 1}
      const int u = s + p - 1;
      const int q = u & 1;                 // two's complement: -1 & 1 == 1
      const int t = ((u - q) >> 1) + 1;
      acc += __bfloat162float(dw2[(((p * cout + co) * 3 + r) * 3 + t) * 64 + q * 32 + c]);
    }
    dw[i] = __float2bfloat16(acc);
  }
}

// stats[j] += a[j] + a[cout + j] (j < cout) for the sum and the sum of squares halves: the BatchNorm statistics of
// the pair-pixel output [.., 2 * cout] folded to the real channels
__global__ void fold_pair_stats_kernel(const float* __restrict__ s2, float* __restrict__ stats, int cout) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < cout) {
    stats[j] += s2[j] + s2[cout + j];
    stats[cout + j] += s2[2 * cout + j] + s2[3 * cout + j];
  }
}

}  // namespace

void pair_weight_expand(const void* w, void* w2, int cout, cudaStream_t s) {
  pair_weight_expand_kernel<<<grid_for((int64_t)2 * cout * 9 * 64), kThreads, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(w), reinterpret_cast<__nv_bfloat16*>(w2), cout);
}
void pair_weight_fold(const void* dw2, void* dw, int cout, bool accumulate, cudaStream_t s) {
  pair_weight_fold_kernel<<<grid_for((int64_t)cout * 9 * 32), kThreads, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dw2), reinterpret_cast<__nv_bfloat16*>(dw), cout, accumulate ? 1 : 0);
}
void fold_pair_stats(const float* s2, float* stats, int cout, cudaStream_t s) {
  fold_pair_stats_kernel<<<(cout + 127) / 128, 128, 0, s>>>(s2, stats, cout);
}

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
