// First stem convolution of ResNet_vd (3 -> 32 channels, 3x3, stride 2, pad 1, NHWC bf16) as a direct CUDA-core kernel
// with the train-mode BatchNorm statistics of its output fused in.
//
// Reference call site: conv_bn_layer(input, 32, 3, stride=2) in example/distill/resnet/models/resnet_vd.py:55-60 (cuDNN
// through Paddle).  K = 27 is far too short for a tensor-core tile and the layer is memory-bound: 9.6 MB of input and
// 25.7 MB of output at batch 32, i.e. ~5 us at HBM speed, against 74 us for the library kernel plus ~15 us for the
// separate statistics pass (profiles/kineto_r1_final.txt).  One thread computes one output pixel (all 32 channels in
// registers, weights broadcast from shared memory), CTAs are persistent (grid-stride over pixels); per pixel group a
// 31-shuffle warp transpose-reduce leaves channel c's sum in lane c, accumulated in two registers and flushed once per CTA.
#include "common.cuh"
#include "kernels.h"

namespace edl {
namespace {

constexpr int kCout = 32;
constexpr int kTaps = 27;          // 3 x 3 x 3
constexpr int kThreads = 256;

// lane j receives the sum over the 32 lanes of v[j] (31 shuffles)
EDL_DEVICE float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = hi ? v[i] : v[i + off];
      const float keep = hi ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(kThreads)
stem_conv3x3s2_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                      __nv_bfloat16* __restrict__ y, float* __restrict__ stats, int N, int H, int W, int Ho, int Wo) {
  __shared__ __align__(16) float ws[kTaps][kCout];          // ws[(r*3+s)*3+ci][co]
  __shared__ __align__(16) float sstat[2 * kCout];
  for (int i = threadIdx.x; i < kTaps * kCout; i += kThreads) {
    const int co = i / kTaps, k = i % kTaps;   // KRSC: w[co][r][s][ci], k = (r*3+s)*3+ci
    ws[k][co] = __bfloat162float(w[i]);
  }
  if (threadIdx.x < 2 * kCout) sstat[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float kZero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float t1 = 0.f, t2 = 0.f;                    // lane c: running sum / sum of squares of channel c over this warp's pixels
  const int64_t total = (int64_t)N * Ho * Wo;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t iters = (total + stride - 1) / stride;                  // same trip count in every thread: the warp
  for (int64_t it = 0; it < iters; ++it) {                              // shuffles below need all 32 lanes
    const int64_t p = it * stride + (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool valid = p < total;
    bf16x8 pk[kCout / 8];
#pragma unroll
    for (int q = 0; q < kCout / 8; ++q) pk[q] = pack8(kZero8);
    if (valid) {
      float acc[kCout];
#pragma unroll
      for (int c = 0; c < kCout; ++c) acc[c] = 0.f;
      const int wo = (int)(p % Wo);
      const int ho = (int)((p / Wo) % Ho);
      const int n = (int)(p / ((int64_t)Wo * Ho));
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int h = 2 * ho + r - 1;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int wi = 2 * wo + s - 1;
          if (h < 0 || h >= H || wi < 0 || wi >= W) continue;             // zero padding
          const __nv_bfloat16* px = x + (((int64_t)n * H + h) * W + wi) * 3;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) {
            const float xv = __bfloat162float(px[ci]);
            const float4* wr = reinterpret_cast<const float4*>(ws[(r * 3 + s) * 3 + ci]);
#pragma unroll
            for (int q = 0; q < kCout / 4; ++q) {
              const float4 wv = wr[q];                                     // same address in every thread: broadcast
              acc[4 * q + 0] = fmaf(xv, wv.x, acc[4 * q + 0]);
              acc[4 * q + 1] = fmaf(xv, wv.y, acc[4 * q + 1]);
              acc[4 * q + 2] = fmaf(xv, wv.z, acc[4 * q + 2]);
              acc[4 * q + 3] = fmaf(xv, wv.w, acc[4 * q + 3]);
            }
          }
        }
      }
      __nv_bfloat16* py = y + p * kCout;
#pragma unroll
      for (int q = 0; q < kCout / 8; ++q) {
        float v8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v8[j] = acc[8 * q + j];
        pk[q] = pack8(v8);
        st_vec(py + 8 * q, pk[q]);
      }
    }
    if (stats != nullptr) {
      // statistics of the STORED bf16 values; two passes over the packed registers keep one 32-float array live
      float f[kCout];
#pragma unroll
      for (int q = 0; q < kCout / 8; ++q) {
        float v8[8];
        unpack8(pk[q], v8);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[8 * q + j] = v8[j];
      }
      t1 += warp_transpose_sum(f, lane);                                   // invalid lanes contribute zeros
#pragma unroll
      for (int q = 0; q < kCout / 8; ++q) {
        float v8[8];
        unpack8(pk[q], v8);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[8 * q + j] = v8[j] * v8[j];
      }
      t2 += warp_transpose_sum(f, lane);
    }
  }
  if (stats != nullptr) {
    atomicAdd(&sstat[lane], t1);
    atomicAdd(&sstat[kCout + lane], t2);
    __syncthreads();
    if (threadIdx.x < 2 * kCout / 4) {
      const float4 v = reinterpret_cast<const float4*>(sstat)[threadIdx.x];
      red_add_v4(stats + 4 * threadIdx.x, v.x, v.y, v.z, v.w);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Weight gradient of the same convolution:  dW[co][r][s][c] = sum over (n, ho, wo) of
//     dY[n, ho, wo, co] * X[n, 2 ho + r - 1, 2 wo + s - 1, c]
// (reference: conv2d_grad through cuDNN; round 1 of this repo used the library's wgrad on the side stream).  864
// outputs, 347 M MACs at batch 32: CUDA cores.  One CTA of 4 warps walks output rows; warp w owns output channels
// [8w, 8w + 8), LANE l < 27 owns tap l = (r, s, c) -- per output pixel a lane reads ITS input value (consecutive
// lanes read consecutive floats of the staged input rows) and the warp's 8 dY values (one broadcast 16-byte read)
// and does 8 FMAs.  Partials go to a tap-major fp32 workspace [27][32] with two 16-byte vector reductions per lane,
// the last CTA converts to bf16 KRSC (+= the gradient bucket) and re-zeroes the workspace.
constexpr int kWgThreads = 128;

__global__ void __launch_bounds__(kWgThreads)
stem_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ ws,
                  int* __restrict__ counter, __nv_bfloat16* __restrict__ dw, int accumulate, int N, int H, int W,
                  int Ho, int Wo) {
  extern __shared__ __align__(16) uint8_t wg_smem[];
  // xs[3][3 + 3 W + 3] floats: three input rows with one zero pixel on either side; dys[Wo][32] bf16
  const int xrow = 3 * W + 6;
  float* xs = reinterpret_cast<float*>(wg_smem);
  __nv_bfloat16* dys = reinterpret_cast<__nv_bfloat16*>(wg_smem + ((3 * xrow * 4 + 15) / 16) * 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap = lane < kTaps ? lane : 0;
  const int r = tap / 9, sc = tap % 9;                 // sc = s * 3 + c
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int rows = N * Ho;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / Ho, ho = row - n * Ho;
    __syncthreads();                                   // the previous row's tiles are no longer read
    for (int i = threadIdx.x; i < 3 * xrow; i += kWgThreads) {
      const int rr = i / xrow, j = i - rr * xrow;      // j = 3 + 3 * w + c  (w = -1 .. W)
      const int h = 2 * ho + rr - 1;
      const int wc = j - 3;
      float v = 0.f;
      if (h >= 0 && h < H && wc >= 0 && wc < 3 * W) v = __bfloat162float(x[((int64_t)n * H + h) * W * 3 + wc]);
      xs[i] = v;
    }
    const int4* src = reinterpret_cast<const int4*>(dy + (int64_t)row * Wo * kCout);
    int4* dst = reinterpret_cast<int4*>(dys);
    for (int i = threadIdx.x; i < Wo * kCout / 8; i += kWgThreads) dst[i] = src[i];
    __syncthreads();
    const float* xr = xs + r * xrow + sc;              // + 6 * wo: input column 2 wo + s - 1, channel c
    const __nv_bfloat16* dr = dys + warp * 8;
#pragma unroll 4
    for (int wo = 0; wo < Wo; ++wo) {
      const float xv = xr[6 * wo];
      float g[8];
      unpack8(*reinterpret_cast<const bf16x8*>(dr + wo * kCout), g);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(g[k], xv, acc[k]);
    }
  }
  if (lane < kTaps) {
    float* dst = ws + tap * kCout + warp * 8;
    red_add_v4(dst, acc[0], acc[1], acc[2], acc[3]);
    red_add_v4(dst + 4, acc[4], acc[5], acc[6], acc[7]);
  }
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int old = atomicAdd(counter, 1);
    s_last = old == (int)gridDim.x - 1;
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int i = threadIdx.x; i < kTaps * kCout; i += kWgThreads) {
      const int t = i / kCout, co = i - t * kCout;
      float v = __ldcg(ws + i);
      ws[i] = 0.f;
      __nv_bfloat16* o = dw + co * kTaps + t;          // KRSC: [co][(r*3+s)*3+c]
      if (accumulate) v += __bfloat162float(*o);
      *o = __float2bfloat16(v);
    }
  }
}

}  // namespace

// x: bf16 NHWC [N, H, W, 3], w: bf16 KRSC [32, 3, 3, 3], y: bf16 NHWC [N, Ho, Wo, 32]; stats (optional): fp32 [64],
// += per-channel sum and sum of squares of y (16-byte aligned).
void stem_conv3x3s2(const void* x, const void* w, void* y, float* stats, int N, int H, int W, cudaStream_t s) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)N * Ho * Wo;
  int64_t blocks = (total + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)kNumSMs * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  stem_conv3x3s2_kernel<<<(int)blocks, kThreads, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(w),
      reinterpret_cast<__nv_bfloat16*>(y), stats, N, H, W, Ho, Wo);
}

const char* stem_wgrad(const void* x, const void* dy, float* ws, int* counter, void* dw, bool accumulate, int N, int H,
                       int W, cudaStream_t stream) {
  if ((H & 1) || (W & 1)) return "stem_wgrad: even input sizes only";
  const int Ho = H / 2, Wo = W / 2;
  const size_t smem = ((3 * (3 * W + 6) * 4 + 15) / 16) * 16 + (size_t)Wo * kCout * 2;
  if (smem > 200 * 1024) return "stem_wgrad: image too wide";
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    attr_set = true;
  }
  int grid = kNumSMs * 4;
  if (grid > N * Ho) grid = N * Ho;
  stem_wgrad_kernel<<<grid, kWgThreads, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), ws, counter,
      reinterpret_cast<__nv_bfloat16*>(dw), accumulate ? 1 : 0, N, H, W, Ho, Wo);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace edl
