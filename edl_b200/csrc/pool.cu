// NHWC pooling kernels for sm_100a: max 3x3/s2/p1, avg 2x2/s2 (ceil, exclusive)
// used by the ResNet_vd shortcut, and global average pooling.  Reference call sites are Paddle
// pool2d library ops (example/distill/resnet/models/resnet_vd.py:97-102,183-189,131-132).
// All kernels move 8 channels (128 bit) per thread; the max-pool forward records the arg-max
// window slot in one byte so the backward is an atomics-free gather.
#include "common.cuh"
#include "kernels.h"

namespace edl {
namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
maxpool3x3s2_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                        uint8_t* __restrict__ idx, int N, int H, int W, int C, int Ho, int Wo) {
  const int cv = C / 8;
  const int64_t total = (int64_t)N * Ho * Wo * cv;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(t % cv);
    int64_t p = t / cv;
    int ow = (int)(p % Wo);
    p /= Wo;
    int oh = (int)(p % Ho);
    int n = (int)(p / Ho);
    float best[8];
    int bi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      best[i] = -INFINITY;
      bi[i] = 0;
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int ih = oh * 2 - 1 + kh;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int iw = ow * 2 - 1 + kw;
        if (iw < 0 || iw >= W) continue;
        float f[8];
        unpack8(ld_vec(x + (((int64_t)n * H + ih) * W + iw) * C + c * 8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (f[i] > best[i]) {
            best[i] = f[i];
            bi[i] = kh * 3 + kw;
          }
        }
      }
    }
    const int64_t o = (((int64_t)n * Ho + oh) * Wo + ow) * C + c * 8;
    st_vec(y + o, pack8(best));
    uint2 packed;
    packed.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    packed.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
    *reinterpret_cast<uint2*>(idx + o) = packed;
  }
}

__global__ void __launch_bounds__(kThreads)
maxpool3x3s2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                        __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int Ho,
                        int Wo) {
  const int cv = C / 8;
  const int64_t total = (int64_t)N * H * W * cv;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(t % cv);
    int64_t p = t / cv;
    int iw = (int)(p % W);
    p /= W;
    int ih = (int)(p % H);
    int n = (int)(p / H);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // outputs whose window covers (ih, iw): oh*2-1+kh == ih  =>  oh = (ih+1-kh)/2
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int th = ih + 1 - kh;
      if (th < 0 || (th & 1)) continue;
      int oh = th >> 1;
      if (oh >= Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int tw = iw + 1 - kw;
        if (tw < 0 || (tw & 1)) continue;
        int ow = tw >> 1;
        if (ow >= Wo) continue;
        const int64_t o = (((int64_t)n * Ho + oh) * Wo + ow) * C + c * 8;
        uint2 packed = *reinterpret_cast<const uint2*>(idx + o);
        float g[8];
        unpack8(ld_vec(dy + o), g);
        const int slot = kh * 3 + kw;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          unsigned b = i < 4 ? (packed.x >> (8 * i)) & 0xff : (packed.y >> (8 * (i - 4))) & 0xff;
          if ((int)b == slot) acc[i] += g[i];
        }
      }
    }
    st_vec(dx + (((int64_t)n * H + ih) * W + iw) * C + c * 8, pack8(acc));
  }
}

// 2x2 stride-2 average pooling, ceil_mode, exclusive (divide by the number of valid taps).
__global__ void __launch_bounds__(kThreads)
avgpool2x2_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N,
                      int H, int W, int C, int Ho, int Wo) {
  const int cv = C / 8;
  const int64_t total = (int64_t)N * Ho * Wo * cv;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(t % cv);
    int64_t p = t / cv;
    int ow = (int)(p % Wo);
    p /= Wo;
    int oh = (int)(p % Ho);
    int n = (int)(p / Ho);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    int cnt = 0;
#pragma unroll
    for (int kh = 0; kh < 2; ++kh) {
      int ih = oh * 2 + kh;
      if (ih >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 2; ++kw) {
        int iw = ow * 2 + kw;
        if (iw >= W) continue;
        float f[8];
        unpack8(ld_stream(x + (((int64_t)n * H + ih) * W + iw) * C + c * 8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];
        ++cnt;
      }
    }
    const float inv = 1.f / (float)cnt;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= inv;
    st_vec(y + (((int64_t)n * Ho + oh) * Wo + ow) * C + c * 8, pack8(acc));
  }
}

__global__ void __launch_bounds__(kThreads)
avgpool2x2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N,
                      int H, int W, int C, int Ho, int Wo) {
  const int cv = C / 8;
  const int64_t total = (int64_t)N * H * W * cv;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(t % cv);
    int64_t p = t / cv;
    int iw = (int)(p % W);
    p /= W;
    int ih = (int)(p % H);
    int n = (int)(p / H);
    int oh = ih >> 1, ow = iw >> 1;
    int cnt = ((oh * 2 + 1 < H) ? 2 : 1) * ((ow * 2 + 1 < W) ? 2 : 1);
    float g[8];
    unpack8(ld_vec(dy + (((int64_t)n * Ho + oh) * Wo + ow) * C + c * 8), g);
    const float inv = 1.f / (float)cnt;
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] *= inv;
    st_vec(dx + (((int64_t)n * H + ih) * W + iw) * C + c * 8, pack8(g));
  }
}

// Global average pool: x [N, HW, C] -> y [N, C].  One block per (n, 256-channel tile).
__global__ void __launch_bounds__(kThreads)
gap_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int HW,
               int C) {
  __shared__ float sh[kThreads * 8];
  const int n = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 channel-vectors x 8 row lanes
  const int cvec = blockIdx.x * 32 + tx;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (cvec * 8 < C) {
    for (int r = ty; r < HW; r += 8) {
      float f[8];
      unpack8(ld_stream(x + ((int64_t)n * HW + r) * C + cvec * 8), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[(ty * 8 + i) * 32 + tx] = acc[i];
  __syncthreads();
  if (ty == 0 && cvec * 8 < C) {
    const float inv = 1.f / (float)HW;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += sh[(j * 8 + i) * 32 + tx];
      acc[i] = s * inv;
    }
    st_vec(y + (int64_t)n * C + cvec * 8, pack8(acc));
  }
}

__global__ void __launch_bounds__(kThreads)
gap_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N, int HW,
               int C) {
  const int cv = C / 8;
  const int64_t total = (int64_t)N * HW * cv;
  const float inv = 1.f / (float)HW;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(t % cv);
    int64_t p = t / cv;
    int n = (int)(p / HW);
    float g[8];
    unpack8(ld_vec(dy + (int64_t)n * C + c * 8), g);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] *= inv;
    st_vec(dx + p * C + c * 8, pack8(g));
  }
}

inline int grid_for(int64_t total) {
  int64_t b = (total + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// Zero-insertion ("dilate by 2"): out[n, 2h, 2w, :] = in[n, h, w, :], everything else 0.  The input gradient of a
// stride-2 convolution is the stride-1 input gradient of its zero-inserted output gradient, so the stride-2 layers
// reuse the tcgen05 stride-1 dgrad kernel (conv3x3.cu / gemm_persist.cu) instead of the library's strided dgrad.
__global__ void __launch_bounds__(kThreads)
dilate2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int N, int H, int W, int C) {
  const int cv = C / 8;
  const int64_t total = (int64_t)N * (2 * H) * (2 * W) * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    int64_t r = i / cv;
    const int w2 = (int)(r % (2 * W));
    r /= 2 * W;
    const int h2 = (int)(r % (2 * H));
    const int n = (int)(r / (2 * H));
    bf16x8 v;
    if (((h2 | w2) & 1) == 0) {
      v = ld_stream(in + (((int64_t)n * H + (h2 >> 1)) * W + (w2 >> 1)) * C + c * 8);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) v.v[k] = __floats2bfloat162_rn(0.f, 0.f);
    }
    st_vec(out + i * 8, v);
  }
}

}  // namespace

#define BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BFW(p) reinterpret_cast<__nv_bfloat16*>(p)

void maxpool3x3s2_fwd(const void* x, void* y, uint8_t* idx, int N, int H, int W, int C,
                      cudaStream_t s) {
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  maxpool3x3s2_fwd_kernel<<<grid_for((int64_t)N * Ho * Wo * (C / 8)), kThreads, 0, s>>>(
      BF(x), BFW(y), idx, N, H, W, C, Ho, Wo);
}
void maxpool3x3s2_bwd(const void* dy, const uint8_t* idx, void* dx, int N, int H, int W, int C,
                      cudaStream_t s) {
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  maxpool3x3s2_bwd_kernel<<<grid_for((int64_t)N * H * W * (C / 8)), kThreads, 0, s>>>(
      BF(dy), idx, BFW(dx), N, H, W, C, Ho, Wo);
}
void avgpool2x2_fwd(const void* x, void* y, int N, int H, int W, int C, cudaStream_t s) {
  int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  avgpool2x2_fwd_kernel<<<grid_for((int64_t)N * Ho * Wo * (C / 8)), kThreads, 0, s>>>(
      BF(x), BFW(y), N, H, W, C, Ho, Wo);
}
void avgpool2x2_bwd(const void* dy, void* dx, int N, int H, int W, int C, cudaStream_t s) {
  int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  avgpool2x2_bwd_kernel<<<grid_for((int64_t)N * H * W * (C / 8)), kThreads, 0, s>>>(
      BF(dy), BFW(dx), N, H, W, C, Ho, Wo);
}
void dilate2(const void* x, void* y, int N, int H, int W, int C, cudaStream_t s) {
  dilate2_kernel<<<grid_for((int64_t)N * 4 * H * W * (C / 8)), kThreads, 0, s>>>(BF(x), BFW(y), N, H, W, C);
}
void gap_fwd(const void* x, void* y, int N, int HW, int C, cudaStream_t s) {
  dim3 grid((C / 8 + 31) / 32, N);
  gap_fwd_kernel<<<grid, kThreads, 0, s>>>(BF(x), BFW(y), HW, C);
}
void gap_bwd(const void* dy, void* dx, int N, int HW, int C, cudaStream_t s) {
  gap_bwd_kernel<<<grid_for((int64_t)N * HW * (C / 8)), kThreads, 0, s>>>(BF(dy), BFW(dx), N, HW,
                                                                          C);
}

}  // namespace edl
