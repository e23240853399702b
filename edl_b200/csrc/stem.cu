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
// (reference: conv2d_grad through cuDNN).  As a GEMM: dW^T-free form  D[32 co x 27 taps] = dY^T[32 x P] * Patch[P x 27]
// with P = 401k output pixels at batch 32 -- far too small an output for tcgen05 tiles (and K = pixels is gathered
// from three input rows), so the warp-level tensor-core path does it: mma.sync m16n8k16, A = dY^T through
// ldmatrix.trans from the staged dY row, B = the im2col patch built on the fly from the staged input rows (4 16-bit
// shared loads per fragment), 8 MMAs per 16 pixels per warp.  The first version of this kernel used CUDA cores
// (lane = tap, 8 FMAs per shared-memory read): 86-92 us, issue- and latency-bound, and -- because the stem's gradient
// is the LAST one backward produces -- exposed 1 : 1 at the end of the step (profiles/timeline_r2).  This one is bound
// by reading dY and X once.  Partials: shared-memory [tap][co] per CTA, then 16-byte vector reductions into a tap-major
// fp32 workspace; the last CTA converts to bf16 KRSC (+= the gradient bucket) and re-zeroes the workspace.
constexpr int kWgThreads = 128;

EDL_DEVICE void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
EDL_DEVICE void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kWgThreads)
stem_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ ws,
                  int* __restrict__ counter, __nv_bfloat16* __restrict__ dw, int accumulate, int N, int H, int W,
                  int Ho, int Wo, int Wo16, int xrow) {
  extern __shared__ __align__(16) uint8_t wg_smem[];
  // xs[3][xrow] bf16: three input rows; the data start at element 8 (16-byte aligned: vector copies), the elements
  // 5..7 in front of it are pixel -1 (zero), zeros behind the row;  pixel w, channel c  ->  element 8 + 3 w + c
  // dys[Wo16][32] bf16: one row of dY, zero beyond Wo;  red[32 taps][32 co] fp32
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(wg_smem);
  __nv_bfloat16* dys = xs + 3 * xrow;
  float* red = reinterpret_cast<float*>(dys + Wo16 * kCout);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * kCout; i += kWgThreads) red[i] = 0.f;
  for (int i = threadIdx.x; i < 3 * xrow; i += kWgThreads) xs[i] = __float2bfloat16(0.f);   // halo + tail stay zero
  for (int i = threadIdx.x; i < Wo16 * kCout; i += kWgThreads) dys[i] = __float2bfloat16(0.f);   // the tail beyond Wo stays 0
  float acc[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
  // this lane's B-fragment geometry: tap n = nt * 8 + lane / 4 -> (r, sc = 3 s + c); pixel k0 = (lane % 4) * 2
  int boff[4];
  bool bval[4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int tap = nt * 8 + (lane >> 2);
    bval[nt] = tap < kTaps;
    const int t = bval[nt] ? tap : 0;
    boff[nt] = (t / 9) * xrow + (t % 9) + 5 + 6 * ((lane & 3) * 2);   // 8 + 3 (2 p + s - 1) + c = 6 p + (3 s + c) + 5
  }
  const uint32_t dys_addr = static_cast<uint32_t>(__cvta_generic_to_shared(dys));
  // ldmatrix.trans lane address: tile t = lane / 8 -> pixels (t >> 1) * 8 + lane % 8, channels (t & 1) * 8
  const uint32_t a_lane = (uint32_t)((((lane >> 4) & 1) * 8 + (lane & 7)) * kCout + ((lane >> 3) & 1) * 8) * 2u;
  const int rows = N * Ho;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / Ho, ho = row - n * Ho;
    __syncthreads();                                   // the previous row's tiles are no longer read
    {
      // 3 W bf16 per input row = a whole number of 16-byte vectors for even W % 8 == 0 rows (224: 84 vectors); the
      // general case copies the remainder element-wise.  Everything outside [8, 8 + 3 W) was zeroed once below.
      const int vec_per_row = (W % 8 == 0) ? (3 * W) / 8 : 0;       // 16-byte aligned rows only
      for (int i = threadIdx.x; i < 3 * vec_per_row; i += kWgThreads) {
        const int rr = i / vec_per_row, j = i - rr * vec_per_row;
        const int h = 2 * ho + rr - 1;
        int4 v = make_int4(0, 0, 0, 0);
        if (h >= 0 && h < H) v = *reinterpret_cast<const int4*>(x + ((int64_t)n * H + h) * W * 3 + j * 8);
        *reinterpret_cast<int4*>(xs + rr * xrow + 8 + j * 8) = v;
      }
      const int rem0 = vec_per_row * 8, rem = 3 * W - rem0;
      for (int i = threadIdx.x; i < 3 * rem; i += kWgThreads) {
        const int rr = i / rem, j = rem0 + (i - rr * rem);
        const int h = 2 * ho + rr - 1;
        xs[rr * xrow + 8 + j] = (h >= 0 && h < H) ? x[((int64_t)n * H + h) * W * 3 + j] : __float2bfloat16(0.f);
      }
    }
    const int4* src = reinterpret_cast<const int4*>(dy + (int64_t)row * Wo * kCout);
    int4* dst = reinterpret_cast<int4*>(dys);
    for (int i = threadIdx.x; i < Wo * kCout / 8; i += kWgThreads) dst[i] = src[i];
    __syncthreads();
    for (int chunk = warp; chunk * 16 < Wo16; chunk += kWgThreads / 32) {
      const int p0 = chunk * 16;
      uint32_t a[2][4];
      ldmatrix_x4_trans(a[0], dys_addr + (uint32_t)(p0 * kCout) * 2u + a_lane);
      ldmatrix_x4_trans(a[1], dys_addr + (uint32_t)(p0 * kCout + 16) * 2u + a_lane);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const unsigned short* xp = reinterpret_cast<const unsigned short*>(xs) + boff[nt] + 6 * p0;
        uint32_t b0 = 0u, b1 = 0u;
        if (bval[nt]) {
          b0 = (uint32_t)xp[0] | ((uint32_t)xp[6] << 16);            // pixels k0, k0 + 1
          b1 = (uint32_t)xp[48] | ((uint32_t)xp[54] << 16);          // pixels k0 + 8, k0 + 9
        }
        mma_bf16_16816(acc[0][nt], a[0], b0, b1);
        mma_bf16_16816(acc[1][nt], a[1], b0, b1);
      }
    }
  }
  // accumulator (mt, nt, i): co = mt * 16 + lane / 4 + (i >= 2 ? 8 : 0), tap = nt * 8 + (lane % 4) * 2 + (i & 1)
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int co = mt * 16 + (lane >> 2) + ((i & 2) ? 8 : 0);
        const int tap = nt * 8 + (lane & 3) * 2 + (i & 1);
        atomicAdd(&red[tap * kCout + co], acc[mt][nt][i]);
      }
  __syncthreads();
  for (int i = threadIdx.x; i < kTaps * kCout / 4; i += kWgThreads) {
    const float4 v = reinterpret_cast<const float4*>(red)[i];
    red_add_v4(ws + i * 4, v.x, v.y, v.z, v.w);
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
  const int Wo16 = (Wo + 15) / 16 * 16;
  int xrow = 6 * Wo16 + 24;                            // covers index 6 * (Wo16 - 1) + 8 + 5 + 54 ... and 8 + 3 W
  if (xrow < 3 * W + 16) xrow = 3 * W + 16;
  xrow = (xrow + 63) / 8 * 8;
  const size_t smem = (size_t)3 * xrow * 2 + (size_t)Wo16 * kCout * 2 + 32 * kCout * 4;
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
      reinterpret_cast<__nv_bfloat16*>(dw), accumulate ? 1 : 0, N, H, W, Ho, Wo, Wo16, xrow);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// ---------------------------------------------------------------------------------------------------------------------
// im2col of a 7x7 / stride 2 / pad 3 convolution on a 3-channel NHWC image (the ResNeXt / ResNet stem of the teacher):
// row m = (n, ho, wo) of A [M, 160] holds the 7 x 7 x 3 = 147 window elements in (r, s, c) order -- for a fixed filter row
// the 21 elements are CONTIGUOUS in the NHWC image -- followed by 13 zeros, so that the tcgen05 GEMM (K blocks of 64, KRSC
// weights padded to 160 columns) computes the convolution with the folded-BN / ReLU epilogue.  One CTA per output row
// (n, ho): the seven input rows go to shared memory with coalesced 32-bit loads, every thread then assembles 16-byte
// pieces of A.  Replaces the library kernel that took 193 us of the teacher's forward (profiles/teacher_c37.txt).
constexpr int kStem7K = 147, kStem7KP = 160, kStem7Threads = 256;

__global__ void __launch_bounds__(kStem7Threads)
stem7_im2col_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ a, int H, int W, int Ho, int Wo, int rowlen) {
  extern __shared__ __align__(16) unsigned char stem7_smem[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(stem7_smem);        // [7][rowlen]: 9 zeros | 3 W elements | zeros
  __shared__ int koff[kStem7KP];
  const int n = blockIdx.x / Ho, ho = blockIdx.x - n * Ho;
  for (int k = threadIdx.x; k < kStem7KP; k += kStem7Threads) koff[k] = k < kStem7K ? (k / 21) * rowlen + (k % 21) : -1;
  const int words = rowlen / 2;                                             // rowlen is even
  for (int r = 0; r < 7; ++r) {
    const int hi = ho * 2 - 3 + r;
    uint32_t* dst = reinterpret_cast<uint32_t*>(xs + r * rowlen);
    if (hi < 0 || hi >= H) {
      for (int i = threadIdx.x; i < words; i += kStem7Threads) dst[i] = 0u;
      continue;
    }
    const __nv_bfloat16* src = x + ((size_t)n * H + hi) * W * 3;
    // data starts at element 9 (odd): element-wise copies of the 3 W values, the pads as zeros
    for (int i = threadIdx.x; i < rowlen; i += kStem7Threads) {
      const int j = i - 9;
      xs[r * rowlen + i] = (j >= 0 && j < 3 * W) ? src[j] : __float2bfloat16(0.f);
    }
  }
  __syncthreads();
  const size_t m0 = ((size_t)n * Ho + ho) * Wo;
  for (int idx = threadIdx.x; idx < Wo * (kStem7KP / 8); idx += kStem7Threads) {
    const int wo = idx / (kStem7KP / 8), q = idx - wo * (kStem7KP / 8);
    const int base = wo * 6;                                                // (wi + 3) * 3 with wi = 2 wo - 3 + s
    alignas(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int o = koff[q * 8 + e];
      v[e] = o >= 0 ? xs[o + base] : __float2bfloat16(0.f);
    }
    *reinterpret_cast<uint4*>(a + (m0 + wo) * kStem7KP + q * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

const char* stem7_im2col(const void* x, void* a, int N, int H, int W, cudaStream_t stream) {
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  int rowlen = 3 * (W + 6) + 2;                    // index (2 (Wo - 1) + 6) * 3 + 2 + 9 ... stays inside
  rowlen = (rowlen + 1) / 2 * 2;
  const size_t smem = (size_t)7 * rowlen * 2;
  if (smem > 96 * 1024) return "stem7_im2col: image too wide";
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(stem7_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    attr_set = true;
  }
  stem7_im2col_kernel<<<N * Ho, kStem7Threads, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                             reinterpret_cast<__nv_bfloat16*>(a), H, W, Ho, Wo, rowlen);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace edl
