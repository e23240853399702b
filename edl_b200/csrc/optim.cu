// Fused flat-buffer optimizers for sm_100a.
//
// The reference runs one Paddle `momentum` op per parameter tensor (167 launches for ResNet50_vd,
// example/distill/resnet/train_with_fleet.py:106-122) plus separate AMP cast / unscale /
// check_finite ops (utils/fp16_utils.py:86-129).  Here all parameters of one dtype live in ONE
// flat buffer, so the whole optimizer step is a single streaming kernel:
//   g  = grad * grad_scale (+ weight_decay * w)     [bf16 or fp32 grad in]
//   v  = mu * v + g                                 [fp32 momentum]
//   w  = w - lr * v                                 [fp32 master]
//   p  = bf16(w)                                    [bf16 model copy out]
// lr / grad_scale / found_inf are read from device memory so the launch is CUDA-graph replayable
// while the host LR schedule (cosine/piecewise, K8) just updates a device scalar.
#include "common.cuh"
#include "kernels.h"

namespace edl {
namespace {

constexpr int kThreads = 256;

// Per-segment hyper-parameters: decay[i] applies to elements in [seg_start[i], seg_start[i+1]).
// For the common "same decay everywhere" case num_segs == 0 and `wd` is used.
struct SgdArgs {
  const float* lr;          // device scalar
  const float* grad_scale;  // device scalar or nullptr (1.0)
  const int* found_inf;     // device flag or nullptr; non-zero => skip the step
  float momentum;
  float wd;
  int nesterov;
};

template <typename GradT>
__global__ void __launch_bounds__(kThreads)
sgd_momentum_kernel(__nv_bfloat16* __restrict__ param_lp, float* __restrict__ master,
                    float* __restrict__ mom, const GradT* __restrict__ grad,
                    const float* __restrict__ wd_mask, int64_t n, SgdArgs a) {
  if (a.found_inf != nullptr && *a.found_inf != 0) return;
  const float lr = *a.lr;
  const float gs = a.grad_scale ? *a.grad_scale : 1.f;
  const int64_t nvec = n / 8;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < nvec; i += nthreads) {
    const int64_t off = i * 8;
    float g[8];
    if constexpr (sizeof(GradT) == 2) {
      unpack8(ld_stream(grad + off), g);
    } else {
      float4 a0 = *reinterpret_cast<const float4*>(grad + off);
      float4 a1 = *reinterpret_cast<const float4*>(grad + off + 4);
      g[0] = a0.x; g[1] = a0.y; g[2] = a0.z; g[3] = a0.w;
      g[4] = a1.x; g[5] = a1.y; g[6] = a1.z; g[7] = a1.w;
    }
    float4 w0 = *reinterpret_cast<const float4*>(master + off);
    float4 w1 = *reinterpret_cast<const float4*>(master + off + 4);
    float4 v0 = *reinterpret_cast<const float4*>(mom + off);
    float4 v1 = *reinterpret_cast<const float4*>(mom + off + 4);
    float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    float wdm[8];
    if (wd_mask != nullptr) {
      float4 m0 = *reinterpret_cast<const float4*>(wd_mask + off);
      float4 m1 = *reinterpret_cast<const float4*>(wd_mask + off + 4);
      wdm[0] = m0.x; wdm[1] = m0.y; wdm[2] = m0.z; wdm[3] = m0.w;
      wdm[4] = m1.x; wdm[5] = m1.y; wdm[6] = m1.z; wdm[7] = m1.w;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float wd = wd_mask != nullptr ? a.wd * wdm[k] : a.wd;
      float gg = fmaf(wd, w[k], g[k] * gs);
      v[k] = fmaf(a.momentum, v[k], gg);
      float upd = a.nesterov ? fmaf(a.momentum, v[k], gg) : v[k];
      w[k] = fmaf(-lr, upd, w[k]);
    }
    *reinterpret_cast<float4*>(master + off) = make_float4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<float4*>(master + off + 4) = make_float4(w[4], w[5], w[6], w[7]);
    *reinterpret_cast<float4*>(mom + off) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(mom + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
    if (param_lp != nullptr) st_vec(param_lp + off, pack8(w));
  }
  // scalar tail
  for (int64_t i = nvec * 8 + tid; i < n; i += nthreads) {
    float g;
    if constexpr (sizeof(GradT) == 2) g = __bfloat162float(grad[i]);
    else g = grad[i];
    float wd = wd_mask != nullptr ? a.wd * wd_mask[i] : a.wd;
    float w = master[i];
    float gg = fmaf(wd, w, g * gs);
    float v = fmaf(a.momentum, mom[i], gg);
    float upd = a.nesterov ? fmaf(a.momentum, v, gg) : v;
    w = fmaf(-lr, upd, w);
    master[i] = w;
    mom[i] = v;
    if (param_lp != nullptr) param_lp[i] = __float2bfloat16(w);
  }
}

struct AdamArgs {
  const float* lr;
  const float* grad_scale;
  const int* found_inf;
  const float* step;  // device scalar: step count (already incremented), float
  float beta1, beta2, eps, wd;
  int decoupled;  // AdamW
};

template <typename GradT>
__global__ void __launch_bounds__(kThreads)
adam_kernel(__nv_bfloat16* __restrict__ param_lp, float* __restrict__ master,
            float* __restrict__ m, float* __restrict__ v, const GradT* __restrict__ grad,
            int64_t n, AdamArgs a) {
  if (a.found_inf != nullptr && *a.found_inf != 0) return;
  const float lr = *a.lr;
  const float gs = a.grad_scale ? *a.grad_scale : 1.f;
  const float t = *a.step;
  const float bc1 = 1.f - powf(a.beta1, t);
  const float bc2 = 1.f - powf(a.beta2, t);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n; i += nthreads) {
    float g;
    if constexpr (sizeof(GradT) == 2) g = __bfloat162float(grad[i]);
    else g = grad[i];
    g *= gs;
    float w = master[i];
    if (!a.decoupled) g = fmaf(a.wd, w, g);
    float mi = fmaf(a.beta1, m[i], (1.f - a.beta1) * g);
    float vi = fmaf(a.beta2, v[i], (1.f - a.beta2) * g * g);
    float mhat = mi / bc1;
    float vhat = vi / bc2;
    if (a.decoupled) w -= lr * a.wd * w;
    w -= lr * mhat / (sqrtf(vhat) + a.eps);
    master[i] = w;
    m[i] = mi;
    v[i] = vi;
    if (param_lp != nullptr) param_lp[i] = __float2bfloat16(w);
  }
}

template <typename GradT>
__global__ void __launch_bounds__(kThreads)
grad_sqnorm_kernel(const GradT* __restrict__ grad, int64_t n, float* __restrict__ out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  float sq = 0.f;
  if constexpr (sizeof(GradT) == 2) {
    const int64_t nvec = n / 8;
    for (int64_t i = tid; i < nvec; i += nthreads) {
      float g[8];
      unpack8(ld_stream(grad + i * 8), g);
#pragma unroll
      for (int k = 0; k < 8; ++k) sq = fmaf(g[k], g[k], sq);
    }
    for (int64_t i = nvec * 8 + tid; i < n; i += nthreads) {
      const float g = __bfloat162float(grad[i]);
      sq = fmaf(g, g, sq);
    }
  } else {
    for (int64_t i = tid; i < n; i += nthreads) sq = fmaf(grad[i], grad[i], sq);
  }
  __shared__ float sh[kThreads / 32];
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// one warp: global-norm clip factor from the per-rank squared-norm partials (SURVEY K9 "clip")
__global__ void clip_scale_kernel(const float* __restrict__ parts, int nparts, int stride, float max_norm,
                                  float* __restrict__ grad_scale, float* __restrict__ norm_out) {
  float s = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 32) s += parts[(int64_t)i * stride];
  s = warp_sum(s);
  if (threadIdx.x == 0) {
    const float norm = sqrtf(s);
    // a non-finite norm leaves the scale at 1: the found_inf machinery (loss scaling) owns that case
    *grad_scale = (norm == norm && norm > max_norm) ? max_norm / (norm + 1e-6f) : 1.f;
    if (norm_out != nullptr) *norm_out = norm;
  }
}

inline int grid_for(int64_t n_items) {
  int64_t blocks = (n_items + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

void sgd_momentum(void* param_lp, float* master, float* mom, const void* grad, bool grad_is_bf16,
                  const float* wd_mask, int64_t n, const float* lr, const float* grad_scale,
                  const int* found_inf, float momentum, float wd, bool nesterov,
                  cudaStream_t stream) {
  SgdArgs a{lr, grad_scale, found_inf, momentum, wd, nesterov ? 1 : 0};
  int grid = grid_for((n + 7) / 8);
  if (grad_is_bf16)
    sgd_momentum_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(
        reinterpret_cast<__nv_bfloat16*>(param_lp), master, mom,
        reinterpret_cast<const __nv_bfloat16*>(grad), wd_mask, n, a);
  else
    sgd_momentum_kernel<float><<<grid, kThreads, 0, stream>>>(
        reinterpret_cast<__nv_bfloat16*>(param_lp), master, mom,
        reinterpret_cast<const float*>(grad), wd_mask, n, a);
}

void grad_sqnorm(const void* grad, bool grad_is_bf16, int64_t n, float* out, cudaStream_t stream) {
  int grid = grid_for((n + 7) / 8);
  if (grid > kNumSMs * 2) grid = kNumSMs * 2;
  if (grad_is_bf16)
    grad_sqnorm_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(grad), n, out);
  else
    grad_sqnorm_kernel<float><<<grid, kThreads, 0, stream>>>(reinterpret_cast<const float*>(grad), n, out);
}

void clip_scale(const float* parts, int nparts, int stride, float max_norm, float* grad_scale, float* norm_out,
                cudaStream_t stream) {
  clip_scale_kernel<<<1, 32, 0, stream>>>(parts, nparts, stride, max_norm, grad_scale, norm_out);
}

void adam_step(void* param_lp, float* master, float* m, float* v, const void* grad,
               bool grad_is_bf16, int64_t n, const float* lr, const float* grad_scale,
               const int* found_inf, const float* step, float beta1, float beta2, float eps,
               float wd, bool decoupled, cudaStream_t stream) {
  AdamArgs a{lr, grad_scale, found_inf, step, beta1, beta2, eps, wd, decoupled ? 1 : 0};
  int grid = grid_for(n);
  if (grad_is_bf16)
    adam_kernel<__nv_bfloat16><<<grid, kThreads, 0, stream>>>(
        reinterpret_cast<__nv_bfloat16*>(param_lp), master, m, v,
        reinterpret_cast<const __nv_bfloat16*>(grad), n, a);
  else
    adam_kernel<float><<<grid, kThreads, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(param_lp),
                                                      master, m, v,
                                                      reinterpret_cast<const float*>(grad), n, a);
}

}  // namespace edl
