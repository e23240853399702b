// Fused distillation / classification loss kernels for sm_100a.
//
// Reference call sites (Paddle library ops): softmax_with_cross_entropy(soft_label=True) on the
// teacher's scores, the hard-label variant, accuracy(k=1,5) and mean
// (example/distill/resnet/train_with_fleet.py:254-275); KL / KL_T temperature losses
// (example/distill/nlp/model.py:54-66).  One kernel computes the row softmax statistics, the loss
// reduction (mean over the batch) and keeps (lse, sum_p) so the backward is one more tiny kernel.
// Targets can be probabilities, raw teacher logits (softmax with temperature applied on the fly --
// this is the form the NVSwitch logit-ship path delivers) or hard int64 labels.
#include "common.cuh"
#include "kernels.h"

namespace edl {
namespace {

constexpr int kRowThreads = 128;

EDL_DEVICE float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < kRowThreads / 32; ++i) r += sh[i];
  return r;
}
EDL_DEVICE float block_max(float v, float* sh) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = -INFINITY;
#pragma unroll
  for (int i = 0; i < kRowThreads / 32; ++i) r = fmaxf(r, sh[i]);
  return r;
}

template <typename T>
EDL_DEVICE float ldf(const T* p, int64_t i) {
  if constexpr (sizeof(T) == 2) return __bfloat162float(p[i]);
  else return (float)p[i];
}

// mode: 0 = target probabilities, 1 = target logits (softmax(t / t_temp)), 2 = hard labels
template <typename LT, typename TT>
__global__ void __launch_bounds__(kRowThreads)
soft_ce_fwd_kernel(const LT* __restrict__ logits, const TT* __restrict__ target,
                   const int64_t* __restrict__ labels, float* __restrict__ loss_out,
                   float* __restrict__ row_stats /* [N,4]: zmax, lse, psum, tmax|tlse */, int N,
                   int C, int mode, float s_temp, float t_temp, float label_smooth, int kl,
                   float loss_scale) {
  __shared__ float sh[kRowThreads / 32];
  const int row = blockIdx.x;
  const LT* z = logits + (int64_t)row * C;
  const float inv_ts = 1.f / s_temp;
  float zmax = -INFINITY;
  for (int j = threadIdx.x; j < C; j += kRowThreads) zmax = fmaxf(zmax, ldf(z, j) * inv_ts);
  zmax = block_max(zmax, sh);
  float zsum = 0.f;
  for (int j = threadIdx.x; j < C; j += kRowThreads) zsum += __expf(ldf(z, j) * inv_ts - zmax);
  zsum = block_sum(zsum, sh);
  const float lse = zmax + __logf(zsum);

  float psum = 0.f, pz = 0.f, plogp = 0.f, tmax = 0.f, tlse = 0.f;
  if (mode == 2) {
    const int64_t lab = labels[row];
    const float zl = ldf(z, lab) * inv_ts;
    if (label_smooth > 0.f) {
      float zs = 0.f;
      for (int j = threadIdx.x; j < C; j += kRowThreads) zs += ldf(z, j) * inv_ts;
      zs = block_sum(zs, sh);
      pz = (1.f - label_smooth) * zl + label_smooth * zs / (float)C;
    } else {
      pz = zl;
    }
    psum = 1.f;
  } else if (mode == 0) {
    const TT* t = target + (int64_t)row * C;
    for (int j = threadIdx.x; j < C; j += kRowThreads) {
      float p = ldf(t, j);
      psum += p;
      pz = fmaf(p, ldf(z, j) * inv_ts, pz);
      if (kl && p > 0.f) plogp = fmaf(p, __logf(p), plogp);
    }
    psum = block_sum(psum, sh);
    pz = block_sum(pz, sh);
    if (kl) plogp = block_sum(plogp, sh);
  } else {
    const TT* t = target + (int64_t)row * C;
    const float inv_tt = 1.f / t_temp;
    tmax = -INFINITY;
    for (int j = threadIdx.x; j < C; j += kRowThreads) tmax = fmaxf(tmax, ldf(t, j) * inv_tt);
    tmax = block_max(tmax, sh);
    float tsum = 0.f;
    for (int j = threadIdx.x; j < C; j += kRowThreads) tsum += __expf(ldf(t, j) * inv_tt - tmax);
    tsum = block_sum(tsum, sh);
    tlse = tmax + __logf(tsum);
    for (int j = threadIdx.x; j < C; j += kRowThreads) {
      float lt = ldf(t, j) * inv_tt - tlse;
      float p = __expf(lt);
      pz = fmaf(p, ldf(z, j) * inv_ts, pz);
      if (kl) plogp = fmaf(p, lt, plogp);
    }
    pz = block_sum(pz, sh);
    if (kl) plogp = block_sum(plogp, sh);
    psum = 1.f;
  }
  if (threadIdx.x == 0) {
    float loss = lse * psum - pz + (kl ? plogp : 0.f);
    atomicAdd(loss_out, loss * loss_scale / (float)N);
    row_stats[row * 4 + 0] = zmax;
    row_stats[row * 4 + 1] = lse;
    row_stats[row * 4 + 2] = psum;
    row_stats[row * 4 + 3] = tlse;
  }
}

template <typename LT, typename TT>
__global__ void __launch_bounds__(kRowThreads)
soft_ce_bwd_kernel(const LT* __restrict__ logits, const TT* __restrict__ target,
                   const int64_t* __restrict__ labels, const float* __restrict__ row_stats,
                   const float* __restrict__ grad_out, LT* __restrict__ dlogits, int N, int C,
                   int mode, float s_temp, float t_temp, float label_smooth, float loss_scale) {
  const int row = blockIdx.x;
  const LT* z = logits + (int64_t)row * C;
  LT* dz = dlogits + (int64_t)row * C;
  const float inv_ts = 1.f / s_temp;
  const float lse = row_stats[row * 4 + 1];
  const float psum = row_stats[row * 4 + 2];
  const float tlse = row_stats[row * 4 + 3];
  const float go = (grad_out ? *grad_out : 1.f) * loss_scale * inv_ts / (float)N;
  const int64_t lab = mode == 2 ? labels[row] : -1;
  const TT* t = mode == 2 ? nullptr : target + (int64_t)row * C;
  const float inv_tt = 1.f / t_temp;
  for (int j = threadIdx.x; j < C; j += kRowThreads) {
    float sm = __expf(ldf(z, j) * inv_ts - lse);
    float p;
    if (mode == 2) p = (j == lab ? 1.f - label_smooth : 0.f) + label_smooth / (float)C;
    else if (mode == 0) p = ldf(t, j);
    else p = __expf(ldf(t, j) * inv_tt - tlse);
    float g = (sm * psum - p) * go;
    if constexpr (sizeof(LT) == 2) dz[j] = __float2bfloat16(g);
    else dz[j] = g;
  }
}

template <typename LT>
__global__ void __launch_bounds__(kRowThreads)
topk_acc_kernel(const LT* __restrict__ logits, const int64_t* __restrict__ labels,
                float* __restrict__ counts /* [2]: top1, top5 */, int N, int C) {
  __shared__ float sh[kRowThreads / 32];
  const int row = blockIdx.x;
  const LT* z = logits + (int64_t)row * C;
  const int64_t lab = labels[row];
  const float zl = ldf(z, lab);
  float greater = 0.f;
  for (int j = threadIdx.x; j < C; j += kRowThreads) {
    float v = ldf(z, j);
    greater += (v > zl || (v == zl && j < lab)) ? 1.f : 0.f;
  }
  greater = block_sum(greater, sh);
  if (threadIdx.x == 0) {
    if (greater < 1.f) atomicAdd(&counts[0], 1.f);
    if (greater < 5.f) atomicAdd(&counts[1], 1.f);
  }
}

}  // namespace

void soft_ce_fwd(const void* logits, bool logits_bf16, const void* target, bool target_bf16,
                 const int64_t* labels, float* loss_out, float* row_stats, int N, int C, int mode,
                 float s_temp, float t_temp, float label_smooth, bool kl, float loss_scale,
                 cudaStream_t stream) {
  using bf = __nv_bfloat16;
#define LAUNCH(LT, TT)                                                                          \
  soft_ce_fwd_kernel<LT, TT><<<N, kRowThreads, 0, stream>>>(                                    \
      reinterpret_cast<const LT*>(logits), reinterpret_cast<const TT*>(target), labels, loss_out, \
      row_stats, N, C, mode, s_temp, t_temp, label_smooth, kl ? 1 : 0, loss_scale)
  if (logits_bf16 && target_bf16) LAUNCH(bf, bf);
  else if (logits_bf16) LAUNCH(bf, float);
  else if (target_bf16) LAUNCH(float, bf);
  else LAUNCH(float, float);
#undef LAUNCH
}

void soft_ce_bwd(const void* logits, bool logits_bf16, const void* target, bool target_bf16,
                 const int64_t* labels, const float* row_stats, const float* grad_out,
                 void* dlogits, int N, int C, int mode, float s_temp, float t_temp,
                 float label_smooth, float loss_scale, cudaStream_t stream) {
  using bf = __nv_bfloat16;
#define LAUNCH(LT, TT)                                                                            \
  soft_ce_bwd_kernel<LT, TT><<<N, kRowThreads, 0, stream>>>(                                      \
      reinterpret_cast<const LT*>(logits), reinterpret_cast<const TT*>(target), labels, row_stats, \
      grad_out, reinterpret_cast<LT*>(dlogits), N, C, mode, s_temp, t_temp, label_smooth,         \
      loss_scale)
  if (logits_bf16 && target_bf16) LAUNCH(bf, bf);
  else if (logits_bf16) LAUNCH(bf, float);
  else if (target_bf16) LAUNCH(float, bf);
  else LAUNCH(float, float);
#undef LAUNCH
}

void topk_acc(const void* logits, bool logits_bf16, const int64_t* labels, float* counts, int N,
              int C, cudaStream_t stream) {
  if (logits_bf16)
    topk_acc_kernel<__nv_bfloat16><<<N, kRowThreads, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(logits), labels, counts, N, C);
  else
    topk_acc_kernel<float><<<N, kRowThreads, 0, stream>>>(reinterpret_cast<const float*>(logits),
                                                          labels, counts, N, C);
}

}  // namespace edl
