// Fused gradient all-reduce over NVSwitch peer memory.
//
// The reference delegates gradient sync to NCCL ring all-reduce inside Paddle fleet
// (example/distill/resnet/train_with_fleet.py:332-333, <=16 MB fused buckets,
// scripts/train_gpu.sh:67-70) and runs AMP unscale / check_finite as separate ops
// (utils/fp16_utils.py:86-129).  Here one kernel per bucket does
//     barrier -> P2P (or in-switch NVLS) reduction in fp32 -> scale -> finite check ->
//     squared-norm partial -> write-back to every peer -> barrier
// Three algorithms, chosen by the Python-side planner per bucket and per world size:
//   one-shot  : every rank reads the whole bucket from every peer (latency-optimal, small buckets)
//   two-shot  : rank r owns slice r: reduce-scatter by peer loads, all-gather by peer stores
//   multimem  : two-shot where both halves run inside the switch (multimem.ld_reduce / .st)
// The kernels take `nblocks` so the planner can restrict them to a few SMs and overlap them with
// backward kernels on another stream.
#include "comm.cuh"
#include "kernels.h"

namespace edl {
namespace {

constexpr int kCommThreads = 512;

struct Epilogue {
  float scale;               // applied to the reduced sum (e.g. 1/world)
  int* found_inf;            // device flag, set to 1 if any reduced value is non-finite (nullable)
  float* sqnorm;             // device accumulator of sum(g^2) over *this rank's* elements (nullable)
};

template <typename T>
struct Vec;  // 16-byte vector of T, accumulate in fp32

template <>
struct Vec<__nv_bfloat16> {
  static constexpr int kElems = 8;
  EDL_DEVICE static void add(float (&acc)[8], const int4& raw) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(&raw), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += f[i];
  }
  EDL_DEVICE static int4 pack(const float (&acc)[8]) {
    bf16x8 p = pack8(acc);
    return *reinterpret_cast<int4*>(&p);
  }
};
template <>
struct Vec<float> {
  static constexpr int kElems = 4;
  EDL_DEVICE static void add(float (&acc)[8], const int4& raw) {
    const float4& f = *reinterpret_cast<const float4*>(&raw);
    acc[0] += f.x; acc[1] += f.y; acc[2] += f.z; acc[3] += f.w;
  }
  EDL_DEVICE static int4 pack(const float (&acc)[8]) {
    float4 f = make_float4(acc[0], acc[1], acc[2], acc[3]);
    return *reinterpret_cast<int4*>(&f);
  }
};

template <typename T>
EDL_DEVICE void apply_epilogue(float (&acc)[8], const Epilogue& e, bool& bad, float& sq) {
#pragma unroll
  for (int i = 0; i < Vec<T>::kElems; ++i) {
    acc[i] *= e.scale;
    bad |= !isfinite(acc[i]);
    sq = fmaf(acc[i], acc[i], sq);
  }
}

EDL_DEVICE void finish_epilogue(const Epilogue& e, bool bad, float sq) {
  if (e.found_inf != nullptr && __syncthreads_or(bad ? 1 : 0)) {
    if (threadIdx.x == 0) atomicExch(e.found_inf, 1);
  }
  if (e.sqnorm != nullptr) {
    __shared__ float sh[kCommThreads / 32];
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = threadIdx.x < kCommThreads / 32 ? sh[threadIdx.x] : 0.f;
      v = warp_sum(v);
      if (threadIdx.x == 0) atomicAdd(e.sqnorm, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// one-shot: out (local, may be non-symmetric) = scale * sum_p data[p][0:n)
template <typename T>
__global__ void __launch_bounds__(kCommThreads)
allreduce_oneshot_kernel(CommCtx c, T* __restrict__ out, int64_t n, Epilogue e) {
  const uint32_t epoch = comm_epoch_begin(c);
  comm_barrier<false>(c, kSigStart, epoch);
  constexpr int VE = Vec<T>::kElems;
  const int64_t nvec = n / VE;
  bool bad = false;
  float sq = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    // Unstaggered peer order + fixed summation order (rank 0..world-1): every rank computes a
    // bit-identical result, and `raw` stays in registers (static indexing only).
    int4 raw[kMaxWorld];
#pragma unroll
    for (int q = 0; q < kMaxWorld; ++q)
      if (q < c.world) raw[q] = ld_peer(reinterpret_cast<const int4*>(c.data[q]) + i);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < kMaxWorld; ++q)
      if (q < c.world) Vec<T>::add(acc, raw[q]);
    apply_epilogue<T>(acc, e, bad, sq);
    reinterpret_cast<int4*>(out)[i] = Vec<T>::pack(acc);
  }
  // scalar tail handled by block 0 / thread 0 (tiny)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = nvec * VE; i < n; ++i) {
      float s = 0.f;
      for (int q = 0; q < c.world; ++q) {
        const T* src = reinterpret_cast<const T*>(c.data[q]);
        if constexpr (sizeof(T) == 2) s += __bfloat162float(src[i]);
        else s += src[i];
      }
      s *= e.scale;
      bad |= !isfinite(s);
      sq = fmaf(s, s, sq);
      if constexpr (sizeof(T) == 2) out[i] = __float2bfloat16(s);
      else out[i] = s;
    }
  }
  finish_epilogue(e, bad, sq);
  comm_barrier<true>(c, kSigEnd, epoch);
  comm_epoch_end(c, epoch);
}

// ---------------------------------------------------------------------------------------------
// two-shot, in place on the symmetric buffer.  n must be a multiple of VE (16 bytes); any world
// size works (elastic 8 -> 6 -> 8): slice r = vectors [r*S, min((r+1)*S, nvec)), S = ceil(nvec/W).
template <typename T, bool kMultimem>
__global__ void __launch_bounds__(kCommThreads)
allreduce_twoshot_kernel(CommCtx c, int64_t n, Epilogue e) {
  const uint32_t epoch = comm_epoch_begin(c);
  comm_barrier<false>(c, kSigStart, epoch);
  constexpr int VE = Vec<T>::kElems;
  const int64_t nvec = n / VE;
  const int64_t slice_cap = (nvec + c.world - 1) / c.world;
  const int64_t base = slice_cap * c.rank;
  int64_t slice_vecs = nvec - base;
  if (slice_vecs > slice_cap) slice_vecs = slice_cap;
  if (slice_vecs < 0) slice_vecs = 0;
  bool bad = false;
  float sq = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slice_vecs;
       i += (int64_t)gridDim.x * blockDim.x) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if constexpr (kMultimem) {
      const char* mc = reinterpret_cast<const char*>(c.mc_data) + (base + i) * 16;
      if constexpr (sizeof(T) == 2) {
        bf16x8 v = multimem_ld_reduce_bf16(mc);
        unpack8(v, acc);
      } else {
        float4 v = multimem_ld_reduce_f32(mc);
        acc[0] = v.x; acc[1] = v.y; acc[2] = v.z; acc[3] = v.w;
      }
    } else {
      int4 raw[kMaxWorld];
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p) {
        if (p < c.world) {
          int peer = c.rank + p;
          if (peer >= c.world) peer -= c.world;
          raw[p] = ld_peer(reinterpret_cast<const int4*>(c.data[peer]) + base + i);
        }
      }
      // only this rank reduces this slice, so the (staggered) order needs no cross-rank agreement
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p)
        if (p < c.world) Vec<T>::add(acc, raw[p]);
    }
    apply_epilogue<T>(acc, e, bad, sq);
    const int4 packed = Vec<T>::pack(acc);
    if constexpr (kMultimem) {
      char* mc = reinterpret_cast<char*>(c.mc_data) + (base + i) * 16;
      if constexpr (sizeof(T) == 2) multimem_st_bf16(mc, *reinterpret_cast<const bf16x8*>(&packed));
      else multimem_st_f32(mc, *reinterpret_cast<const float4*>(&packed));
    } else {
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p) {
        if (p < c.world) {
          int peer = c.rank + p;
          if (peer >= c.world) peer -= c.world;
          st_peer(reinterpret_cast<int4*>(c.data[peer]) + base + i, packed);
        }
      }
    }
  }
  finish_epilogue(e, bad, sq);
  comm_barrier<true>(c, kSigEnd, epoch);
  comm_epoch_end(c, epoch);
}

// ---------------------------------------------------------------------------------------------
// Fused gradient reduce-scatter -> SGD-momentum -> parameter all-gather, ONE kernel per bucket (SURVEY K7 + K10).
//
// The reference runs NCCL all-reduce, then one Paddle `momentum` op per tensor on every rank
// (example/distill/resnet/train_with_fleet.py:106-122,332-333): every rank applies the same update to the same
// 25.6 M parameters.  Here rank r owns slice r of the bucket: it pulls the summed gradient of its slice out of the
// switch (multimem.ld_reduce, or W peer loads), updates ITS slice of the fp32 master weights and momentum -- the
// optimizer work and state traffic drop by the world size -- and pushes the new bf16 parameters of that slice into
// every rank's parameter buffer (multimem.st / peer stores).  The NVLink traffic equals a two-shot all-reduce
// (gradients in, parameters out); the separate optimizer pass over the bucket and the end-of-backward join
// before it disappear.  Master / momentum outside the owned slice go stale on purpose; the host consolidates them
// (ElasticDataParallel.consolidate_optimizer_state) before a checkpoint or a stage change.
// A bucket is skipped as a whole (no update, no parameter store) when this rank's error word is set: a peer died
// and the sum would be partial -- the elastic hot-recovery path takes over from the last good parameters.
struct FusedSgd {
  void* pdata[kMaxWorld];    // every rank's bf16 parameter window (bucket offset applied)
  void* mc_param;            // multicast alias of that window or nullptr
  float* master;             // local fp32 master weights of the bucket
  float* mom;                // local fp32 momentum of the bucket
  const float* wd_mask;      // optional per-element weight-decay mask of the bucket
  const float* lr;           // device scalar
  const int* found_inf;      // optional device flag: non-zero => skip the update
  float momentum, wd;
  int nesterov;
};

template <bool kMultimem>
__global__ void __launch_bounds__(kCommThreads)
allreduce_sgd_kernel(CommCtx c, int64_t n, Epilogue e, FusedSgd f) {
  using T = __nv_bfloat16;
  const uint32_t epoch = comm_epoch_begin(c);
  comm_barrier<false>(c, kSigStart, epoch);
  const bool broken = *reinterpret_cast<volatile uint32_t*>(c.sig[c.rank] + kSigError) != 0u ||
                      (f.found_inf != nullptr && *f.found_inf != 0);
  constexpr int VE = 8;
  const int64_t nvec = n / VE;
  const int64_t slice_cap = (nvec + c.world - 1) / c.world;
  const int64_t base = slice_cap * c.rank;
  int64_t slice_vecs = nvec - base;
  if (slice_vecs > slice_cap) slice_vecs = slice_cap;
  if (slice_vecs < 0 || broken) slice_vecs = 0;
  const float lr = *f.lr;
  bool bad = false;
  float sq = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slice_vecs;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = base + i;
    // optimizer state of this vector: issued before the (long-latency) cross-GPU gradient fetch
    const float4 w0 = *reinterpret_cast<const float4*>(f.master + v * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(f.master + v * 8 + 4);
    const float4 m0 = *reinterpret_cast<const float4*>(f.mom + v * 8);
    const float4 m1 = *reinterpret_cast<const float4*>(f.mom + v * 8 + 4);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if constexpr (kMultimem) {
      bf16x8 g = multimem_ld_reduce_bf16(reinterpret_cast<const char*>(c.mc_data) + v * 16);
      unpack8(g, acc);
    } else {
      int4 raw[kMaxWorld];
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p) {
        if (p < c.world) {
          int peer = c.rank + p;
          if (peer >= c.world) peer -= c.world;
          raw[p] = ld_peer(reinterpret_cast<const int4*>(c.data[peer]) + v);
        }
      }
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p)
        if (p < c.world) Vec<T>::add(acc, raw[p]);
    }
    apply_epilogue<T>(acc, e, bad, sq);
    float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    float mv[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    float wdm[8];
    if (f.wd_mask != nullptr) {
      const float4 k0 = *reinterpret_cast<const float4*>(f.wd_mask + v * 8);
      const float4 k1 = *reinterpret_cast<const float4*>(f.wd_mask + v * 8 + 4);
      wdm[0] = k0.x; wdm[1] = k0.y; wdm[2] = k0.z; wdm[3] = k0.w;
      wdm[4] = k1.x; wdm[5] = k1.y; wdm[6] = k1.z; wdm[7] = k1.w;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float wd = f.wd_mask != nullptr ? f.wd * wdm[k] : f.wd;
      const float gg = fmaf(wd, w[k], acc[k]);
      mv[k] = fmaf(f.momentum, mv[k], gg);
      const float upd = f.nesterov ? fmaf(f.momentum, mv[k], gg) : mv[k];
      w[k] = fmaf(-lr, upd, w[k]);
    }
    *reinterpret_cast<float4*>(f.master + v * 8) = make_float4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<float4*>(f.master + v * 8 + 4) = make_float4(w[4], w[5], w[6], w[7]);
    *reinterpret_cast<float4*>(f.mom + v * 8) = make_float4(mv[0], mv[1], mv[2], mv[3]);
    *reinterpret_cast<float4*>(f.mom + v * 8 + 4) = make_float4(mv[4], mv[5], mv[6], mv[7]);
    const bf16x8 packed = pack8(w);
    if constexpr (kMultimem) {
      multimem_st_bf16(reinterpret_cast<char*>(f.mc_param) + v * 16, packed);
    } else {
      const int4& raw = *reinterpret_cast<const int4*>(&packed);
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p) {
        if (p < c.world) {
          int peer = c.rank + p;
          if (peer >= c.world) peer -= c.world;
          st_peer(reinterpret_cast<int4*>(f.pdata[peer]) + v, raw);
        }
      }
    }
  }
  finish_epilogue(e, bad, sq);
  comm_barrier<true>(c, kSigEnd, epoch);
  comm_epoch_end(c, epoch);
}

// ---------------------------------------------------------------------------------------------
// broadcast from `root` into every rank's symmetric buffer (state hand-off to joiners).
__global__ void __launch_bounds__(kCommThreads)
broadcast_kernel(CommCtx c, int root, int64_t nbytes) {
  const uint32_t epoch = comm_epoch_begin(c);
  comm_barrier<false>(c, kSigStart, epoch);
  if (c.rank != root) {
    const int64_t nvec = nbytes / 16;
    const int4* src = reinterpret_cast<const int4*>(c.data[root]);
    int4* dst = reinterpret_cast<int4*>(c.data[c.rank]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
         i += (int64_t)gridDim.x * blockDim.x)
      dst[i] = ld_peer(src + i);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      const char* s = reinterpret_cast<const char*>(c.data[root]);
      char* d = reinterpret_cast<char*>(c.data[c.rank]);
      for (int64_t i = nvec * 16; i < nbytes; ++i) d[i] = s[i];
    }
  }
  comm_barrier<true>(c, kSigEnd, epoch);
  comm_epoch_end(c, epoch);
}

// all-gather of 8 floats per rank through the signal-pad scratch area (norm partials, counters).
__global__ void __launch_bounds__(kCommThreads)
allgather_scalars_kernel(CommCtx c, const float* __restrict__ in, float* __restrict__ out,
                         int count /* <= 8 */) {
  const uint32_t epoch = comm_epoch_begin(c);
  comm_barrier<false>(c, kSigStart, epoch);
  if ((int)threadIdx.x < c.world * count) {
    int peer = threadIdx.x / count, k = threadIdx.x % count;
    float* dst = reinterpret_cast<float*>(c.sig[peer] + kSigScratch) + c.rank * 8 + k;
    *dst = in[k];
  }
  comm_barrier<true>(c, kSigEnd, epoch);
  if ((int)threadIdx.x < c.world * count) {
    const float* src = reinterpret_cast<const float*>(c.sig[c.rank] + kSigScratch);
    int r = threadIdx.x / count, k = threadIdx.x % count;
    out[r * count + k] = src[r * 8 + k];
  }
  comm_epoch_end(c, epoch);
}

inline int clamp_blocks(int nblocks) {
  if (nblocks < 1) nblocks = 1;
  if (nblocks > kMaxCommBlocks) nblocks = kMaxCommBlocks;
  return nblocks;
}

}  // namespace

int comm_sig_words() { return kSigWords; }
int comm_max_world() { return kMaxWorld; }
int comm_error_word_offset() { return kSigError; }

static CommCtx make_ctx(const CommHandles& h) {
  CommCtx c;
  for (int i = 0; i < kMaxWorld; ++i) {
    c.data[i] = i < h.world ? h.data[i] : nullptr;
    c.sig[i] = i < h.world ? reinterpret_cast<uint32_t*>(h.sig[i]) : nullptr;
  }
  c.mc_data = h.mc_data;
  c.rank = h.rank;
  c.world = h.world;
  c.timeout_ns = h.timeout_ns;
  return c;
}

void allreduce_oneshot(const CommHandles& h, void* out, bool is_bf16, int64_t n, float scale,
                       int* found_inf, float* sqnorm, int nblocks, cudaStream_t stream) {
  CommCtx c = make_ctx(h);
  Epilogue e{scale, found_inf, sqnorm};
  nblocks = clamp_blocks(nblocks);
  if (is_bf16)
    allreduce_oneshot_kernel<__nv_bfloat16><<<nblocks, kCommThreads, 0, stream>>>(
        c, reinterpret_cast<__nv_bfloat16*>(out), n, e);
  else
    allreduce_oneshot_kernel<float><<<nblocks, kCommThreads, 0, stream>>>(
        c, reinterpret_cast<float*>(out), n, e);
}

void allreduce_twoshot(const CommHandles& h, bool is_bf16, int64_t n, float scale, int* found_inf,
                       float* sqnorm, bool multimem, int nblocks, cudaStream_t stream) {
  CommCtx c = make_ctx(h);
  Epilogue e{scale, found_inf, sqnorm};
  nblocks = clamp_blocks(nblocks);
  if (multimem && h.mc_data != nullptr) {
    if (is_bf16)
      allreduce_twoshot_kernel<__nv_bfloat16, true><<<nblocks, kCommThreads, 0, stream>>>(c, n, e);
    else
      allreduce_twoshot_kernel<float, true><<<nblocks, kCommThreads, 0, stream>>>(c, n, e);
  } else {
    if (is_bf16)
      allreduce_twoshot_kernel<__nv_bfloat16, false><<<nblocks, kCommThreads, 0, stream>>>(c, n, e);
    else
      allreduce_twoshot_kernel<float, false><<<nblocks, kCommThreads, 0, stream>>>(c, n, e);
  }
}

void allreduce_sgd(const CommHandles& h, void* const* param_ptrs, void* mc_param, float* master, float* mom,
                   const float* wd_mask, int64_t n, float scale, int* found_inf_out, float* sqnorm,
                   const float* lr, const int* skip_flag, float momentum, float wd, bool nesterov, bool multimem,
                   int nblocks, cudaStream_t stream) {
  CommCtx c = make_ctx(h);
  Epilogue e{scale, found_inf_out, sqnorm};
  FusedSgd f{};
  for (int i = 0; i < kMaxWorld; ++i) f.pdata[i] = i < h.world ? param_ptrs[i] : nullptr;
  f.mc_param = mc_param;
  f.master = master;
  f.mom = mom;
  f.wd_mask = wd_mask;
  f.lr = lr;
  f.found_inf = skip_flag;
  f.momentum = momentum;
  f.wd = wd;
  f.nesterov = nesterov ? 1 : 0;
  nblocks = clamp_blocks(nblocks);
  if (multimem && h.mc_data != nullptr && mc_param != nullptr)
    allreduce_sgd_kernel<true><<<nblocks, kCommThreads, 0, stream>>>(c, n, e, f);
  else
    allreduce_sgd_kernel<false><<<nblocks, kCommThreads, 0, stream>>>(c, n, e, f);
}

void comm_broadcast(const CommHandles& h, int root, int64_t nbytes, int nblocks,
                    cudaStream_t stream) {
  CommCtx c = make_ctx(h);
  broadcast_kernel<<<clamp_blocks(nblocks), kCommThreads, 0, stream>>>(c, root, nbytes);
}

void comm_allgather_scalars(const CommHandles& h, const float* in, float* out, int count,
                            cudaStream_t stream) {
  CommCtx c = make_ctx(h);
  allgather_scalars_kernel<<<1, kCommThreads, 0, stream>>>(c, in, out, count);
}

}  // namespace edl
