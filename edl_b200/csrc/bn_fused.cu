// SM-resident fused batch-norm kernels: the whole activation is staged ONCE into the aggregate
// shared memory of the chip (148 SMs x ~190 KB = 28 MB), the per-channel reduction is finished with
// a grid-wide barrier, and the second pass (normalise / input-gradient) reads the staged copy -- so
// each tensor crosses HBM exactly once and the two BN passes cost one launch instead of two.
//
//   forward : stats(x) -> barrier -> y = act(x*scale + shift (+res))           (layers whose conv
//             does not already deliver the statistics from its GEMM epilogue)
//   backward: (dbeta, dgamma) -> barrier -> dx (+ dres)
//
// Applies when  NT * numel * 2 B  fits the resident capacity (NT = number of staged operands); larger
// tensors use the two-kernel streaming path (bn_stream.cu).  All chunk loads are issued up front with
// cp.async.bulk + one mbarrier per chunk (maximum memory-level parallelism, no ring reuse).
// Launched cooperatively so that all CTAs are co-resident (required by the grid barrier).
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace edl {
namespace {

constexpr int kConsumers = 512;
constexpr int kThreads = kConsumers;            // every thread consumes; thread 0 also issues the loads
constexpr int kChunkElems = 8192;               // 16 KB per operand chunk
constexpr int kChunkBytes = kChunkElems * 2;
constexpr int kVecPerThread = kChunkElems / 8 / kConsumers;  // 2
constexpr int kMaxResidentBytes = 192 * 1024;   // per CTA, staged operands

EDL_DEVICE void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
      : "memory");
}

EDL_DEVICE unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs of the (cooperatively launched) grid.  `counter` is zero on entry of the kernel.
EDL_DEVICE void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (ld_acquire_gpu(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

struct Layout {
  int chunks_per_cta;   // K
  int64_t n_chunks;
  int64_t total;
};

template <int NT>
struct Stage {
  uint8_t* data;     // [K][NT][kChunkBytes]
  uint64_t* bars;    // [K]
  float* scratch;    // kConsumers * 16 floats (reduction), separate from the staged data
  EDL_DEVICE uint8_t* buf(int k, int t) const { return data + ((size_t)k * NT + t) * kChunkBytes; }
};

template <int NT>
EDL_DEVICE Stage<NT> stage_init(uint8_t* smem_raw, int K) {
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  Stage<NT> s;
  s.data = smem;
  s.bars = reinterpret_cast<uint64_t*>(smem + (size_t)K * NT * kChunkBytes);
  s.scratch = reinterpret_cast<float*>(s.bars + K);
  if (threadIdx.x == 0) {
    for (int k = 0; k < K; ++k) ptx::mbar_init(&s.bars[k], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  return s;
}

template <int NT>
size_t stage_smem_bytes(int K) {
  return (size_t)K * NT * kChunkBytes + (size_t)K * 8 + (size_t)kConsumers * 16 * 4 + 256;
}

// thread 0: issue every load of this CTA's chunk range
template <int NT>
EDL_DEVICE void issue_all(const Stage<NT>& s, const __nv_bfloat16* const (&src)[NT], const Layout& L) {
  if (threadIdx.x != 0) return;
  const int64_t first = (int64_t)blockIdx.x * L.chunks_per_cta;
  for (int k = 0; k < L.chunks_per_cta; ++k) {
    const int64_t ch = first + k;
    if (ch >= L.n_chunks) break;
    const int64_t off = ch * kChunkElems;
    const int64_t rem = L.total - off;
    const uint32_t bytes = (uint32_t)((rem < kChunkElems ? rem : kChunkElems) * 2);
    ptx::mbar_arrive_expect_tx(&s.bars[k], bytes * NT);
#pragma unroll
    for (int t = 0; t < NT; ++t) bulk_load(s.buf(k, t), src[t] + off, bytes, &s.bars[k]);
  }
}

// visit every staged 16-byte vector owned by this thread; `wait` = first pass (chunks may still be landing)
template <int NT, class Body>
EDL_DEVICE void for_each_vec(const Stage<NT>& s, const Layout& L, bool wait, Body body) {
  const int64_t first = (int64_t)blockIdx.x * L.chunks_per_cta;
  for (int k = 0; k < L.chunks_per_cta; ++k) {
    const int64_t ch = first + k;
    if (ch >= L.n_chunks) break;
    if (wait) ptx::mbar_wait(&s.bars[k], 0);
    const int64_t off = ch * kChunkElems;
    const int64_t rem = L.total - off;
    const int valid = (int)((rem < kChunkElems ? rem : kChunkElems) / 8);
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
      const int v = threadIdx.x + j * kConsumers;
      if (v < valid) {
        const bf16x8* ptrs[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) ptrs[t] = reinterpret_cast<const bf16x8*>(s.buf(k, t)) + v;
        body(off + (int64_t)v * 8, ptrs);
      }
    }
  }
}

template <int NV>
EDL_DEVICE void reduce_groups(float (&acc)[NV], int cvecs, float* scratch) {
  const int tid = threadIdx.x;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) scratch[k * kConsumers + tid] = acc[k];
  __syncthreads();
  for (int stride = kConsumers >> 1; stride >= cvecs; stride >>= 1) {
    if (tid < stride) {
#pragma unroll
      for (int k = 0; k < NV; ++k) scratch[k * kConsumers + tid] += scratch[k * kConsumers + tid + stride];
    }
    __syncthreads();
  }
  if (tid < cvecs) {
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = scratch[k * kConsumers + tid];
  }
}

// ---------------------------------------------------------------------------------------------
template <bool HAS_RES>
__global__ void __launch_bounds__(kThreads, 1)
bn_fwd_fused_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                    __nv_bfloat16* __restrict__ y, float* __restrict__ sums, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float* __restrict__ running_mean,
                    float* __restrict__ running_var, float* __restrict__ saved_mean,
                    float* __restrict__ saved_rstd, Layout L, int C, float eps, float momentum, int relu,
                    unsigned int* __restrict__ sync_counter) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int NT = HAS_RES ? 2 : 1;
  Stage<NT> s = stage_init<NT>(smem_raw, L.chunks_per_cta);
  {
    const __nv_bfloat16* src[NT];
    src[0] = x;
    if (HAS_RES) src[NT - 1] = res;
    issue_all<NT>(s, src, L);
  }
  const int tid = threadIdx.x;
  const int cvecs = C / 8;
  const int c0 = (tid % cvecs) * 8;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for_each_vec<NT>(s, L, true, [&](int64_t, const bf16x8* const (&p)[NT]) {
    float f[8];
    unpack8(*p[0], f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += f[i];
      acc[8 + i] = fmaf(f[i], f[i], acc[8 + i]);
    }
  });
  reduce_groups<16>(acc, cvecs, s.scratch);
  if (tid < cvecs) {
    red_add8(&sums[tid * 8], acc);
    red_add8(&sums[C + tid * 8], acc + 8);
  }
  grid_barrier(sync_counter, gridDim.x);
  const int64_t M = L.total / C;
  const float inv_m = 1.f / (float)M;
  float scale[8], shift[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float mean = __ldcg(&sums[c0 + i]) * inv_m;
    const float var = fmaxf(__ldcg(&sums[C + c0 + i]) * inv_m - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    scale[i] = gamma[c0 + i] * rstd;
    shift[i] = beta[c0 + i] - mean * scale[i];
    if (blockIdx.x == 0 && tid < cvecs) {
      saved_mean[c0 + i] = mean;
      saved_rstd[c0 + i] = rstd;
      if (running_mean != nullptr) {
        const float unbiased = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
        running_mean[c0 + i] = (1.f - momentum) * running_mean[c0 + i] + momentum * mean;
        running_var[c0 + i] = (1.f - momentum) * running_var[c0 + i] + momentum * unbiased;
      }
    }
  }
  for_each_vec<NT>(s, L, false, [&](int64_t goff, const bf16x8* const (&p)[NT]) {
    float f[8];
    unpack8(*p[0], f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], scale[i], shift[i]);
    if (HAS_RES) {
      float g[8];
      unpack8(*p[NT - 1], g);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += g[i];
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    st_vec(y + goff, pack8(f));
  });
}

struct Coef {
  float mean[8], rstd[8], scale[8], shift[8];
};

template <bool RELU, bool HAS_Y>
EDL_DEVICE void masked_grad(float (&g)[8], const float (&f)[8], const bf16x8* yv, const Coef& k) {
  if (!RELU) return;
  if (HAS_Y) {
    float o[8];
    unpack8(*yv, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = o[i] > 0.f ? g[i] : 0.f;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = fmaf(f[i], k.scale[i], k.shift[i]) > 0.f ? g[i] : 0.f;
  }
}

template <bool RELU, bool HAS_Y>
__global__ void __launch_bounds__(kThreads, 1)
bn_bwd_fused_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                    const __nv_bfloat16* __restrict__ y, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ saved_mean,
                    const float* __restrict__ saved_rstd, float* __restrict__ dsums,
                    __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dres,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, Layout L, int C, int accumulate,
                    unsigned int* __restrict__ sync_counter) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int NT = HAS_Y ? 3 : 2;
  Stage<NT> s = stage_init<NT>(smem_raw, L.chunks_per_cta);
  {
    const __nv_bfloat16* src[NT];
    src[0] = dy;
    src[1] = x;
    if (HAS_Y) src[NT - 1] = y;
    issue_all<NT>(s, src, L);
  }
  const int tid = threadIdx.x;
  const int cvecs = C / 8;
  const int c0 = (tid % cvecs) * 8;
  Coef k;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    k.mean[i] = saved_mean[c0 + i];
    k.rstd[i] = saved_rstd[c0 + i];
    k.scale[i] = gamma[c0 + i] * k.rstd[i];
    k.shift[i] = beta[c0 + i] - k.mean[i] * k.scale[i];
  }
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for_each_vec<NT>(s, L, true, [&](int64_t, const bf16x8* const (&p)[NT]) {
    float g[8], f[8];
    unpack8(*p[0], g);
    unpack8(*p[1], f);
    masked_grad<RELU, HAS_Y>(g, f, p[NT - 1], k);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += g[i];
      acc[8 + i] = fmaf(g[i], (f[i] - k.mean[i]) * k.rstd[i], acc[8 + i]);
    }
  });
  reduce_groups<16>(acc, cvecs, s.scratch);
  if (tid < cvecs) {
    red_add8(&dsums[tid * 8], acc);
    red_add8(&dsums[C + tid * 8], acc + 8);
  }
  grid_barrier(sync_counter, gridDim.x);
  const float inv_m = 1.f / (float)(L.total / C);
  float k1[8], k2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float db = __ldcg(&dsums[c0 + i]), dg = __ldcg(&dsums[C + c0 + i]);
    k1[i] = db * inv_m;
    k2[i] = dg * inv_m;
    if (blockIdx.x == 0 && tid < cvecs && dgamma != nullptr) {
      if (accumulate) {
        dgamma[c0 + i] += dg;
        dbeta[c0 + i] += db;
      } else {
        dgamma[c0 + i] = dg;
        dbeta[c0 + i] = db;
      }
    }
  }
  for_each_vec<NT>(s, L, false, [&](int64_t goff, const bf16x8* const (&p)[NT]) {
    float g[8], f[8];
    unpack8(*p[0], g);
    unpack8(*p[1], f);
    masked_grad<RELU, HAS_Y>(g, f, p[NT - 1], k);
    if (dres != nullptr) st_vec(dres + goff, pack8(g));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xhat = (f[i] - k.mean[i]) * k.rstd[i];
      f[i] = k.scale[i] * (g[i] - k1[i] - xhat * k2[i]);
    }
    st_vec(dx + goff, pack8(f));
  });
}

bool plan(int64_t total, int nt, Layout* L, int* grid, size_t* smem) {
  const int64_t n_chunks = (total + kChunkElems - 1) / kChunkElems;
  int g = (int)(n_chunks < kNumSMs ? n_chunks : kNumSMs);
  const int k = (int)((n_chunks + g - 1) / g);
  if ((size_t)k * nt * kChunkBytes > (size_t)kMaxResidentBytes) return false;
  g = (int)((n_chunks + k - 1) / k);
  L->chunks_per_cta = k;
  L->n_chunks = n_chunks;
  L->total = total;
  *grid = g;
  *smem = (size_t)k * nt * kChunkBytes + (size_t)k * 8 + (size_t)kConsumers * 16 * 4 + 256;
  return true;
}

template <class K, class... Args>
const char* launch_coop(K kern, int grid, size_t smem, cudaStream_t stream, Args... args) {
  static thread_local const void* done[16];
  static thread_local int n = 0;
  bool found = false;
  for (int i = 0; i < n; ++i) found |= done[i] == (const void*)kern;
  if (!found) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    apply_carveout((const void*)kern);
    if (n < 16) done[n++] = (const void*)kern;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

#define BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BFW(p) reinterpret_cast<__nv_bfloat16*>(p)

bool bn_fused_fits(int64_t M, int C, int num_operands) {
  if (C < 8 || C > 4096 || (C & (C - 1)) != 0) return false;
  Layout L;
  int grid;
  size_t smem;
  return plan(M * C, num_operands, &L, &grid, &smem);
}

const char* bn_fwd_fused(const void* x, const void* res, void* y, float* sums, const float* gamma,
                         const float* beta, float* running_mean, float* running_var, float* saved_mean,
                         float* saved_rstd, int64_t M, int C, float eps, float momentum, bool relu,
                         unsigned int* sync_counter, cudaStream_t s) {
  Layout L;
  int grid;
  size_t smem;
  if (!plan(M * C, res != nullptr ? 2 : 1, &L, &grid, &smem)) return "tensor does not fit the resident capacity";
  if (res != nullptr)
    return launch_coop(bn_fwd_fused_kernel<true>, grid, smem, s, BF(x), BF(res), BFW(y), sums, gamma, beta,
                       running_mean, running_var, saved_mean, saved_rstd, L, C, eps, momentum, relu ? 1 : 0,
                       sync_counter);
  return launch_coop(bn_fwd_fused_kernel<false>, grid, smem, s, BF(x), BF(res), BFW(y), sums, gamma, beta,
                     running_mean, running_var, saved_mean, saved_rstd, L, C, eps, momentum, relu ? 1 : 0,
                     sync_counter);
}

const char* bn_bwd_fused(const void* dy, const void* x, const void* y, const float* gamma, const float* beta,
                         const float* saved_mean, const float* saved_rstd, float* dsums, void* dx, void* dres,
                         float* dgamma, float* dbeta, int64_t M, int C, bool relu, bool accumulate,
                         unsigned int* sync_counter, cudaStream_t s) {
  Layout L;
  int grid;
  size_t smem;
  const bool has_y = relu && y != nullptr;
  if (!plan(M * C, has_y ? 3 : 2, &L, &grid, &smem)) return "tensor does not fit the resident capacity";
#define LAUNCH(R, Y)                                                                                       \
  return launch_coop(bn_bwd_fused_kernel<R, Y>, grid, smem, s, BF(dy), BF(x), BF(y), gamma, beta, saved_mean, \
                     saved_rstd, dsums, BFW(dx), BFW(dres), dgamma, dbeta, L, C, accumulate ? 1 : 0,        \
                     sync_counter)
  if (!relu) LAUNCH(false, false);
  if (has_y) LAUNCH(true, true);
  LAUNCH(true, false);
#undef LAUNCH
}

}  // namespace edl
