// FP8 (e4m3 x e4m3 -> fp32 in TMEM -> bf16) GEMM for the teacher's 1x1 convolutions / classifier on
// sm_100a: tcgen05.mma kind::f8f6f4 (UMMA_K = 32), twice the tensor throughput of bf16 and half the
// operand bytes ("fp8 activations, tcgen05/TMEM tiles fed by TMA").  Same warp-specialised
// structure as gemm.cu: warp 0 TMA producer, warp 1 TMEM allocator + single-thread MMA issuer, warps 2..5
// epilogue (dequantisation scale * folded-BN scale, shift, ReLU, bf16, swizzled staging, TMA store).
// A K block is 128 fp8 = one 128-byte swizzle span, so the shared-memory geometry and the operand
// descriptors are those of the bf16 kernel.  Also hosts the bf16 -> e4m3 quantiser.
#include <cuda.h>
#include <cuda_fp8.h>
#include <cstdio>

#include "gemm.h"
#include "kernels.h"
#include "ptx.cuh"

namespace edl {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockKBytes = 128;   // = 128 e4m3 elements
constexpr int kUmmaKBytes = 32;
constexpr int kThreads = 192;

struct Fp8Params {
  int M, N, K;
  const float* col_scale;
  const float* col_shift;
  int relu;
};

template <int BLOCK_N, int STAGES>
struct Smem {
  static constexpr int kABytes = kBlockM * kBlockKBytes;
  static constexpr int kBBytes = BLOCK_N * kBlockKBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kDBytes = kBlockM * BLOCK_N * 2;
  static constexpr int kTileBytes = STAGES * kStageBytes > kDBytes ? STAGES * kStageBytes : kDBytes;
  static constexpr int kBarOffset = kTileBytes;
  static constexpr int kTotal = kTileBytes + 256 + 1024;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kThreads, 2)
gemm_fp8_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmD, const Fp8Params p) {
  using L = Smem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int m0 = (blockIdx.x / tiles_n) * kBlockM;
  const int n0 = (blockIdx.x % tiles_n) * BLOCK_N;
  const int num_kb = (p.K + kBlockKBytes - 1) / kBlockKBytes;
  constexpr uint32_t kTmemCols = BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmD);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // warp-uniform loops, one elected lane issues (see gemm_persist.cu)
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      ptx::mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* sa = smem + s * L::kStageBytes;
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&full_bar[s], L::kStageBytes);
        ptx::tma_load_2d(sa, &tmA, &full_bar[s], i * kBlockKBytes, m0);
        ptx::tma_load_2d(sa + L::kABytes, &tmB, &full_bar[s], i * kBlockKBytes, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = ptx::make_idesc(0, 0, kBlockM, BLOCK_N, 0, 0);   // e4m3 x e4m3, K-major
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      ptx::mbar_wait(&full_bar[s], ph);
      ptx::tc_fence_after();
      const uint32_t sa = ptx::smem_u32(smem + s * L::kStageBytes);
      const uint64_t da0 = ptx::make_smem_desc(sa, 16, 1024);
      const uint64_t db0 = ptx::make_smem_desc(sa + L::kABytes, 16, 1024);
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < kBlockKBytes / kUmmaKBytes; ++k)     // start-address field counts 16-byte units
          ptx::umma_f8(tmem_base, da0 + (uint64_t)(k * (kUmmaKBytes >> 4)), db0 + (uint64_t)(k * (kUmmaKBytes >> 4)), idesc,
                       (i | k) != 0 ? 1u : 0u);
        ptx::umma_commit(&empty_bar[s]);
        if (i == num_kb - 1) ptx::umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* sd = smem;
#pragma unroll 1
    for (int c32 = 0; c32 < BLOCK_N / 32; ++c32) {
      uint32_t rg[32];
      ptx::tmem_ld_32x32(taddr + c32 * 32, rg);
      ptx::tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = n0 + c32 * 32 + j;
        const bool in = col < p.N;
        const float sc = (p.col_scale != nullptr && in) ? p.col_scale[col] : 1.f;
        const float sh = (p.col_shift != nullptr && in) ? p.col_shift[col] : 0.f;
        float v = fmaf(__uint_as_float(rg[j]), sc, sh);
        f[j] = p.relu ? fmaxf(v, 0.f) : v;
      }
      uint8_t* rowp = sd + (c32 >> 1) * (kBlockM * 128) + row * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int chunk = (c32 & 1) * 4 + c;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = f[c * 8 + j];
        st_vec(rowp + ((chunk ^ (row & 7)) << 4), pack8(v));
      }
    }
    ptx::fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (et == 0) {
#pragma unroll
      for (int h = 0; h < (BLOCK_N + 63) / 64; ++h)
        if (n0 + h * 64 < p.N) ptx::tma_store_2d(&tmD, sd + h * (kBlockM * 128), n0 + h * 64, m0);
      ptx::tma_store_commit();
      ptx::tma_store_wait_read0();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N, int STAGES>
const char* launch(const GemmFp8Args& g, cudaStream_t stream) {
  using L = Smem<BLOCK_N, STAGES>;
  alignas(64) CUtensorMap tmA, tmB, tmD;
  {
    const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.M};
    const uint64_t st[1] = {(uint64_t)g.lda};
    const uint32_t box[2] = {kBlockKBytes, kBlockM};
    if (const char* e = encode_tmap(&tmA, g.A, 2, dims, st, box, 1)) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.N};
    const uint64_t st[1] = {(uint64_t)g.ldb};
    const uint32_t box[2] = {kBlockKBytes, (uint32_t)BLOCK_N};
    if (const char* e = encode_tmap(&tmB, g.B, 2, dims, st, box, 1)) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)g.N, (uint64_t)g.M};
    const uint64_t st[1] = {(uint64_t)g.ldd * 2};
    const uint32_t box[2] = {64, kBlockM};
    if (const char* e = encode_tmap(&tmD, g.D, 2, dims, st, box, 2)) return e;
  }
  auto kern = gemm_fp8_tcgen05_kernel<BLOCK_N, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    apply_carveout((const void*)kern);
    attr_set = true;
  }
  Fp8Params p{g.M, g.N, g.K, g.col_scale, g.col_shift, g.relu ? 1 : 0};
  const int tiles = ((g.M + kBlockM - 1) / kBlockM) * ((g.N + BLOCK_N - 1) / BLOCK_N);
  kern<<<tiles, kThreads, L::kTotal, stream>>>(tmA, tmB, tmD, p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

__global__ void __launch_bounds__(256)
quantize_e4m3_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q, int64_t n,
                     const float* __restrict__ scale, float* __restrict__ amax_out) {
  const float inv = 1.f / *scale;
  float amax = 0.f;
  const int64_t nvec = n / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const bf16x8 v = ld_vec(x + i * 8);
    float f[8];
    unpack8(v, f);
    uint8_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      amax = fmaxf(amax, fabsf(f[j]));
      o[j] = (uint8_t)__nv_cvt_float_to_fp8(f[j] * inv, __NV_SATFINITE, __NV_E4M3);
    }
    *reinterpret_cast<uint2*>(q + i * 8) = *reinterpret_cast<const uint2*>(o);
  }
  if (blockIdx.x == 0)
    for (int64_t i = nvec * 8 + threadIdx.x; i < n; i += blockDim.x) {
      const float f = __bfloat162float(x[i]);
      amax = fmaxf(amax, fabsf(f));
      q[i] = (uint8_t)__nv_cvt_float_to_fp8(f * inv, __NV_SATFINITE, __NV_E4M3);
    }
  if (amax_out != nullptr) {
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(amax_out), __float_as_int(amax));  // amax >= 0
  }
}

}  // namespace

const char* gemm_fp8(const GemmFp8Args& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return "empty GEMM";
  if (g.K % 16 != 0 || g.lda % 16 != 0 || g.ldb % 16 != 0 || g.ldd % 8 != 0) return "fp8 GEMM needs 16-byte aligned rows";
  if (g.device >= 0) {
    cudaError_t e = cudaSetDevice(g.device);
    if (e != cudaSuccess) return cudaGetErrorString(e);
  }
  return g.N <= 64 ? launch<64, 4>(g, stream) : launch<128, 3>(g, stream);
}

void quantize_e4m3(const void* x, void* q, int64_t n, const float* scale, float* amax_out, cudaStream_t s) {
  int64_t blocks = (n / 8 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  quantize_e4m3_kernel<<<(int)blocks, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                    reinterpret_cast<uint8_t*>(q), n, scale, amax_out);
}

}  // namespace edl
