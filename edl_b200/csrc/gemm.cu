// tcgen05 / TMEM / TMA GEMM for sm_100a with fused epilogues (SURVEY K1, K5 and the conv-as-GEMM
// half of K2).  The reference gets these from cuDNN/cuBLAS via Paddle
// (example/distill/resnet/models/resnet_vd.py:153-162 conv2d, :135-141 fc); this is an
// independent Blackwell-native implementation.
//
//   D[M,N] = A[M,K] * B[N,K]^T          (NHWC 1x1 conv fwd, FC)      A K-major,  B K-major
//   D[M,N] = A[M,K] * B[K,N]            (1x1 conv dgrad)             A K-major,  B MN-major
//   D[M,N] = A[K,M]^T * B[K,N]          (1x1 conv wgrad, split-K)    A MN-major, B MN-major
//
// Structure (one CTA per 128 x BLOCK_N output tile, 2 CTAs resident per SM):
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (fp32 accumulators in TMEM)
//   warps 2..5  : epilogue: tcgen05.ld TMEM -> registers -> per-channel scale/shift/ReLU ->
//                 bf16 -> swizzled smem -> TMA store; per-channel sum / sum^2 of the stored values
//                 (train-mode BatchNorm statistics) reduced from the staged tile and pushed with
//                 fp32 atomics, so BN needs no extra pass over the conv output.
//                 Split-K variant: fp32 vector reductions (red.global.add.v4.f32) instead.
#include <cuda.h>
#include <cstdio>
#include <mutex>

#include "gemm.h"
#include "kernels.h"
#include "ptx.cuh"

namespace edl {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle span
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 192;
constexpr int kEpiThreads = 128;

struct GemmParams {
  int M, N, K;
  int kb_per_split;
  const float* col_scale;
  const float* col_shift;
  int relu;
  float* col_stats;  // [2N]
  float* out_f32;    // split-K accumulation target, row-major [M, N]
  // fused split-K finalize: the last CTA of each output tile converts the fp32 tile to bf16 into
  // out_bf16 (+= when accumulate) and re-zeroes the fp32 workspace + its tile counter
  __nv_bfloat16* out_bf16;
  long long ldo;
  int* tile_counters;
  int accumulate;
  float* part;       // split-K partial tiles [split][tiles][BLOCK_M x BLOCK_N] (no atomics) or nullptr
  // GEMM -> peer ship (see gemm.h)
  uint32_t* ship_flag;
  const uint32_t* ship_seq_ptr;
  uint32_t ship_seq_imm;
  unsigned int* ship_done;
};

template <int BLOCK_N, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kDBytes = BLOCK_M * BLOCK_N * 2;
  static constexpr int kTileBytes =
      STAGES * kStageBytes > kDBytes ? STAGES * kStageBytes : kDBytes;
  static constexpr int kBarOffset = kTileBytes;
  static constexpr int kTotal = kTileBytes + 256 + 1024;  // barriers + alignment slack
};

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmD, const GemmParams p) {
  using L = SmemLayout<BLOCK_N, STAGES>;
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
  const int tile = blockIdx.x;
  const int m0 = (tile / tiles_n) * BLOCK_M;
  const int n0 = (tile % tiles_n) * BLOCK_N;
  const int total_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > total_kb) kb_end = total_kb;
  const int num_kb = kb_end - kb_begin;

  constexpr uint32_t kTmemCols = BLOCK_N <= 32 ? 32 : (BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256));

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if (EPI == 0) ptx::prefetch_tmap(&tmD);
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

  if (num_kb > 0) {
    if (warp == 0) {
      // ------------------------------------------------------------ TMA producer
      // warp-uniform loop, one elected lane issues (see gemm_persist.cu: loops under `if (lane == 0)` cost a vote +
      // broadcast sequence per uniform-register operand)
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * L::kStageBytes;
        uint8_t* sb = sa + L::kABytes;
        const int k0 = (kb_begin + i) * BLOCK_K;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full_bar[s], L::kStageBytes);
          if (!A_MN) {
            ptx::tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
          } else {
#pragma unroll
            for (int h = 0; h < BLOCK_M / 64; ++h)
              ptx::tma_load_2d(sa + h * 8192, &tmA, &full_bar[s], m0 + h * 64, k0);
          }
          if (!B_MN) {
            ptx::tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
          } else {
#pragma unroll
            for (int h = 0; h < BLOCK_N / 64; ++h)
              ptx::tma_load_2d(sb + h * 8192, &tmB, &full_bar[s], n0 + h * 64, k0);
          }
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------ MMA issuer (one elected lane, warp-uniform loop)
      constexpr uint32_t idesc = ptx::make_idesc(1, 1, BLOCK_M, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + s * L::kStageBytes);
        const uint32_t sb = sa + L::kABytes;
        // one descriptor per operand and stage; a k step adds 32 bytes (K-major) or 2048 bytes (MN-major) to the start
        // address field, which counts 16-byte units
        const uint64_t da0 = A_MN ? ptx::make_smem_desc(sa, 8192, 1024) : ptx::make_smem_desc(sa, 16, 1024);
        const uint64_t db0 = B_MN ? ptx::make_smem_desc(sb, 8192, 1024) : ptx::make_smem_desc(sb, 16, 1024);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            ptx::umma_f16(tmem_base, da0 + (uint64_t)(k * (A_MN ? 128 : 2)), db0 + (uint64_t)(k * (B_MN ? 128 : 2)), idesc,
                          (i | k) != 0 ? 1u : 0u);
          ptx::umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs have read it
          if (i == num_kb - 1) ptx::umma_commit(tmem_full_bar);  // accumulator complete
        }
        __syncwarp();
      }
    } else {
      // ------------------------------------------------------------ epilogue (warps 2..5)
      const int q = warp & 3;             // TMEM lane quarter this warp may access
      const int row = q * 32 + lane;      // accumulator row within the tile
      const int et = threadIdx.x - 64;    // 0..127 epilogue thread id
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      if (EPI == 0) {
        uint8_t* sd = smem;  // pipeline stages are drained: reuse them as the store staging tile
#pragma unroll 1
        for (int c32 = 0; c32 < BLOCK_N / 32; ++c32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(taddr + c32 * 32, r);
          ptx::tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]);
          if (p.col_scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = n0 + c32 * 32 + j;
              const float sc = col < p.N ? p.col_scale[col] : 0.f;
              f[j] *= sc;
            }
          }
          if (p.col_shift != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = n0 + c32 * 32 + j;
              f[j] += col < p.N ? p.col_shift[col] : 0.f;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          const int half = c32 >> 1;
          uint8_t* rowp = sd + half * (BLOCK_M * 128) + row * 128;
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
            if (n0 + h * 64 < p.N) ptx::tma_store_2d(&tmD, sd + h * (BLOCK_M * 128), n0 + h * 64, m0);
          ptx::tma_store_commit();
        }
        if (p.col_stats != nullptr) {
          int rows_valid = p.M - m0;
          if (rows_valid > BLOCK_M) rows_valid = BLOCK_M;
          for (int col = et; col < BLOCK_N; col += kEpiThreads) {
            const bool valid = n0 + col < p.N;     // no early exit: the whole warp shuffles below
            const int half = col >> 6, cc = col & 63, chunk = cc >> 3, within = cc & 7;
            const uint8_t* base = sd + half * (BLOCK_M * 128) + within * 2;
            float s = 0.f, sq = 0.f;
#pragma unroll 8
            for (int rr = 0; rr < (valid ? rows_valid : 0); ++rr) {
              const __nv_bfloat16 hv = *reinterpret_cast<const __nv_bfloat16*>(
                  base + rr * 128 + ((chunk ^ (rr & 7)) << 4));
              const float v = __bfloat162float(hv);
              s += v;
              sq = fmaf(v, v, sq);
            }
            // four neighbouring columns -> one vector reduction (lane 4i collects lanes 4i..4i+3)
            const float s1 = __shfl_down_sync(0xffffffffu, s, 1), s2 = __shfl_down_sync(0xffffffffu, s, 2),
                        s3 = __shfl_down_sync(0xffffffffu, s, 3);
            const float q1 = __shfl_down_sync(0xffffffffu, sq, 1), q2 = __shfl_down_sync(0xffffffffu, sq, 2),
                        q3 = __shfl_down_sync(0xffffffffu, sq, 3);
            float* ps = &p.col_stats[n0 + col];
            float* pq = &p.col_stats[p.N + n0 + col];
            const int c4 = col & ~3;   // the decision is per group of four columns, identical in its 4 lanes
            const bool vec = (p.N % 4 == 0) && n0 + c4 + 3 < p.N &&
                             ((reinterpret_cast<uintptr_t>(&p.col_stats[n0 + c4]) & 15) == 0);
            if (vec) {
              if ((col & 3) == 0) {
                red_add_v4(ps, s, s1, s2, s3);
                red_add_v4(pq, sq, q1, q2, q3);
              }
            } else if (valid) {
              atomicAdd(ps, s);
              atomicAdd(pq, sq);
            }
          }
        }
        if (et == 0) {
          if (p.ship_flag == nullptr) {
            ptx::tma_store_wait_read0();
          } else {
            // the tile went to a peer GPU: wait until the bulk store has fully completed (not just
            // read its smem source), order it system-wide, count this CTA in; the last one raises
            // the consumer's flag.
            ptx::tma_store_wait0();
            __threadfence_system();
            const unsigned int prev = atomicAdd(p.ship_done, 1u);
            if (prev == gridDim.x - 1) {
              *p.ship_done = 0u;
              __threadfence_system();
              const uint32_t seq = p.ship_seq_ptr != nullptr ? *p.ship_seq_ptr : p.ship_seq_imm;
              asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.ship_flag), "r"(seq) : "memory");
            }
          }
        }
      } else {
        // split-K: stage the fp32 partial tile in smem (the pipeline stages are drained), then push
        // it to the fp32 workspace with COALESCED vector reductions: consecutive threads cover
        // consecutive 16-byte pieces of a row (row-per-thread reds touch 32 cache lines per warp
        // instruction and measured 5-20x slower, profiles/kernels_r1.txt).
        constexpr int kPitch = BLOCK_N * 4 + 16;        // bytes; +16 keeps v4 stores off one bank
        constexpr int kVecPerRow = BLOCK_N / 4;         // float4 per row
        uint8_t* sf = smem;
#pragma unroll 1
        for (int c32 = 0; c32 < BLOCK_N / 32; ++c32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(taddr + c32 * 32, r);
          ptx::tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(sf + row * kPitch + c32 * 128);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            dst[c] = make_float4(__uint_as_float(r[c * 4]), __uint_as_float(r[c * 4 + 1]),
                                 __uint_as_float(r[c * 4 + 2]), __uint_as_float(r[c * 4 + 3]));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        int rows_valid = p.M - m0;
        if (rows_valid > BLOCK_M) rows_valid = BLOCK_M;
        int cols_valid = p.N - n0;
        if (cols_valid > BLOCK_N) cols_valid = BLOCK_N;
        const bool vec_ok = (p.N % 4) == 0;
        if (gridDim.z == 1 && p.out_bf16 != nullptr) {
          // no split: this CTA owns the tile -> write bf16 (+= sink) straight from the staged tile,
          // coalesced, without touching the fp32 workspace, atomics or tile counters
          for (int f = et; f < rows_valid * kVecPerRow; f += kEpiThreads) {
            const int rr = f / kVecPerRow, c4 = (f % kVecPerRow) * 4;
            if (c4 + 3 >= cols_valid) continue;
            float4 t = *reinterpret_cast<const float4*>(sf + rr * kPitch + c4 * 4);
            __nv_bfloat16* o = p.out_bf16 + (int64_t)(m0 + rr) * p.ldo + n0 + c4;
            if (p.accumulate) {
              const uint2 old = *reinterpret_cast<const uint2*>(o);
              const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&old.x));
              const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&old.y));
              t.x += a.x; t.y += a.y; t.z += b.x; t.w += b.y;
            }
            const __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y);
            const __nv_bfloat162 hi = __floats2bfloat162_rn(t.z, t.w);
            uint2 packed;
            packed.x = *reinterpret_cast<const uint32_t*>(&lo);
            packed.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(o) = packed;
          }
        } else if (p.part != nullptr) {
          // split-K partial tile: plain coalesced 16-byte stores, summed by splitk_reduce_kernel afterwards
          float* dst = p.part + ((int64_t)blockIdx.z * gridDim.x + tile) * (BLOCK_M * BLOCK_N);
          for (int f = et; f < BLOCK_M * kVecPerRow; f += kEpiThreads) {
            const int rr = f / kVecPerRow, c4 = (f % kVecPerRow) * 4;
            *reinterpret_cast<float4*>(dst + rr * BLOCK_N + c4) = *reinterpret_cast<const float4*>(sf + rr * kPitch + c4 * 4);
          }
        } else {
        for (int f = et; f < rows_valid * kVecPerRow; f += kEpiThreads) {
          const int rr = f / kVecPerRow, c4 = (f % kVecPerRow) * 4;
          if (c4 >= cols_valid) continue;
          const float4 v = *reinterpret_cast<const float4*>(sf + rr * kPitch + c4 * 4);
          float* dst = p.out_f32 + (int64_t)(m0 + rr) * p.N + n0 + c4;
          if (vec_ok && c4 + 3 < cols_valid) {
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(v.x), "f"(v.y),
                         "f"(v.z), "f"(v.w)
                         : "memory");
          } else {
            const float vv[4] = {v.x, v.y, v.z, v.w};
            for (int j = 0; j < 4; ++j)
              if (c4 + j < cols_valid) atomicAdd(dst + j, vv[j]);
          }
        }
        }
        if (p.tile_counters != nullptr && gridDim.z > 1 && p.part == nullptr) {
          // ---- fused finalize by the last-arriving CTA of this output tile (coalesced) ----
          uint32_t* s_last = tmem_slot + 1;
          __threadfence();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            const int old = atomicAdd(&p.tile_counters[tile], 1);
            const int last = old == (int)gridDim.z - 1;
            if (last) p.tile_counters[tile] = 0;
            *s_last = (uint32_t)last;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (*s_last != 0u) {
            __threadfence();
            const int total = rows_valid * kVecPerRow;
            constexpr int U = 4;
            for (int f0 = et; f0 < total; f0 += kEpiThreads * U) {
              float4 v[U];
              bool ok[U];
#pragma unroll
              for (int u = 0; u < U; ++u) {
                const int f = f0 + u * kEpiThreads;
                const int rr = f / kVecPerRow, c4 = (f % kVecPerRow) * 4;
                ok[u] = f < total && c4 + 3 < cols_valid;
                if (ok[u])
                  v[u] = __ldcg(reinterpret_cast<const float4*>(
                      p.out_f32 + (int64_t)(m0 + rr) * p.N + n0 + c4));
              }
#pragma unroll
              for (int u = 0; u < U; ++u) {
                if (!ok[u]) continue;
                const int f = f0 + u * kEpiThreads;
                const int rr = f / kVecPerRow, c4 = (f % kVecPerRow) * 4;
                float* w = p.out_f32 + (int64_t)(m0 + rr) * p.N + n0 + c4;
                __nv_bfloat16* o = p.out_bf16 + (int64_t)(m0 + rr) * p.ldo + n0 + c4;
                *reinterpret_cast<float4*>(w) = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 t = v[u];
                if (p.accumulate) {
                  const uint2 old = *reinterpret_cast<const uint2*>(o);
                  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&old.x));
                  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&old.y));
                  t.x += a.x; t.y += a.y; t.z += b.x; t.w += b.y;
                }
                const __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y);
                const __nv_bfloat162 hi = __floats2bfloat162_rn(t.z, t.w);
                uint2 packed;
                packed.x = *reinterpret_cast<const uint32_t*>(&lo);
                packed.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(o) = packed;
              }
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// -------------------------------------------------------------------------------------------
// host side

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

thread_local char g_tmap_err[256];

// 2-D bf16 tensor map: dims {inner, outer}, row pitch in elements, box {box_inner, box_outer}.
bool make_tmap_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer,
                  uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    snprintf(g_tmap_err, sizeof(g_tmap_err), "cuTensorMapEncodeTiled entry point not found");
    return false;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    snprintf(g_tmap_err, sizeof(g_tmap_err),
             "cuTensorMapEncodeTiled failed: CUresult=%d ptr=%p dims={%llu,%llu} pitch=%llu box={%u,%u}",
             (int)r, ptr, (unsigned long long)inner, (unsigned long long)outer,
             (unsigned long long)pitch_elems, box_inner, box_outer);
  return r == CUDA_SUCCESS;
}

}  // namespace

// N-d bf16 tensor map (128B swizzle, zero OOB fill) for the other TMA kernels (conv3x3.cu).
// dims / box are innermost-first; strides_bytes has rank-1 entries (dims 1..rank-1).
const char* encode_tmap_bf16(void* out, const void* ptr, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_tmap(out, ptr, rank, dims, strides_bytes, box, 2);
}

const char* encode_tmap(void* out, const void* ptr, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, int elem_bytes) {
  return encode_tmap_strided(out, ptr, rank, dims, strides_bytes, box, nullptr, elem_bytes);
}

// elem_strides (optional): traversal stride per dimension.  With stride t the unit loads ceil(box / t) elements
// along that dimension (every t-th one), i.e. box[i] = n * t selects n elements -- how a stride-2 convolution reads
// its input without an im2col or a subsampling pass.
const char* encode_tmap_strided(void* out, const void* ptr, int rank, const uint64_t* dims,
                                const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                                int elem_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return "cuTensorMapEncodeTiled entry point not found";
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = elem_strides != nullptr ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out),
                  elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                  const_cast<void*>(ptr), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_tmap_err, sizeof(g_tmap_err),
             "cuTensorMapEncodeTiled(rank %d) failed: CUresult=%d ptr=%p dims={%llu,%llu,..} box={%u,%u,..}",
             rank, (int)r, ptr, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return g_tmap_err;
  }
  return nullptr;
}

namespace {

template <int BLOCK_N, int STAGES, bool A_MN, bool B_MN, int EPI>
const char* launch_variant(const GemmArgs& g, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, STAGES>;
  CUtensorMap tmA, tmB, tmD;
  bool ok = true;
  // A: K-major -> global [M rows][K] ; MN-major -> global [K rows][M]
  if (!A_MN) ok &= make_tmap_2d(&tmA, g.A, g.K, g.M, g.lda, BLOCK_K, BLOCK_M);
  else ok &= make_tmap_2d(&tmA, g.A, g.M, g.K, g.lda, 64, BLOCK_K);
  if (!B_MN) ok &= make_tmap_2d(&tmB, g.B, g.K, g.N, g.ldb, BLOCK_K, BLOCK_N);
  else ok &= make_tmap_2d(&tmB, g.B, g.N, g.K, g.ldb, 64, BLOCK_K);
  if (EPI == 0) ok &= make_tmap_2d(&tmD, g.D, g.N, g.M, g.ldd, BLOCK_N < 64 ? BLOCK_N : 64, BLOCK_M);
  else tmD = tmA;
  if (!ok) return g_tmap_err;
  auto kern = gemm_tcgen05_kernel<BLOCK_N, STAGES, A_MN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    apply_carveout((const void*)kern);
    attr_set = true;
  }
  GemmParams p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  const int total_kb = (g.K + BLOCK_K - 1) / BLOCK_K;
  int split = EPI == 1 ? (g.split_k < 1 ? 1 : g.split_k) : 1;
  if (split > total_kb) split = total_kb;
  p.kb_per_split = (total_kb + split - 1) / split;
  split = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.col_scale = g.col_scale;
  p.col_shift = g.col_shift;
  p.relu = g.relu ? 1 : 0;
  p.col_stats = g.col_stats;
  p.out_f32 = g.out_f32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(g.out_bf16);
  p.ldo = g.ldo;
  p.tile_counters = g.tile_counters;
  p.accumulate = g.accumulate_out ? 1 : 0;
  p.part = nullptr;
  p.ship_flag = EPI == 0 ? reinterpret_cast<uint32_t*>(g.ship_flag) : nullptr;
  p.ship_seq_ptr = reinterpret_cast<const uint32_t*>(g.ship_seq_ptr);
  p.ship_seq_imm = g.ship_seq_imm;
  p.ship_done = reinterpret_cast<unsigned int*>(g.ship_done);
  const int tiles_m = (g.M + BLOCK_M - 1) / BLOCK_M;
  const int tiles_n = (g.N + BLOCK_N - 1) / BLOCK_N;
  dim3 grid(tiles_m * tiles_n, 1, split);
  const bool use_part = EPI == 1 && split > 1 && g.partials != nullptr && g.out_bf16 != nullptr && (g.N % 4) == 0 &&
                        g.partials_elems >= (int64_t)split * grid.x * BLOCK_M * BLOCK_N;
  if (use_part) p.part = g.partials;
  kern<<<grid, kGemmThreads, L::kTotal, stream>>>(tmA, tmB, tmD, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cudaGetErrorString(e);
  if (use_part) {
    splitk_reduce(g.partials, split, (int)grid.x, tiles_n, BLOCK_M, BLOCK_N, 1, 0, g.M, g.N, g.out_bf16, g.ldo,
                  g.accumulate_out, stream);
    e = cudaGetLastError();
  }
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// Sum of the split-K partial tiles -> bf16 output (+= when accumulating).  A CTA of 256 threads covers 64 consecutive
// float4 of a tile row; its 4 thread groups each sum every 4th split (independent 16-byte loads, unrolled so that
// several are in flight), a shared-memory pass adds the 4 partial sums.  The kernel is latency-bound by construction
// (a few MB out of L2): the z-parallel layout keeps it at 2-4 us instead of split x load latency.
constexpr int kRedVec = 64;     // float4 per CTA

template <int kRedZ>            // split groups per CTA: 4, or 16 for deep splits (few tiles, 32+ partials each)
__global__ void __launch_bounds__(kRedVec * kRedZ)
splitk_reduce_kernel(const float* __restrict__ part, int split, int ctas, int tiles_n, int rows, int cols, int taps,
                     int cin, int M, int N, __nv_bfloat16* __restrict__ out, long long ldo, int accumulate) {
  __shared__ float4 sm[kRedZ][kRedVec];
  const int vec_per_row = cols / 4;
  const int64_t per_tile = (int64_t)rows * vec_per_row;
  const int64_t total = (int64_t)ctas * per_tile;
  const int64_t tile_elems = (int64_t)rows * cols;
  const int64_t zstride = (int64_t)ctas * tile_elems;
  const int vl = threadIdx.x % kRedVec, q = threadIdx.x / kRedVec;
  for (int64_t base = (int64_t)blockIdx.x * kRedVec; base < total; base += (int64_t)gridDim.x * kRedVec) {
    const int64_t i = base + vl;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    bool ok = i < total;
    int m = 0, n = 0;
    if (ok) {
      const int t = (int)(i / per_tile);
      const int rem = (int)(i - (int64_t)t * per_tile);
      const int rr = rem / vec_per_row, c4 = (rem - rr * vec_per_row) * 4;
      if (taps == 1) {
        m = (t / tiles_n) * rows + rr;
        n = (t % tiles_n) * cols + c4;
      } else {
        const int tile = t / 3, r = t - tile * 3, bn = cols / 3;
        const int s = c4 / bn, cc = c4 - s * bn;
        m = (tile / tiles_n) * rows + rr;
        n = (r * 3 + s) * cin + (tile % tiles_n) * bn + cc;
        ok = (tile % tiles_n) * bn + cc + 3 < cin;
      }
      ok = ok && m < M && n + 3 < N;
      if (ok) {
        const float* src = part + (int64_t)t * tile_elems + (int64_t)rr * cols + c4;
        int z = q;
        for (; z + 3 * kRedZ < split; z += 4 * kRedZ) {
          const float4 a = __ldcg(reinterpret_cast<const float4*>(src + (int64_t)z * zstride));
          const float4 b = __ldcg(reinterpret_cast<const float4*>(src + (int64_t)(z + kRedZ) * zstride));
          const float4 c = __ldcg(reinterpret_cast<const float4*>(src + (int64_t)(z + 2 * kRedZ) * zstride));
          const float4 d = __ldcg(reinterpret_cast<const float4*>(src + (int64_t)(z + 3 * kRedZ) * zstride));
          acc.x += (a.x + b.x) + (c.x + d.x); acc.y += (a.y + b.y) + (c.y + d.y);
          acc.z += (a.z + b.z) + (c.z + d.z); acc.w += (a.w + b.w) + (c.w + d.w);
        }
        for (; z < split; z += kRedZ) {
          const float4 a = __ldcg(reinterpret_cast<const float4*>(src + (int64_t)z * zstride));
          acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        }
      }
    }
    sm[q][vl] = acc;
    __syncthreads();
    if (q == 0 && ok) {
#pragma unroll
      for (int k = 1; k < kRedZ; ++k) {
        const float4 o = sm[k][vl];
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
      }
      __nv_bfloat16* o = out + (int64_t)m * ldo + n;
      if (accumulate) {
        const uint2 old = *reinterpret_cast<const uint2*>(o);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&old.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&old.y));
        acc.x += a.x; acc.y += a.y; acc.z += b.x; acc.w += b.y;
      }
      const __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y);
      const __nv_bfloat162 hi = __floats2bfloat162_rn(acc.z, acc.w);
      uint2 packed;
      packed.x = *reinterpret_cast<const uint32_t*>(&lo);
      packed.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(o) = packed;
    }
    __syncthreads();
  }
}

}  // namespace

void splitk_reduce(const float* partials, int split, int ctas, int tiles_n, int rows, int cols, int taps, int cin, int M,
                   int N, void* out_bf16, int64_t ldo, bool accumulate, cudaStream_t stream) {
  const int64_t total = (int64_t)ctas * rows * (cols / 4);
  int blocks = (int)((total + kRedVec - 1) / kRedVec);
  if (blocks < 1) blocks = 1;
  auto* out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (split >= 32) {
    if (blocks > kNumSMs * 2) blocks = kNumSMs * 2;
    splitk_reduce_kernel<16><<<blocks, kRedVec * 16, 0, stream>>>(partials, split, ctas, tiles_n, rows, cols, taps, cin,
                                                                 M, N, out, (long long)ldo, accumulate ? 1 : 0);
  } else {
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    splitk_reduce_kernel<4><<<blocks, kRedVec * 4, 0, stream>>>(partials, split, ctas, tiles_n, rows, cols, taps, cin, M,
                                                               N, out, (long long)ldo, accumulate ? 1 : 0);
  }
}

const char* gemm_bf16(const GemmArgs& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return "empty GEMM";
  // cuTensorMapEncodeTiled is a driver call and needs a current context on THIS thread.  Autograd
  // worker threads may not have one yet (torch binds lazily; observed CUDA_ERROR_INVALID_CONTEXT
  // = 201 when a dgrad GEMM was the first CUDA work of the backward thread).
  if (g.device >= 0) {
    cudaError_t e = cudaSetDevice(g.device);
    if (e != cudaSuccess) return cudaGetErrorString(e);
  }
  const bool n64 = g.N <= 64;
  if (g.out_f32 == nullptr && !g.a_mn_major && g.ship_flag == nullptr && persistent_gemm_enabled())
    return gemm_bf16_persistent(g, stream);
  if (g.add_src != nullptr) return "gemm add_src is implemented by the persistent kernel only";
  if (g.out_f32 != nullptr) {
    // split-K fp32 accumulation (wgrad): both operands MN-major or both K-major
    if (g.a_mn_major && g.b_mn_major)
      return n64 ? launch_variant<64, 4, true, true, 1>(g, stream)
                 : launch_variant<128, 3, true, true, 1>(g, stream);
    if (!g.a_mn_major && !g.b_mn_major)
      return n64 ? launch_variant<64, 4, false, false, 1>(g, stream)
                 : launch_variant<128, 3, false, false, 1>(g, stream);
    return "unsupported split-K operand layout";
  }
  if (!g.a_mn_major && !g.b_mn_major)
    return n64 ? launch_variant<64, 4, false, false, 0>(g, stream)
               : launch_variant<128, 3, false, false, 0>(g, stream);
  if (!g.a_mn_major && g.b_mn_major)
    return n64 ? launch_variant<64, 4, false, true, 0>(g, stream)
               : launch_variant<128, 3, false, true, 0>(g, stream);
  return "unsupported operand layout";
}

}  // namespace edl
