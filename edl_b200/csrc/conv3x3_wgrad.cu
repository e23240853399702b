// Weight gradient of the 3x3 / stride 1 / pad 1 NHWC convolution on tcgen05 (wgrad third).
// The reference gets it from cuDNN through Paddle's conv2d_grad (example/distill/resnet/models/
// resnet_vd.py:153-162); round 1 of this repo also used the library kernel on the side stream.
//
//   dW[co, r, s, ci] = sum over pixels (n, h, w) of  dY[n, h, w, co] * X[n, h + r - 1, w + s - 1, ci]
//
// As a GEMM per filter tap: M = Cout, N = Cin, K = pixels, both operands "MN-major" (the channel dimension is
// the contiguous one in NHWC, the reduction runs over pixels):
//   A tile = 4-D TMA box {64 ch, W, BH, NB} of dY          -> smem [KB pixel rows][64 ch = 128 B], 128B swizzle
//   B tile = the SAME box of X shifted by (s - 1, r - 1)    -> out-of-image pixels are zero-filled by the TMA
//                                                              unit (= the padding; no im2col, no index math)
// One CTA owns a 128 (Cout) x 64 (Cin) tile of ONE FILTER ROW r: the dY tile of a pixel block is loaded once
// and multiplied with the three shifted X tiles (s = 0, 1, 2) into three TMEM accumulators, so dY crosses
// L2 -> SM once per row instead of once per tap.  The pixel reduction is split over gridDim.z CTAs; partial
// tiles are pushed with coalesced red.global.add.v4.f32 into an all-zero fp32 workspace laid out like the
// KRSC weight, and the LAST CTA of a tile converts it to bf16 straight into the flat gradient bucket
// (+= when accumulating) and re-zeroes the workspace -- the same fused finalize as the 1x1 wgrad GEMM (gemm.cu).
//
//   warp 0      : TMA producer (2 + 3 box loads per pixel block into a 3-stage ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (3 accumulators x 64 fp32 columns)
//   warps 2..5  : epilogue (tcgen05.ld -> smem staging -> vector reductions / finalize)
#include <cuda.h>

#include "gemm.h"
#include "kernels.h"
#include "ptx.cuh"

namespace edl {
namespace {

constexpr int kBlockM = 128;      // Cout rows per tile
constexpr int kBlockN = 64;       // Cin columns per tile (one 128-byte swizzle span)
constexpr int kMaxKB = 112;       // pixel rows per K block (multiple of 16; 112 = 2x56 = 4x28 = 8x14 = 16x7)
constexpr int kUmmaK = 16;
constexpr int kStages = 3;
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kChunkBytes = kMaxKB * 128;                 // one [KB][64 ch] operand chunk (stride in the ring)
constexpr int kStageBytes = (2 + 3) * kChunkBytes;        // A: 2 chunks of 64 Cout; B: 3 taps x 64 Cin
constexpr int kBarOffset = kStages * kStageBytes;
constexpr int kSmemTotal = kBarOffset + 256 + 1024;
constexpr uint32_t kTmemCols = 256;                       // 3 x 64 accumulator columns, power of two
constexpr int kPitch = kBlockN * 4 + 16;                  // fp32 staging row pitch (bytes)
constexpr int kVecPerRow = kBlockN / 4;

struct WgradParams {
  int Cout, Cin;
  int BH, NB, HB;            // box rows / images, row blocks per image
  int KB;                    // pixel rows per block = W * BH * NB (multiple of 16, <= kMaxKB)
  int total_kb, kb_per_split;
  int tiles_n;
  float* ws;                 // fp32 [Cout][9*Cin], all zero on entry and on exit
  __nv_bfloat16* out;        // bf16 [Cout][9*Cin] (KRSC)
  int* counters;             // [tiles * 3], zero on entry and on exit
  int accumulate;
  int stride;                // 1, or 2: X rows 2*h + r - 1 (the columns are subsampled by the tensor map)
  float* part;               // split-K partial tiles [split][ctas][128 x 192] (no atomics) or nullptr
};

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                     const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // blockIdx.x = (tile * 3 + r); tile = mt * tiles_n + nt
  const int r = blockIdx.x % 3;
  const int tile = blockIdx.x / 3;
  const int m0 = (tile / p.tiles_n) * kBlockM;
  const int n0 = (tile % p.tiles_n) * kBlockN;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > p.total_kb) kb_end = p.total_kb;
  const int num_kb = kb_end - kb_begin;
  const bool two_chunks = m0 + 64 < p.Cout;     // Cout == 64: rows 64..127 of the accumulators are never stored
  const uint32_t chunk_tx = (uint32_t)p.KB * 128u;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmDy);
    ptx::prefetch_tmap(&tmX);
    for (int s = 0; s < kStages; ++s) {
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
      // warp-uniform loop, one elected lane issues (gemm_persist.cu explains why not `if (lane == 0)` around the loop)
      for (int i = 0; i < num_kb; ++i) {
        const int st = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        ptx::mbar_wait(&empty_bar[st], ph ^ 1);
        uint8_t* sa = smem + st * kStageBytes;
        uint8_t* sb = sa + 2 * kChunkBytes;
        const int kb = kb_begin + i;
        const int h0 = (kb % p.HB) * p.BH;
        const int img0 = (kb / p.HB) * p.NB;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full_bar[st], chunk_tx * (two_chunks ? 5u : 4u));
          ptx::tma_load_4d(sa, &tmDy, &full_bar[st], m0, 0, h0, img0);
          if (two_chunks) ptx::tma_load_4d(sa + kChunkBytes, &tmDy, &full_bar[st], m0 + 64, 0, h0, img0);
#pragma unroll
          for (int s = 0; s < 3; ++s)
            ptx::tma_load_4d(sb + s * kChunkBytes, &tmX, &full_bar[st], n0, s - 1, h0 * p.stride + r - 1, img0);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------ MMA issuer (one thread)
      constexpr uint32_t idesc = ptx::make_idesc(1, 1, kBlockM, kBlockN, 1, 1);
      const int ksteps = p.KB / kUmmaK;
      for (int i = 0; i < num_kb; ++i) {
        const int st = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        ptx::mbar_wait(&full_bar[st], ph);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + st * kStageBytes);
        const uint32_t sb = sa + 2 * kChunkBytes;
        // MN-major, 128B swizzle: 16 pixel rows = 2048 B per UMMA_K step (128 in the descriptor's 16-byte units),
        // 8-row groups 1024 B apart, the second 64-channel chunk of A one ring chunk further
        const uint64_t da0 = ptx::make_smem_desc(sa, kChunkBytes, 1024);
        const uint64_t db0 = ptx::make_smem_desc(sb, kChunkBytes, 1024);
        if (ptx::elect_one()) {
          for (int k = 0; k < ksteps; ++k) {
#pragma unroll
            for (int s = 0; s < 3; ++s)
              ptx::umma_f16(tmem_base + s * kBlockN, da0 + (uint64_t)(k * 128),
                            db0 + (uint64_t)(s * (kChunkBytes >> 4) + k * 128), idesc, (i | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[st]);
          if (i == num_kb - 1) ptx::umma_commit(tmem_full_bar);
        }
        __syncwarp();
      }
    } else {
      // ------------------------------------------------------------ epilogue (warps 2..5)
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const int et = threadIdx.x - 64;
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after();
      uint8_t* sf = smem;                        // the ring is drained: fp32 staging tile
      int rows_valid = p.Cout - m0;
      if (rows_valid > kBlockM) rows_valid = kBlockM;
      const int64_t ldw = (int64_t)9 * p.Cin;
      const bool direct = gridDim.z == 1;
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kBlockN);
#pragma unroll 1
        for (int c32 = 0; c32 < kBlockN / 32; ++c32) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(taddr + c32 * 32, v);
          ptx::tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(sf + row * kPitch + c32 * 128);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            dst[c] = make_float4(__uint_as_float(v[c * 4]), __uint_as_float(v[c * 4 + 1]),
                                 __uint_as_float(v[c * 4 + 2]), __uint_as_float(v[c * 4 + 3]));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int64_t col0 = (int64_t)(r * 3 + s) * p.Cin + n0;
        for (int f = et; f < rows_valid * kVecPerRow; f += kEpiThreads) {
          const int rr = f / kVecPerRow, c4 = (f % kVecPerRow) * 4;
          float4 t = *reinterpret_cast<const float4*>(sf + rr * kPitch + c4 * 4);
          const int64_t off = (int64_t)(m0 + rr) * ldw + col0 + c4;
          if (direct) {
            // this CTA saw every pixel: bf16 (+= bucket) straight from the staged tile
            __nv_bfloat16* o = p.out + off;
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
          } else if (p.part != nullptr) {
            *reinterpret_cast<float4*>(p.part + ((int64_t)blockIdx.z * gridDim.x + blockIdx.x) * (kBlockM * 3 * kBlockN) +
                                       rr * (3 * kBlockN) + s * kBlockN + c4) = t;
          } else {
            red_add_v4(p.ws + off, t.x, t.y, t.z, t.w);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // staging tile is reused by the next tap
      }
      if (!direct && p.part == nullptr) {
        // ---- fused finalize by the last-arriving CTA of this (tile, filter row) ----
        uint32_t* s_last = tmem_slot + 1;
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          const int old = atomicAdd(&p.counters[blockIdx.x], 1);
          const int last = old == (int)gridDim.z - 1;
          if (last) p.counters[blockIdx.x] = 0;
          *s_last = (uint32_t)last;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (*s_last != 0u) {
          __threadfence();
          const int total = rows_valid * (3 * kVecPerRow);      // the three taps of a row are adjacent in KRSC
          // when n0 spans the whole Cin they are contiguous; in general tap s starts at (r*3+s)*Cin + n0
          for (int f = et; f < total; f += kEpiThreads) {
            const int rr = f / (3 * kVecPerRow);
            const int rem = f % (3 * kVecPerRow);
            const int s = rem / kVecPerRow, c4 = (rem % kVecPerRow) * 4;
            const int64_t off = (int64_t)(m0 + rr) * ldw + (int64_t)(r * 3 + s) * p.Cin + n0 + c4;
            float4 t = __ldcg(reinterpret_cast<const float4*>(p.ws + off));
            *reinterpret_cast<float4*>(p.ws + off) = make_float4(0.f, 0.f, 0.f, 0.f);
            __nv_bfloat16* o = p.out + off;
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
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// =====================================================================================================
// Version 2 (round 2): ONE X load per (pixel block, filter row) serves the three taps s = 0, 1, 2.
//
// Version 1 above is L2->SM bound: 5 operand chunks per pixel block for 21 MMAs (ncu: 101-154 MB through the
// crossbar for 3 MB of inputs, tensor pipe 11-28 % busy, 43-50 us per layer against 14-30 us for the library).
// Here both operands are loaded as boxes {64 ch, Wb, BH, NB} that START AT COLUMN w = -1 (Wb >= W + 1): the TMA
// unit zero-fills column -1 (and everything right of the image), so every image row is followed by at least one
// zero pixel.  In shared memory the pixel rows (K index) of both tiles are 128 B apart, and the X pixel a tap
// (r, s) needs for the dY pixel at K row k sits at K row k + (s - 1) of the SAME X tile: the three taps are the
// same tile read through descriptors whose start address is shifted by -128 / 0 / +128 bytes.  Where the shift
// crosses an image row it lands on a zero pixel of one of the operands (dY's column -1 / X's column -1); the rows
// just outside the tile are zeroed once at kernel start.  Per pixel block: 2 dY chunks + BLOCK_N/64 X chunks
// for 3 * KB/16 MMAs of N = BLOCK_N -- with BLOCK_N = 128 that is 2.6x fewer bytes per FLOP than version 1.
constexpr int kMaxKB2 = 128;
constexpr int kChunkA2 = kMaxKB2 * 128;                    // 16 KB per 64-channel dY chunk
constexpr int kPad2 = 1024;                                // zero rows before / after an X tile (8 rows)
constexpr int kChunkB2 = kChunkA2 + 2 * kPad2;
constexpr int kStages2 = 3;

template <int BLOCK_N>
struct W2Smem {
  static constexpr int kNCh = BLOCK_N / 64;
  static constexpr int kStage = 2 * kChunkA2 + kNCh * kChunkB2;
  static constexpr int kBar = kStages2 * kStage;
  static constexpr int kTotal = kBar + 256 + 1024;
  static constexpr int kPitch = BLOCK_N * 4 + 16;
  static_assert(128 * kPitch <= kBar, "fp32 staging tile must fit in the drained ring");
};

struct Wgrad2Params {
  int Cout, Cin;
  int BH, NB, HB, KB;
  int total_kb, kb_per_split;
  int tiles_n;
  int base_offset_mode;      // 0: descriptor base offset 0 (address-based swizzle); 1: (addr >> 7) & 7
  float* ws;
  __nv_bfloat16* out;
  int* counters;
  int accumulate;
  float* part;               // split-K partial tiles [split][ctas][128 x 3*BLOCK_N] (no atomics) or nullptr
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_wgrad2_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                      const Wgrad2Params p) {
  using L = W2Smem<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBar);
  uint64_t* empty_bar = full_bar + kStages2;
  uint64_t* tmem_full_bar = empty_bar + kStages2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x % 3;
  const int tile = blockIdx.x / 3;
  const int m0 = (tile / p.tiles_n) * kBlockM;
  const int n0 = (tile % p.tiles_n) * BLOCK_N;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > p.total_kb) kb_end = p.total_kb;
  const int num_kb = kb_end - kb_begin;
  const bool two_chunks = m0 + 64 < p.Cout;
  const uint32_t chunk_tx = (uint32_t)p.KB * 128u;

  // the X regions of the ring are zeroed ONCE: the TMA only ever writes rows [0, KB) of a tile, the rows around it
  // (read by the shifted descriptors at the first / last K row) stay zero for the whole kernel
  for (int s = 0; s < kStages2; ++s) {
    uint4* zb = reinterpret_cast<uint4*>(smem + s * L::kStage + 2 * kChunkA2);
    for (int i = threadIdx.x; i < L::kNCh * kChunkB2 / 16; i += kThreads) zb[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  ptx::fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmDy);
    ptx::prefetch_tmap(&tmX);
    for (int s = 0; s < kStages2; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (num_kb > 0) {
    if (warp == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int st = i % kStages2;
        const uint32_t ph = (i / kStages2) & 1;
        ptx::mbar_wait(&empty_bar[st], ph ^ 1);
        uint8_t* sa = smem + st * L::kStage;
        uint8_t* sb = sa + 2 * kChunkA2;
        const int kb = kb_begin + i;
        const int h0 = (kb % p.HB) * p.BH;
        const int img0 = (kb / p.HB) * p.NB;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full_bar[st], chunk_tx * (uint32_t)((two_chunks ? 2 : 1) + L::kNCh));
          ptx::tma_load_4d(sa, &tmDy, &full_bar[st], m0, -1, h0, img0);
          if (two_chunks) ptx::tma_load_4d(sa + kChunkA2, &tmDy, &full_bar[st], m0 + 64, -1, h0, img0);
#pragma unroll
          for (int c = 0; c < L::kNCh; ++c)
            ptx::tma_load_4d(sb + c * kChunkB2 + kPad2, &tmX, &full_bar[st], n0 + c * 64, -1, h0 + r - 1, img0);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = ptx::make_idesc(1, 1, kBlockM, BLOCK_N, 1, 1);
      const int ksteps = p.KB / kUmmaK;
      for (int i = 0; i < num_kb; ++i) {
        const int st = i % kStages2;
        const uint32_t ph = (i / kStages2) & 1;
        ptx::mbar_wait(&full_bar[st], ph);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + st * L::kStage);
        const uint32_t sb = sa + 2 * kChunkA2 + kPad2;
        if (ptx::elect_one()) {
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = ptx::make_smem_desc(sa + k * 2048, kChunkA2, 1024);
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const uint32_t start = sb + k * 2048 + (s - 1) * 128;
              const uint64_t db = ptx::make_smem_desc_bo(start, kChunkB2, 1024,
                                                         p.base_offset_mode ? ((start >> 7) & 7u) : 0u);
              ptx::umma_f16(tmem_base + s * BLOCK_N, da, db, idesc, (i | k) != 0 ? 1u : 0u);
            }
          }
          ptx::umma_commit(&empty_bar[st]);
          if (i == num_kb - 1) ptx::umma_commit(tmem_full_bar);
        }
        __syncwarp();
      }
    } else {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const int et = threadIdx.x - 64;
      constexpr int kVec = BLOCK_N / 4;
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after();
      uint8_t* sf = smem;
      int rows_valid = p.Cout - m0;
      if (rows_valid > kBlockM) rows_valid = kBlockM;
      const int64_t ldw = (int64_t)9 * p.Cin;
      const bool direct = gridDim.z == 1;
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * BLOCK_N);
#pragma unroll 1
        for (int c32 = 0; c32 < BLOCK_N / 32; ++c32) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(taddr + c32 * 32, v);
          ptx::tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(sf + row * L::kPitch + c32 * 128);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            dst[c] = make_float4(__uint_as_float(v[c * 4]), __uint_as_float(v[c * 4 + 1]),
                                 __uint_as_float(v[c * 4 + 2]), __uint_as_float(v[c * 4 + 3]));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int64_t col0 = (int64_t)(r * 3 + s) * p.Cin + n0;
        for (int f = et; f < rows_valid * kVec; f += kEpiThreads) {
          const int rr = f / kVec, c4 = (f % kVec) * 4;
          float4 t = *reinterpret_cast<const float4*>(sf + rr * L::kPitch + c4 * 4);
          const int64_t off = (int64_t)(m0 + rr) * ldw + col0 + c4;
          if (direct) {
            __nv_bfloat16* o = p.out + off;
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
          } else if (p.part != nullptr) {
            *reinterpret_cast<float4*>(p.part + ((int64_t)blockIdx.z * gridDim.x + blockIdx.x) * (kBlockM * 3 * BLOCK_N) +
                                       rr * (3 * BLOCK_N) + s * BLOCK_N + c4) = t;
          } else {
            red_add_v4(p.ws + off, t.x, t.y, t.z, t.w);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (!direct && p.part == nullptr) {
        uint32_t* s_last = tmem_slot + 1;
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          const int old = atomicAdd(&p.counters[blockIdx.x], 1);
          const int last = old == (int)gridDim.z - 1;
          if (last) p.counters[blockIdx.x] = 0;
          *s_last = (uint32_t)last;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (*s_last != 0u) {
          __threadfence();
          const int total = rows_valid * (3 * kVec);
          for (int f = et; f < total; f += kEpiThreads) {
            const int rr = f / (3 * kVec);
            const int rem = f % (3 * kVec);
            const int s = rem / kVec, c4 = (rem % kVec) * 4;
            const int64_t off = (int64_t)(m0 + rr) * ldw + (int64_t)(r * 3 + s) * p.Cin + n0 + c4;
            float4 t = __ldcg(reinterpret_cast<const float4*>(p.ws + off));
            *reinterpret_cast<float4*>(p.ws + off) = make_float4(0.f, 0.f, 0.f, 0.f);
            __nv_bfloat16* o = p.out + off;
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
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

struct Wgrad2Geometry {
  int WB = 0, BH = 0, NB = 0, KB = 0;
  bool ok = false;
};

// Box {Wb >= W + 1 columns starting at w = -1, BH rows, NB images}: Wb * BH * NB a multiple of 16 and <= kMaxKB2;
// fewest wasted (zero) K rows first, then the largest block.
Wgrad2Geometry plan_wgrad2(int N, int H, int W) {
  Wgrad2Geometry best;
  double best_eff = 0.0;
  if (W < 1 || H < 1 || N < 1) return best;
  for (int wb = W + 1; wb <= kMaxKB2 && wb <= W + 16; ++wb) {
    for (int nb = 1; nb <= 16 && nb <= N; nb *= 2) {
      for (int bh = 1; bh <= H; ++bh) {
        const int kb = wb * bh * nb;
        if (kb > kMaxKB2) break;
        if (kb % kUmmaK != 0) continue;
        const int hb = (H + bh - 1) / bh, ng = (N + nb - 1) / nb;
        const double eff = (double)H * N * W / ((double)hb * bh * ng * nb * wb);
        if (eff > best_eff + 1e-9 || (eff > best_eff - 1e-9 && kb > best.KB)) {
          best_eff = eff;
          best.WB = wb; best.BH = bh; best.NB = nb; best.KB = kb; best.ok = true;
        }
      }
    }
  }
  return best;
}

int g_wgrad3_version = [] {
  const char* e = getenv("EDL_WGRAD3_V");
  return (e != nullptr && e[0] == '1') ? 1 : 2;
}();
int g_wgrad3_bo = [] {
  const char* e = getenv("EDL_WGRAD3_BO");
  return (e != nullptr && e[0] == '1') ? 1 : 0;
}();

template <int BLOCK_N>
const char* launch_wgrad2(const Conv3x3WgradArgs& a, const Wgrad2Geometry& geo, cudaStream_t stream) {
  using L = W2Smem<BLOCK_N>;
  alignas(64) CUtensorMap tmDy, tmX;
  {
    const uint64_t dims[4] = {(uint64_t)a.Cout, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
    const uint64_t st[3] = {(uint64_t)a.Cout * 2, (uint64_t)a.W * a.Cout * 2, (uint64_t)a.H * a.W * a.Cout * 2};
    const uint32_t box[4] = {64, (uint32_t)geo.WB, (uint32_t)geo.BH, (uint32_t)geo.NB};
    if (const char* e = encode_tmap_bf16(&tmDy, a.dY, 4, dims, st, box)) return e;
  }
  {
    const uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
    const uint64_t st[3] = {(uint64_t)a.Cin * 2, (uint64_t)a.W * a.Cin * 2, (uint64_t)a.H * a.W * a.Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)geo.WB, (uint32_t)geo.BH, (uint32_t)geo.NB};
    if (const char* e = encode_tmap_bf16(&tmX, a.X, 4, dims, st, box)) return e;
  }
  auto kern = conv3x3_wgrad2_kernel<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    apply_carveout((const void*)kern);
    attr_set = true;
  }
  Wgrad2Params p;
  p.Cout = a.Cout; p.Cin = a.Cin;
  p.BH = geo.BH; p.NB = geo.NB; p.KB = geo.KB;
  p.HB = (a.H + geo.BH - 1) / geo.BH;
  p.total_kb = p.HB * ((a.N + geo.NB - 1) / geo.NB);
  int split = a.split_k < 1 ? 1 : a.split_k;
  if (split > p.total_kb) split = p.total_kb;
  p.kb_per_split = (p.total_kb + split - 1) / split;
  split = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.tiles_n = a.Cin / BLOCK_N;
  p.base_offset_mode = g_wgrad3_bo;
  const int tiles_m = (a.Cout + kBlockM - 1) / kBlockM;
  p.ws = a.ws;
  p.out = reinterpret_cast<__nv_bfloat16*>(a.dW);
  p.counters = a.counters;
  p.accumulate = a.accumulate ? 1 : 0;
  dim3 grid(tiles_m * p.tiles_n * 3, 1, split);
  const bool use_part = split > 1 && a.partials != nullptr &&
                        a.partials_elems >= (int64_t)split * grid.x * kBlockM * 3 * BLOCK_N;
  p.part = use_part ? a.partials : nullptr;
  kern<<<grid, kThreads, L::kTotal, stream>>>(tmDy, tmX, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cudaGetErrorString(e);
  if (use_part) {
    splitk_reduce(a.partials, split, (int)grid.x, p.tiles_n, kBlockM, 3 * BLOCK_N, 3, a.Cin, a.Cout, 9 * a.Cin, a.dW,
                  (int64_t)9 * a.Cin, a.accumulate, stream);
    e = cudaGetLastError();
  }
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

struct WgradGeometry {
  int BH = 0, NB = 0, KB = 0;
  bool ok = false;
};

// Pick the pixel box {W, BH rows, NB images}: W*BH*NB a multiple of 16 (UMMA_K) and <= kMaxKB, as few
// zero-filled (out-of-image) rows as possible, then as many pixels per block as possible.
WgradGeometry plan_wgrad(int N, int H, int W) {
  WgradGeometry best;
  double best_eff = 0.0;
  if (W < 1 || W > kMaxKB || H < 1 || N < 1) return best;
  for (int nb = 1; nb <= 16 && nb <= N; nb *= 2) {
    for (int bh = 1; bh <= H; ++bh) {      // boxes never exceed the tensor extent (partial last blocks are zero-filled)
      const int kb = W * bh * nb;
      if (kb > kMaxKB) break;
      if (kb % kUmmaK != 0) continue;
      const int hb = (H + bh - 1) / bh, ng = (N + nb - 1) / nb;
      const double eff = (double)H * N / ((double)hb * bh * ng * nb);
      if (eff > best_eff + 1e-9 || (eff > best_eff - 1e-9 && kb > best.KB)) {
        best_eff = eff;
        best.BH = bh; best.NB = nb; best.KB = kb; best.ok = true;
      }
    }
  }
  return best;
}

}  // namespace

bool conv3x3_wgrad_supported(int N, int H, int W, int Cin, int Cout) {
  if (Cin % 64 != 0 || Cout % 64 != 0) return false;
  return g_wgrad3_version == 2 ? plan_wgrad2(N, H, W).ok : plan_wgrad(N, H, W).ok;
}

void set_wgrad3_version(int version, int base_offset_mode) {
  g_wgrad3_version = version == 1 ? 1 : 2;
  g_wgrad3_bo = base_offset_mode ? 1 : 0;
}
int get_wgrad3_version() { return g_wgrad3_version; }

bool conv3x3_wgrad_s2_supported(int N, int Ho, int Wo, int Cin, int Cout) {
  if (Cin % 64 != 0 || Cout % 64 != 0 || 2 * Wo > 256) return false;
  const WgradGeometry g = plan_wgrad(N, Ho, Wo);
  return g.ok && 2 * g.BH <= 256;
}

const char* conv3x3_wgrad_bf16(const Conv3x3WgradArgs& a, cudaStream_t stream) {
  const bool s2 = a.stride == 2;
  if (s2 ? !conv3x3_wgrad_s2_supported(a.N, a.H, a.W, a.Cin, a.Cout)
         : !conv3x3_wgrad_supported(a.N, a.H, a.W, a.Cin, a.Cout))
    return "conv3x3_wgrad: unsupported shape";
  if (a.ws == nullptr || a.counters == nullptr || a.dW == nullptr) return "conv3x3_wgrad: missing buffers";
  if (a.device >= 0) {
    cudaError_t e = cudaSetDevice(a.device);   // tensor-map encoding needs a bound context (see gemm.cu)
    if (e != cudaSuccess) return cudaGetErrorString(e);
  }
  if (g_wgrad3_version == 2 && !s2) {
    const Wgrad2Geometry g2 = plan_wgrad2(a.N, a.H, a.W);
    return a.Cin % 128 == 0 ? launch_wgrad2<128>(a, g2, stream) : launch_wgrad2<64>(a, g2, stream);
  }
  const WgradGeometry geo = plan_wgrad(a.N, a.H, a.W);
  alignas(64) CUtensorMap tmDy, tmX;
  {
    const uint64_t dims[4] = {(uint64_t)a.Cout, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
    const uint64_t st[3] = {(uint64_t)a.Cout * 2, (uint64_t)a.W * a.Cout * 2, (uint64_t)a.H * a.W * a.Cout * 2};
    const uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)geo.BH, (uint32_t)geo.NB};
    if (const char* e = encode_tmap_bf16(&tmDy, a.dY, 4, dims, st, box)) return e;
  }
  if (!s2) {
    const uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
    const uint64_t st[3] = {(uint64_t)a.Cin * 2, (uint64_t)a.W * a.Cin * 2, (uint64_t)a.H * a.W * a.Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)geo.BH, (uint32_t)geo.NB};
    if (const char* e = encode_tmap_bf16(&tmX, a.X, 4, dims, st, box)) return e;
  } else {
    // the input is [N, 2H, 2W, Cin]: a box spanning 2W x 2BH elements with traversal stride 2 delivers the W x BH
    // input pixels that tap (r, s) pairs with the dY box -- same K-row order, no im2col, no subsample pass
    const uint64_t H2 = 2ull * a.H, W2 = 2ull * a.W;
    const uint64_t dims[4] = {(uint64_t)a.Cin, W2, H2, (uint64_t)a.N};
    const uint64_t st[3] = {(uint64_t)a.Cin * 2, W2 * a.Cin * 2, H2 * W2 * a.Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)(2 * a.W), (uint32_t)(2 * geo.BH), (uint32_t)geo.NB};
    const uint32_t es[4] = {1, 2, 2, 1};
    if (const char* e = encode_tmap_strided(&tmX, a.X, 4, dims, st, box, es, 2)) return e;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    apply_carveout((const void*)conv3x3_wgrad_kernel);
    attr_set = true;
  }
  WgradParams p;
  p.Cout = a.Cout; p.Cin = a.Cin;
  p.BH = geo.BH; p.NB = geo.NB; p.KB = geo.KB;
  p.HB = (a.H + geo.BH - 1) / geo.BH;
  p.total_kb = p.HB * ((a.N + geo.NB - 1) / geo.NB);
  int split = a.split_k < 1 ? 1 : a.split_k;
  if (split > p.total_kb) split = p.total_kb;
  p.kb_per_split = (p.total_kb + split - 1) / split;
  split = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.tiles_n = a.Cin / kBlockN;
  const int tiles_m = (a.Cout + kBlockM - 1) / kBlockM;
  p.ws = a.ws;
  p.out = reinterpret_cast<__nv_bfloat16*>(a.dW);
  p.counters = a.counters;
  p.accumulate = a.accumulate ? 1 : 0;
  p.stride = s2 ? 2 : 1;
  dim3 grid(tiles_m * p.tiles_n * 3, 1, split);
  const bool use_part = split > 1 && a.partials != nullptr &&
                        a.partials_elems >= (int64_t)split * grid.x * kBlockM * 3 * kBlockN;
  p.part = use_part ? a.partials : nullptr;
  conv3x3_wgrad_kernel<<<grid, kThreads, kSmemTotal, stream>>>(tmDy, tmX, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cudaGetErrorString(e);
  if (use_part) {
    splitk_reduce(a.partials, split, (int)grid.x, p.tiles_n, kBlockM, 3 * kBlockN, 3, a.Cin, a.Cout, 9 * a.Cin, a.dW,
                  (int64_t)9 * a.Cin, a.accumulate, stream);
    e = cudaGetLastError();
  }
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// upper bound over both kernel versions (sizes the tile-counter array)
int conv3x3_wgrad_tiles(int Cin, int Cout) { return ((Cout + kBlockM - 1) / kBlockM) * (Cin / kBlockN) * 3; }
// (tile, filter row) CTAs per split of the active version: what the split-K heuristic divides the SMs by
int conv3x3_wgrad_ctas(int Cin, int Cout) {
  const int bn = (g_wgrad3_version == 2 && Cin % 128 == 0) ? 128 : 64;
  return ((Cout + kBlockM - 1) / kBlockM) * (Cin / bn) * 3;
}

void conv3x3_wgrad_plan(int N, int H, int W, int* bh, int* nb, int* kb) {
  if (g_wgrad3_version == 2) {
    const Wgrad2Geometry g2 = plan_wgrad2(N, H, W);
    *bh = g2.ok ? g2.BH : 0;
    *nb = g2.ok ? g2.NB : 0;
    *kb = g2.ok ? g2.KB : 0;
    return;
  }
  const WgradGeometry g = plan_wgrad(N, H, W);
  *bh = g.ok ? g.BH : 0;
  *nb = g.ok ? g.NB : 0;
  *kb = g.ok ? g.KB : 0;
}

// stride-2 layers always run on the version-1 kernel: pixel blocks / CTAs per split for the split-K heuristic
int conv3x3_wgrad_s2_kblocks(int N, int Ho, int Wo) {
  const WgradGeometry g = plan_wgrad(N, Ho, Wo);
  if (!g.ok) return 0;
  return ((Ho + g.BH - 1) / g.BH) * ((N + g.NB - 1) / g.NB);
}

int conv3x3_wgrad_kblocks(int N, int H, int W) {
  if (g_wgrad3_version == 2) {
    const Wgrad2Geometry g2 = plan_wgrad2(N, H, W);
    if (!g2.ok) return 0;
    return ((H + g2.BH - 1) / g2.BH) * ((N + g2.NB - 1) / g2.NB);
  }
  const WgradGeometry g = plan_wgrad(N, H, W);
  if (!g.ok) return 0;
  return ((H + g.BH - 1) / g.BH) * ((N + g.NB - 1) / g.NB);
}

}  // namespace edl
