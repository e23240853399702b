// CTA-pair tcgen05 GEMM for sm_100a: D[M, N] = A[M, K] * B[N, K]^T (+ addend) with the scale / shift / ReLU epilogue,
// 256 x 256 output tiles computed by TWO SMs of one TPC with `tcgen05.mma.cta_group::2`.
//
// Why: profiles/teacher_c22_halo.txt -- the teacher's 1x1 convolutions (26 GFLOP each, N = 256 .. 4096, K = 256 .. 4096)
// run at ~720 TFLOP/s on the 128 x 256 single-CTA tiles of gemm_persist.cu.  A 128 x 256 tile pulls 48 KB of operands
// per 512 tensor-core cycles = 96 B / clk / SM, and the L2 -> SM fabric sustains about 6300 B / clk for the whole chip
// (~42 B / clk / SM, B300_MICROARCH.md): the kernel is fed at less than half the rate the tensor cores consume.  In a
// CTA pair every SM loads its own 128 rows of A and only HALF of the B tile (128 of the 256 rows); the MMA unit of either
// SM reads both halves (the peer's through the pair's shared-memory path), so a 256 x 256 tile costs 32 KB per SM and
// k-block: 64 B / clk / SM for the same math rate.
//
// Structure (per CTA; cluster = 2 CTAs, rank 0 is the leader):
//   warp 0      TMA producer: own A tile + own half of B per stage, `cp.async.bulk.tensor...cta_group::2` with the
//               LEADER's full barrier as completion target (the leader expects the bytes of both CTAs)
//   warp 1      TMEM allocation (cta_group::2, both CTAs); in the leader: the single thread that issues the pair's MMAs
//               (UMMA 256 x 256 x 16) and multicasts the commits to BOTH CTAs' empty / accumulator-full barriers
//   warps 2..9  epilogue of the CTA's own 128 accumulator rows: tcgen05.ld -> (+ addend) -> shift -> ReLU -> bf16 ->
//               swizzled staging tile -> TMA store; both CTAs' epilogue warps release the accumulator on the leader's
//               barrier (remote mbarrier arrive)
//   Two accumulators (2 x 256 TMEM columns per SM) so that the MMAs of tile i + 1 overlap the epilogue of tile i.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>

#include "gemm.h"
#include "kernels.h"
#include "ptx.cuh"

namespace edl {
namespace {

constexpr int kBM = 128;              // rows per CTA (256 per pair)
constexpr int kBN = 256;              // columns per pair; every CTA stages kBN / 2 rows of B
constexpr int kBK = 64;
constexpr int kUK = 16;
constexpr int kStages = 5;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;

struct PairParams {
  int M, N, K;
  int tiles_m, tiles_n, num_kb;       // tiles of 256 x 256
  const float* col_scale;
  const float* col_shift;
  int relu;
  int has_add;
};

struct PairSmem {
  static constexpr int kABytes = kBM * kBK * 2;          // 16 KB
  static constexpr int kBBytes = (kBN / 2) * kBK * 2;    // 16 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kDBytes = kBM * kBN * 2;          // 64 KB staging tile (4 column blocks of 64)
  static constexpr int kBarOffset = kRingBytes + kDBytes;
  static constexpr int kTotal = kBarOffset + 512 + 1024;
};
static_assert(PairSmem::kTotal <= 227 * 1024, "shared memory budget");

EDL_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
EDL_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` in CTA `rank` of the cluster
EDL_DEVICE uint32_t map_to_rank(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
EDL_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// pair TMA load: data into THIS CTA's shared memory, completion bytes onto the barrier at `bar_cluster_addr` (leader)
EDL_DEVICE void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <uint32_t kCols>
EDL_DEVICE void tmem_alloc_pair(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_dst)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
EDL_DEVICE void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
EDL_DEVICE void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once the pair's previously issued MMAs have completed
EDL_DEVICE void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmAdd, const PairParams p) {
  using L = PairSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sd = smem + L::kRingBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);   // used in the leader only
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2], used in the leader only (both CTAs' epilogues arrive there)
  uint64_t* add_bar = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(add_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int total_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmD);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 2 * kEpiWarps);
    }
    ptx::mbar_init(add_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  ptx::tc_fence_before();
  cluster_sync_all();          // the peer's barriers exist before anything is signalled across the pair
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    // warp-uniform loops, one elected lane issues (operands stay in uniform registers)
    uint32_t it = 0;
    for (int t = pair; t < total_tiles; t += num_pairs) {
      const int tile_m = t / p.tiles_n;
      const int n0 = (t - tile_m * p.tiles_n) * kBN;
      const int m0 = tile_m * (2 * kBM) + (int)rank * kBM;
      for (int i = 0; i < p.num_kb; ++i, ++it) {
        const int s = it % kStages;
        ptx::mbar_wait(&empty_bar[s], ((it / kStages) & 1) ^ 1);
        uint8_t* sa = smem + s * L::kStageBytes;
        uint8_t* sb = sa + L::kABytes;
        const uint32_t bar = map_to_rank(ptx::smem_u32(&full_bar[s]), 0);
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * L::kStageBytes);
          tma_load_2d_pair(sa, &tmA, bar, i * kBK, m0);
          tma_load_2d_pair(sb, &tmB, bar, i * kBK, n0 + (int)rank * (kBN / 2));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc(1, 1, 2 * kBM, kBN, 0, 0);
      uint32_t it = 0, tc = 0;
      for (int t = pair; t < total_tiles; t += num_pairs, ++tc) {
        const uint32_t slot = tc & 1, aph = (tc >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[slot], aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t acc = tmem_base + slot * kBN;
        for (int i = 0; i < p.num_kb; ++i, ++it) {
          const int s = it % kStages;
          ptx::mbar_wait(&full_bar[s], (it / kStages) & 1);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + s * L::kStageBytes);
          const uint64_t da0 = ptx::make_smem_desc(sa, 16, 1024);
          const uint64_t db0 = ptx::make_smem_desc(sa + L::kABytes, 16, 1024);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / kUK; ++k)    // 32 bytes per k step = 2 units of the start-address field
              umma_f16_pair(acc, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc, (i | k) != 0 ? 1u : 0u);
            umma_commit_pair(&empty_bar[s]);
            if (i == p.num_kb - 1) umma_commit_pair(&tmem_full[slot]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    const int ew = warp - 2;
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int grp = ew >> 2;                 // column half owned by this warpgroup
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    constexpr int kColsPerGrp = kBN / 2;
    uint32_t tc = 0;
    for (int t = pair; t < total_tiles; t += num_pairs, ++tc) {
      const uint32_t slot = tc & 1, aph = (tc >> 1) & 1;
      const int tile_m = t / p.tiles_n;
      const int n0 = (t - tile_m * p.tiles_n) * kBN;
      const int m0 = tile_m * (2 * kBM) + (int)rank * kBM;
      const bool rows_exist = m0 < p.M;            // the second half of the last M tile may lie beyond the matrix
      if (et == 0) ptx::tma_store_wait_read0();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const bool has_add = p.has_add != 0 && rows_exist;
      if (has_add) {
        if (et == 0) {
          ptx::mbar_arrive_expect_tx(add_bar, 4 * (kBM * 128));
#pragma unroll
          for (int hh = 0; hh < kBN / 64; ++hh) ptx::tma_load_2d(sd + hh * (kBM * 128), &tmAdd, add_bar, n0 + hh * 64, m0);
        }
      }
      uint32_t add_phase = 0;
      if (p.has_add != 0) {
        // phase of add_bar = number of addend loads this CTA has issued before, which is tc unless a tile was skipped
        // (only the very last tile of rank 1 can be): tiles with rows_exist are a prefix of the CTA's sequence
        add_phase = tc & 1;
        if (has_add) ptx::mbar_wait(add_bar, add_phase);
      }
      ptx::mbar_wait(&tmem_full[slot], aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + slot * kBN + grp * kColsPerGrp + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c32 = 0; c32 < kColsPerGrp / 32; ++c32) {
        uint32_t rg[32];
        ptx::tmem_ld_32x32(taddr + c32 * 32, rg);
        ptx::tmem_ld_wait();
        if (c32 == kColsPerGrp / 32 - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(map_to_rank(ptx::smem_u32(&tmem_empty[slot]), 0));
        }
        const int cbase = grp * kColsPerGrp + c32 * 32;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(rg[j]);
        if (has_add) {
          const uint32_t arow = ptx::smem_u32(sd) + (cbase >> 6) * (kBM * 128) + row * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int chunk = ((cbase >> 5) & 1) * 4 + c;
            uint32_t w4[4];
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(w4[0]), "=r"(w4[1]), "=r"(w4[2]), "=r"(w4[3])
                         : "r"(arow + ((chunk ^ (row & 7)) << 4)));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[j]));
              f[c * 8 + 2 * j] += t2.x;
              f[c * 8 + 2 * j + 1] += t2.y;
            }
          }
        }
        if (p.col_scale != nullptr) {      // 16-byte aligned (checked on the host), N % 256 == 0
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v4 = __ldg(reinterpret_cast<const float4*>(p.col_scale + n0 + cbase) + j);
            f[4 * j] *= v4.x; f[4 * j + 1] *= v4.y; f[4 * j + 2] *= v4.z; f[4 * j + 3] *= v4.w;
          }
        }
        if (p.col_shift != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v4 = __ldg(reinterpret_cast<const float4*>(p.col_shift + n0 + cbase) + j);
            f[4 * j] += v4.x; f[4 * j + 1] += v4.y; f[4 * j + 2] += v4.z; f[4 * j + 3] += v4.w;
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        const uint32_t rowp = ptx::smem_u32(sd) + (cbase >> 6) * (kBM * 128) + row * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int chunk = ((cbase >> 5) & 1) * 4 + c;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = f[c * 8 + j];
          const bf16x8 pk = pack8(v);
          const uint4 u = *reinterpret_cast<const uint4*>(&pk);
          ptx::sts128(rowp + ((chunk ^ (row & 7)) << 4), u.x, u.y, u.z, u.w);
        }
      }
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et == 0 && rows_exist) {
#pragma unroll
        for (int hh = 0; hh < kBN / 64; ++hh) ptx::tma_store_2d(&tmD, sd + hh * (kBM * 128), n0 + hh * 64, m0);
        ptx::tma_store_commit();
      }
    }
    if (et == 0) ptx::tma_store_wait_read0();
  }
  // neither CTA may leave while the other can still signal its barriers or read its shared memory
  ptx::tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

bool g_pair_gemm = [] {
  const char* e = getenv("EDL_GEMM_PAIR");
  return !(e != nullptr && e[0] == '0');
}();

const char* tmap2(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems, uint32_t box_inner,
                  uint32_t box_outer) {
  const uint64_t dims[2] = {inner, outer};
  const uint64_t st[1] = {pitch_elems * 2};
  const uint32_t box[2] = {box_inner, box_outer};
  return encode_tmap_bf16(out, ptr, 2, dims, st, box);
}

}  // namespace

void set_pair_gemm(bool on) { g_pair_gemm = on; }
bool get_pair_gemm() { return g_pair_gemm; }

// Shapes the pair kernel takes: plain bf16 GEMM (B K-major), N a multiple of 256, enough tiles to fill the 74 pairs
bool gemm_pair_supported(const GemmArgs& g) {
  if (!g_pair_gemm || g.a_mn_major || g.b_mn_major || g.bn.x != nullptr || g.col_stats != nullptr || g.out_f32 != nullptr ||
      g.split_k > 1 || g.ship_flag != nullptr)
    return false;
  if (g.N % kBN != 0 || g.K % 8 != 0 || g.K < 256) return false;
  if ((reinterpret_cast<uintptr_t>(g.col_scale) & 15) != 0 || (reinterpret_cast<uintptr_t>(g.col_shift) & 15) != 0) return false;
  const int64_t tiles = (int64_t)((g.M + 2 * kBM - 1) / (2 * kBM)) * (g.N / kBN);
  return tiles >= kNumSMs / 2;
}

const char* gemm_bf16_pair(const GemmArgs& g, cudaStream_t stream) {
  alignas(64) CUtensorMap tmA, tmB, tmD, tmAdd;
  if (const char* e = tmap2(&tmA, g.A, g.K, g.M, g.lda, kBK, kBM)) return e;
  if (const char* e = tmap2(&tmB, g.B, g.K, g.N, g.ldb, kBK, kBN / 2)) return e;
  if (const char* e = tmap2(&tmD, g.D, g.N, g.M, g.ldd, 64, kBM)) return e;
  PairParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.tiles_m = (g.M + 2 * kBM - 1) / (2 * kBM);
  p.tiles_n = g.N / kBN;
  p.num_kb = (g.K + kBK - 1) / kBK;
  p.col_scale = g.col_scale; p.col_shift = g.col_shift; p.relu = g.relu ? 1 : 0;
  p.has_add = g.add_src != nullptr ? 1 : 0;
  if (g.add_src != nullptr) {
    if (g.ld_add % 8 != 0 || (reinterpret_cast<uintptr_t>(g.add_src) & 15) != 0) return "gemm add_src needs 16-byte aligned rows";
    if (const char* e = tmap2(&tmAdd, g.add_src, g.N, g.M, g.ld_add, 64, kBM)) return e;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PairSmem::kTotal);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    apply_carveout((const void*)gemm_pair_kernel);
    attr_set = true;
  }
  const int total = p.tiles_m * p.tiles_n;
  const int pairs = total < kNumSMs / 2 ? total : kNumSMs / 2;
  gemm_pair_kernel<<<dim3(2 * pairs), dim3(kThreads), PairSmem::kTotal, stream>>>(tmA, tmB, tmD, g.add_src != nullptr ? tmAdd : tmD,
                                                                                 p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace edl
