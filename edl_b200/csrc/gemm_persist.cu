// Persistent tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
// gemm.cu / conv3x3.cu launch one CTA per output tile.  The layers of ResNet50_vd at batch 32 have SHORT
// K loops (1..36 k-blocks), so a CTA's fixed work (barrier init, TMEM allocation, tensor-map fetch,
// pipeline fill, epilogue, teardown) dominated: profiles/prof_conv3_56 shows 10 us of CTA lifetime for
// nine k-steps and the tensor pipe 16 % busy.  This kernel keeps ONE CTA per SM alive and streams tiles
// through it:
//
//   warp 0      TMA producer: one smem ring shared by all tiles of the CTA, never drained between tiles
//   warp 1      TMEM allocator + tcgen05.mma issuer; TWO accumulators in TMEM (2 x BLOCK_N columns), so
//               the MMAs of tile i+1 run while tile i is still being read out
//   warps 2..9  epilogue (two warpgroups, each owns half of the tile's columns): tcgen05.ld -> scale /
//               shift / ReLU -> bf16 -> swizzled staging tile -> TMA store, BatchNorm statistics of the
//               stored values -> vector reductions; releases the accumulator as soon as it is in registers
//
// Modes: 0 = D = A * B^T (B K-major: 1x1 conv fprop, FC), 1 = D = A * B (B MN-major: 1x1 conv dgrad),
//        2 = 3x3/s1/p1 conv fprop, 3 = 3x3/s1/p1 conv dgrad (see conv3x3.cu for the shifted-box trick).
#include <cuda.h>
#include <cstdio>
#include <cstdlib>

#include "gemm.h"
#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace edl {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
// Epilogue warps per CTA: 8 (two per TMEM lane quarter, each owns half of the tile's columns) or 16 (four per quarter, a
// quarter of the columns each).  The short-K layers of the student are bound by the epilogue's instruction LATENCY, not by
// its issue slots (profiles/prof_c25_fwd_s1.ncu.txt: 35 % of the issue slots used, 0.47 eligible warps per scheduler and
// cycle with 2.5 resident warps per scheduler), so the 128- and 256-column kernels run 16.

// BatchNorm statistics are accumulated per CTA in shared memory over ALL the tiles it processes and
// pushed to global memory once at the end: with one reduction request per tile and column group the
// L2 retires ~12 requests/ns into the few hot lines, which made the 1568-tile 1x1 convs atomic-bound
// (400k requests = 33 us; profiles/prof_persist_fwd2).  N <= kMaxStatsN, else direct reductions.
constexpr int kMaxStatsN = 2048;

struct PersistParams {
  // GEMM view
  int M, N, K;
  int tiles_m, tiles_n;
  int num_kb;                 // k-blocks per tile (conv: 9 * kc_blocks)
  const float* col_scale;
  const float* col_shift;
  int relu;
  float* col_stats;
  const __nv_bfloat16* add_src;   // optional addend tile source (GEMM modes)
  long long ld_add;
  // fused BatchNorm-backward reduction (BNR kernels): D is the gradient dy of a BN layer's output
  const float* bn_mean;
  const float* bn_rstd;
  const float* bn_gamma;
  const float* bn_beta;
  float* bn_dsums;                // [2N]: += sum(dy_masked), sum(dy_masked * xhat)
  int bn_relu, bn_has_y;
  // conv view (modes 2, 3)
  int n_img, H, W, kc_blocks, c_in_w, BH, BN, tiles_h;
  int Wb;                         // halo layout: columns per image row of the A tile (W + 1)
  int b_res;                      // halo layout, one k-block per tap: the nine B tap tiles stay resident (see kRes)
  int tc_stats;                   // BatchNorm statistics of the stored tile on the tensor cores (see kTcStats)
  int cin_g, cout_g;              // grouped fprop: channels per group (cout_g == N for a dense conv)
  int stride;                     // conv fprop: 1, or 2 (input rows 2*h + r - 1; the columns come from the tensor map)
  // explicit tap list (conv modes; 0 = the nine standard taps): k-block i belongs to tap i / kc_blocks, which reads the
  // input box shifted by (tap_dh, tap_dw) and the weight tap tap_w.  Used by the stride-2 dgrad, which is four small
  // stride-1 convolutions over the output gradient -- one per parity class of the input pixel -- with 1, 2, 2 and 4 taps.
  int ntaps;
  signed char tap_dh[9], tap_dw[9], tap_w[9];
  // debug: per-phase clock64() stamps of CTA 0 (tools/trace_persist.py), [3 roles][kTraceTiles][8 phases]
  long long* trace;
};
constexpr int kTraceTiles = 16;
#define EDL_TRACE(role, tile, phase)                                                                    \
  do {                                                                                                  \
    if (p.trace != nullptr && blockIdx.x == 0 && (tile) < kTraceTiles)                                  \
      p.trace[((role) * kTraceTiles + (tile)) * 8 + (phase)] = clock64();                               \
  } while (0)

// ASTAGES > 0 selects the "haloed A tile" layout of the 3x3 convolution modes (see the kernel): the A tiles (one per filter
// ROW and k-block, 128 rows + a 1 KB zero pad on either side) and the B tiles (one per filter TAP) then live in two
// separate rings of ASTAGES and STAGES slots.
template <int BLOCK_N, int STAGES, int BNR = 0, int ASTAGES = 0>
struct PSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kPad = 1024;
  static constexpr int kARegion = kABytes + 2 * kPad;              // halo layout: pad | tile | pad
  static constexpr int kARingBytes = ASTAGES * kARegion;
  static constexpr int kDBytes = kBlockM * BLOCK_N * 2;
  static constexpr int kRingBytes = ASTAGES > 0 ? kARingBytes + STAGES * kBBytes : STAGES * kStageBytes;
  static constexpr int kStatsBytes = 2 * kMaxStatsN * 4;   // CTA-local per-channel (sum, sum^2) accumulators
  static constexpr int kXOffset = kRingBytes + kDBytes;           // BNR: BN input tile x, then BN output tile y
  static constexpr int kXYBytes = BNR ? 2 * kDBytes : 0;
  static constexpr int kConstOffset = kXOffset + kXYBytes;        // BNR: mean, rstd, scale, shift of the tile's columns
  static constexpr int kConstBytes = BNR ? 4 * BLOCK_N * 4 : 0;    // float4 {mean, rstd, scale, shift} per column
  static constexpr int kStatsOffset = kConstOffset + kConstBytes;
  static constexpr int kOnesOffset = kStatsOffset + kStatsBytes;   // bf16 ones [16 x 128], the B operand of the column sums
  static constexpr int kOnesBytes = (BLOCK_N == 128 && BNR == 0) ? 4096 : 0;
  static constexpr int kBarOffset = kOnesOffset + kOnesBytes;
  static constexpr int kTotal = kBarOffset + 512 + 1024;
};

EDL_DEVICE void tma_store_4d_p(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// 32 x 32 transpose-reduce: every lane passes its 32 per-column values, lane j gets column j summed over
// the 32 lanes (= 32 accumulator rows) with 31 shuffles instead of 160.
EDL_DEVICE float warp_colsum32(float (&v)[32], int lane) {
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

// BNR: TMA loads of the BN input (and output) tile of output tile (tm, nn0).  Called by a WHOLE warp (warp-uniform
// arguments); one elected lane issues.
template <int BLOCK_N, bool CONV>
EDL_DEVICE void issue_bn_tiles(const PersistParams& p, int tm, int nn0, int rows_tile, uint8_t* sx, uint8_t* sy,
                               const CUtensorMap* tmX, const CUtensorMap* tmY, uint64_t* bar) {
  const int mm0 = tm * kBlockM;
  const int im0 = CONV ? (tm / p.tiles_h) * p.BN : 0;
  const int hh0 = CONV ? (tm % p.tiles_h) * p.BH : 0;
  uint32_t halves = 0;
#pragma unroll
  for (int hh = 0; hh < (BLOCK_N + 63) / 64; ++hh) halves += (nn0 + hh * 64 < p.N) ? 1u : 0u;
  const uint32_t tile_bytes = CONV ? (uint32_t)rows_tile * 128u : (uint32_t)(kBlockM * 128);
  if (ptx::elect_one()) {
    ptx::mbar_arrive_expect_tx(bar, halves * tile_bytes * (p.bn_has_y ? 2u : 1u));
#pragma unroll
    for (int hh = 0; hh < (BLOCK_N + 63) / 64; ++hh) {
      if (nn0 + hh * 64 >= p.N) continue;
      if (!CONV) {
        ptx::tma_load_2d(sx + hh * (kBlockM * 128), tmX, bar, nn0 + hh * 64, mm0);
        if (p.bn_has_y) ptx::tma_load_2d(sy + hh * (kBlockM * 128), tmY, bar, nn0 + hh * 64, mm0);
      } else {
        ptx::tma_load_4d(sx + hh * (kBlockM * 128), tmX, bar, nn0 + hh * 64, 0, hh0, im0);
        if (p.bn_has_y) ptx::tma_load_4d(sy + hh * (kBlockM * 128), tmY, bar, nn0 + hh * 64, 0, hh0, im0);
      }
    }
  }
  __syncwarp();
}

template <int BLOCK_N, int STAGES, int MODE, int BNR, int ASTAGES, int EPIW>
__global__ void __launch_bounds__(64 + EPIW * 32, 1)
gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmAdd,
                    const __grid_constant__ CUtensorMap tmBnX, const __grid_constant__ CUtensorMap tmBnY,
                    const PersistParams p) {
  using L = PSmem<BLOCK_N, STAGES, BNR, ASTAGES>;
  constexpr int kEpiWarps = EPIW;
  constexpr int kThreads = 64 + kEpiWarps * 32;
  constexpr int kEpiThreads = kEpiWarps * 32;
  constexpr int kColGroups = kEpiWarps / 4;              // warps per TMEM lane quarter = column groups of the tile
  static_assert(EPIW == 8 || EPIW == 16, "epilogue warps");
  static_assert(BLOCK_N / kColGroups >= 32 && (BLOCK_N / kColGroups) % 32 == 0, "a column group is whole 32-column chunks");
  static_assert(BNR != 1 || EPIW == 8, "the legacy in-register BN reduction assumes 8 epilogue warps");
  constexpr bool kConv = MODE >= 2;
  constexpr bool kBMN = MODE == 1 || MODE == 3;
  constexpr bool kDgrad = MODE == 3;
  // Haloed A tiles (3x3 / stride 1 convolutions): the pixel box of a tile is loaded with one extra column on the LEFT
  // (Wb = W + 1 columns starting at column -1, zero-filled by TMA), so tile row m = image_row * Wb + (w + 1) and the left /
  // right neighbour of a pixel is simply tile row m -/+ 1: the right neighbour of the last pixel of an image row is the
  // halo (zero) column of the next one.  The three taps of a filter row then read the SAME shared-memory tile through
  // UMMA descriptors whose start address is shifted by -128 / 0 / +128 bytes (the 128-byte swizzle is a function of the
  // shared-memory address, so whole-row shifts stay consistent with what TMA wrote; zero pads absorb the rows before
  // the first / after the last).  One A tile per filter row instead of one per tap: a third of the A traffic from L2.
  // The accumulator rows of the halo column are garbage and are dropped when the epilogue compacts the tile into the
  // dense [BN][BH][W] staging layout, so statistics, the fused BN-backward reduction and the TMA store are unchanged.
  constexpr bool kHalo = ASTAGES > 0;
  static_assert(!kHalo || (kConv && BNR != 1), "halo tiles: conv modes, BNR 0 or 2");
  // Resident B (64-channel layers / groups: one k-block per tap, the nine tap tiles of an N tile are 72 KB = the whole B
  // ring): the tiles are ordered N-major and every CTA takes a CONTIGUOUS run of them, so the weights of an N tile (a
  // group) are loaded once per run instead of once per tile -- slot j of the ring holds tap j, its full / empty barriers
  // flip once per N-tile change.  profiles/teacher_c22_*.txt: the grouped 14 x 14 layers of the teacher moved 229 MB
  // from L2 per launch, 147 MB of it the same 2.4 MB of weights fetched by each of the 64 pixel tiles.
  constexpr bool kRes = kHalo && BLOCK_N == 64 && STAGES == 9;
  // BatchNorm statistics on the tensor cores (128-column tiles, no BNR): the bf16 tile Y that the epilogue packs into the
  // staging buffer ([128 rows][64 columns x 128 B] halves, 128-byte swizzle) IS a valid MN-major UMMA operand with the
  // tile ROWS as the contraction dimension.  Two small MMA batches per tile, issued by an epilogue thread right after
  // the TMA store:
  //     G[n, n'] = sum_r Y[r, n] Y[r, n']   (A = B = Y, both MN-major; 128 x 128 x 128)   -> diag G = sum of squares
  //     S[n, j]  = sum_r Y[r, n] * 1        (A = Y, B = a 16 x 128 tile of ones, K-major) -> column sums
  // accumulated in spare TMEM columns ACROSS the tiles of a CTA that share the N tile and read back once per run -- exact
  // fp32 sums of the stored bf16 values, in place of the column-pair loop over the staging tile that took 2070 of the
  // 3850 cycles a short-K tile spends in the epilogue (profiles/trace_persist_c28.txt).
  constexpr bool kTcStats = BLOCK_N == 128 && BNR == 0 && MODE != 1 && MODE != 3;
  constexpr uint32_t kGramCol = 2 * BLOCK_N, kSumCol = 3 * BLOCK_N;     // TMEM columns behind the two accumulators
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sd = smem + L::kRingBytes;                    // dedicated store-staging tile
  float* sstats = reinterpret_cast<float*>(smem + L::kStatsOffset);
  uint8_t* sx = smem + L::kXOffset;                       // BNR only
  uint8_t* sy = sx + L::kDBytes;
  float* sconst = reinterpret_cast<float*>(smem + L::kConstOffset);
  float* const stats_dst = BNR ? p.bn_dsums : p.col_stats;
  const bool local_stats = stats_dst != nullptr && p.N <= kMaxStatsN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;              // [2]
  uint64_t* tmem_empty = tmem_full + 2;                  // [2]
  uint64_t* add_bar = tmem_empty + 2;                    // addend tile landed in the staging buffer
  uint64_t* bn_bar = add_bar + 1;                        // BN x (/ y) tiles landed
  uint64_t* afull_bar = bn_bar + 1;                      // halo layout: the A ring's barriers
  uint64_t* aempty_bar = afull_bar + (kHalo ? ASTAGES : 0);
  uint64_t* stat_bar = aempty_bar + (kHalo ? ASTAGES : 0);   // statistics MMAs of a tile have read the staging buffer
  uint64_t* stage_rdy = stat_bar + 1;                        // the staging buffer holds a complete tile (epilogue -> MMA warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_rdy + 1);
  static_assert((2 * STAGES + 8 + 2 * ASTAGES) * 8 + 4 <= 512, "barrier block");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n;
  const bool res = kRes && p.b_res != 0;
  int t_begin = blockIdx.x, t_end = total_tiles, t_step = gridDim.x;
  if (res) {
    // balanced contiguous runs: every CTA gets floor or ceil(total / grid) tiles (all SMs keep loads in flight)
    t_begin = (int)((long long)blockIdx.x * total_tiles / (int)gridDim.x);
    t_end = (int)((long long)(blockIdx.x + 1) * total_tiles / (int)gridDim.x);
    t_step = 1;
  }
  // tile index -> (M tile, first column): M-major normally (neighbouring CTAs share the A tile in L2), N-major when
  // B is resident
  // Incremental form for the role loops (an integer division by a run-time value is ~100 cycles of dependent instructions,
  // and the epilogue needs the coordinates of this tile and the next: ~440 cycles per tile, profiles/trace_persist_c34.txt)
  struct TileIt { int tm, ni; };
  const int step_m = res ? 0 : t_step / p.tiles_n, step_n = res ? 0 : t_step - (t_step / p.tiles_n) * p.tiles_n;
  auto tile_init = [&](int t) -> TileIt {
    if (res) {
      const int nt = t / p.tiles_m;
      return TileIt{t - nt * p.tiles_m, nt};
    }
    const int tm = t / p.tiles_n;
    return TileIt{tm, t - tm * p.tiles_n};
  };
  auto tile_next = [&](TileIt c) -> TileIt {
    if (res) {
      if (++c.tm >= p.tiles_m) { c.tm = 0; ++c.ni; }
    } else {
      c.tm += step_m;
      c.ni += step_n;
      if (c.ni >= p.tiles_n) { c.ni -= p.tiles_n; ++c.tm; }
    }
    return c;
  };
  auto tile_coords = [&](int t, int& tm, int& n0) {
    if (res) {
      const int nt = t / p.tiles_m;
      tm = t - nt * p.tiles_m;
      n0 = nt * BLOCK_N;
    } else {
      tm = t / p.tiles_n;
      n0 = (t - tm * p.tiles_n) * BLOCK_N;
    }
  };
  constexpr uint32_t kTmemCols =
      kTcStats ? 512
               : (2 * BLOCK_N <= 32 ? 32 : (2 * BLOCK_N <= 64 ? 64 : (2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512))));
  // (conv: only the halo layout keeps rows behind the tile at zero; no epilogue constants, so rows outside the matrix are 0)
  const bool tcs = kTcStats && p.tc_stats != 0 && local_stats && (!kConv || kHalo) && p.col_scale == nullptr &&
                   p.col_shift == nullptr && p.relu == 0 && p.add_src == nullptr;
  static_assert(2 * BLOCK_N <= 512, "two accumulators must fit the 512 TMEM columns");

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmD);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], kEpiWarps);
    }
    ptx::mbar_init(add_bar, 1);
    ptx::mbar_init(bn_bar, 1);
    ptx::mbar_init(stat_bar, 1);
    ptx::mbar_init(stage_rdy, 1);
    for (int s = 0; s < ASTAGES; ++s) {
      ptx::mbar_init(&afull_bar[s], 1);
      ptx::mbar_init(&aempty_bar[s], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  if (local_stats)
    for (int i = threadIdx.x; i < 2 * p.N; i += kThreads) sstats[i] = 0.f;
  if (tcs) {
    // ones tile; and a zeroed staging buffer: rows the epilogue never writes (halo layout: rows behind the tile) are
    // part of the contraction
    for (int i = threadIdx.x; i < L::kOnesBytes / 16; i += kThreads)
      ptx::sts128(ptx::smem_u32(smem + L::kOnesOffset) + i * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    for (int i = threadIdx.x; i < L::kDBytes / 16; i += kThreads) ptx::sts128(ptx::smem_u32(sd) + i * 16, 0u, 0u, 0u, 0u);
    ptx::fence_proxy_async_smem();
  }
  if (kHalo) {
    // pads AND tiles: TMA only ever writes the first rows_in rows of a tile, the rows behind them must read as zero
    for (int i = threadIdx.x; i < L::kARingBytes / 16; i += kThreads) ptx::sts128(ptx::smem_u32(smem) + i * 16, 0u, 0u, 0u, 0u);
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // everything above is CTA-local set-up; under programmatic dependent launch it overlapped the tail of the
  // previous kernel in the stream.  From here on global memory is touched.
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;
  const int rows_tile = kConv ? p.BN * p.BH * p.W : kBlockM;            // rows of the (dense) output tile
  const int rows_in = kHalo ? p.BN * p.BH * p.Wb : rows_tile;          // rows of the A tile
  const uint32_t a_bytes = kConv ? (uint32_t)rows_in * 128u : (uint32_t)L::kABytes;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The WHOLE warp runs the loops and waits on the barriers; one elected lane issues the copies.  With the loops
    // inside `if (lane == 0)` every tensor-map / barrier / coordinate operand is a per-thread value that the compiler has to
    // move into uniform registers with a vote + broadcast + retry sequence per operand (profiles/README.md, "issue-bound").
    {
      uint32_t it = 0;
      [[maybe_unused]] uint32_t ita = 0;
      [[maybe_unused]] int cur_n0 = -1;
      [[maybe_unused]] uint32_t gcount = 0;
      TileIt cur = tile_init(t_begin);
      for (int t = t_begin; t < t_end; t += t_step, cur = tile_next(cur)) {
        const int tile_m = cur.tm, n0 = cur.ni * BLOCK_N;
        const int m0 = tile_m * kBlockM;
        const int img0 = kConv ? (tile_m / p.tiles_h) * p.BN : 0;
        const int h0 = kConv ? (tile_m % p.tiles_h) * p.BH : 0;
        if constexpr (kHalo) {
          // num_kb counts A steps here: (filter row r, k-block kc); each is followed by its three B tap tiles
          const int cbase_in = (MODE == 2) ? (n0 / p.cout_g) * p.cin_g : 0;
          const bool newg = n0 != cur_n0;
          if (newg) { cur_n0 = n0; ++gcount; }
          for (int i = 0; i < p.num_kb; ++i, ++ita) {
            const int r = i / p.kc_blocks, kc = i - r * p.kc_blocks;
            const int sa_i = ita % ASTAGES;
            ptx::mbar_wait(&aempty_bar[sa_i], ((ita / ASTAGES) & 1) ^ 1);
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(&afull_bar[sa_i], a_bytes);
              ptx::tma_load_4d(smem + sa_i * L::kARegion + L::kPad, &tmA, &afull_bar[sa_i], cbase_in + kc * kBlockK, -1,
                               h0 + (kDgrad ? 1 - r : r - 1), img0);
            }
            __syncwarp();
            if (res && !newg) continue;            // the taps of this N tile are already in their slots
            for (int sft = 0; sft < 3; ++sft, ++it) {
              const int s = res ? r * 3 + sft : (int)(it % STAGES);
              ptx::mbar_wait(&empty_bar[s], res ? ((gcount - 1) & 1) ^ 1 : ((it / STAGES) & 1) ^ 1);
              uint8_t* sb = smem + L::kARingBytes + s * L::kBBytes;
              const int tap = r * 3 + sft;
              if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&full_bar[s], L::kBBytes);
                if (!kBMN) {
                  ptx::tma_load_2d(sb, &tmB, &full_bar[s], tap * p.c_in_w + kc * kBlockK, n0);
                } else {
#pragma unroll
                  for (int hh = 0; hh < BLOCK_N / 64; ++hh)
                    ptx::tma_load_2d(sb + hh * 8192, &tmB, &full_bar[s], tap * p.c_in_w + n0 + hh * 64, kc * kBlockK);
                }
              }
              __syncwarp();
            }
          }
          continue;
        }
        for (int i = 0; i < p.num_kb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(&empty_bar[s], ph ^ 1);
          if (i == 0 && lane == 0) EDL_TRACE(0, (t - t_begin) / t_step, 0);
          uint8_t* sa = smem + s * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if (!kConv) {
            const int k0 = i * kBlockK;
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(&full_bar[s], a_bytes + L::kBBytes);
              ptx::tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
              if (!kBMN) {
                ptx::tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
              } else {
#pragma unroll
                for (int hh = 0; hh < BLOCK_N / 64; ++hh)
                  ptx::tma_load_2d(sb + hh * 8192, &tmB, &full_bar[s], n0 + hh * 64, k0);
              }
            }
            __syncwarp();
          } else {
            int tap = i / p.kc_blocks;
            const int kc = i - tap * p.kc_blocks;
            int dh, dw;
            if (p.ntaps > 0) {
              dh = p.tap_dh[tap];
              dw = p.tap_dw[tap];
              tap = p.tap_w[tap];
            } else {
              const int r = tap / 3, sft = tap - r * 3;
              dh = kDgrad ? 1 - r : r - 1;
              dw = kDgrad ? 1 - sft : sft - 1;
            }
            // grouped fprop: the N tile lies in one group; its input channels start at group * cin_g
            const int cbase_in = (MODE == 2) ? (n0 / p.cout_g) * p.cin_g : 0;
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(&full_bar[s], a_bytes + L::kBBytes);
              ptx::tma_load_4d(sa, &tmA, &full_bar[s], cbase_in + kc * kBlockK, dw, h0 * p.stride + dh, img0);
              if (!kBMN) {
                ptx::tma_load_2d(sb, &tmB, &full_bar[s], tap * p.c_in_w + kc * kBlockK, n0);
              } else {
#pragma unroll
                for (int hh = 0; hh < BLOCK_N / 64; ++hh)
                  ptx::tma_load_2d(sb + hh * 8192, &tmB, &full_bar[s], tap * p.c_in_w + n0 + hh * 64, kc * kBlockK);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Warp-uniform loops, one elected lane issues (always the same one: tcgen05.commit covers the MMAs of ITS thread).
    // The descriptors are built once per stage in uniform code; advancing k inside the 128-byte swizzle atom adds 32 bytes
    // (K-major) or 2048 bytes (MN-major) to the start-address field, i.e. 2 or 128 in the descriptor's 16-byte units.
    {
      constexpr uint32_t idesc = ptx::make_idesc(1, 1, kBlockM, BLOCK_N, 0, kBMN ? 1 : 0);
      uint32_t it = 0, tc = 0;
      [[maybe_unused]] uint32_t ita = 0;
      [[maybe_unused]] int cur_n0 = -1;
      [[maybe_unused]] uint32_t gcount = 0;
      // Tensor-core statistics of tile i (see kTcStats), issued from this warp AFTER the main MMAs of tile i + 1 so that
      // neither the epilogue warps nor the accumulator pipeline wait for them: the batch reads ~100 KB of shared memory
      // (~2000 cycles at the rate MN-major operands are fetched, profiles/trace_persist_c30.txt).
      [[maybe_unused]] auto issue_stats = [&](uint32_t i) {
        if constexpr (kTcStats) {
          int tm_i, n0_i, tm_p, n0_p = -1;
          tile_coords(t_begin + (int)i * t_step, tm_i, n0_i);
          if (i > 0) tile_coords(t_begin + (int)(i - 1) * t_step, tm_p, n0_p);
          const uint32_t acc_flag = n0_p == n0_i ? 1u : 0u;     // earlier tiles of the same N tile are in the TMEM sums
          ptx::mbar_wait(stage_rdy, i & 1);
          ptx::tc_fence_after();
          constexpr uint32_t idesc_g = ptx::make_idesc(1, 1, kBlockM, BLOCK_N, 1, 1);   // Y^T Y
          constexpr uint32_t idesc_s = ptx::make_idesc(1, 1, kBlockM, 16, 1, 0);        // Y^T ones
          // Y as MN-major operand: 64-column halves kBlockM * 128 bytes apart (LBO), 8-row groups 1024 bytes apart (SBO),
          // one UMMA_K step = 16 tile rows = 2048 bytes = 128 units of the start-address field
          const uint64_t dy0 = ptx::make_smem_desc(ptx::smem_u32(sd), kBlockM * 128, 1024);
          // ones tile, K-major: [16 rows][64 k] halves 2048 bytes apart, 32 bytes (2 units) per k step inside a half
          const uint64_t d10 = ptx::make_smem_desc(ptx::smem_u32(smem + L::kOnesOffset), 16, 1024);
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < kBlockM / kUmmaK; ++j) {
              const uint64_t dy = dy0 + (uint64_t)(j * 128);
              const uint64_t d1 = d10 + (uint64_t)((j >> 2) * 128 + (j & 3) * 2);
              ptx::umma_f16(tmem_base + kGramCol, dy, dy, idesc_g, j != 0 ? 1u : acc_flag);
              ptx::umma_f16(tmem_base + kSumCol, dy, d1, idesc_s, j != 0 ? 1u : acc_flag);
            }
            ptx::umma_commit(stat_bar);
          }
          __syncwarp();
        }
      };
      TileIt cur = tile_init(t_begin);
      for (int t = t_begin; t < t_end; t += t_step, ++tc, cur = tile_next(cur)) {
        const uint32_t slot = tc & 1, aph = (tc >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[slot], aph ^ 1);      // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t acc = tmem_base + slot * BLOCK_N;
        if constexpr (kHalo) {
          bool newg = false, lastg = false;
          if (res) {
            const int n0_ = cur.ni * BLOCK_N;
            const int n02_ = t + 1 < t_end ? tile_next(cur).ni * BLOCK_N : -1;
            newg = n0_ != cur_n0;
            lastg = n02_ != n0_;                     // the slots are handed back after the last tile of the N tile
            if (newg) { cur_n0 = n0_; ++gcount; }
          }
          for (int i = 0; i < p.num_kb; ++i, ++ita) {
            const int sa_i = ita % ASTAGES;
            ptx::mbar_wait(&afull_bar[sa_i], (ita / ASTAGES) & 1);
            const uint32_t sa = ptx::smem_u32(smem + sa_i * L::kARegion + L::kPad);
            for (int sft = 0; sft < 3; ++sft, ++it) {
              const int s = res ? i * 3 + sft : (int)(it % STAGES);
              if (!res) ptx::mbar_wait(&full_bar[s], (it / STAGES) & 1);
              else if (newg) ptx::mbar_wait(&full_bar[s], (gcount - 1) & 1);
              ptx::tc_fence_after();
              const uint32_t sb = ptx::smem_u32(smem + L::kARingBytes + s * L::kBBytes);
              const int shift = (kDgrad ? 1 - sft : sft - 1) * 128;      // tile rows are 128 bytes
              const uint64_t da0 = ptx::make_smem_desc(sa + shift, 16, 1024);
              const uint64_t db0 = kBMN ? ptx::make_smem_desc(sb, 8192, 1024) : ptx::make_smem_desc(sb, 16, 1024);
              if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  ptx::umma_f16(acc, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * (kBMN ? 128 : 2)), idesc,
                                (i | sft | k) != 0 ? 1u : 0u);
                if (!res || lastg) ptx::umma_commit(&empty_bar[s]);
                if (sft == 2) {
                  ptx::umma_commit(&aempty_bar[sa_i]);
                  if (i == p.num_kb - 1) ptx::umma_commit(&tmem_full[slot]);
                }
              }
              __syncwarp();
            }
          }
          if (kTcStats && tcs && tc > 0) issue_stats(tc - 1);
          continue;
        }
        if (lane == 0) EDL_TRACE(1, tc, 0);
        for (int i = 0; i < p.num_kb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          if (lane == 0 && i == 0) EDL_TRACE(1, tc, 1);
          if (lane == 0 && i == p.num_kb - 1) EDL_TRACE(1, tc, 2);
          const uint32_t sa = ptx::smem_u32(smem + s * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
          const uint64_t da0 = ptx::make_smem_desc(sa, 16, 1024);
          const uint64_t db0 = kBMN ? ptx::make_smem_desc(sb, 8192, 1024) : ptx::make_smem_desc(sb, 16, 1024);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              ptx::umma_f16(acc, da0 + (uint64_t)(k * 2), db0 + (uint64_t)(k * (kBMN ? 128 : 2)), idesc, (i | k) != 0 ? 1u : 0u);
            ptx::umma_commit(&empty_bar[s]);
            if (i == p.num_kb - 1) ptx::umma_commit(&tmem_full[slot]);
          }
          __syncwarp();
        }
        if (kTcStats && tcs && tc > 0) issue_stats(tc - 1);
      }
      if (kTcStats && tcs && tc > 0) issue_stats(tc - 1);      // the last tile's
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int ew = warp - 2;                 // 0..7
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int grp = ew >> 2;                 // column group owned by this warpgroup
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;         // 0..255
    constexpr int kColsPerGrp = BLOCK_N / kColGroups;
    int const_n0 = -1;
    uint32_t tc = 0;
    // Column sums of the statistics / BN-backward loops stay in registers across the tiles of a run that share the N tile
    // (a thread always owns the same column pair) and go to the CTA-local accumulators once per run: fp32 atomicAdd on
    // shared memory is a compare-and-swap loop (ATOMS.CAST.SPIN), and with kSplit threads per column hitting the same
    // word four times per tile it cost several hundred cycles of every tile (profiles/trace_persist_c32.txt).
    [[maybe_unused]] float run_s0 = 0.f, run_s1 = 0.f, run_q0 = 0.f, run_q1 = 0.f;
    // Software-pipelined accumulator read-out (one or two 32-column chunks per warp): TMEM is read at 64 B / clk / SM, i.e.
    // >= 1024 cycles for a 128 x 128 fp32 tile (profiles/trace_persist_c33.txt: 1200 of a short-K tile's 3470 cycles).
    // tcgen05.ld is asynchronous until tcgen05.wait::ld, so the loads of tile t + 1 are issued BEFORE the statistics /
    // BN-backward loop of tile t (if that tile's MMAs are already complete, which they are in the short-K layers) and
    // collected at the top of the next iteration.
    constexpr int kNCh = kColsPerGrp / 32;
    constexpr bool kPipeLd = kNCh <= 2 && BNR != 1;
    uint32_t rgn[kPipeLd ? kNCh * 32 : 32];            // kNCh chunks of 32 columns, read with one instruction
    bool ld_inflight = false;
    [[maybe_unused]] uint32_t stat_batches = 0;      // statistic MMA batches committed so far (phase of stat_bar)
    TileIt cur = tile_init(t_begin);
    // halo layout: accumulator row -> row of the dense staging tile (fixed per thread; the halo column's rows are dropped)
    int srow = row;
    bool store_ok = true;
    [[maybe_unused]] int halo_bi = 0, halo_hr = 0;
    if (kHalo) {
      const int ir = row / p.Wb, cw = row - ir * p.Wb;
      store_ok = cw != 0 && row < rows_in;
      srow = ir * p.W + cw - 1;
      halo_bi = ir / p.BH;
      halo_hr = ir - halo_bi * p.BH;
    }
    for (int t = t_begin; t < t_end; t += t_step, ++tc) {
      const uint32_t slot = tc & 1, aph = (tc >> 1) & 1;
      const int tile_m = cur.tm, n0 = cur.ni * BLOCK_N;
      const bool has_next = t + t_step < t_end;
      cur = tile_next(cur);
      const int next_tm = cur.tm, next_n0 = cur.ni * BLOCK_N;
      const int m0 = tile_m * kBlockM;
      const int img0 = kConv ? (tile_m / p.tiles_h) * p.BN : 0;
      const int h0 = kConv ? (tile_m % p.tiles_h) * p.BH : 0;
      // the staging tile must have been read by the previous TMA store and by every stats thread
      if (et == 0) EDL_TRACE(2, tc, 0);
      // (bulk groups belong to the thread that issued the stores: always the elected lane of epilogue warp 0)
      if (ew == 0) {
        if (ptx::elect_one()) ptx::tma_store_wait_read0();
        __syncwarp();
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (et == 0) EDL_TRACE(2, tc, 1);
      // addend (if any): TMA-load its tile into the staging buffer (same swizzled layout as the output)
      // while the MMAs of this tile are still running; every thread later adds its own 16-byte pieces
      const bool has_add = !kConv && p.add_src != nullptr;
      if (has_add) {
        if (ew == 0) {
          uint32_t halves = 0;
#pragma unroll
          for (int hh = 0; hh < (BLOCK_N + 63) / 64; ++hh) halves += (n0 + hh * 64 < p.N) ? 1u : 0u;
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(add_bar, halves * (kBlockM * 128));
#pragma unroll
            for (int hh = 0; hh < (BLOCK_N + 63) / 64; ++hh)
              if (n0 + hh * 64 < p.N) ptx::tma_load_2d(sd + hh * (kBlockM * 128), &tmAdd, add_bar, n0 + hh * 64, m0);
          }
          __syncwarp();
        }
        ptx::mbar_wait(add_bar, tc & 1);
      }
      // rows of this tile that are real pixels of the tensor (conv: partial patches, tail images)
      bool row_ok = true;
      if (BNR) {
        if (!kConv) {
          row_ok = m0 + row < p.M;
        } else {
          const int rows_per_img = p.BH * p.W;
          const int bi = row / rows_per_img, rr = row - bi * rows_per_img;
          row_ok = row < rows_tile && img0 + bi < p.n_img && h0 + rr / p.W < p.H;
        }
        if (tc == 0 && ew == 0)                               // later tiles are prefetched one tile ahead
          issue_bn_tiles<BLOCK_N, kConv>(p, tile_m, n0, rows_tile, sx, sy, &tmBnX, &tmBnY, bn_bar);
        if (n0 != const_n0) {
          // per-column constants of the BN layer for this N tile (a CTA usually keeps its N tile)
          if (tc != 0) asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory");   // everybody is done with the old ones
          if (et < BLOCK_N) {
            const int col = n0 + et;
            const bool in = col < p.N;
            const float mean = in ? p.bn_mean[col] : 0.f, rstd = in ? p.bn_rstd[col] : 0.f;
            const float scale = in ? p.bn_gamma[col] * rstd : 0.f;
            reinterpret_cast<float4*>(sconst)[et] = make_float4(mean, rstd, scale, in ? p.bn_beta[col] - mean * scale : 0.f);
          }
          asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory");
          const_n0 = n0;
        }
        if (BNR == 1) ptx::mbar_wait(bn_bar, tc & 1);     // mode 2 waits where it reads x / y
      }
      // a staging row outside the image is clipped by the store, but it must not reach the tensor-core statistics
      [[maybe_unused]] const bool zero_row = kHalo && (img0 + halo_bi >= p.n_img || h0 + halo_hr >= p.H);
      const uint32_t lane_col = grp * kColsPerGrp + ((uint32_t)(q * 32) << 16);
      const uint32_t taddr = tmem_base + slot * BLOCK_N + lane_col;
      if constexpr (kPipeLd) {
        if (!ld_inflight) {
          ptx::mbar_wait(&tmem_full[slot], aph);
          ptx::tc_fence_after();
          if constexpr (kNCh == 2) ptx::tmem_ld_32x64(taddr, rgn);
          else ptx::tmem_ld_32x32(taddr, rgn);
        }
        ptx::tmem_ld_wait();
        ld_inflight = false;
        // accumulator fully in registers: hand it back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[slot]);
      } else {
        ptx::mbar_wait(&tmem_full[slot], aph);
        ptx::tc_fence_after();
      }
      if (et == 0) EDL_TRACE(2, tc, 2);
#pragma unroll
      for (int c32 = 0; c32 < kColsPerGrp / 32; ++c32) {
        uint32_t rg1[32];
        if constexpr (!kPipeLd) {
          ptx::tmem_ld_32x32(taddr + c32 * 32, rg1);
          ptx::tmem_ld_wait();
          if (c32 == kColsPerGrp / 32 - 1) {
            // accumulator fully in registers: hand it back to the MMA warp before the slow part
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[slot]);
          }
        }
        const uint32_t* rg = kPipeLd ? &rgn[kPipeLd ? c32 * 32 : 0] : rg1;
        const int cbase = grp * kColsPerGrp + c32 * 32;   // first tile column of this chunk
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(rg[j]);
        if (has_add) {
          const uint32_t arow = ptx::smem_u32(sd) + (cbase >> 6) * (kBlockM * 128) + row * 128;
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
        if (!BNR) {
          // 32 consecutive per-column constants: eight 16-byte loads when the chunk lies inside the matrix and is aligned
          // (the scalar form costs 32 load instructions per chunk and thread)
          const bool vec_cols = n0 + cbase + 32 <= p.N;
          if (p.col_scale != nullptr) {
            const float* ps = p.col_scale + n0 + cbase;
            if (vec_cols && (reinterpret_cast<uintptr_t>(ps) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 v4 = __ldg(reinterpret_cast<const float4*>(ps) + j);
                f[4 * j] *= v4.x; f[4 * j + 1] *= v4.y; f[4 * j + 2] *= v4.z; f[4 * j + 3] *= v4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int col = n0 + cbase + j;
                f[j] *= col < p.N ? p.col_scale[col] : 0.f;
              }
            }
          }
          if (p.col_shift != nullptr) {
            const float* ps = p.col_shift + n0 + cbase;
            if (vec_cols && (reinterpret_cast<uintptr_t>(ps) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 v4 = __ldg(reinterpret_cast<const float4*>(ps) + j);
                f[4 * j] += v4.x; f[4 * j + 1] += v4.y; f[4 * j + 2] += v4.z; f[4 * j + 3] += v4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int col = n0 + cbase + j;
                f[j] += col < p.N ? p.col_shift[col] : 0.f;
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (kTcStats && kHalo && tcs && zero_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 0.f;
          }
        }
        if (BNR == 1) {
          // dy = the bf16 value this tile stores; masked by the ReLU of the BN layer, reduced per channel
          const uint32_t xrow = ptx::smem_u32(sx) + (cbase >> 6) * (kBlockM * 128) + row * 128;
          const uint32_t yrow = ptx::smem_u32(sy) + (cbase >> 6) * (kBlockM * 128) + row * 128;
          // two passes (sum dy_m, then sum dy_m * xhat) so that only one 32-value array is live at a time
          float rsum[2];
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
            float g[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int chunk = ((cbase >> 5) & 1) * 4 + c;
              const uint32_t off = (chunk ^ (row & 7)) << 4;
              uint32_t xw[4], yw[4] = {0, 0, 0, 0};
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                           : "=r"(xw[0]), "=r"(xw[1]), "=r"(xw[2]), "=r"(xw[3]) : "r"(xrow + off));
              if (p.bn_has_y)
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(yw[0]), "=r"(yw[1]), "=r"(yw[2]), "=r"(yw[3]) : "r"(yrow + off));
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 xv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw[j]));
                const float2 yv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yw[j]));
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int jj = c * 8 + 2 * j + e;
                  const int cc = cbase + jj;                                        // column inside the tile
                  const float xe = e == 0 ? xv.x : xv.y, ye = e == 0 ? yv.x : yv.y;
                  const float dyv = __bfloat162float(__float2bfloat16_rn(f[jj]));
                  const float4 kc4 = reinterpret_cast<const float4*>(sconst)[cc];   // mean, rstd, scale, shift
                  bool keep = row_ok;
                  if (p.bn_relu) keep = keep && (p.bn_has_y ? ye > 0.f : fmaf(xe, kc4.z, kc4.w) > 0.f);
                  // rows beyond the tensor hold stale shared memory (possibly NaN): select, never multiply
                  g[jj] = keep ? (pass == 0 ? dyv : dyv * ((xe - kc4.x) * kc4.y)) : 0.f;
                }
              }
            }
            rsum[pass] = warp_colsum32(g, lane);
          }
          const float r1 = rsum[0], r2 = rsum[1];
          const int col = n0 + cbase + lane;
          if (col < p.N) {
            if (local_stats) {
              atomicAdd(&sstats[col], r1);
              atomicAdd(&sstats[p.N + col], r2);
            } else {
              atomicAdd(&p.bn_dsums[col], r1);
              atomicAdd(&p.bn_dsums[p.N + col], r2);
            }
          }
        }
        // the previous tile's statistics MMAs read the staging buffer: they must be done before it is overwritten
        if (kTcStats && tcs && c32 == 0 && stat_batches > 0) ptx::mbar_wait(stat_bar, (stat_batches - 1) & 1);
        const int half = cbase >> 6;
        const uint32_t rowp = ptx::smem_u32(sd) + half * (kBlockM * 128) + srow * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int chunk = ((cbase >> 5) & 1) * 4 + c;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = f[c * 8 + j];
          const bf16x8 pk = pack8(v);
          const uint4 u = *reinterpret_cast<const uint4*>(&pk);
          if (store_ok) ptx::sts128(rowp + ((chunk ^ (srow & 7)) << 4), u.x, u.y, u.z, u.w);
        }
      }
      if (et == 0) EDL_TRACE(2, tc, 3);
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (et == 0) EDL_TRACE(2, tc, 4);
      if (BNR == 1 && ew == 0 && has_next)   // x / y buffers are free: prefetch the next tile's
        issue_bn_tiles<BLOCK_N, kConv>(p, next_tm, next_n0, rows_tile, sx, sy, &tmBnX, &tmBnY, bn_bar);
      if (ew == 0) {
        if (ptx::elect_one()) {
#pragma unroll
          for (int hh = 0; hh < (BLOCK_N + 63) / 64; ++hh) {
            if (n0 + hh * 64 >= p.N) continue;
            if (!kConv) ptx::tma_store_2d(&tmD, sd + hh * (kBlockM * 128), n0 + hh * 64, m0);
            else tma_store_4d_p(&tmD, sd + hh * (kBlockM * 128), n0 + hh * 64, 0, h0, img0);
          }
          ptx::tma_store_commit();
        }
        __syncwarp();
      }
      if constexpr (kPipeLd) {
        if (has_next) {
          const uint32_t nslot = (tc + 1) & 1, naph = ((tc + 1) >> 1) & 1;
          // only if the next accumulator is complete NOW: blocking here would put the loops below behind its MMAs
          if (__all_sync(0xffffffffu, ptx::mbar_test_wait(&tmem_full[nslot], naph))) {
            ptx::tc_fence_after();
            const uint32_t naddr = tmem_base + nslot * BLOCK_N + lane_col;
            if constexpr (kNCh == 2) ptx::tmem_ld_32x64(naddr, rgn);
            else ptx::tmem_ld_32x32(naddr, rgn);
            ld_inflight = true;
          }
        }
      }
      if constexpr (kTcStats) {
        if (tcs) {
          if (ew == 0) {
            // every epilogue thread has written its part of the tile (bar.sync above, generic -> async proxy fence before
            // it): hand the staging buffer to the MMA warp
            if (ptx::elect_one()) ptx::mbar_arrive(stage_rdy);
            __syncwarp();
          }
          ++stat_batches;
          if (!has_next || next_n0 != n0) {
            // last tile of this N tile in the CTA's sequence: read the statistics back into the CTA-local accumulators
            ptx::mbar_wait(stat_bar, (stat_batches - 1) & 1);
            ptx::tc_fence_after();
            if (grp == 0) {
              const int nrow = q * 32 + lane;                 // TMEM lane = column of the tile
              const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
              uint32_t g[32];
              ptx::tmem_ld_32x32(lane_addr + kGramCol + q * 32, g);     // the 32 x 32 diagonal block of this quarter
              const uint32_t s1 = ptx::tmem_ld_32x1(lane_addr + kSumCol);
              ptx::tmem_ld_wait();
              uint32_t gd = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) gd = (j == lane) ? g[j] : gd;
              if (n0 + nrow < p.N) {
                atomicAdd(&sstats[n0 + nrow], __uint_as_float(s1));
                atomicAdd(&sstats[p.N + n0 + nrow], __uint_as_float(gd));
              }
            }
            ptx::tc_fence_before();       // ordered before the next batch (accumulate = 0) by the bar.sync at the tile top
          }
        }
      }
      if (BNR == 2) {
        // BatchNorm-backward reduction of the tile just staged: per channel sum(dy_m) and sum(dy_m * xhat), dy_m = the
        // STORED bf16 gradient masked by the layer's ReLU.  Same scheme as the forward statistics below -- a thread owns
        // a pair of adjacent columns and every kSplit-th row, eight independent row loads in flight -- reading the
        // gradient from the staging tile and x (/ y) from the tiles the producer fetched by TMA.  (The first version
        // reduced in registers with a 31-shuffle transpose per 32 columns and pass: 46-59 us per short-K dgrad kernel
        // instead of 15 us, profiles/README.md.)
        ptx::mbar_wait(bn_bar, tc & 1);
        if (et == 0) EDL_TRACE(2, tc, 5);
        constexpr int kPairs = BLOCK_N / 2;
        constexpr int kSplit = kEpiThreads / kPairs;
        const int pair = et % kPairs;
        const int part = et / kPairs;
        const int col = pair * 2;
        const bool valid = n0 + col < p.N;
        const int half = col >> 6, cc = col & 63, chunk = cc >> 3, within = cc & 7;
        const uint32_t toff = half * (kBlockM * 128) + within * 2;
        const uint32_t based = ptx::smem_u32(sd) + toff, basex = ptx::smem_u32(sx) + toff, basey = ptx::smem_u32(sy) + toff;
        const float4 k0 = reinterpret_cast<const float4*>(sconst)[col];          // mean, rstd, scale, shift
        const float4 k1 = reinterpret_cast<const float4*>(sconst)[col + 1];
        const bool relu = p.bn_relu != 0, has_y = p.bn_has_y != 0;
        // s = sum(dy_m), q = sum(dy_m * x) over this thread's rows; sum(dy_m * xhat) = rstd * (q - mean * s) is formed once
        // per thread and tile below (two instructions per element less than normalising x in the loop).  With the BN
        // OUTPUT at hand the ReLU mask is a packed bf16 compare and a multiplication by 1.0 / 0.0 (exact).
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
        auto row_update = [&](uint32_t wd, uint32_t wx, uint32_t wy) {
          __nv_bfloat162 d2 = *reinterpret_cast<const __nv_bfloat162*>(&wd);
          const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wx));
          float2 d;
          if (relu && has_y) {
            d2 = __hmul2(d2, __hgt2(*reinterpret_cast<const __nv_bfloat162*>(&wy), zero2));
            d = __bfloat1622float2(d2);
          } else {
            d = __bfloat1622float2(d2);
            if (relu) {
              d.x = fmaf(x.x, k0.z, k0.w) > 0.f ? d.x : 0.f;
              d.y = fmaf(x.y, k1.z, k1.w) > 0.f ? d.y : 0.f;
            }
          }
          s0 += d.x; q0 = fmaf(d.x, x.x, q0);
          s1 += d.y; q1 = fmaf(d.y, x.y, q1);
        };
        // Row r of a staged tile sits at r * 128 + ((chunk ^ (r & 7)) << 4).  A thread visits rows r0, r0 + kSplit, ...:
        // with kSplit a multiple of 4 the swizzle term takes two values (even / odd steps), so two base addresses plus
        // compile-time offsets replace five address instructions per row and load.
        constexpr bool kFastRows = kSplit % 4 == 0;      // (the 256-column kernels carry no reduction; they only compile this)
        auto accum_rows = [&](int r_begin, int r_end) {
          int rr = r_begin + part;
          uint32_t o0 = rr * 128 + ((chunk ^ (rr & 7)) << 4), o1 = (rr + kSplit) * 128 + ((chunk ^ ((rr + kSplit) & 7)) << 4);
          for (; rr + 7 * kSplit < r_end; rr += 8 * kSplit, o0 += 8 * kSplit * 128, o1 += 8 * kSplit * 128) {
            uint32_t wd[8], wx[8], wy[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int r = rr + u * kSplit;
              const uint32_t o = kFastRows ? ((u & 1) ? o1 : o0) + (u >> 1) * (2 * kSplit * 128)
                                           : (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4));
              wd[u] = ptx::lds32(based + o);
              wx[u] = ptx::lds32(basex + o);
              wy[u] = has_y ? ptx::lds32(basey + o) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) row_update(wd[u], wx[u], wy[u]);
          }
#pragma unroll 1
          for (int u = 0; rr < r_end; rr += kSplit, ++u) {
            const uint32_t o = kFastRows ? ((u & 1) ? o1 : o0) + (u >> 1) * (2 * kSplit * 128)
                                         : (uint32_t)(rr * 128 + ((chunk ^ (rr & 7)) << 4));
            row_update(ptx::lds32(based + o), ptx::lds32(basex + o), has_y ? ptx::lds32(basey + o) : 0u);
          }
        };
        if (valid) {
          // only rows that are real pixels of the tensor: the others hold stale shared memory
          if (!kConv) {
            int rows_valid = p.M - m0;
            if (rows_valid > kBlockM) rows_valid = kBlockM;
            accum_rows(0, rows_valid);
          } else {
            const int rows_per_img = p.BH * p.W;
            int hv = p.H - h0;
            if (hv > p.BH) hv = p.BH;
            for (int b = 0; b < p.BN && img0 + b < p.n_img; ++b) accum_rows(b * rows_per_img, b * rows_per_img + hv * p.W);
          }
          run_s0 += s0; run_q0 += (q0 - k0.x * s0) * k0.y;          // k = {mean, rstd, scale, shift}
          run_s1 += s1; run_q1 += (q1 - k1.x * s1) * k1.y;
          if (!has_next || next_n0 != n0) {
            if (local_stats) {
              atomicAdd(&sstats[n0 + col], run_s0);
              atomicAdd(&sstats[p.N + n0 + col], run_q0);
              if (n0 + col + 1 < p.N) {
                atomicAdd(&sstats[n0 + col + 1], run_s1);
                atomicAdd(&sstats[p.N + n0 + col + 1], run_q1);
              }
            } else {
              atomicAdd(&p.bn_dsums[n0 + col], run_s0);
              atomicAdd(&p.bn_dsums[p.N + n0 + col], run_q0);
              if (n0 + col + 1 < p.N) {
                atomicAdd(&p.bn_dsums[n0 + col + 1], run_s1);
                atomicAdd(&p.bn_dsums[p.N + n0 + col + 1], run_q1);
              }
            }
          }
        }
        if (!has_next || next_n0 != n0) run_s0 = run_s1 = run_q0 = run_q1 = 0.f;
        // everybody is done with this tile's x / y: fetch the next tile's while its MMAs run
        if (et == 0) EDL_TRACE(2, tc, 6);
        asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads) : "memory");
        if (et == 0) EDL_TRACE(2, tc, 7);
        if (ew == 0 && has_next)
          issue_bn_tiles<BLOCK_N, kConv>(p, next_tm, next_n0, rows_tile, sx, sy, &tmBnX, &tmBnY, bn_bar);
      }
      if (!BNR && et == 0) EDL_TRACE(2, tc, 5);
      if (!BNR && p.col_stats != nullptr && !tcs) {
        // Per-channel sum / sum of squares of the STORED bf16 values, from the staged tile.  A thread owns
        // a PAIR of adjacent columns (one 32-bit shared load per row) and every kSplit-th row; eight
        // independent loads are in flight per thread (the naive one-column, one-accumulator loop was the
        // bottleneck of short-K tiles: ~1 us of exposed LDS latency per tile, profiles/prof_persist_fwd).
        constexpr int kPairs = BLOCK_N / 2;                    // 64 (N=128) or 32 (N=64)
        constexpr int kSplit = kEpiThreads / kPairs;           // 4 or 8 row subsets
        const int pair = et % kPairs;
        const int part = et / kPairs;
        const int col = pair * 2;
        const bool valid = n0 + col < p.N;
        const int half = col >> 6, cc = col & 63, chunk = cc >> 3, within = cc & 7;
        const uint32_t base = ptx::smem_u32(sd) + half * (kBlockM * 128) + within * 2;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        constexpr bool kFastRows = kSplit % 4 == 0;
        auto accum_rows = [&](int r_begin, int r_end) {      // addressing: see the BN-backward loop above
          int rr = r_begin + part;
          uint32_t o0 = base + rr * 128 + ((chunk ^ (rr & 7)) << 4);
          uint32_t o1 = base + (rr + kSplit) * 128 + ((chunk ^ ((rr + kSplit) & 7)) << 4);
          for (; rr + 7 * kSplit < r_end; rr += 8 * kSplit, o0 += 8 * kSplit * 128, o1 += 8 * kSplit * 128) {
            uint32_t w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int r = rr + u * kSplit;
              w[u] = ptx::lds32(kFastRows ? ((u & 1) ? o1 : o0) + (u >> 1) * (2 * kSplit * 128)
                                          : base + r * 128 + ((chunk ^ (r & 7)) << 4));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[u]));
              s0 += v.x; q0 = fmaf(v.x, v.x, q0);
              s1 += v.y; q1 = fmaf(v.y, v.y, q1);
            }
          }
#pragma unroll 1
          for (int u = 0; rr < r_end; rr += kSplit, ++u) {
            const uint32_t w = ptx::lds32(kFastRows ? ((u & 1) ? o1 : o0) + (u >> 1) * (2 * kSplit * 128)
                                                    : base + rr * 128 + ((chunk ^ (rr & 7)) << 4));
            const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
            s0 += v.x; q0 = fmaf(v.x, v.x, q0);
            s1 += v.y; q1 = fmaf(v.y, v.y, q1);
          }
        };
        if (valid) {
          if (!kConv) {
            int rows_valid = p.M - m0;
            if (rows_valid > kBlockM) rows_valid = kBlockM;
            accum_rows(0, rows_valid);
          } else {
            const int rows_per_img = p.BH * p.W;
            int hv = p.H - h0;
            if (hv > p.BH) hv = p.BH;
            for (int b = 0; b < p.BN && img0 + b < p.n_img; ++b) accum_rows(b * rows_per_img, b * rows_per_img + hv * p.W);
          }
        }
        if (et == 0) EDL_TRACE(2, tc, 7);
        run_s0 += s0; run_s1 += s1; run_q0 += q0; run_q1 += q1;
        const bool run_end = !has_next || next_n0 != n0;
        if (run_end) { s0 = run_s0; s1 = run_s1; q0 = run_q0; q1 = run_q1; run_s0 = run_s1 = run_q0 = run_q1 = 0.f; }
        if (!run_end) {
          // the sums travel on in registers
        } else if (local_stats) {
          if (valid) {
            atomicAdd(&sstats[n0 + col], s0);
            atomicAdd(&sstats[p.N + n0 + col], q0);
            if (n0 + col + 1 < p.N) {
              atomicAdd(&sstats[n0 + col + 1], s1);
              atomicAdd(&sstats[p.N + n0 + col + 1], q1);
            }
          }
        } else {
          // lanes 2i, 2i+1 hold four neighbouring columns -> one vector reduction each for sum and sum^2
          const float s2 = __shfl_down_sync(0xffffffffu, s0, 1), s3 = __shfl_down_sync(0xffffffffu, s1, 1);
          const float q2 = __shfl_down_sync(0xffffffffu, q0, 1), q3 = __shfl_down_sync(0xffffffffu, q1, 1);
          const int c4 = col & ~3;
          float* ps4 = &p.col_stats[n0 + c4];
          const bool vec = (p.N % 4 == 0) && n0 + c4 + 3 < p.N && ((reinterpret_cast<uintptr_t>(ps4) & 15) == 0);
          if (vec) {
            if ((pair & 1) == 0) {
              red_add_v4(ps4, s0, s1, s2, s3);
              red_add_v4(ps4 + p.N, q0, q1, q2, q3);
            }
          } else if (valid) {
            atomicAdd(&p.col_stats[n0 + col], s0);
            atomicAdd(&p.col_stats[p.N + n0 + col], q0);
            if (n0 + col + 1 < p.N) {
              atomicAdd(&p.col_stats[n0 + col + 1], s1);
              atomicAdd(&p.col_stats[p.N + n0 + col + 1], q1);
            }
          }
        }
        if (et == 0) EDL_TRACE(2, tc, 6);
      }
    }
    if (local_stats) {
      // one flush per CTA: only the column groups this CTA touched are non-zero
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      const bool vec_ok = (p.N % 4 == 0) && ((reinterpret_cast<uintptr_t>(stats_dst) & 15) == 0);
      if (vec_ok) {
        for (int i = et * 4; i < 2 * p.N; i += kEpiThreads * 4) {
          const float4 v = *reinterpret_cast<const float4*>(&sstats[i]);
          if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add_v4(&stats_dst[i], v.x, v.y, v.z, v.w);
        }
      } else {
        for (int i = et; i < 2 * p.N; i += kEpiThreads)
          if (sstats[i] != 0.f) atomicAdd(&stats_dst[i], sstats[i]);
      }
    }
    if (ew == 0) {
      if (ptx::elect_one()) ptx::tma_store_wait_read0();
      __syncwarp();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

bool g_persistent = true;
long long* g_trace = nullptr;
bool g_wide_tiles = [] {
  const char* e = getenv("EDL_GEMM_WIDE");
  return !(e != nullptr && e[0] == '0');
}();
// fused BatchNorm-backward reduction in the dgrad epilogue: 2 (default) = column-pair loop over the staged tiles,
// 1 = the round-1 in-register shuffle transpose (epilogue-bound; EDL_BNR_MODE=1 keeps it selectable for A/B)
// haloed A tiles for the 3x3 / stride 1 convolutions (EDL_CONV_HALO=0 keeps the one-box-per-tap version for A/B)
bool g_conv_halo = [] {
  const char* e = getenv("EDL_CONV_HALO");
  return !(e != nullptr && e[0] == '0');
}();

// BatchNorm statistics of the forward tiles on the tensor cores (EDL_TC_STATS=1).  Exact and validated
// (tests/test_persist_gpu.py), but NOT faster than the column-pair loop: the two MMA batches read ~100 KB of shared memory
// per tile as MN-major operands and compete with the TMA store and the next tile's pack for the same shared memory
// (3600 vs 3850 cycles per short-K tile, the student step 4.403 vs 4.402 ms: profiles/trace_persist_c31.txt), so the loop
// stays the default.
bool g_tc_stats = [] {
  const char* e = getenv("EDL_TC_STATS");
  return e != nullptr && e[0] == '1';
}();
bool g_conv_bres = [] {
  const char* e = getenv("EDL_CONV_BRES");
  return !(e != nullptr && e[0] == '0');
}();

// tile geometry with Wb = W + 1 columns per image row (conv3x3.cu's planner with the row width replaced)
bool halo_geometry(int N, int H, int W, int* BH, int* BN, int* tiles_h, int* tiles_img) {
  const int wb = W + 1;
  if (W < 1 || wb > kBlockM || H < 1) return false;
  int bh = kBlockM / wb;
  if (bh > H) bh = H;
  int best = bh;
  for (int d = bh; d >= 1; --d)
    if (H % d == 0) { best = d; break; }
  if (best * 4 < bh * 3) best = bh;
  int bn = 1;
  if (best == H) {
    bn = kBlockM / (H * wb);
    if (bn < 1) bn = 1;
    if (bn > N) bn = N;
  }
  *BH = best; *BN = bn;
  *tiles_h = (H + best - 1) / best;
  *tiles_img = (N + bn - 1) / bn;
  return true;
}

int g_bnr_mode = [] {
  const char* e = getenv("EDL_BNR_MODE");
  return (e != nullptr && e[0] == '1') ? 1 : 2;
}();

// 16 epilogue warps are selectable (EDL_EPI_WARPS=16) but not the default: measured neutral on the teacher (3.95 vs 3.97 ms)
// and 1 % slower on the student step (4.341 vs 4.302 ms, profiles/bench_runs.json b26_*): in the captured step these
// kernels wait for DRAM, not for instruction latency as the cold single-kernel ncu capture suggested.
bool g_epi16 = [] {
  const char* e = getenv("EDL_EPI_WARPS");
  return e != nullptr && e[0] == '1' && e[1] == '6';
}();

template <int BLOCK_N, int STAGES, int MODE, int BNR, int ASTAGES, int EPIW>
const char* launch_pe(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const PersistParams& p,
                     cudaStream_t stream, const CUtensorMap* tmAdd = nullptr, const CUtensorMap* tmBnX = nullptr,
                     const CUtensorMap* tmBnY = nullptr) {
  using L = PSmem<BLOCK_N, STAGES, BNR, ASTAGES>;
  static_assert(L::kTotal <= 227 * 1024, "shared memory budget");
  constexpr int kThreads = 64 + EPIW * 32;
  auto kern = gemm_persist_kernel<BLOCK_N, STAGES, MODE, BNR, ASTAGES, EPIW>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    apply_carveout((const void*)kern);
    attr_set = true;
  }
  const int total = p.tiles_m * p.tiles_n;
  const int grid = total < kNumSMs ? total : kNumSMs;
  PersistParams pt = p;
  pt.trace = g_trace;
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(kThreads), (size_t)L::kTotal, stream, tmA, tmB, tmD,
                             tmAdd != nullptr ? *tmAdd : tmD, tmBnX != nullptr ? *tmBnX : tmD,
                             tmBnY != nullptr ? *tmBnY : tmD, pt);
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

template <int BLOCK_N, int STAGES, int MODE, int BNR = 0, int ASTAGES = 0>
const char* launch_p(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const PersistParams& p,
                     cudaStream_t stream, const CUtensorMap* tmAdd = nullptr, const CUtensorMap* tmBnX = nullptr,
                     const CUtensorMap* tmBnY = nullptr) {
  if constexpr (BLOCK_N >= 128 && BNR != 1) {
    if (g_epi16) return launch_pe<BLOCK_N, STAGES, MODE, BNR, ASTAGES, 16>(tmA, tmB, tmD, p, stream, tmAdd, tmBnX, tmBnY);
  }
  return launch_pe<BLOCK_N, STAGES, MODE, BNR, ASTAGES, 8>(tmA, tmB, tmD, p, stream, tmAdd, tmBnX, tmBnY);
}

int bnr_mode() { return g_bnr_mode; }

const char* tmap2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                   uint32_t box_inner, uint32_t box_outer) {
  const uint64_t dims[2] = {inner, outer};
  const uint64_t st[1] = {pitch_elems * 2};
  const uint32_t box[2] = {box_inner, box_outer};
  return encode_tmap_bf16(out, ptr, 2, dims, st, box);
}

}  // namespace

namespace {
void fill_bn(PersistParams& p, const BnBwdFuse& bn) {
  p.bn_mean = bn.mean; p.bn_rstd = bn.rstd; p.bn_gamma = bn.gamma; p.bn_beta = bn.beta;
  p.bn_dsums = bn.dsums;
  p.bn_relu = bn.relu ? 1 : 0;
  p.bn_has_y = (bn.relu && bn.y != nullptr) ? 1 : 0;
}
}  // namespace

void set_bnr_mode(int mode) { g_bnr_mode = mode == 2 ? 2 : 1; }
int get_bnr_mode() { return g_bnr_mode; }
void set_persistent_gemm(bool on) { g_persistent = on; }
void set_wide_gemm_tiles(bool on) { g_wide_tiles = on; }
bool persistent_gemm_enabled() { return g_persistent; }

// GEMM front end (EPI 0 semantics of gemm.cu; A K-major)
const char* gemm_bf16_persistent(const GemmArgs& g, cudaStream_t stream) {
  if (gemm_pair_supported(g)) return gemm_bf16_pair(g, stream);
  alignas(64) CUtensorMap tmA, tmB, tmD;
  const bool n64 = g.N <= 64;
  // 128 x 256 tiles for the compute-heavy GEMMs (the teacher's 1x1 convolutions: N = 512 .. 4096, K = 256 .. 4096):
  // a 128 x 128 tile needs 32 KB of operands per 256 tensor-core cycles = 128 B / clk / SM, about twice what the
  // L2 -> SM path delivers when every SM pulls (profiles/teacher_r1.txt: 35-40 % of the bf16 peak); with N = 256 the
  // A tile is reused twice as often and the requirement drops to 96 B / clk.  Both accumulators then fill the TMEM.
  const bool wide = g_wide_tiles && !g.b_mn_major && g.bn.x == nullptr && g.col_stats == nullptr && g.N % 256 == 0 &&
                    g.K >= 256 &&
                    (int64_t)((g.M + kBlockM - 1) / kBlockM) * (g.N / 256) >= kNumSMs;
  const int bn = n64 ? 64 : (wide ? 256 : 128);
  if (const char* e = tmap2d(&tmA, g.A, g.K, g.M, g.lda, kBlockK, kBlockM)) return e;
  if (!g.b_mn_major) {
    if (const char* e = tmap2d(&tmB, g.B, g.K, g.N, g.ldb, kBlockK, bn)) return e;
  } else {
    if (const char* e = tmap2d(&tmB, g.B, g.N, g.K, g.ldb, 64, kBlockK)) return e;
  }
  if (const char* e = tmap2d(&tmD, g.D, g.N, g.M, g.ldd, 64, kBlockM)) return e;
  PersistParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.tiles_m = (g.M + kBlockM - 1) / kBlockM;
  p.tiles_n = (g.N + bn - 1) / bn;
  p.num_kb = (g.K + kBlockK - 1) / kBlockK;
  p.col_scale = g.col_scale; p.col_shift = g.col_shift; p.relu = g.relu ? 1 : 0;
  p.col_stats = g.col_stats;
  p.tc_stats = g_tc_stats ? 1 : 0;
  p.add_src = reinterpret_cast<const __nv_bfloat16*>(g.add_src);
  p.ld_add = g.ld_add;
  alignas(64) CUtensorMap tmAdd;
  const CUtensorMap* padd = nullptr;
  if (g.add_src != nullptr) {
    if (g.N % 8 != 0 || g.ld_add % 8 != 0 || (reinterpret_cast<uintptr_t>(g.add_src) & 15) != 0)
      return "gemm add_src needs N % 8 == 0 and 16-byte aligned rows";
    if (const char* e = tmap2d(&tmAdd, g.add_src, g.N, g.M, g.ld_add, 64, kBlockM)) return e;
    padd = &tmAdd;
  }
  if (g.bn.x != nullptr) {
    // dgrad whose output is the gradient of a BatchNorm output: reduce it in the epilogue
    if (!g.b_mn_major || g.col_stats != nullptr) return "fused BN-backward reduction is a dgrad (B MN-major) feature";
    alignas(64) CUtensorMap tmBx, tmBy;
    if (const char* e = tmap2d(&tmBx, g.bn.x, g.N, g.M, g.ldd, 64, kBlockM)) return e;
    const bool has_y = g.bn.relu && g.bn.y != nullptr;
    if (has_y)
      if (const char* e = tmap2d(&tmBy, g.bn.y, g.N, g.M, g.ldd, 64, kBlockM)) return e;
    fill_bn(p, g.bn);
    if (bnr_mode() == 2)
      return n64 ? launch_p<64, 4, 1, 2>(tmA, tmB, tmD, p, stream, padd, &tmBx, has_y ? &tmBy : nullptr)
                 : launch_p<128, 3, 1, 2>(tmA, tmB, tmD, p, stream, padd, &tmBx, has_y ? &tmBy : nullptr);
    return n64 ? launch_p<64, 4, 1, 1>(tmA, tmB, tmD, p, stream, padd, &tmBx, has_y ? &tmBy : nullptr)
               : launch_p<128, 3, 1, 1>(tmA, tmB, tmD, p, stream, padd, &tmBx, has_y ? &tmBy : nullptr);
  }
  if (!g.b_mn_major) {
    if (wide) return launch_p<256, 3, 0>(tmA, tmB, tmD, p, stream, padd);
    return n64 ? launch_p<64, 6, 0>(tmA, tmB, tmD, p, stream, padd) : launch_p<128, 5, 0>(tmA, tmB, tmD, p, stream, padd);
  }
  return n64 ? launch_p<64, 6, 1>(tmA, tmB, tmD, p, stream, padd) : launch_p<128, 5, 1>(tmA, tmB, tmD, p, stream, padd);
}

// conv front end: geometry comes from conv3x3.cu's planner
const char* conv3x3_bf16_persistent(const Conv3x3Args& a, int BH, int BN, int tiles_h, int tiles_img,
                                    cudaStream_t stream) {
  const bool dg = a.dgrad;
  const int cx = dg ? a.Cout : a.Cin, cy = dg ? a.Cin : a.Cout;
  const int groups = dg ? 1 : a.groups;
  const int cin_g = cx / groups, cout_g = cy / groups;
  const bool n64 = cout_g <= 64;
  const int bn = n64 ? 64 : 128;
  alignas(64) CUtensorMap tmX, tmW, tmY;
  const int sd = a.stride == 2 ? 2 : 1;
  if (sd == 2 && (dg || 2 * a.W > 256 || 2 * BH > 256)) return "conv3x3 stride 2: fprop only, output width <= 128";
  const bool halo = g_conv_halo && sd == 1 && !(dg && a.bn.x != nullptr && bnr_mode() != 2) &&
                    halo_geometry(a.N, a.H, a.W, &BH, &BN, &tiles_h, &tiles_img);
  if (halo) {
    const uint64_t dims[4] = {(uint64_t)cx, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
    const uint64_t st[3] = {(uint64_t)cx * 2, (uint64_t)a.W * cx * 2, (uint64_t)a.H * a.W * cx * 2};
    const uint32_t box[4] = {64, (uint32_t)(a.W + 1), (uint32_t)BH, (uint32_t)BN};
    if (const char* e = encode_tmap_bf16(&tmX, a.X, 4, dims, st, box)) return e;
  } else {
    // stride 2: the input is [N, 2H, 2W, C]; a box that spans 2W x 2BH elements with traversal stride 2 in w and h
    // delivers the W x BH pixels a filter tap needs, starting at column s - 1 and row 2*h0 + r - 1
    const uint64_t dims[4] = {(uint64_t)cx, (uint64_t)a.W * sd, (uint64_t)a.H * sd, (uint64_t)a.N};
    const uint64_t st[3] = {(uint64_t)cx * 2, (uint64_t)a.W * sd * cx * 2, (uint64_t)a.H * sd * a.W * sd * cx * 2};
    const uint32_t box[4] = {64, (uint32_t)(a.W * sd), (uint32_t)(BH * sd), (uint32_t)BN};
    const uint32_t es[4] = {1, (uint32_t)sd, (uint32_t)sd, 1};
    if (const char* e = encode_tmap_strided(&tmX, a.X, 4, dims, st, box, es, 2)) return e;
  }
  {
    const uint64_t dims[4] = {(uint64_t)cy, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
    const uint64_t st[3] = {(uint64_t)cy * 2, (uint64_t)a.W * cy * 2, (uint64_t)a.H * a.W * cy * 2};
    const uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)BH, (uint32_t)BN};
    if (const char* e = encode_tmap_bf16(&tmY, a.Y, 4, dims, st, box)) return e;
  }
  const int w_cin = a.Cin / a.groups;      // channel extent of the KRSC weight tensor
  if (const char* e = tmap2d(&tmW, a.Wt, (uint64_t)9 * w_cin, a.Cout, (uint64_t)9 * w_cin, 64, dg ? kBlockK : bn))
    return e;
  PersistParams p{};
  p.M = tiles_img * tiles_h * kBlockM; p.N = cy; p.K = 9 * cx;
  p.tiles_m = tiles_img * tiles_h;
  p.tiles_n = (cy + bn - 1) / bn;
  p.kc_blocks = cin_g / kBlockK;
  p.num_kb = (halo ? 3 : 9) * p.kc_blocks;
  p.Wb = halo ? a.W + 1 : a.W;
  p.b_res = (halo && n64 && p.kc_blocks == 1 && g_conv_bres) ? 1 : 0;
  p.tc_stats = (halo && g_tc_stats) ? 1 : 0;
  p.col_stats = dg ? nullptr : a.col_stats;
  p.col_scale = dg ? nullptr : a.col_scale;
  p.col_shift = dg ? nullptr : a.col_shift;
  p.relu = (!dg && a.relu) ? 1 : 0;
  p.cin_g = cin_g; p.cout_g = cout_g;
  p.n_img = a.N; p.H = a.H; p.W = a.W; p.c_in_w = w_cin; p.BH = BH; p.BN = BN; p.tiles_h = tiles_h;
  p.stride = sd;
  if (dg && a.bn.x != nullptr) {
    alignas(64) CUtensorMap tmBx, tmBy;
    const uint64_t dims[4] = {(uint64_t)cy, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
    const uint64_t st[3] = {(uint64_t)cy * 2, (uint64_t)a.W * cy * 2, (uint64_t)a.H * a.W * cy * 2};
    const uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)BH, (uint32_t)BN};
    if (const char* e = encode_tmap_bf16(&tmBx, a.bn.x, 4, dims, st, box)) return e;
    const bool has_y = a.bn.relu && a.bn.y != nullptr;
    if (has_y)
      if (const char* e = encode_tmap_bf16(&tmBy, a.bn.y, 4, dims, st, box)) return e;
    fill_bn(p, a.bn);
    if (halo)
      return n64 ? launch_p<64, 9, 3, 2, 3>(tmX, tmW, tmY, p, stream, nullptr, &tmBx, has_y ? &tmBy : nullptr)
                 : launch_p<128, 4, 3, 2, 2>(tmX, tmW, tmY, p, stream, nullptr, &tmBx, has_y ? &tmBy : nullptr);
    if (bnr_mode() == 2)
      return n64 ? launch_p<64, 4, 3, 2>(tmX, tmW, tmY, p, stream, nullptr, &tmBx, has_y ? &tmBy : nullptr)
                 : launch_p<128, 3, 3, 2>(tmX, tmW, tmY, p, stream, nullptr, &tmBx, has_y ? &tmBy : nullptr);
    return n64 ? launch_p<64, 4, 3, 1>(tmX, tmW, tmY, p, stream, nullptr, &tmBx, has_y ? &tmBy : nullptr)
               : launch_p<128, 3, 3, 1>(tmX, tmW, tmY, p, stream, nullptr, &tmBx, has_y ? &tmBy : nullptr);
  }
  if (halo) {
    // 64-column tiles: six A slots (with resident weights the A ring is all that is in flight: L2 -> SM throughput is
    // bytes in flight / ~3 us, profiles/prof_c25_conv3g.ncu.txt)
    if (!dg)
      return n64 ? launch_p<64, 9, 2, 0, 6>(tmX, tmW, tmY, p, stream) : launch_p<128, 7, 2, 0, 3>(tmX, tmW, tmY, p, stream);
    return n64 ? launch_p<64, 9, 3, 0, 6>(tmX, tmW, tmY, p, stream) : launch_p<128, 7, 3, 0, 3>(tmX, tmW, tmY, p, stream);
  }
  if (!dg) return n64 ? launch_p<64, 6, 2>(tmX, tmW, tmY, p, stream) : launch_p<128, 5, 2>(tmX, tmW, tmY, p, stream);
  return n64 ? launch_p<64, 6, 3>(tmX, tmW, tmY, p, stream) : launch_p<128, 5, 3>(tmX, tmW, tmY, p, stream);
}

void set_epilogue_warps(int n) { g_epi16 = n >= 16; }
void set_tc_stats(bool on) { g_tc_stats = on; }
void set_persist_trace(long long* buf) { g_trace = buf; }
// the tile geometry of the halo layout, for the CPU replay of its index math (tests/test_halo_plan_cpu.py)
bool conv3x3_halo_plan(int N, int H, int W, int* BH, int* BN, int* tiles_h, int* tiles_img) {
  return halo_geometry(N, H, W, BH, BN, tiles_h, tiles_img);
}
void set_conv_halo(bool on) { g_conv_halo = on; }
void set_conv_resident_weights(bool on) { g_conv_bres = on; }
bool get_conv_halo() { return g_conv_halo; }

// Input gradient of the 3x3 / pad 1 / STRIDE 2 convolution (round 2).  dX[n, h, w] = sum over the taps (r, s) with
// h + 1 - r and w + 1 - s even of dY[n, (h + 1 - r) / 2, (w + 1 - s) / 2] * W[:, r, s, :]: for a fixed parity (h & 1, w & 1)
// that is a stride-1 convolution of dY with 1, 2, 2 or 4 taps whose result lands on every second pixel of dX.  Four
// launches of the dgrad mode with an explicit tap list and a strided output tensor map -- 9 tap-MMAs per output-gradient
// pixel in total, exactly the work of the forward pass (zero-insertion + stride-1 dgrad costs four times that).
// a.X = dY [N, Ho, Wo, Cout], a.Y = dX [N, 2 Ho, 2 Wo, Cin], a.N / H / W = the dY geometry.
const char* conv3x3_dgrad_s2_persistent(const Conv3x3Args& a, int BH, int BN, int tiles_h, int tiles_img,
                                        cudaStream_t stream) {
  const int cx = a.Cout, cy = a.Cin;
  const bool n64 = cy <= 64;
  const int bn = n64 ? 64 : 128;
  alignas(64) CUtensorMap tmX, tmW;
  {
    const uint64_t dims[4] = {(uint64_t)cx, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
    const uint64_t st[3] = {(uint64_t)cx * 2, (uint64_t)a.W * cx * 2, (uint64_t)a.H * a.W * cx * 2};
    const uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)BH, (uint32_t)BN};
    if (const char* e = encode_tmap_bf16(&tmX, a.X, 4, dims, st, box)) return e;
  }
  if (const char* e = tmap2d(&tmW, a.Wt, (uint64_t)9 * a.Cin, a.Cout, (uint64_t)9 * a.Cin, 64, kBlockK)) return e;
  const uint64_t W2 = 2ull * a.W, H2 = 2ull * a.H;
  for (int ph = 0; ph < 2; ++ph) {
    for (int pw = 0; pw < 2; ++pw) {
      alignas(64) CUtensorMap tmY;
      {
        // the pixels (2 i + ph, 2 j + pw) of dX as a dense-looking [N, Ho, Wo, Cin] tensor with doubled pixel strides
        const uint64_t dims[4] = {(uint64_t)cy, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.N};
        const uint64_t st[3] = {2ull * cy * 2, 2ull * W2 * cy * 2, H2 * W2 * cy * 2};
        const uint32_t box[4] = {64, (uint32_t)a.W, (uint32_t)BH, (uint32_t)BN};
        const char* base = reinterpret_cast<const char*>(a.Y) + ((uint64_t)ph * W2 + pw) * cy * 2;
        if (const char* e = encode_tmap_bf16(&tmY, base, 4, dims, st, box)) return e;
      }
      PersistParams p{};
      p.M = tiles_img * tiles_h * kBlockM; p.N = cy; p.K = 9 * cx;
      p.tiles_m = tiles_img * tiles_h;
      p.tiles_n = (cy + bn - 1) / bn;
      p.kc_blocks = cx / kBlockK;
      p.cin_g = cx; p.cout_g = cy;
      p.n_img = a.N; p.H = a.H; p.W = a.W; p.c_in_w = a.Cin; p.BH = BH; p.BN = BN; p.tiles_h = tiles_h;
      p.stride = 1;
      int nt = 0;
      for (int r = 0; r < 3; ++r) {
        if (((ph + 1 - r) & 1) != 0) continue;
        for (int sft = 0; sft < 3; ++sft) {
          if (((pw + 1 - sft) & 1) != 0) continue;
          p.tap_dh[nt] = (signed char)((ph + 1 - r) / 2);      // exact: the numerator is even (0 or 2)
          p.tap_dw[nt] = (signed char)((pw + 1 - sft) / 2);
          p.tap_w[nt] = (signed char)(r * 3 + sft);
          ++nt;
        }
      }
      p.ntaps = nt;
      p.num_kb = nt * p.kc_blocks;
      const char* e = n64 ? launch_p<64, 6, 3>(tmX, tmW, tmY, p, stream) : launch_p<128, 5, 3>(tmX, tmW, tmY, p, stream);
      if (e != nullptr) return e;
    }
  }
  return nullptr;
}

}  // namespace edl
