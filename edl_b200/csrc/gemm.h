// Host API of the tcgen05 GEMM / implicit-GEMM conv kernels (gemm.cu, conv.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edl {

// Fused BatchNorm-backward reduction (persistent dgrad kernels): the GEMM / conv output D is the gradient
// dy of a BN layer's output; the epilogue accumulates dsums[0:C] += sum(dy_m), dsums[C:2C] += sum(dy_m * xhat)
// (dy_m = dy masked by the layer's ReLU) so that the BN backward needs no reduction pass of its own.
struct BnBwdFuse {
  const void* x = nullptr;       // bf16, BN input (conv output), same shape / layout as D
  const void* y = nullptr;       // bf16, BN output: only for ReLU after a residual add (mask = y > 0)
  const float* mean = nullptr;
  const float* rstd = nullptr;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float* dsums = nullptr;        // fp32 [2C], pre-zeroed
  bool relu = false;
};

struct GemmArgs {
  const void* A = nullptr;  // bf16
  const void* B = nullptr;  // bf16
  void* D = nullptr;        // bf16 [M, N] (ignored when out_f32 != nullptr)
  int M = 0, N = 0, K = 0;
  int64_t lda = 0, ldb = 0, ldd = 0;  // row pitches in elements of the *stored* 2-D arrays
  bool a_mn_major = false;            // A stored as [K, M] (M contiguous)
  bool b_mn_major = false;            // B stored as [K, N] (N contiguous); default B is [N, K]
  const float* col_scale = nullptr;   // per-output-column scale (folded BN) or nullptr
  const float* col_shift = nullptr;   // per-output-column shift / bias or nullptr
  bool relu = false;
  float* col_stats = nullptr;         // [2N] fp32: += sum, sum of squares of the stored outputs
  float* out_f32 = nullptr;           // split-K target: fp32 [M, N], += A*B
  int split_k = 1;
  // fused split-K finalize (optional, with out_f32 as an all-zero workspace on entry):
  // the last CTA per tile writes bf16(out_f32 tile) (+= if accumulate_out) to out_bf16 [M, N] with
  // row pitch ldo, then re-zeroes the workspace tile and its counter.
  void* out_bf16 = nullptr;
  int64_t ldo = 0;
  int* tile_counters = nullptr;       // >= ceil(M/128)*ceil(N/BLOCK_N) zero-initialised ints
  bool accumulate_out = false;
  // split-K WITHOUT atomics (preferred when given and large enough): every CTA stores its fp32 partial tile into
  // partials[split][tile][128 x BLOCK_N] with plain coalesced stores and a second small kernel sums the splits into
  // out_bf16.  fp32 reductions into L2 retire at ~12 requests/ns GPU-wide (profiles/prof_wgrad.ncu.txt: SMs active
  // 36 % of the kernel, the rest is the reduction queue draining); plain stores do not queue.
  float* partials = nullptr;
  int64_t partials_elems = 0;
  int device = -1;                    // CUDA device ordinal of the operands (binds the context)
  // D = A*B + add_src (bf16 [M, N], row pitch ld_add): fuses the gradient accumulation of a tensor with two
  // consumers (residual branch + 1x1 conv) into the dgrad epilogue.  Persistent kernel only.
  const void* add_src = nullptr;
  int64_t ld_add = 0;
  BnBwdFuse bn;                        // bn.x != nullptr enables the fused BN-backward reduction (row pitch = ldd)
  // fused "GEMM -> peer ship" (EPI 0 only): D may be a PEER GPU's buffer (NVLink-mapped address);
  // every CTA TMA-stores its tile there, and the CTA that completes last publishes
  // *ship_flag (peer address) = seq with release/system semantics -- no separate copy kernel.
  void* ship_flag = nullptr;
  const void* ship_seq_ptr = nullptr;  // device uint32 (graph-replayable) or nullptr -> ship_seq_imm
  uint32_t ship_seq_imm = 0;
  void* ship_done = nullptr;           // local zero-initialised uint32 counter (re-armed by the kernel)
};

// Returns nullptr on success, else a static error string.
const char* gemm_bf16(const GemmArgs& args, cudaStream_t stream);

// out[m0 + r][col(c)] (+)= sum over splits of partials[z][tile][r][c]; tiles are [rows x cols] fp32, tile t covers output
// rows (t / tiles_n) * rows.. and -- taps == 1 -- columns (t % tiles_n) * cols..; taps == 3 (3x3 wgrad): t = tile * 3 + r,
// column c of the tile is tap s = c / (cols / 3) of filter row r: output column (r * 3 + s) * cin + n0 + c % (cols / 3).
void splitk_reduce(const float* partials, int split, int ctas, int tiles_n, int rows, int cols, int taps, int cin,
                   int M, int N, void* out_bf16, int64_t ldo, bool accumulate, cudaStream_t stream);

// 3x3 / stride 1 / pad 1 NHWC convolution as an implicit GEMM (conv3x3.cu): nine shifted TMA tile loads
// per input-channel block (out-of-bounds rows/columns are zero-filled by the TMA unit = the padding)
// accumulate into one TMEM tile.  `dgrad` computes the data gradient with the same kernel (mirrored
// taps, weight tile read MN-major).
struct Conv3x3Args {
  const void* X = nullptr;   // bf16 NHWC [N, H, W, Cx]  (fwd: input; dgrad: dY)
  const void* Wt = nullptr;  // bf16 KRSC [Cout, 3, 3, Cin]
  void* Y = nullptr;         // bf16 NHWC [N, H, W, Cy]  (fwd: output; dgrad: dX)
  int N = 0, H = 0, W = 0, Cin = 0, Cout = 0;
  bool dgrad = false;
  float* col_stats = nullptr;  // fwd only: [2*Cout] += per-channel sum / sum of squares of Y
  BnBwdFuse bn;                // dgrad only: fused BN-backward reduction over Y (= dX)
  // inference epilogue (fprop, persistent kernel): Y = relu?(conv * col_scale[c] + col_shift[c]) (folded BN)
  const float* col_scale = nullptr;
  const float* col_shift = nullptr;
  bool relu = false;
  int groups = 1;              // grouped fprop: Cin/groups % 64 == 0, Cout/groups == 64 or % 128 == 0; Wt is
                               // [Cout, 3, 3, Cin/groups]
  int stride = 1;              // 1 or 2 (fprop on the persistent kernel only).  N, H, W are the OUTPUT geometry; with
                               // stride 2 the input X is [N, 2H, 2W, Cin] and is sampled by the TMA traversal stride
  int device = -1;
};
bool conv3x3_supported(int N, int H, int W, int Cin, int Cout, bool dgrad, int groups = 1);
const char* conv3x3_bf16(const Conv3x3Args& args, cudaStream_t stream);
// input gradient of the stride-2 convolution: args.X = dY [N, H, W, Cout], args.Y = dX [N, 2H, 2W, Cin], N/H/W of dY
bool conv3x3_dgrad_s2_supported(int N, int Ho, int Wo, int Cin, int Cout);
const char* conv3x3_dgrad_s2_bf16(const Conv3x3Args& args, cudaStream_t stream);

// Weight gradient of the same convolution (conv3x3_wgrad.cu): dW[Cout,3,3,Cin] (+)= sum_pixels dY x shifted X, pixel
// reduction split over `split_k` CTAs with the fused fp32 -> bf16 finalize of the 1x1 wgrad GEMM.
struct Conv3x3WgradArgs {
  const void* X = nullptr;    // bf16 NHWC [N, H, W, Cin]   (the convolution's input)
  const void* dY = nullptr;   // bf16 NHWC [N, H, W, Cout]
  void* dW = nullptr;         // bf16 KRSC [Cout, 3, 3, Cin] (e.g. a window of the flat gradient bucket)
  float* ws = nullptr;        // fp32 [Cout * 9 * Cin] all-zero workspace (left all-zero)
  int* counters = nullptr;    // >= conv3x3_wgrad_tiles(Cin, Cout) zero ints (left zero)
  int N = 0, H = 0, W = 0, Cin = 0, Cout = 0;   // H, W: OUTPUT (= dY) size; the input is stride times as large
  int split_k = 1;
  bool accumulate = false;    // dW += result instead of dW = result
  int device = -1;
  int stride = 1;             // 2: X is [N, 2H, 2W, Cin], read through the TMA traversal stride (version-1 kernel)
  float* partials = nullptr;  // split-K without atomics (see GemmArgs::partials): [split][ctas][128 x 3*BLOCK_N]
  int64_t partials_elems = 0;
};
bool conv3x3_wgrad_supported(int N, int H, int W, int Cin, int Cout);
bool conv3x3_wgrad_s2_supported(int N, int Ho, int Wo, int Cin, int Cout);
int conv3x3_wgrad_s2_kblocks(int N, int Ho, int Wo);
int conv3x3_wgrad_tiles(int Cin, int Cout);
int conv3x3_wgrad_ctas(int Cin, int Cout);        // (tile, filter row) CTAs per K split of the active kernel version
void set_wgrad3_version(int version, int base_offset_mode);   // 1 = three X loads per block, 2 = one haloed X load (default)
int get_wgrad3_version();
int conv3x3_wgrad_kblocks(int N, int H, int W);   // pixel blocks of the reduction (upper bound for split_k)
void conv3x3_wgrad_plan(int N, int H, int W, int* bh, int* nb, int* kb);   // pixel box {W, bh rows, nb images}, kb = W*bh*nb
const char* conv3x3_wgrad_bf16(const Conv3x3WgradArgs& args, cudaStream_t stream);

// Persistent variants (gemm_persist.cu): one CTA per SM streams tiles, double-buffered TMEM accumulators,
// epilogue overlapped with the next tile's MMAs.  Used by gemm_bf16 / conv3x3_bf16 when enabled (default).
void set_bnr_mode(int mode);   // fused BN-backward reduction: 1 = shuffle transpose in registers, 2 = column loop over staged tiles
int get_bnr_mode();
void set_persistent_gemm(bool on);
bool persistent_gemm_enabled();
// CTA-pair (cta_group::2) 256 x 256 tiles for the inference GEMMs (gemm_2cta.cu; default on, EDL_GEMM_PAIR=0)
bool gemm_pair_supported(const GemmArgs& args);
const char* gemm_bf16_pair(const GemmArgs& args, cudaStream_t stream);
void set_pair_gemm(bool on);
bool get_pair_gemm();
void set_persist_trace(long long* buf);   // debug: int64 [3 * 16 * 8] device buffer for clock64() phase stamps of CTA 0, or nullptr
void set_tc_stats(bool on);          // forward BN statistics as tensor-core Gram / ones products of the staged tile (EDL_TC_STATS)
void set_epilogue_warps(int n);      // 8 or 16 epilogue warps for the 128 / 256 column persistent kernels (EDL_EPI_WARPS)
bool conv3x3_halo_plan(int N, int H, int W, int* BH, int* BN, int* tiles_h, int* tiles_img);   // host only
void set_conv_halo(bool on);         // haloed A tiles for the 3x3 / stride 1 fprop and dgrad (default on; EDL_CONV_HALO=0)
bool get_conv_halo();
void set_conv_resident_weights(bool on);   // 64-channel layers / groups: weights resident across a CTA's run of tiles (EDL_CONV_BRES=0)
void set_wide_gemm_tiles(bool on);   // 128 x 256 tiles for large-N inference GEMMs (default on; EDL_GEMM_WIDE=0)
const char* gemm_bf16_persistent(const GemmArgs& args, cudaStream_t stream);
const char* conv3x3_dgrad_s2_persistent(const Conv3x3Args& args, int BH, int BN, int tiles_h, int tiles_img,
                                        cudaStream_t stream);
const char* conv3x3_bf16_persistent(const Conv3x3Args& args, int BH, int BN, int tiles_h, int tiles_img,
                                    cudaStream_t stream);

// internal: N-d bf16 tensor-map encoder shared by the TMA kernels (defined in gemm.cu)
const char* encode_tmap_bf16(void* out, const void* ptr, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box);
const char* encode_tmap(void* out, const void* ptr, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, int elem_bytes);
const char* encode_tmap_strided(void* out, const void* ptr, int rank, const uint64_t* dims,
                                const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                                int elem_bytes);

// FP8 (e4m3) inference GEMM (gemm_fp8.cu): D[M,N] bf16 = relu?((A[M,K] * B[N,K]^T) * col_scale[n] + col_shift[n])
// A, B e4m3 bytes, K-major; col_scale carries act_scale * weight_scale[n] (* folded BN scale).
struct GemmFp8Args {
  const void* A = nullptr;
  const void* B = nullptr;
  void* D = nullptr;
  int M = 0, N = 0, K = 0;
  int64_t lda = 0, ldb = 0, ldd = 0;   // row pitches in elements
  const float* col_scale = nullptr;
  const float* col_shift = nullptr;
  bool relu = false;
  int device = -1;
};
const char* gemm_fp8(const GemmFp8Args& args, cudaStream_t stream);

// bf16 -> e4m3 with a per-tensor scale (q = sat(x / scale)); amax_out (optional) receives max|x| for
// calibration / delayed scaling.
void quantize_e4m3(const void* x_bf16, void* q, int64_t n, const float* scale, float* amax_out, cudaStream_t s);

}  // namespace edl
