// torch bindings for the tcgen05 GEMM (gemm.cu).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "gemm.h"

namespace {
using torch::Tensor;

template <typename T>
inline T* optp(const c10::optional<Tensor>& t) {
  return t.has_value() && t->defined() ? reinterpret_cast<T*>(t->data_ptr()) : nullptr;
}

// bn = [x, y | None, mean, rstd, gamma, beta, dsums] (+ relu flag): see gemm.h BnBwdFuse
edl::BnBwdFuse parse_bn(const c10::optional<std::vector<c10::optional<Tensor>>>& bn, bool relu, int64_t numel,
                        int64_t channels) {
  edl::BnBwdFuse f;
  if (!bn.has_value()) return f;
  const auto& v = *bn;
  TORCH_CHECK(v.size() == 7, "bn fuse list needs 7 entries");
  TORCH_CHECK(v[0].has_value() && v[0]->scalar_type() == at::kBFloat16 && v[0]->numel() == numel);
  f.x = v[0]->data_ptr();
  if (v[1].has_value() && v[1]->defined()) {
    TORCH_CHECK(v[1]->scalar_type() == at::kBFloat16 && v[1]->numel() == numel);
    f.y = v[1]->data_ptr();
  }
  for (int i = 2; i < 7; ++i)
    TORCH_CHECK(v[i].has_value() && v[i]->scalar_type() == at::kFloat && v[i]->is_contiguous() &&
                v[i]->numel() >= (i == 6 ? 2 : 1) * channels, "bn fuse tensor ", i);
  f.mean = v[2]->data_ptr<float>();
  f.rstd = v[3]->data_ptr<float>();
  f.gamma = v[4]->data_ptr<float>();
  f.beta = v[5]->data_ptr<float>();
  f.dsums = v[6]->data_ptr<float>();
  f.relu = relu;
  return f;
}

// D[M,N] = A * B with the operand layouts described in gemm.h.  A, B, D are 2-D bf16 tensors whose
// last dimension is contiguous (row pitch = stride(0)).
void gemm_bf16(const Tensor& A, const Tensor& B, const c10::optional<Tensor>& D, bool a_mn_major,
               bool b_mn_major, const c10::optional<Tensor>& col_scale,
               const c10::optional<Tensor>& col_shift, bool relu,
               const c10::optional<Tensor>& col_stats, const c10::optional<Tensor>& out_f32,
               int64_t split_k, const c10::optional<Tensor>& out_bf16,
               const c10::optional<Tensor>& tile_counters, bool accumulate_out,
               const c10::optional<Tensor>& add_src,
               const c10::optional<std::vector<c10::optional<Tensor>>>& bn, bool bn_relu,
               const c10::optional<Tensor>& partials) {
  TORCH_CHECK(A.is_cuda() && B.is_cuda() && A.dim() == 2 && B.dim() == 2);
  TORCH_CHECK(A.scalar_type() == at::kBFloat16 && B.scalar_type() == at::kBFloat16);
  TORCH_CHECK(A.stride(1) == 1 && B.stride(1) == 1, "operands need a contiguous last dim");
  edl::GemmArgs g;
  g.A = A.data_ptr();
  g.B = B.data_ptr();
  g.a_mn_major = a_mn_major;
  g.b_mn_major = b_mn_major;
  g.M = a_mn_major ? A.size(1) : A.size(0);
  g.K = a_mn_major ? A.size(0) : A.size(1);
  g.N = b_mn_major ? B.size(1) : B.size(0);
  const int64_t kb = b_mn_major ? B.size(0) : B.size(1);
  TORCH_CHECK(kb == g.K, "K mismatch: ", kb, " vs ", g.K);
  g.lda = A.stride(0);
  g.ldb = B.stride(0);
  TORCH_CHECK(g.lda % 8 == 0 && g.ldb % 8 == 0, "row pitch must be a multiple of 16 bytes");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(g.A) % 16 == 0 && reinterpret_cast<uintptr_t>(g.B) % 16 == 0);
  if (out_f32.has_value() && out_f32->defined()) {
    TORCH_CHECK(out_f32->scalar_type() == at::kFloat && out_f32->is_contiguous());
    TORCH_CHECK(out_f32->numel() >= (int64_t)g.M * g.N);
    g.out_f32 = out_f32->data_ptr<float>();
    g.split_k = (int)split_k;
    if (out_bf16.has_value() && out_bf16->defined()) {
      TORCH_CHECK(out_bf16->scalar_type() == at::kBFloat16 && out_bf16->dim() == 2);
      TORCH_CHECK(out_bf16->size(0) == g.M && out_bf16->size(1) == g.N && out_bf16->stride(1) == 1);
      TORCH_CHECK(g.N % 4 == 0 && out_bf16->stride(0) % 4 == 0, "fused finalize needs N % 4 == 0");
      TORCH_CHECK(tile_counters.has_value() && tile_counters->scalar_type() == at::kInt);
      TORCH_CHECK(tile_counters->numel() >= (int64_t)((g.M + 127) / 128) * ((g.N + 63) / 64));
      g.out_bf16 = out_bf16->data_ptr();
      g.ldo = out_bf16->stride(0);
      g.tile_counters = tile_counters->data_ptr<int>();
      g.accumulate_out = accumulate_out;
    }
  } else {
    TORCH_CHECK(D.has_value() && D->defined() && D->scalar_type() == at::kBFloat16);
    TORCH_CHECK(D->dim() == 2 && D->size(0) == g.M && D->size(1) == g.N && D->stride(1) == 1);
    g.D = D->data_ptr();
    g.ldd = D->stride(0);
    TORCH_CHECK(g.ldd % 8 == 0 && reinterpret_cast<uintptr_t>(g.D) % 16 == 0);
  }
  g.col_scale = optp<float>(col_scale);
  g.col_shift = optp<float>(col_shift);
  g.relu = relu;
  g.col_stats = optp<float>(col_stats);
  if (add_src.has_value() && add_src->defined()) {
    TORCH_CHECK(add_src->scalar_type() == at::kBFloat16 && add_src->dim() == 2 && add_src->stride(1) == 1);
    TORCH_CHECK(add_src->size(0) == g.M && add_src->size(1) == g.N);
    g.add_src = add_src->data_ptr();
    g.ld_add = add_src->stride(0);
  }
  if (bn.has_value()) {
    TORCH_CHECK(g.D != nullptr && g.ldd == g.N, "fused BN reduction needs a dense bf16 output");
    g.bn = parse_bn(bn, bn_relu, (int64_t)g.M * g.N, g.N);
  }
  if (partials.has_value() && partials->defined()) {
    TORCH_CHECK(partials->scalar_type() == at::kFloat && partials->is_contiguous() &&
                reinterpret_cast<uintptr_t>(partials->data_ptr()) % 16 == 0);
    g.partials = partials->data_ptr<float>();
    g.partials_elems = partials->numel();
  }
  g.device = A.device().index();
  c10::cuda::CUDAGuard guard(A.device());
  const char* err = edl::gemm_bf16(g, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(err == nullptr, "edl gemm_bf16 failed: ", err);
}

// Fused "linear -> peer ship": D = A[M,K] * B[N,K]^T (+ bias) TMA-stored straight into a peer GPU's
// buffer at `d_ptr` (row pitch ldd elements); the last CTA releases `flag_ptr` (peer) = seq.
void gemm_bf16_ship(const Tensor& A, const Tensor& B, int64_t d_ptr, int64_t ldd,
                    const c10::optional<Tensor>& col_shift, int64_t flag_ptr,
                    const c10::optional<Tensor>& seq, int64_t seq_imm, const Tensor& done) {
  TORCH_CHECK(A.is_cuda() && B.is_cuda() && A.dim() == 2 && B.dim() == 2);
  TORCH_CHECK(A.scalar_type() == at::kBFloat16 && B.scalar_type() == at::kBFloat16);
  TORCH_CHECK(A.stride(1) == 1 && B.stride(1) == 1 && A.size(1) == B.size(1));
  TORCH_CHECK(done.is_cuda() && done.scalar_type() == at::kInt && done.numel() >= 1);
  edl::GemmArgs g;
  g.A = A.data_ptr();
  g.B = B.data_ptr();
  g.M = A.size(0);
  g.K = A.size(1);
  g.N = B.size(0);
  g.lda = A.stride(0);
  g.ldb = B.stride(0);
  g.D = reinterpret_cast<void*>(d_ptr);
  g.ldd = ldd;
  TORCH_CHECK(g.lda % 8 == 0 && g.ldb % 8 == 0 && g.ldd % 8 == 0 && d_ptr % 16 == 0,
              "row pitches / base must be 16-byte aligned");
  g.col_shift = optp<float>(col_shift);
  g.ship_flag = reinterpret_cast<void*>(flag_ptr);
  g.ship_seq_ptr = seq.has_value() && seq->defined() ? seq->data_ptr() : nullptr;
  g.ship_seq_imm = (uint32_t)seq_imm;
  g.ship_done = done.data_ptr();
  g.device = A.device().index();
  c10::cuda::CUDAGuard guard(A.device());
  const char* err = edl::gemm_bf16(g, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(err == nullptr, "edl gemm_bf16_ship failed: ", err);
}

// 3x3 / stride 1 / pad 1 convolution (conv3x3.cu).  x, y: logical NCHW tensors in channels_last memory
// (= NHWC); w: KRSC [Cout, 3, 3, Cin].  dgrad: x is dY [N, Cout, H, W], y is dX [N, Cin, H, W].
void conv3x3(const Tensor& x, const Tensor& w, Tensor& y, bool dgrad, const c10::optional<Tensor>& col_stats,
             const c10::optional<std::vector<c10::optional<Tensor>>>& bn, bool bn_relu) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 4 && y.dim() == 4 && w.dim() == 4);
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && y.scalar_type() == at::kBFloat16 &&
              w.scalar_type() == at::kBFloat16);
  TORCH_CHECK(x.is_contiguous(at::MemoryFormat::ChannelsLast) && y.is_contiguous(at::MemoryFormat::ChannelsLast),
              "conv3x3 needs channels_last activations");
  TORCH_CHECK(w.is_contiguous() && w.size(1) == 3 && w.size(2) == 3, "weight must be KRSC [Cout,3,3,Cin]");
  edl::Conv3x3Args a;
  a.X = x.data_ptr();
  a.Wt = w.data_ptr();
  a.Y = y.data_ptr();
  a.N = x.size(0);
  a.H = x.size(2);
  a.W = x.size(3);
  a.Cout = w.size(0);
  a.Cin = w.size(3);
  a.dgrad = dgrad;
  TORCH_CHECK(x.size(1) == (dgrad ? a.Cout : a.Cin) && y.size(1) == (dgrad ? a.Cin : a.Cout));
  TORCH_CHECK(y.size(0) == a.N && y.size(2) == a.H && y.size(3) == a.W);
  a.col_stats = optp<float>(col_stats);
  if (bn.has_value()) {
    TORCH_CHECK(dgrad, "fused BN reduction belongs to the dgrad launch");
    a.bn = parse_bn(bn, bn_relu, y.numel(), y.size(1));
  }
  a.device = x.device().index();
  c10::cuda::CUDAGuard guard(x.device());
  const char* err = edl::conv3x3_bf16(a, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(err == nullptr, "edl conv3x3 failed: ", err);
}

// dx [N,Cin,2H,2W] = input gradient of the 3x3 / pad 1 / stride 2 convolution for dy [N,Cout,H,W]; w KRSC [Cout,3,3,Cin]
void conv3x3_dgrad_s2(const Tensor& dy, const Tensor& w, Tensor& dx) {
  TORCH_CHECK(dy.is_cuda() && dy.dim() == 4 && dx.dim() == 4 && w.dim() == 4);
  TORCH_CHECK(dy.scalar_type() == at::kBFloat16 && dx.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16);
  TORCH_CHECK(dy.is_contiguous(at::MemoryFormat::ChannelsLast) && dx.is_contiguous(at::MemoryFormat::ChannelsLast));
  TORCH_CHECK(w.is_contiguous() && w.size(1) == 3 && w.size(2) == 3, "weight must be KRSC [Cout,3,3,Cin]");
  edl::Conv3x3Args a;
  a.X = dy.data_ptr();
  a.Wt = w.data_ptr();
  a.Y = dx.data_ptr();
  a.N = dy.size(0);
  a.H = dy.size(2);
  a.W = dy.size(3);
  a.Cout = w.size(0);
  a.Cin = w.size(3);
  a.dgrad = true;
  TORCH_CHECK(dy.size(1) == a.Cout && dx.size(1) == a.Cin && dx.size(0) == a.N && dx.size(2) == 2 * a.H &&
              dx.size(3) == 2 * a.W, "conv3x3_dgrad_s2: shapes do not belong together");
  a.device = dy.device().index();
  c10::cuda::CUDAGuard guard(dy.device());
  const char* err = edl::conv3x3_dgrad_s2_bf16(a, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(err == nullptr, "edl conv3x3_dgrad_s2 failed: ", err);
}
bool conv3x3_dgrad_s2_supported(int64_t n, int64_t ho, int64_t wo, int64_t cin, int64_t cout) {
  return edl::conv3x3_dgrad_s2_supported((int)n, (int)ho, (int)wo, (int)cin, (int)cout);
}

// D bf16 [M,N] = relu?((A8[M,K] * B8[N,K]^T) * col_scale + col_shift); A8/B8 are e4m3 bytes (uint8 / float8 tensors)
void gemm_fp8(const Tensor& A, const Tensor& B, Tensor& D, const c10::optional<Tensor>& col_scale,
              const c10::optional<Tensor>& col_shift, bool relu) {
  TORCH_CHECK(A.is_cuda() && B.is_cuda() && A.dim() == 2 && B.dim() == 2 && D.dim() == 2);
  TORCH_CHECK(A.element_size() == 1 && B.element_size() == 1 && D.scalar_type() == at::kBFloat16);
  TORCH_CHECK(A.stride(1) == 1 && B.stride(1) == 1 && D.stride(1) == 1 && A.size(1) == B.size(1));
  TORCH_CHECK(D.size(0) == A.size(0) && D.size(1) == B.size(0));
  edl::GemmFp8Args g;
  g.A = A.data_ptr();
  g.B = B.data_ptr();
  g.D = D.data_ptr();
  g.M = A.size(0);
  g.K = A.size(1);
  g.N = B.size(0);
  g.lda = A.stride(0);
  g.ldb = B.stride(0);
  g.ldd = D.stride(0);
  g.col_scale = optp<float>(col_scale);
  g.col_shift = optp<float>(col_shift);
  g.relu = relu;
  g.device = A.device().index();
  c10::cuda::CUDAGuard guard(A.device());
  const char* err = edl::gemm_fp8(g, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(err == nullptr, "edl gemm_fp8 failed: ", err);
}

void quantize_e4m3(const Tensor& x, Tensor& q, const Tensor& scale, const c10::optional<Tensor>& amax) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.is_contiguous());
  TORCH_CHECK(q.is_cuda() && q.element_size() == 1 && q.is_contiguous() && q.numel() == x.numel());
  TORCH_CHECK(scale.scalar_type() == at::kFloat && scale.numel() >= 1);
  c10::cuda::CUDAGuard guard(x.device());
  edl::quantize_e4m3(x.data_ptr(), q.data_ptr(), x.numel(), scale.data_ptr<float>(), optp<float>(amax),
                     at::cuda::getCurrentCUDAStream().stream());
}

bool conv3x3_supported(int64_t n, int64_t h, int64_t w, int64_t cin, int64_t cout, bool dgrad, int64_t groups) {
  return edl::conv3x3_supported((int)n, (int)h, (int)w, (int)cin, (int)cout, dgrad, (int)groups);
}

// dW (+)= wgrad of the 3x3 / stride 1 / pad 1 conv.  x [N,Cin,H,W], dy [N,Cout,H,W] channels_last; dw KRSC
// [Cout,3,3,Cin] bf16 contiguous; ws fp32 zeros (>= dw.numel()), counters int32 zeros.
void conv3x3_wgrad(const Tensor& x, const Tensor& dy, Tensor& dw, Tensor& ws, Tensor& counters, int64_t split_k,
                   bool accumulate, const c10::optional<Tensor>& partials) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 4 && dy.dim() == 4 && dw.dim() == 4);
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && dy.scalar_type() == at::kBFloat16 && dw.scalar_type() == at::kBFloat16);
  TORCH_CHECK(x.is_contiguous(at::MemoryFormat::ChannelsLast) && dy.is_contiguous(at::MemoryFormat::ChannelsLast),
              "conv3x3_wgrad needs channels_last activations");
  TORCH_CHECK(dw.is_contiguous() && dw.size(1) == 3 && dw.size(2) == 3, "dw must be KRSC [Cout,3,3,Cin]");
  TORCH_CHECK(ws.scalar_type() == at::kFloat && ws.is_contiguous() && ws.numel() >= dw.numel());
  edl::Conv3x3WgradArgs a;
  a.X = x.data_ptr();
  a.dY = dy.data_ptr();
  a.dW = dw.data_ptr();
  a.ws = ws.data_ptr<float>();
  a.N = x.size(0);
  a.H = dy.size(2);
  a.W = dy.size(3);
  a.Cin = x.size(1);
  a.Cout = dy.size(1);
  // stride 1: x and dy have the same size; stride 2: x is twice as large in both directions
  a.stride = (x.size(2) == 2 * a.H && x.size(3) == 2 * a.W && a.H != x.size(2)) ? 2 : 1;
  TORCH_CHECK(dy.size(0) == a.N && x.size(2) == a.stride * a.H && x.size(3) == a.stride * a.W &&
              dw.size(0) == a.Cout && dw.size(3) == a.Cin, "conv3x3_wgrad: x / dy / dw shapes do not belong together");
  TORCH_CHECK(counters.scalar_type() == at::kInt && counters.is_contiguous() &&
              counters.numel() >= edl::conv3x3_wgrad_tiles(a.Cin, a.Cout));
  TORCH_CHECK(reinterpret_cast<uintptr_t>(a.dW) % 8 == 0, "dw window must be 8-byte aligned");
  a.counters = counters.data_ptr<int>();
  a.split_k = (int)split_k;
  a.accumulate = accumulate;
  if (partials.has_value() && partials->defined()) {
    TORCH_CHECK(partials->scalar_type() == at::kFloat && partials->is_contiguous() &&
                reinterpret_cast<uintptr_t>(partials->data_ptr()) % 16 == 0);
    a.partials = partials->data_ptr<float>();
    a.partials_elems = partials->numel();
  }
  a.device = x.device().index();
  c10::cuda::CUDAGuard guard(x.device());
  const char* err = edl::conv3x3_wgrad_bf16(a, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(err == nullptr, "edl conv3x3_wgrad failed: ", err);
}

// 3x3 / pad 1 / STRIDE 2 fprop (dense or grouped) on the persistent kernel: x [N, Cin, 2H, 2W], y [N, Cout, H, W]
// channels_last.  Training use: col_stats (BN statistics of y); inference use: col_scale / col_shift / relu.
void conv3x3_s2(const Tensor& x, const Tensor& w, Tensor& y, const c10::optional<Tensor>& col_stats,
                const c10::optional<Tensor>& col_scale, const c10::optional<Tensor>& col_shift, bool relu,
                int64_t groups) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 4 && y.dim() == 4 && w.dim() == 4);
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && y.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16);
  TORCH_CHECK(x.is_contiguous(at::MemoryFormat::ChannelsLast) && y.is_contiguous(at::MemoryFormat::ChannelsLast));
  TORCH_CHECK(w.is_contiguous() && w.size(1) == 3 && w.size(2) == 3);
  edl::Conv3x3Args a;
  a.X = x.data_ptr();
  a.Wt = w.data_ptr();
  a.Y = y.data_ptr();
  a.N = y.size(0);
  a.H = y.size(2);
  a.W = y.size(3);
  a.Cin = x.size(1);
  a.Cout = w.size(0);
  a.groups = (int)groups;
  a.stride = 2;
  TORCH_CHECK(x.size(0) == a.N && x.size(2) == 2 * a.H && x.size(3) == 2 * a.W, "stride-2 conv needs an even input: x is 2H x 2W");
  TORCH_CHECK(w.size(3) * groups == a.Cin && y.size(1) == a.Cout);
  a.col_stats = optp<float>(col_stats);
  a.col_scale = optp<float>(col_scale);
  a.col_shift = optp<float>(col_shift);
  a.relu = relu;
  a.device = x.device().index();
  c10::cuda::CUDAGuard guard(x.device());
  const char* err = edl::conv3x3_bf16(a, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(err == nullptr, "edl conv3x3_s2 failed: ", err);
}

std::vector<int64_t> conv3x3_wgrad_plan(int64_t n, int64_t h, int64_t w) {
  int bh = 0, nb = 0, kb = 0;
  edl::conv3x3_wgrad_plan((int)n, (int)h, (int)w, &bh, &nb, &kb);
  // box width: W for version 1; for version 2 the box starts at column -1 and spans kb / (bh * nb) >= W + 1 columns
  const int wb = (bh > 0 && nb > 0) ? kb / (bh * nb) : 0;
  return {bh, nb, kb, wb};
}

bool conv3x3_wgrad_supported(int64_t n, int64_t h, int64_t w, int64_t cin, int64_t cout) {
  return edl::conv3x3_wgrad_supported((int)n, (int)h, (int)w, (int)cin, (int)cout);
}
bool conv3x3_wgrad_s2_supported(int64_t n, int64_t ho, int64_t wo, int64_t cin, int64_t cout) {
  return edl::conv3x3_wgrad_s2_supported((int)n, (int)ho, (int)wo, (int)cin, (int)cout);
}

// inference 3x3 conv (optionally grouped) with the folded-BN scale / shift / ReLU epilogue
void conv3x3_infer(const Tensor& x, const Tensor& w, Tensor& y, const c10::optional<Tensor>& col_scale,
                   const c10::optional<Tensor>& col_shift, bool relu, int64_t groups) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 4 && y.dim() == 4 && w.dim() == 4);
  TORCH_CHECK(x.scalar_type() == at::kBFloat16 && y.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16);
  TORCH_CHECK(x.is_contiguous(at::MemoryFormat::ChannelsLast) && y.is_contiguous(at::MemoryFormat::ChannelsLast));
  TORCH_CHECK(w.is_contiguous() && w.size(1) == 3 && w.size(2) == 3);
  edl::Conv3x3Args a;
  a.X = x.data_ptr();
  a.Wt = w.data_ptr();
  a.Y = y.data_ptr();
  a.N = x.size(0);
  a.H = x.size(2);
  a.W = x.size(3);
  a.Cin = x.size(1);
  a.Cout = w.size(0);
  a.groups = (int)groups;
  TORCH_CHECK(w.size(3) * groups == a.Cin && y.size(1) == a.Cout && y.size(0) == a.N && y.size(2) == a.H && y.size(3) == a.W);
  a.col_scale = optp<float>(col_scale);
  a.col_shift = optp<float>(col_shift);
  a.relu = relu;
  a.device = x.device().index();
  c10::cuda::CUDAGuard guard(x.device());
  const char* err = edl::conv3x3_bf16(a, at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(err == nullptr, "edl conv3x3_infer failed: ", err);
}
}  // namespace

void register_gemm_bindings(pybind11::module_& m) {
  m.def("gemm_bf16", &gemm_bf16);
  m.def("conv3x3", &conv3x3);
  m.def("set_persistent_gemm", &edl::set_persistent_gemm);
  m.def("set_bnr_mode", &edl::set_bnr_mode);
  m.def("get_bnr_mode", &edl::get_bnr_mode);
  m.def("persistent_gemm_enabled", &edl::persistent_gemm_enabled);
  m.def("gemm_fp8", &gemm_fp8);
  m.def("quantize_e4m3", &quantize_e4m3);
  m.def("conv3x3_supported", &conv3x3_supported);
  m.def("conv3x3_infer", &conv3x3_infer);
  m.def("gemm_bf16_ship", &gemm_bf16_ship);
  m.def("conv3x3_dgrad_s2", &conv3x3_dgrad_s2);
  m.def("conv3x3_dgrad_s2_supported", &conv3x3_dgrad_s2_supported);
  m.def("conv3x3_wgrad", &conv3x3_wgrad);
  m.def("conv3x3_wgrad_supported", &conv3x3_wgrad_supported);
  m.def("conv3x3_wgrad_s2_supported", &conv3x3_wgrad_s2_supported);
  m.def("conv3x3_wgrad_s2_kblocks", &edl::conv3x3_wgrad_s2_kblocks);
  m.def("conv3x3_wgrad_tiles", &edl::conv3x3_wgrad_tiles);
  m.def("conv3x3_wgrad_kblocks", &edl::conv3x3_wgrad_kblocks);
  m.def("conv3x3_wgrad_ctas", &edl::conv3x3_wgrad_ctas);
  m.def("set_wide_gemm_tiles", &edl::set_wide_gemm_tiles);
  m.def("set_pair_gemm", &edl::set_pair_gemm);
  m.def("get_pair_gemm", &edl::get_pair_gemm);
  m.def("set_persist_trace", [](c10::optional<Tensor> t) {
    if (!t.has_value()) { edl::set_persist_trace(nullptr); return; }
    TORCH_CHECK(t->is_cuda() && t->scalar_type() == at::kLong && t->numel() >= 3 * 16 * 8 && t->is_contiguous(),
                "trace buffer: contiguous CUDA int64 tensor with at least 384 elements");
    edl::set_persist_trace(reinterpret_cast<long long*>(t->data_ptr<int64_t>()));
  });
  m.def("set_tc_stats", &edl::set_tc_stats);
  m.def("set_epilogue_warps", &edl::set_epilogue_warps);
  m.def("conv3x3_halo_plan", [](int64_t n, int64_t h, int64_t w) -> std::vector<int64_t> {
    int bh = 0, bn = 0, th = 0, ti = 0;
    if (!edl::conv3x3_halo_plan((int)n, (int)h, (int)w, &bh, &bn, &th, &ti)) return {};
    return {bh, bn, th, ti};
  });
  m.def("set_conv_halo", &edl::set_conv_halo);
  m.def("get_conv_halo", &edl::get_conv_halo);
  m.def("set_conv_resident_weights", &edl::set_conv_resident_weights);
  m.def("set_wgrad3_version", &edl::set_wgrad3_version);
  m.def("get_wgrad3_version", &edl::get_wgrad3_version);
  m.def("conv3x3_wgrad_plan", &conv3x3_wgrad_plan);
  m.def("conv3x3_s2", &conv3x3_s2);
}
