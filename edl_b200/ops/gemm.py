"""tcgen05 GEMM front-end (csrc/gemm.cu): plain GEMM, Linear and NHWC 1x1 convolution with the
train-mode BatchNorm statistics fused into the GEMM epilogue.

Reference call sites are library ops: ``fluid.layers.conv2d`` (cuDNN) and ``fluid.layers.fc``
(cuBLAS) in example/distill/resnet/models/resnet_vd.py:153-162,135-141."""
from __future__ import annotations

from typing import Optional

import torch

from .bn import _cl

_NUM_SMS = 148


def gemm_bf16(a, b, out=None, a_mn_major=False, b_mn_major=False, col_scale=None, col_shift=None,
              relu=False, col_stats=None, out_f32=None, split_k=1, out_bf16=None, tile_counters=None,
              accumulate_out=False, add=None, bn=None, partials=None):
    """D = op(A) @ op(B): see csrc/gemm.h for the operand conventions.

    default        : A [M, K], B [N, K]  -> D [M, N] = A @ B^T
    b_mn_major     : B [K, N]            -> D = A @ B
    a_mn_major     : A [K, M]            -> D = A^T @ op(B)
    out_f32 given  : fp32 [M, N] += result, split over K across CTAs (no bf16 output).
    out_bf16 given : (with out_f32 as an all-zero workspace) the last CTA of every output tile writes
                     bf16(result) (+= if accumulate_out) into out_bf16 and re-zeroes the workspace.
    add given      : D = result + add (bf16 [M, N]) in the epilogue (gradient accumulation fused into dgrad).
    bn given       : a ``BNBackwardHook`` (ops/bn.py): D is the gradient of that BatchNorm's output and the
                     epilogue accumulates its backward reduction (sum dy, sum dy * xhat) into hook.dsums."""
    from . import native, count_launch

    m = a.shape[1] if a_mn_major else a.shape[0]
    n = b.shape[1] if b_mn_major else b.shape[0]
    if not a.is_cuda or not _tma_compatible(a, b, out, n):
        # CPU tensors, or CUDA shapes TMA cannot describe (row pitch / base not 16-byte aligned):
        # plain PyTorch math with identical semantics -- counted and logged when it happens on a GPU
        if a.is_cuda:
            from . import count_fallback
            count_fallback("gemm_bf16: operand not TMA-describable (pitch / base not 16-byte aligned or not bf16): "
                           "fp32 matmul on the library, %s x %s" % (tuple(a.shape), tuple(b.shape)))
        af = a.float().t() if a_mn_major else a.float()
        bf = b.float() if b_mn_major else b.float().t()
        d = af @ bf
        if add is not None:
            d = d + add.float()
        if out_bf16 is not None:
            out_bf16.copy_((out_bf16.float() + d) if accumulate_out else d)
            return out_bf16
        if out_f32 is not None:
            out_f32.add_(d)
            return out_f32
        if col_scale is not None:
            d = d * col_scale.float()
        if col_shift is not None:
            d = d + col_shift.float()
        if relu:
            d = torch.relu(d)
        d = d.to(torch.bfloat16)
        if col_stats is not None:
            col_stats[:n] += d.float().sum(0)
            col_stats[n:] += (d.float() ** 2).sum(0)
        if out is not None:
            out.copy_(d)
            return out
        return d
    if out_f32 is None and out is None:
        out = torch.empty((m, n), device=a.device, dtype=torch.bfloat16)
    if add is not None and not (native().persistent_gemm_enabled() and n % 8 == 0 and add.stride(-1) == 1
                                and add.stride(0) % 8 == 0 and add.data_ptr() % 16 == 0 and out_f32 is None):
        # the fused addend lives in the persistent kernel's epilogue; otherwise one extra elementwise pass
        native().gemm_bf16(a, b, out, a_mn_major, b_mn_major, col_scale, col_shift, relu, col_stats,
                           out_f32, int(split_k), out_bf16, tile_counters, bool(accumulate_out), None, None, False,
                           partials)
        count_launch()
        out.add_(add)
        return out
    bn_list, bn_relu = (bn.as_list(m, n), bn.relu) if bn is not None else (None, False)
    native().gemm_bf16(a, b, out, a_mn_major, b_mn_major, col_scale, col_shift, relu, col_stats,
                       out_f32, int(split_k), out_bf16, tile_counters, bool(accumulate_out), add, bn_list, bn_relu,
                       partials)
    count_launch()
    if bn is not None:
        bn.done = True
    if out_bf16 is not None:
        return out_bf16
    return out if out_f32 is None else out_f32


def _tma_compatible(a, b, out, n):
    for t in (a, b, out):
        if t is None:
            continue
        if t.dtype != torch.bfloat16 or t.stride(-1) != 1 or t.stride(0) % 8 != 0 or t.data_ptr() % 16 != 0:
            return False
    return out is not None or n % 8 == 0


def _split_k_for(m, n, k):
    """Split-K factor of the wgrad GEMM.  fp32 atomics cost ~2 us per million element-updates at L2
    (measured, profiles/), so the split is chosen to fill ~one wave of SMs and never more: outputs
    with >= 148 tiles are not split at all (direct bf16 store), tiny outputs (stage-2 layers, 1-8
    tiles with K ~ 100k) take the full 148-way split because they have few elements to update."""
    tiles = ((m + 127) // 128) * ((n + 127) // 128 if n > 64 else 1)
    kb = (k + 63) // 64
    if tiles >= 48:
        # a third of the SMs already have a tile: no split, hence no partial tiles and no reduce kernel.  The weight
        # gradients run on the low-priority side stream and are not the critical path: a kernel that keeps 48-147 SMs
        # busy a little longer leaves the other SMs to the dgrad / BatchNorm chain, which is what limits the step
        return 1
    want = max(1, _NUM_SMS // tiles)
    return max(1, min(want, max(1, kb // 2)))


_WS = {}
_SIDE = {"stream": None, "streams": [], "next": 0, "keep": []}


def set_wgrad_stream(stream):
    """Run weight-gradient GEMMs on ``stream`` (the DP engine passes its communication stream): they
    then overlap the dgrad / BN-backward chain of the main stream, and the bucket all-reduces that
    follow on the same stream are naturally ordered behind them.  ``None`` disables the overlap."""
    if isinstance(stream, (list, tuple)):
        _SIDE["streams"] = [st for st in stream if st is not None]
        _SIDE["stream"] = _SIDE["streams"][0] if _SIDE["streams"] else None
    else:
        _SIDE["stream"] = stream
        _SIDE["streams"] = [stream] if stream is not None else []
    _SIDE["next"] = 0
    _SIDE["keep"].clear()


def _side_stream():
    """The next weight-gradient stream (round-robin over the engine's side streams): consecutive layers' weight
    gradients are independent, two of them in flight shorten the side-stream tail that remains after the last
    dgrad of backward (profiles/timeline_r2: conv1_2 wgrad -> stem wgrad -> last optimizer bucket, 190 us serial)."""
    sts = _SIDE["streams"]
    if not sts:
        return None
    st = sts[_SIDE["next"] % len(sts)]
    _SIDE["next"] += 1
    return st


def release_wgrad_keepalive():
    """Drop the references that kept side-stream operands alive (call after the streams were joined)."""
    _SIDE["keep"].clear()



def _splitk_workspace(device, numel, tiles):
    """Persistent all-zero fp32 workspace + tile counters (per device and stream): the fused
    split-K finalize leaves both zeroed again, so no per-call memset is needed."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws[0].numel() < numel or ws[1].numel() < tiles:
        ws = (torch.zeros(max(numel, 4 << 20), device=device, dtype=torch.float32),
              torch.zeros(max(tiles, 4096), device=device, dtype=torch.int32))
        _WS[key] = ws
    return ws


# TIMING EXPERIMENT ONLY (never set in a real run: the model does not learn): EDL_DEBUG_SKIP_WGRAD=1 drops every
# weight-gradient kernel of the tcgen05 / library conv paths, which bounds what faster wgrad kernels could buy.
_SKIP_WGRAD = __import__("os").environ.get("EDL_DEBUG_SKIP_WGRAD", "0") == "1"


_PART = {}
# split-K without atomics (csrc/gemm.h GemmArgs::partials): EDL_SPLITK_PARTIALS=0 goes back to fp32 reductions
SPLITK_PARTIALS = __import__("os").environ.get("EDL_SPLITK_PARTIALS", "1") == "1"


def _splitk_partials(device):
    """Per (device, stream) scratch for the split-K partial tiles: 148-296 CTAs x (128 x 384) fp32 at most."""
    if not SPLITK_PARTIALS:
        return None
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _PART.get(key)
    if buf is None:
        buf = _PART[key] = torch.empty(16 << 20, device=device, dtype=torch.float32)      # 64 MB
    return buf


def _wgrad(dy2, x2, weight_shape, sink, ready):
    """dW[Cout, Cin] = dy2[M, Cout]^T @ x2[M, Cin]: fp32 split-K accumulation whose last CTA per
    tile converts to bf16 straight into the (flat gradient bucket) sink -- no memset / cast / add
    kernels around it."""
    if _SKIP_WGRAD and sink is not None:
        if ready is not None:
            ready()
        return None
    cout, cin = dy2.shape[1], x2.shape[1]
    split = _split_k_for(cout, cin, dy2.shape[0])
    side = _side_stream() if (dy2.is_cuda and sink is not None) else None
    if side is not None:
        cur = torch.cuda.current_stream(dy2.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        side.wait_event(ev)
        _SIDE["keep"].append((dy2, x2))          # operands must outlive the side-stream kernel
        with torch.cuda.stream(side):
            return _wgrad_impl(dy2, x2, weight_shape, sink, ready, cout, cin, split)
    return _wgrad_impl(dy2, x2, weight_shape, sink, ready, cout, cin, split)


def _wgrad_impl(dy2, x2, weight_shape, sink, ready, cout, cin, split):
    if dy2.is_cuda and cin % 4 == 0:
        tiles = ((cout + 127) // 128) * ((cin + 63) // 64)
        ws, counters = _splitk_workspace(dy2.device, cout * cin, tiles)
        if sink is not None:
            out, acc = sink.view(cout, cin), True
        else:
            out, acc = torch.empty((cout, cin), device=dy2.device, dtype=torch.bfloat16), False
        gemm_bf16(dy2, x2, a_mn_major=True, b_mn_major=True, out_f32=ws[:cout * cin].view(cout, cin),
                  split_k=split, out_bf16=out, tile_counters=counters, accumulate_out=acc,
                  partials=_splitk_partials(dy2.device) if split > 1 else None)
        if sink is not None:
            if ready is not None:
                ready()
            return None
        return out.view(weight_shape)
    acc = torch.zeros((cout, cin), device=dy2.device, dtype=torch.float32)
    gemm_bf16(dy2, x2, a_mn_major=True, b_mn_major=True, out_f32=acc, split_k=split)
    if sink is not None:
        sink.view(cout, cin).add_(acc)
        if ready is not None:
            ready()
        return None
    return acc.to(torch.bfloat16).view(weight_shape)


class _Conv1x1Fn(torch.autograd.Function):
    """``fork=True`` also returns the input as a second output: a tensor that feeds this convolution AND a
    second consumer (the residual branch) then receives ONE gradient, summed inside the dgrad epilogue,
    instead of two gradients that autograd adds with an extra elementwise kernel."""

    @staticmethod
    def forward(ctx, x, w, stats, sink, ready, fork=False, bn_hook=None):
        ctx.bn_hook = bn_hook
        x = _cl(x)
        n, cin, h, wd = x.shape
        cout = w.shape[0]
        x2 = x.permute(0, 2, 3, 1).reshape(-1, cin)
        w2 = w.reshape(cout, cin)
        y = torch.empty((n, cout, h, wd), device=x.device, dtype=x.dtype,
                        memory_format=torch.channels_last)
        y2 = y.permute(0, 2, 3, 1).reshape(-1, cout)
        gemm_bf16(x2, w2, out=y2, col_stats=stats)
        ctx.save_for_backward(x, w)
        ctx.sink, ctx.ready = sink, ready
        if fork:
            return y, x.view_as(x)
        return y

    @staticmethod
    def backward(ctx, dy, dfork=None):
        x, w = ctx.saved_tensors
        if dy is None:      # only the forked alias was used downstream
            return dfork, None, None, None, None, None, None
        dy = _cl(dy)
        n, cin, h, wd = x.shape
        cout = w.shape[0]
        dy2 = dy.permute(0, 2, 3, 1).reshape(-1, cout)
        x2 = x.permute(0, 2, 3, 1).reshape(-1, cin)
        w2 = w.reshape(cout, cin)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            add = None
            if dfork is not None:
                add = _cl(dfork).permute(0, 2, 3, 1).reshape(-1, cin)
            # dx is the gradient of the BatchNorm output that fed this conv: its backward reduction
            # rides in this epilogue when the producing BN left a hook on the tensor
            hook = ctx.bn_hook if _bn_fusable(ctx.bn_hook, x, cin) else None
            gemm_bf16(dy2, w2, out=dx.permute(0, 2, 3, 1).reshape(-1, cin), b_mn_major=True, add=add, bn=hook)
        dw = _wgrad(dy2, x2, w.shape, ctx.sink, ctx.ready) if ctx.needs_input_grad[1] else None
        return dx, dw, None, None, None, None, None


def conv1x1(x, weight, stats: Optional[torch.Tensor] = None, fork: bool = False):
    """NHWC 1x1 stride-1 convolution as a tcgen05 GEMM.  ``weight``: [Cout, 1, 1, Cin] or
    [Cout, Cin] bf16.  If ``stats`` (pre-zeroed fp32 [2*Cout]) is given, the epilogue accumulates
    per-channel sum / sum-of-squares of the output for the following train-mode BatchNorm."""
    sink = getattr(weight, "_edl_grad_sink", None) if torch.is_grad_enabled() else None
    ready = getattr(weight, "_edl_grad_ready", None) if sink is not None else None
    return _Conv1x1Fn.apply(x, weight, stats, sink, ready, fork, getattr(x, "_edl_bn_hook", None))


def _bn_fusable(hook, x, channels) -> bool:
    """The consumer's dgrad may carry the BN-backward reduction iff the hook belongs to exactly this input,
    the persistent kernels are on and the layout is what the fused epilogue expects."""
    from . import native

    return (hook is not None and FUSE_BN_BWD and not hook.done and x.is_cuda and hook.x.shape == x.shape
            and channels % 8 == 0 and native().persistent_gemm_enabled()
            and x.numel() * x.element_size() <= FUSE_BN_BWD_MAX_BYTES)


# ON by default since round 2 (EDL_FUSE_BN_BWD=0 turns it off): the BatchNorm-backward reduction of a conv's input
# gradient rides in that conv's dgrad epilogue as a column-pair loop over the staged tiles (csrc/gemm_persist.cu,
# bnr mode 2).  Validated on B200 at kernel level and at model level against the stand-alone reduction kernels with
# the run-to-run atomics noise as yardstick (tests/test_persist_gpu.py); removes 44 bn_bwd_reduce launches per
# ResNet50_vd step: 4.746 -> 4.618 ms (profiles/bench_runs.json).  The round-1 shuffle version (mode 1, 46-59 us per
# short-K dgrad kernel) stays selectable with EDL_BNR_MODE=1.
FUSE_BN_BWD = __import__("os").environ.get("EDL_FUSE_BN_BWD", "1") == "1"
# Upper bound on the activation size for which the reduction is fused (EDL_FUSE_BN_BWD_MAX_MB).  The fused epilogue
# prefetches the BN input / output tiles of ONE tile ahead (64 KB per SM in flight), which is latency-bound on the
# largest activations (51 MB at 56 x 56 x 256 x 32: 63-75 us per dgrad, kineto_r2_c24 trace); the stand-alone streaming
# reduction reads them at full bandwidth.
FUSE_BN_BWD_MAX_BYTES = float(__import__("os").environ.get("EDL_FUSE_BN_BWD_MAX_MB", "1e9")) * (1 << 20)


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, relu, sink, ready, bias_sink, bias_ready):
        x = x.contiguous()
        y = gemm_bf16(x, w, col_shift=bias.float() if bias is not None else None, relu=relu)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.has_bias = bias is not None
        ctx.bias_dtype = bias.dtype if bias is not None else None
        ctx.sink, ctx.ready = sink, ready
        ctx.bias_sink, ctx.bias_ready = bias_sink, bias_ready
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = dy.contiguous()
        if y is not None:
            dy = dy * (y > 0)
        dx = gemm_bf16(dy, w, b_mn_major=True) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dy, x, w.shape, ctx.sink, ctx.ready)
        db = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum(0).to(ctx.bias_dtype)
            if ctx.bias_sink is not None:
                # straight into the flat gradient bucket: no autograd AccumulateGrad node runs for this
                # parameter (those are pinned to the stream of their first use, which breaks graph capture
                # on another stream after eager steps)
                ctx.bias_sink.add_(db)
                if ctx.bias_ready is not None:
                    ctx.bias_ready()
                db = None
        return dx, dw, db, None, None, None, None, None


def linear_bf16(x, weight, bias=None, relu=False):
    """y = act(x @ weight^T + bias) on the tcgen05 GEMM (bias/activation fused in the epilogue).
    x [M, K] bf16, weight [N, K] bf16, bias [N] (any float dtype)."""
    grad = torch.is_grad_enabled()
    sink = getattr(weight, "_edl_grad_sink", None) if grad else None
    ready = getattr(weight, "_edl_grad_ready", None) if sink is not None else None
    bsink = getattr(bias, "_edl_grad_sink", None) if (grad and bias is not None) else None
    bready = getattr(bias, "_edl_grad_ready", None) if bsink is not None else None
    return _LinearFn.apply(x, weight, bias, relu, sink, ready, bsink, bready)


# EDL_PAIR_WGRAD=0: the library's weight gradient for the 32-channel stem convolutions as well
PAIR_WGRAD = __import__("os").environ.get("EDL_PAIR_WGRAD", "1") == "1"


def _pair_wgrad_ok(x, dy, w, cfg) -> bool:
    from . import native

    if not (PAIR_WGRAD and x.is_cuda and x.dtype == torch.bfloat16 and tuple(cfg) == (1, 1, 1) and w.dim() == 4):
        return False
    cout, kh, kw, cin = w.shape
    n, c, h, wd = x.shape
    return (kh == 3 and kw == 3 and cin == 32 and c == 32 and wd % 2 == 0 and cout % 32 == 0
            and native().conv3x3_wgrad_supported(n, h, wd // 2, 64, 2 * cout))


def _count_lib_conv(x, what, w_krsc, stride, groups):
    if x.is_cuda:
        from . import count_fallback
        cout, kh, kw, cin = w_krsc.shape
        count_fallback("conv %s on cuDNN: %dx%d stride %d groups %d, %d -> %d channels" % (what, kh, kw, stride, groups,
                                                                                      cin * groups, cout))


class _ConvLibFn(torch.autograd.Function):
    """k x k / strided / grouped convolution on the library kernel (cuDNN) with the backward split in
    two: dgrad stays on the critical path, wgrad runs on the side stream next to the tcgen05 wgrad
    GEMMs and lands in the flat gradient bucket (sink) without going through ``param.grad``.
    ``w`` is stored KRSC ([Cout, kh, kw, Cin/groups]); the library sees a zero-copy NCHW-shaped view."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, groups, sink, ready):
        wv = w.permute(0, 3, 1, 2)
        _count_lib_conv(x, "fprop", w, stride, groups)
        y = torch.ops.aten.convolution(x, wv, None, [stride, stride], [padding, padding], [1, 1], False, [0, 0], groups)
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, groups)
        ctx.sink, ctx.ready = sink, ready
        return y

    @staticmethod
    def _bwd(dy, x, wv, cfg, mask):
        stride, padding, groups = cfg
        _count_lib_conv(x, "dgrad" if mask[0] else "wgrad", wv.permute(0, 2, 3, 1), stride, groups)
        return torch.ops.aten.convolution_backward(dy, x, wv, None, [stride, stride], [padding, padding], [1, 1],
                                                   False, [0, 0], groups, mask)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        wv = w.permute(0, 3, 1, 2)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _ConvLibFn._bwd(dy, x, wv, ctx.cfg, [True, False, False])[0]
        dw = None
        if ctx.needs_input_grad[1]:
            sink, ready = ctx.sink, ctx.ready
            side = _side_stream() if (dy.is_cuda and sink is not None) else None

            def wgrad():
                if _SKIP_WGRAD and sink is not None:
                    if ready is not None:
                        ready()
                    return None
                if _pair_wgrad_ok(x, dy, w, ctx.cfg):
                    # 32-input-channel 3x3 / stride 1 layer (the stem's conv1_2 / conv1_3): weight gradient in pixel-pair
                    # form on the tcgen05 kernel even when fprop / dgrad stay on the library (the library's wgrad of
                    # these two layers takes 100-110 us each and is the END of backward: profiles/timeline_r2)
                    from . import native, count_launch
                    xc, dyc = _cl(x), _cl(dy)
                    x2, dy2 = _pair_view(xc), _pair_view(dyc)
                    dw2 = conv3x3_wgrad(x2, dy2, (2 * w.shape[0], 3, 3, 64), None)
                    out = sink.view(w.shape) if sink is not None else torch.empty_like(w)
                    native().pair_weight_fold(dw2, out, sink is not None)
                    count_launch()
                    if sink is None:
                        return out
                    if ready is not None:
                        ready()
                    return None
                gw = _ConvLibFn._bwd(dy, x, wv, ctx.cfg, [False, True, False])[1].permute(0, 2, 3, 1)
                if sink is None:
                    return gw.contiguous()
                sink.view(w.shape).add_(gw)
                if ready is not None:
                    ready()
                return None

            if side is not None:
                cur = torch.cuda.current_stream(dy.device)
                ev = torch.cuda.Event()
                ev.record(cur)
                side.wait_event(ev)
                _SIDE["keep"].append((dy, x))
                with torch.cuda.stream(side):
                    dw = wgrad()
                    # temporaries of the side stream must not be recycled by main-stream allocations
                    # before the streams are joined
                    _SIDE["keep"].append(None)
            else:
                dw = wgrad()
        return dx, dw, None, None, None, None, None


def conv_lib(x, weight_krsc, stride=1, padding=0, groups=1):
    sink = getattr(weight_krsc, "_edl_grad_sink", None) if torch.is_grad_enabled() else None
    ready = getattr(weight_krsc, "_edl_grad_ready", None) if sink is not None else None
    return _ConvLibFn.apply(x, weight_krsc, stride, padding, groups, sink, ready)


def conv3x3_supported(x, weight_krsc, dgrad=False) -> bool:
    from . import native

    if not (x.is_cuda and x.dtype == torch.bfloat16 and weight_krsc.dtype == torch.bfloat16):
        return False
    n, _, h, w = x.shape
    cout, kh, kw, cin = weight_krsc.shape
    return kh == 3 and kw == 3 and native().conv3x3_supported(n, h, w, cin, cout, dgrad, 1)


class _Conv3x3Fn(torch.autograd.Function):
    """3x3 / stride 1 / pad 1 convolution on the tcgen05 implicit-GEMM kernel (csrc/conv3x3.cu): fprop
    with the BatchNorm statistics in its epilogue, dgrad on the same kernel with mirrored taps; the
    weight gradient uses the library kernel on the side stream (off the critical path)."""

    @staticmethod
    def forward(ctx, x, w, stats, sink, ready, bn_hook=None):
        from . import native, count_launch

        ctx.bn_hook = bn_hook
        x = _cl(x)
        n, _, h, wd = x.shape
        y = torch.empty((n, w.shape[0], h, wd), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        native().conv3x3(x, w, y, False, stats, None, False)
        count_launch()
        ctx.save_for_backward(x, w)
        ctx.sink, ctx.ready = sink, ready
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import native, count_launch

        x, w = ctx.saved_tensors
        dy = _cl(dy)
        wv = w.permute(0, 3, 1, 2)
        cfg = (1, 1, 1)
        dx = None
        if ctx.needs_input_grad[0]:
            if conv3x3_supported(dy, w, dgrad=True):
                dx = torch.empty_like(x)
                hook = ctx.bn_hook if _bn_fusable(ctx.bn_hook, x, x.shape[1]) else None
                if hook is not None:
                    native().conv3x3(dy, w, dx, True, None, hook.as_list(dx.numel() // dx.shape[1], dx.shape[1]), hook.relu)
                    hook.done = True
                else:
                    native().conv3x3(dy, w, dx, True, None, None, False)
                count_launch()
            else:
                dx = _ConvLibFn._bwd(dy, x, wv, cfg, [True, False, False])[0]
        dw = None
        if ctx.needs_input_grad[1]:
            sink, ready = ctx.sink, ctx.ready
            side = _side_stream() if sink is not None else None
            own = OWN_WGRAD3 and conv3x3_wgrad_supported(x, w)

            def wgrad():
                if _SKIP_WGRAD and sink is not None:
                    if ready is not None:
                        ready()
                    return None
                if own:
                    out = conv3x3_wgrad(x, dy, w.shape, sink)
                    if sink is None:
                        return out
                    if ready is not None:
                        ready()
                    return None
                gw = _ConvLibFn._bwd(dy, x, wv, cfg, [False, True, False])[1].permute(0, 2, 3, 1)
                if sink is None:
                    return gw.contiguous()
                sink.view(w.shape).add_(gw)
                if ready is not None:
                    ready()
                return None

            if side is not None:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dy.device))
                side.wait_event(ev)
                _SIDE["keep"].append((dy, x))
                with torch.cuda.stream(side):
                    dw = wgrad()
            else:
                dw = wgrad()
        return dx, dw, None, None, None, None


# EXPERIMENTAL until it has been validated on a GPU (written after the round-1 GPU budget was spent):
# EDL_OWN_WGRAD3=1 routes the 3x3 weight gradient through csrc/conv3x3_wgrad.cu instead of the library kernel.
OWN_WGRAD3 = __import__("os").environ.get("EDL_OWN_WGRAD3", "0") == "1"


def conv3x3_wgrad_supported(x, weight_krsc) -> bool:
    from . import native

    if not (x.is_cuda and x.dtype == torch.bfloat16 and weight_krsc.dtype == torch.bfloat16 and x.dim() == 4):
        return False
    n, c, h, w = x.shape
    cout, kh, kw, cin = weight_krsc.shape
    return kh == 3 and kw == 3 and cin == c and native().conv3x3_wgrad_supported(n, h, w, cin, cout)


def conv3x3_wgrad(x, dy, weight_shape, sink=None, split_k: Optional[int] = None):
    """dW [Cout, 3, 3, Cin] of the 3x3 / stride 1 / pad 1 convolution on the tcgen05 kernel
    (csrc/conv3x3_wgrad.cu).  With ``sink`` (a flat bf16 window of the gradient bucket) the result is ADDED
    into it by the kernel's fused finalize and ``None`` is returned; otherwise a new tensor is returned."""
    from . import native, count_launch

    x, dy = _cl(x), _cl(dy)
    cout, _, _, cin = weight_shape
    n, _, h, w = dy.shape                 # stride 2: x is twice as large; the pixel blocks tile dy
    s2 = x.shape[2] != h
    tiles = native().conv3x3_wgrad_tiles(cin, cout)
    ws, counters = _splitk_workspace(x.device, cout * 9 * cin, tiles)
    if split_k is None:
        if s2:                           # always the version-1 kernel (its pixel-block plan, 64-wide Cin tiles)
            kblocks, ctas = native().conv3x3_wgrad_s2_kblocks(n, h, w), tiles
        else:
            kblocks = native().conv3x3_wgrad_kblocks(n, h, w)
            ctas = native().conv3x3_wgrad_ctas(cin, cout)
        # one wave of CTAs, at least two pixel blocks per CTA; no split once a third of the SMs have a CTA (same rule
        # as the 1x1 wgrad split)
        split_k = 1 if ctas >= 48 else max(1, min(max(1, _NUM_SMS // ctas), max(1, kblocks // 2)))
    if sink is not None:
        out, acc = sink.view(weight_shape), True
    else:
        out, acc = torch.empty(tuple(weight_shape), device=x.device, dtype=torch.bfloat16), False
    native().conv3x3_wgrad(x, dy, out, ws, counters, int(split_k), acc,
                           _splitk_partials(x.device) if split_k > 1 else None)
    count_launch(2 if (split_k > 1 and SPLITK_PARTIALS) else 1)
    return None if sink is not None else out


def conv3x3(x, weight_krsc, stats: Optional[torch.Tensor] = None):
    """NHWC 3x3 stride-1 pad-1 convolution; ``stats`` (pre-zeroed fp32 [2*Cout]) receives the per-channel
    sum / sum of squares of the output for the following train-mode BatchNorm."""
    sink = getattr(weight_krsc, "_edl_grad_sink", None) if torch.is_grad_enabled() else None
    ready = getattr(weight_krsc, "_edl_grad_ready", None) if sink is not None else None
    return _Conv3x3Fn.apply(x, weight_krsc, stats, sink, ready, getattr(x, "_edl_bn_hook", None))


# 3x3 / pad 1 / STRIDE 2 forward convolutions (student: first block of stages 2-4; teacher: the grouped ones, which
# round 1 ran as "stride 1 then subsample" = 4x the MMA work) on the persistent tcgen05 kernel; the input is sampled
# by the TMA traversal stride (csrc/gemm_persist.cu).  Validated on B200 in round 2, on by default (EDL_CONV3_S2=0: off).
CONV3_S2 = __import__("os").environ.get("EDL_CONV3_S2", "1") == "1"


def conv3x3_s2_supported(x, weight_krsc, groups=1) -> bool:
    from . import native

    if not (x.is_cuda and x.dtype == torch.bfloat16 and weight_krsc.dtype == torch.bfloat16 and x.dim() == 4):
        return False
    n, c, h, w = x.shape
    cout, kh, kw, cin_g = weight_krsc.shape
    return (kh == 3 and kw == 3 and cin_g * groups == c and h % 2 == 0 and w % 2 == 0 and w // 2 <= 128
            and native().persistent_gemm_enabled()
            and native().conv3x3_supported(n, h // 2, w // 2, c, cout, False, groups))


def _conv3x3_s2_launch(x, weight_krsc, stats, scale, shift, relu, groups):
    from . import native, count_launch

    x = _cl(x)
    n, _, h, w = x.shape
    y = torch.empty((n, weight_krsc.shape[0], h // 2, w // 2), device=x.device, dtype=x.dtype,
                    memory_format=torch.channels_last)
    native().conv3x3_s2(x, weight_krsc, y, stats, scale, shift, relu, groups)
    count_launch()
    return y


@torch.no_grad()
def conv3x3_s2_infer(x, weight_krsc, scale=None, shift=None, relu=False, groups=1):
    """Inference 3x3 / pad 1 / stride 2 convolution (dense or grouped) with the folded-BN epilogue."""
    return _conv3x3_s2_launch(x, weight_krsc, None, scale, shift, relu, groups)


# Backward of the stride-2 3x3 convolutions on our own kernels (EDL_OWN_S2_BWD=1; validated on B200, OFF by default:
# the three layers cost 0.12 ms / step more than the library's strided kernels -- profiles/bench_runs.json "s2lib"):
#   dgrad: four stride-1 tcgen05 launches over dy, one per parity class of the input pixel (1 + 2 + 2 + 4 taps = the
#          nine tap-MMAs of the forward pass), each storing through a strided tensor map.  (First version: zero-insert
#          dy with csrc/pool.cu:dilate2_kernel and run the plain stride-1 dgrad: 4x the MMAs, +0.12 ms/step.);
#   wgrad: the version-1 weight-gradient kernel with X read through the TMA traversal stride (csrc/conv3x3_wgrad.cu).
OWN_S2_BWD = __import__("os").environ.get("EDL_OWN_S2_BWD", "0") == "1"


def _s2_dgrad_supported(x, w) -> bool:
    from . import native

    n, c, h, wd = x.shape
    cout = w.shape[0]
    return (OWN_S2_BWD and x.is_cuda and h % 2 == 0 and wd % 2 == 0
            and native().conv3x3_dgrad_s2_supported(n, h // 2, wd // 2, c, cout))


def _s2_wgrad_supported(x, w) -> bool:
    from . import native

    n, c, h, wd = x.shape
    return (OWN_S2_BWD and x.is_cuda and h % 2 == 0 and wd % 2 == 0
            and native().conv3x3_wgrad_s2_supported(n, h // 2, wd // 2, c, w.shape[0]))


class _Conv3x3S2Fn(torch.autograd.Function):
    """Training form of the stride-2 3x3 convolution: forward on the tcgen05 kernel (BN statistics in its epilogue),
    input gradient = stride-1 tcgen05 dgrad of the zero-inserted output gradient, weight gradient on the tcgen05
    wgrad kernel through a strided tensor map (side stream, straight into the gradient bucket)."""

    @staticmethod
    def forward(ctx, x, w, stats, sink, ready):
        x = _cl(x)
        y = _conv3x3_s2_launch(x, w, stats, None, None, False, 1)
        ctx.save_for_backward(x, w)
        ctx.cfg = (2, 1, 1)
        ctx.sink, ctx.ready = sink, ready
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import native, count_launch

        x, w = ctx.saved_tensors
        dy = _cl(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            if _s2_dgrad_supported(x, w):
                # four stride-1 launches (1 + 2 + 2 + 4 taps), one per parity class of the input pixel, each storing
                # through a strided tensor map (csrc/gemm_persist.cu:conv3x3_dgrad_s2_persistent)
                dx = torch.empty_like(x)
                native().conv3x3_dgrad_s2(dy, w, dx)
                count_launch(4)
            else:
                dx = _ConvLibFn._bwd(dy, x, w.permute(0, 3, 1, 2), ctx.cfg, [True, False, False])[0]
        if ctx.needs_input_grad[1]:
            if _s2_wgrad_supported(x, w):
                sink, ready = ctx.sink, ctx.ready
                side = _side_stream() if sink is not None else None

                def wgrad():
                    out = conv3x3_wgrad(x, dy, w.shape, sink)
                    if sink is not None and ready is not None:
                        ready()
                    return out

                if side is not None:
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(dy.device))
                    side.wait_event(ev)
                    _SIDE["keep"].append((dy, x))
                    with torch.cuda.stream(side):
                        dw = wgrad()
                else:
                    dw = wgrad()
            else:
                ctx2 = ctx
                saved = ctx.needs_input_grad
                dw = _ConvLibFn.backward(_NeedsOnlyWeight(ctx2, saved), dy)[1]
        return dx, dw, None, None, None


class _NeedsOnlyWeight:
    """View of an autograd ctx that asks _ConvLibFn.backward for the weight gradient only."""

    def __init__(self, ctx, needs):
        self._ctx = ctx
        self.needs_input_grad = (False, needs[1]) + tuple(needs[2:])

    def __getattr__(self, k):
        return getattr(self._ctx, k)


def conv3x3_s2(x, weight_krsc, stats: Optional[torch.Tensor] = None):
    sink = getattr(weight_krsc, "_edl_grad_sink", None) if torch.is_grad_enabled() else None
    ready = getattr(weight_krsc, "_edl_grad_ready", None) if sink is not None else None
    return _Conv3x3S2Fn.apply(x, weight_krsc, stats, sink, ready)


# The first stem convolution (3 -> 32 channels, 3x3 / stride 2) runs on the direct kernel of csrc/stem.cu with the
# BatchNorm statistics fused in.  Validated on B200 in round 2, on by default (EDL_OWN_STEM1=0: library kernel).
OWN_STEM1 = __import__("os").environ.get("EDL_OWN_STEM1", "1") == "1"


def stem7_supported(x, weight_krsc) -> bool:
    """7x7 / stride 2 / pad 3 convolution of a 3-channel image (the teacher's stem) as im2col + tcgen05 GEMM."""
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] == 3
            and weight_krsc.dtype == torch.bfloat16 and tuple(weight_krsc.shape[1:]) == (7, 7, 3)
            and weight_krsc.shape[0] % 8 == 0)


_STEM7_W = {}


@torch.no_grad()
def stem7_infer(x, weight_krsc, scale=None, shift=None, relu=False):
    """y = act(conv7x7/s2/p3(x, w) * scale + shift) for x [N,3,H,W] (channels_last bf16), w KRSC [Cout,7,7,3]:
    ``stem7_im2col`` writes the windows as rows of a [N*Ho*Wo, 160] matrix ((r, s, c) order = the KRSC weight order, 13 zero
    columns of padding) and the persistent GEMM applies the folded BatchNorm and the ReLU in its epilogue.  The reference's
    teacher runs this layer through the serving library (start_local_teacher.sh:24-30)."""
    from . import native, count_launch

    x = _cl(x)
    n, _, h, w = x.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    cout = weight_krsc.shape[0]
    key = (weight_krsc.data_ptr(), weight_krsc._version, str(weight_krsc.device))
    wp = _STEM7_W.get(key)
    if wp is None:
        _STEM7_W.clear()
        wp = torch.zeros((cout, 160), device=x.device, dtype=torch.bfloat16)
        wp[:, :147] = weight_krsc.reshape(cout, 147)
        _STEM7_W[key] = wp
    a = torch.empty((n * ho * wo, 160), device=x.device, dtype=torch.bfloat16)
    native().stem7_im2col(x, a)
    count_launch()
    y = torch.empty((n, cout, ho, wo), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    gemm_bf16(a, wp, out=y.permute(0, 2, 3, 1).reshape(-1, cout), col_scale=scale, col_shift=shift, relu=relu)
    return y


def stem_conv_supported(x, weight_krsc) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and weight_krsc.dtype == torch.bfloat16 and x.dim() == 4
            and x.shape[1] == 3 and tuple(weight_krsc.shape) == (32, 3, 3, 3))


class _StemConvFn(torch.autograd.Function):
    """conv3x3 / stride 2 / pad 1, 3 -> 32 channels: forward on the direct kernel (+ BN statistics); the weight gradient
    (and, if anybody asks for it, the input gradient) on the library kernels like every other k x k convolution."""

    @staticmethod
    def forward(ctx, x, w, stats, sink, ready):
        from . import native, count_launch

        x = _cl(x)
        n, _, h, wd = x.shape
        y = torch.empty((n, 32, (h - 1) // 2 + 1, (wd - 1) // 2 + 1), device=x.device, dtype=x.dtype,
                        memory_format=torch.channels_last)
        native().stem_conv3x3s2(x, w, y, stats)
        count_launch()
        ctx.save_for_backward(x, w)
        ctx.cfg = (2, 1, 1)
        ctx.sink, ctx.ready = sink, ready
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import native, count_launch

        x, w = ctx.saved_tensors
        dy = _cl(dy)
        dx = None
        if ctx.needs_input_grad[0]:      # nobody asks for the gradient of the images; kept for completeness
            dx = _ConvLibFn._bwd(dy, x, w.permute(0, 3, 1, 2), ctx.cfg, [True, False, False])[0]
        dw = None
        if ctx.needs_input_grad[1]:
            if not (x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 and x.shape[3] <= 4096):
                dw = _ConvLibFn.backward(_NeedsOnlyWeight(ctx, ctx.needs_input_grad), dy)[1]
                return dx, dw, None, None, None
            sink, ready = ctx.sink, ctx.ready
            side = _side_stream() if sink is not None else None

            def wgrad():
                if _SKIP_WGRAD and sink is not None:
                    if ready is not None:
                        ready()
                    return None
                ws, counters = _splitk_workspace(x.device, 1024, 1)
                out = sink.view(w.shape) if sink is not None else torch.empty_like(w)
                native().stem_wgrad(x, dy, out, ws, counters, sink is not None)
                count_launch()
                if sink is not None:
                    if ready is not None:
                        ready()
                    return None
                return out

            if side is not None:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dy.device))
                side.wait_event(ev)
                _SIDE["keep"].append((dy, x))
                with torch.cuda.stream(side):
                    dw = wgrad()
            else:
                dw = wgrad()
        return dx, dw, None, None, None


def stem_conv(x, weight_krsc, stats: Optional[torch.Tensor] = None):
    sink = getattr(weight_krsc, "_edl_grad_sink", None) if torch.is_grad_enabled() else None
    ready = getattr(weight_krsc, "_edl_grad_ready", None) if sink is not None else None
    return _StemConvFn.apply(x, weight_krsc, stats, sink, ready)


# ---------------------------------------------------------------------------------------------------------------
# 3x3 / stride 1 convolutions with 32 input channels (the stem's conv1_2 / conv1_3) in "pixel-pair" form on the
# 64-channel tcgen05 kernels (csrc/misc.cu:pair_weight_expand_kernel explains the transform).  EDL_OWN_STEM23=1.
OWN_STEM23 = __import__("os").environ.get("EDL_OWN_STEM23", "0") == "1"


def _pair_view(t):
    """NHWC [N, H, W, C] seen as [N, H, W/2, 2C] (zero-copy), returned in the NCHW-shaped channels_last form."""
    n, c, h, w = t.shape
    return t.as_strided((n, 2 * c, h, w // 2), (h * w * c, 1, w * c, 2 * c))


def _unpair_view(t2):
    n, c2, h, w2 = t2.shape
    return t2.as_strided((n, c2 // 2, h, 2 * w2), (h * w2 * c2, 1, w2 * c2, c2 // 2))


def conv3x3_pair_supported(x, weight_krsc) -> bool:
    from . import native

    if not (OWN_STEM23 and x.is_cuda and x.dtype == torch.bfloat16 and weight_krsc.dtype == torch.bfloat16 and x.dim() == 4):
        return False
    n, c, h, w = x.shape
    cout, kh, kw, cin = weight_krsc.shape
    return (kh == 3 and kw == 3 and cin == 32 and c == 32 and w % 2 == 0 and cout % 32 == 0
            and native().conv3x3_supported(n, h, w // 2, 64, 2 * cout, False, 1)
            and native().conv3x3_supported(n, h, w // 2, 2 * cout, 64, True, 1))


class _Conv3x3PairFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stats, sink, ready):
        from . import native, count_launch

        C = native()
        x = _cl(x)
        n, _, h, wd = x.shape
        cout = w.shape[0]
        w2 = torch.empty((2 * cout, 3, 3, 64), device=x.device, dtype=x.dtype)
        C.pair_weight_expand(w.contiguous(), w2)
        x2 = _pair_view(x)
        y2 = torch.empty((n, 2 * cout, h, wd // 2), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        s2 = torch.zeros(4 * cout, device=x.device, dtype=torch.float32) if stats is not None else None
        C.conv3x3(x2, w2, y2, False, s2, None, False)
        count_launch(2)
        if stats is not None:
            C.fold_pair_stats(s2, stats)
            count_launch()
        ctx.save_for_backward(x, w, w2)
        ctx.sink, ctx.ready = sink, ready
        return _unpair_view(y2)

    @staticmethod
    def backward(ctx, dy):
        from . import native, count_launch

        C = native()
        x, w, w2 = ctx.saved_tensors
        dy = _cl(dy)
        x2, dy2 = _pair_view(x), _pair_view(dy)
        dx = None
        if ctx.needs_input_grad[0]:
            dx2 = torch.empty_like(x2)
            C.conv3x3(dy2, w2, dx2, True, None, None, False)
            count_launch()
            dx = _unpair_view(dx2)
        dw = None
        if ctx.needs_input_grad[1]:
            sink, ready = ctx.sink, ctx.ready
            side = _side_stream() if sink is not None else None

            def wgrad():
                if _SKIP_WGRAD and sink is not None:
                    if ready is not None:
                        ready()
                    return None
                dw2 = conv3x3_wgrad(x2, dy2, w2.shape, None)
                out = sink.view(w.shape) if sink is not None else torch.empty_like(w)
                C.pair_weight_fold(dw2, out, sink is not None)
                count_launch()
                if sink is not None:
                    if ready is not None:
                        ready()
                    return None
                return out

            if side is not None:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dy.device))
                side.wait_event(ev)
                _SIDE["keep"].append((dy, x))
                with torch.cuda.stream(side):
                    dw = wgrad()
                    _SIDE["keep"].append(None)
            else:
                dw = wgrad()
        return dx, dw, None, None, None


def conv3x3_pair(x, weight_krsc, stats: Optional[torch.Tensor] = None):
    sink = getattr(weight_krsc, "_edl_grad_sink", None) if torch.is_grad_enabled() else None
    ready = getattr(weight_krsc, "_edl_grad_ready", None) if sink is not None else None
    return _Conv3x3PairFn.apply(x, weight_krsc, stats, sink, ready)


def conv3x3_infer_supported(x, weight_krsc, groups=1) -> bool:
    from . import native

    if not (x.is_cuda and x.dtype == torch.bfloat16 and weight_krsc.dtype == torch.bfloat16 and x.dim() == 4):
        return False
    n, c, h, w = x.shape
    cout, kh, kw, cin_g = weight_krsc.shape
    return (kh == 3 and kw == 3 and cin_g * groups == c
            and native().conv3x3_supported(n, h, w, c, cout, False, groups))


@torch.no_grad()
def conv3x3_infer(x, weight_krsc, scale=None, shift=None, relu=False, groups=1):
    """Inference 3x3 / stride 1 / pad 1 convolution (dense or grouped) on the persistent tcgen05 kernel with
    the folded-BatchNorm scale / shift and the ReLU in its epilogue.  The grouped form is what the
    ResNeXt teacher needs: cuDNN's grouped bf16 NHWC kernel takes 26 ms for ONE 7x7x4096 (g=32) layer at
    batch 32 (profiles/teacher_r1.txt)."""
    from . import native, count_launch

    x = _cl(x)
    n, _, h, w = x.shape
    y = torch.empty((n, weight_krsc.shape[0], h, w), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    native().conv3x3_infer(x, weight_krsc, y, scale, shift, relu, groups)
    count_launch()
    return y
