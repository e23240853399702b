"""Persistent tcgen05 GEMM / conv kernel (csrc/gemm_persist.cu) against the one-tile-per-CTA kernels and
the fp32 reference: many tiles per CTA (accumulator double-buffering, ring wrap-around across tiles)."""
import pytest
import torch
import torch.nn.functional as F

from edl_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.fixture(autouse=True)
def _restore():
    yield
    ops.native().set_persistent_gemm(True)


@pytest.mark.parametrize("m,n,k", [(100352, 64, 64), (100352, 256, 64), (25088, 512, 128), (6272, 1024, 256),
                                   (1568, 2048, 512), (40000, 192, 72), (300, 128, 64), (128 * 148 * 3 + 5, 128, 64)])
@pytest.mark.parametrize("b_mn", [False, True])
def test_persistent_gemm_matches_reference_and_stats(m, n, k, b_mn):
    torch.manual_seed(0)
    a = torch.randn(m, k, device=DEV).bfloat16()
    b = (torch.randn(k, n, device=DEV) if b_mn else torch.randn(n, k, device=DEV)).bfloat16()
    ref = a.float() @ (b.float() if b_mn else b.float().t())
    outs = []
    for persistent in (True, False):
        ops.native().set_persistent_gemm(persistent)
        stats = torch.zeros(2 * n, device=DEV)
        d = ops.gemm_bf16(a, b, b_mn_major=b_mn, col_stats=None if b_mn else stats)
        assert _rel(d, ref) < 1e-2, persistent
        if not b_mn:
            assert _rel(stats[:n], d.float().sum(0)) < 2e-3
            assert _rel(stats[n:], (d.float() ** 2).sum(0)) < 2e-3
        outs.append(d)
    assert torch.equal(outs[0], outs[1])          # same MMA order => bit-identical results


@pytest.mark.parametrize("m,n,k,res", [(25088, 512, 512, False), (6272, 2048, 1024, True), (100352, 512, 256, True),
                                       (1568, 4096, 2048, False)])
def test_wide_tile_gemm_matches_the_128_wide_tiles(m, n, k, res):
    """128 x 256 tiles (both TMEM accumulators 256 columns wide) for the teacher's large-N 1x1 convolutions: same
    results as the 128 x 128 tiles, with the folded-BN epilogue and the TMA-fetched residual addend."""
    torch.manual_seed(2)
    a = torch.randn(m, k, device=DEV).bfloat16()
    b = (torch.randn(n, k, device=DEV) * 0.05).bfloat16()
    sh = torch.randn(n, device=DEV)
    add = torch.randn(m, n, device=DEV).bfloat16() if res else None
    ref = a.float() @ b.float().t() + sh + (add.float() if res else 0.0)
    outs = []
    try:
        ops.native().set_pair_gemm(False)          # these shapes would otherwise take the CTA-pair kernel
        for wide in (True, False):
            ops.native().set_wide_gemm_tiles(wide)
            d = ops.gemm_bf16(a, b, col_shift=sh, relu=True, add=add)
            assert _rel(d, torch.relu(ref)) < 1e-2, wide
            outs.append(d)
    finally:
        ops.native().set_wide_gemm_tiles(True)
        ops.native().set_pair_gemm(True)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("m,n,k,res,scale", [(25088, 512, 512, False, False), (6272, 2048, 1024, True, False),
                                             (100352, 512, 256, True, True), (1568, 4096, 2048, False, True),
                                             (19000, 256, 320, True, False), (9601, 768, 4096, False, False)])
def test_cta_pair_gemm_matches_the_single_cta_tiles(m, n, k, res, scale):
    """256 x 256 tiles on two SMs (tcgen05.mma.cta_group::2, csrc/gemm_2cta.cu): same k order per output element as the
    single-CTA kernel, so the results are bit-identical; M tails that end inside the first / second CTA of a pair, K tails
    (k = 320: the last k-block is half out of bounds), the addend and the scale / shift / ReLU epilogue."""
    torch.manual_seed(3)
    nat = ops.native()
    a = torch.randn(m, k, device=DEV).bfloat16()
    b = (torch.randn(n, k, device=DEV) * 0.05).bfloat16()
    sh = torch.randn(n, device=DEV)
    sc = (torch.rand(n, device=DEV) + 0.5) if scale else None
    add = torch.randn(m, n, device=DEV).bfloat16() if res else None
    ref = a.float() @ b.float().t() + (add.float() if res else 0.0)
    ref = torch.relu((ref * sc if scale else ref) + sh)
    assert nat.get_pair_gemm()
    outs = []
    try:
        for pair in (True, False):
            nat.set_pair_gemm(pair)
            d = ops.gemm_bf16(a, b, col_scale=sc, col_shift=sh, relu=True, add=add)
            assert _rel(d, ref) < 1e-2, pair
            outs.append(d)
    finally:
        nat.set_pair_gemm(True)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("m,n,k", [(100352, 256, 64), (25088, 128, 512), (5000, 384, 192)])
def test_sixteen_epilogue_warps_match_eight(m, n, k):
    """The 128- / 256-column kernels can run 16 epilogue warps (a quarter of the tile's columns per warpgroup;
    EDL_EPI_WARPS=16 / set_epilogue_warps(16)) instead of 8: identical outputs, statistics within summation-order noise, for
    the forward GEMM with BN statistics, the dgrad with the fused BN-backward reduction and the 3x3 convolution."""
    torch.manual_seed(7)
    nat = ops.native()
    a = torch.randn(m, k, device=DEV).bfloat16()
    b = (torch.randn(n, k, device=DEV) * 0.05).bfloat16()
    wmn = (torch.randn(k, n, device=DEV) * 0.05).bfloat16()
    x = torch.randn(m, n, device=DEV).bfloat16()
    xc = torch.randn(8, 128, 28, 28, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    wc = (torch.randn(128, 3, 3, 128, device=DEV) * 0.05).bfloat16()
    res = []
    h = _hook(x, None, n, True)
    try:
        for warps in (16, 8):
            nat.set_epilogue_warps(warps)
            st = torch.zeros(2 * n, device=DEV)
            d = ops.gemm_bf16(a, b, col_stats=st)
            h.dsums = torch.zeros(2 * n, device=DEV)
            g = ops.gemm_bf16(a, wmn, b_mn_major=True, bn=h)
            cst = torch.zeros(256, device=DEV)
            yc = ops.conv3x3(xc, wc, cst)
            res.append((d, st, g, h.dsums, yc, cst))
    finally:
        nat.set_epilogue_warps(8)
    for i in (0, 2, 4):
        assert torch.equal(res[0][i], res[1][i]), i
    for i in (1, 3, 5):
        assert _rel(res[0][i], res[1][i]) < 1e-4, i
    assert _rel(res[0][0], a.float() @ b.float().t()) < 1e-2
    assert _rel(res[0][3], _bn_ref_sums(res[0][2], x, None, h.mean, h.rstd, h.gamma, h.beta, True)) < 5e-3


@pytest.mark.parametrize("kind,shape", [("gemm", (100352, 256, 64)), ("gemm", (5000, 192, 192)), ("gemm", (6272, 1024, 256)),
                                        ("gemm", (300, 128, 64)), ("conv", (32, 128, 128, 28, 28)), ("conv", (5, 128, 256, 11, 20)),
                                        ("conv", (9, 256, 256, 14, 14)), ("conv", (7, 512, 512, 7, 7))])
def test_tensor_core_bn_statistics_match_the_column_loop(kind, shape):
    """Forward BatchNorm statistics as Gram / ones products of the staged bf16 tile on the tensor cores (128-column kernels,
    EDL_TC_STATS=1; exact but not faster, so not the default) against the column-pair loop and against fp32 sums of the stored output: M tails, a
    partial N tile (192), alternating N tiles per CTA (1024 columns), conv tiles with partial patches and tail images."""
    torch.manual_seed(11)
    nat = ops.native()
    if kind == "gemm":
        m, n, k = shape
        a = torch.randn(m, k, device=DEV).bfloat16()
        b = (torch.randn(n, k, device=DEV) * 0.05).bfloat16()
    else:
        nb, c, n, h, w = shape
        x = torch.randn(nb, c, h, w, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
        wt = (torch.randn(n, 3, 3, c, device=DEV) * 0.05).bfloat16()
    res = []
    try:
        for tcs in (True, False):
            nat.set_tc_stats(tcs)
            st = torch.zeros(2 * n, device=DEV)
            y = ops.gemm_bf16(a, b, col_stats=st) if kind == "gemm" else ops.conv3x3(x, wt, st)
            yf = y.float()
            dims = 0 if kind == "gemm" else (0, 2, 3)
            res.append((y, st, torch.cat([yf.sum(dims), (yf * yf).sum(dims)])))
    finally:
        nat.set_tc_stats(False)
    assert torch.equal(res[0][0], res[1][0])
    for y, st, ref in res:
        assert _rel(st[n:], ref[n:]) < 1e-4          # sums of squares: no cancellation
        assert (st[:n] - ref[:n]).abs().max().item() < 2e-3 * ref[n:].sqrt().max().item() + 1e-3
    assert _rel(res[0][1][n:], res[1][1][n:]) < 1e-4


def test_persistent_gemm_epilogue_scale_shift_relu():
    torch.manual_seed(1)
    m, n, k = 5000, 256, 192
    a = torch.randn(m, k, device=DEV).bfloat16()
    b = torch.randn(n, k, device=DEV).bfloat16()
    sc, sh = torch.rand(n, device=DEV) + 0.5, torch.randn(n, device=DEV)
    d = ops.gemm_bf16(a, b, col_scale=sc, col_shift=sh, relu=True)
    assert _rel(d, torch.relu((a.float() @ b.float().t()) * sc + sh)) < 1e-2


@pytest.mark.parametrize("n,cin,cout,h,w", [(32, 64, 64, 56, 56), (32, 128, 128, 28, 28), (32, 256, 256, 14, 14),
                                            (32, 512, 512, 7, 7), (5, 64, 192, 11, 20)])
def test_persistent_conv3x3_matches_tile_kernel(n, cin, cout, h, w):
    torch.manual_seed(2)
    x = torch.randn(n, cin, h, w, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, 3, 3, cin, device=DEV) * 0.05).bfloat16()
    res = []
    for persistent in (True, False):
        ops.native().set_persistent_gemm(persistent)
        st = torch.zeros(2 * cout, device=DEV)
        y = ops.conv3x3(x, wt, st)
        res.append((y, st))
    # the haloed-tile version accumulates in (filter row, k-block, tap) order, the tile kernel in (tap, k-block) order
    assert _rel(res[0][0], res[1][0]) < 2e-3
    assert _rel(res[0][1], res[1][1]) < 2e-3
    ref = F.conv2d(x.float(), wt.permute(0, 3, 1, 2).float(), None, 1, 1)
    assert _rel(res[0][0], ref) < 1e-2
    if cin <= 64 or cin % 128 == 0:
        dy = torch.randn_like(res[0][0])
        dxs = []
        for persistent in (True, False):
            ops.native().set_persistent_gemm(persistent)
            dx = torch.empty_like(x)
            ops.native().conv3x3(dy, wt, dx, True, None, None, False)
            dxs.append(dx)
        assert _rel(dxs[0], dxs[1]) < 2e-3


@pytest.mark.parametrize("n,cin,cout,h,w,groups", [(32, 64, 64, 56, 56, 1), (8, 128, 128, 28, 28, 1), (9, 256, 256, 14, 14, 1),
                                                   (7, 512, 512, 7, 7, 1), (5, 64, 192, 11, 20, 1), (3, 128, 64, 9, 30, 1),
                                                   (2, 64, 64, 5, 127, 1), (3, 512, 512, 14, 14, 8), (2, 256, 512, 6, 6, 2),
                                                   (32, 2048, 2048, 14, 14, 32), (40, 192, 192, 7, 7, 3)])
def test_haloed_conv3x3_tiles_match_the_shifted_box_version(n, cin, cout, h, w, groups):
    """One A tile per filter row read through three shifted descriptors (EDL_CONV_HALO, the default) against one TMA box per
    tap, and both against the fp32 convolution: fprop with statistics / folded-BN epilogue, dgrad, grouped fprop."""
    torch.manual_seed(5)
    nat = ops.native()
    x = torch.randn(n, cin, h, w, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, 3, 3, cin // groups, device=DEV) * 0.05).bfloat16()
    scale = torch.rand(cout, device=DEV) + 0.5
    shift = torch.randn(cout, device=DEV) * 0.1
    assert nat.get_conv_halo()
    outs = []
    try:
        for halo, resident in ((True, True), (True, False), (False, False)):
            nat.set_conv_halo(halo)
            nat.set_conv_resident_weights(resident)     # 64-channel layers / groups: weights loaded once per run of tiles
            if groups == 1:
                st = torch.zeros(2 * cout, device=DEV)
                y = ops.conv3x3(x, wt, st)
            else:
                st = None
                y = ops.conv3x3_infer(x, wt, scale, shift, relu=True, groups=groups)
            dx = None
            if groups == 1 and (cin <= 64 or cin % 128 == 0):
                dy = torch.randn_like(y)
                dx = torch.empty_like(x)
                nat.conv3x3(dy, wt, dx, True, None, None, False)
            outs.append((y, st, dx, dy if dx is not None else None))
    finally:
        nat.set_conv_halo(True)
        nat.set_conv_resident_weights(True)
    wf = wt.permute(0, 3, 1, 2).float()
    ref = F.conv2d(x.float(), wf, None, 1, 1, 1, groups)
    if groups > 1:
        ref = torch.relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    assert _rel(outs[0][0], ref) < 1e-2
    assert _rel(outs[0][0], outs[2][0]) < 2e-3
    assert torch.equal(outs[0][0], outs[1][0])          # same accumulation order, only the tile schedule differs
    if outs[0][1] is not None:
        assert _rel(outs[0][1], outs[1][1]) < 1e-4
        yf = outs[0][0].float()
        assert _rel(outs[0][1][:cout], yf.sum((0, 2, 3))) < 2e-3
        assert _rel(outs[0][1][cout:], (yf * yf).sum((0, 2, 3))) < 2e-3
    if outs[0][2] is not None:
        for y, _, dx, dy in outs:
            dref = torch.nn.grad.conv2d_input(x.shape, wf, dy.float(), 1, 1)
            assert _rel(dx, dref) < 1e-2


def test_conv1x1_fork_sums_both_gradients_in_the_dgrad_epilogue():
    torch.manual_seed(3)
    x = torch.randn(8, 256, 14, 14, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(64, 1, 1, 256, device=DEV) * 0.05).bfloat16().requires_grad_(True)
    r = torch.randn(8, 256, 14, 14, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    dy = torch.randn(8, 64, 14, 14, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    y, xa = ops.conv1x1(x, w, None, fork=True)
    torch.autograd.backward([y, xa], [dy, r])
    xr = x.detach().float().requires_grad_(True)
    wr = w.detach().float().requires_grad_(True)
    yr = F.conv2d(xr, wr.view(64, 256, 1, 1))
    torch.autograd.backward([yr, xr * 1.0], [dy.float(), r.float()])
    assert _rel(x.grad, xr.grad) < 1e-2
    assert _rel(w.grad, wr.grad.view_as(w)) < 1e-2
    # alias unused downstream: plain dgrad
    x2 = x.detach().clone().requires_grad_(True)
    y2, _ = ops.conv1x1(x2, w, None, fork=True)
    y2.backward(dy)
    xr2 = x.detach().float().requires_grad_(True)
    F.conv2d(xr2, wr.detach().view(64, 256, 1, 1)).backward(dy.float())
    assert _rel(x2.grad, xr2.grad) < 1e-2


def _bn_ref_sums(dy, x, y, mean, rstd, gamma, beta, relu):
    dyf, xf = dy.float(), x.float()
    if relu:
        mask = (y.float() > 0) if y is not None else (torch.addcmul(beta - mean * gamma * rstd, xf, gamma * rstd) > 0)
        dyf = dyf * mask
    return torch.cat([dyf.sum(0), (dyf * ((xf - mean) * rstd)).sum(0)])


def _hook(x, y, n, relu):
    from edl_b200.ops.bn import BNBackwardHook

    h = BNBackwardHook()
    h.x, h.y, h.relu = x, y, relu
    h.mean, h.rstd = torch.randn(n, device=DEV) * 0.1, torch.rand(n, device=DEV) + 0.5
    h.gamma, h.beta = torch.rand(n, device=DEV) + 0.5, torch.randn(n, device=DEV) * 0.2
    h.dsums = torch.zeros(2 * n, device=DEV)
    return h


@pytest.mark.parametrize("m,k,n", [(2048, 64, 32), (5000, 64, 128), (6272, 512, 256), (300, 64, 1024)])
@pytest.mark.parametrize("relu,has_y", [(True, False), (True, True), (False, False)])
def test_dgrad_epilogue_bn_backward_reduction_gemm(m, k, n, relu, has_y):
    """(experimental path, off by default) the 1x1 dgrad epilogue reduces sum(dy_m), sum(dy_m * xhat) of the tile."""
    torch.manual_seed(5)
    a = torch.randn(m, k, device=DEV).bfloat16()
    w = (torch.randn(k, n, device=DEV) * 0.1).bfloat16()
    x = torch.randn(m, n, device=DEV).bfloat16()
    y = torch.randn(m, n, device=DEV).bfloat16() if has_y else None
    h = _hook(x, y, n, relu)
    d = ops.gemm_bf16(a, w, b_mn_major=True, bn=h)
    assert h.done and _rel(d, a.float() @ w.float()) < 1e-2
    assert _rel(h.dsums, _bn_ref_sums(d, x, y, h.mean, h.rstd, h.gamma, h.beta, relu)) < 1e-4


@pytest.mark.parametrize("nb,c,hh,ww", [(8, 64, 16, 16), (32, 256, 14, 14), (5, 64, 11, 20), (8, 256, 2, 2)])
def test_dgrad_epilogue_bn_backward_reduction_conv3x3(nb, c, hh, ww):
    torch.manual_seed(6)
    dy = torch.randn(nb, c, hh, ww, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(c, 3, 3, c, device=DEV) * 0.05).bfloat16()
    x = torch.randn(nb, c, hh, ww, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    h = _hook(x, None, c, True)
    dx = torch.empty_like(x)
    ops.native().conv3x3(dy, wt, dx, True, None, h.as_list(nb * hh * ww, c), True)
    d2, x2 = dx.permute(0, 2, 3, 1).reshape(-1, c), x.permute(0, 2, 3, 1).reshape(-1, c)
    assert _rel(h.dsums, _bn_ref_sums(d2, x2, None, h.mean, h.rstd, h.gamma, h.beta, True)) < 1e-4


@pytest.mark.parametrize("n,c,cout,h,w,groups", [(4, 2048, 2048, 14, 14, 32), (4, 4096, 4096, 7, 7, 32), (2, 256, 256, 12, 12, 4),
                                                 (3, 128, 256, 9, 9, 1)])
def test_grouped_conv3x3_inference_with_folded_bn_epilogue(n, c, cout, h, w, groups):
    torch.manual_seed(4)
    x = torch.randn(n, c, h, w, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, 3, 3, c // groups, device=DEV) * 0.05).bfloat16()
    scale, shift = torch.rand(cout, device=DEV) + 0.5, torch.randn(cout, device=DEV) * 0.1
    assert ops.conv3x3_infer_supported(x, wt, groups)
    y = ops.conv3x3_infer(x, wt, scale, shift, True, groups)
    ref = F.conv2d(x.float(), wt.permute(0, 3, 1, 2).float(), None, 1, 1, 1, groups)
    ref = torch.relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    assert _rel(y, ref) < 1e-2


def test_phase_trace_of_the_persistent_kernel():
    """set_persist_trace: CTA 0 stamps clock64() at the phase boundaries of its first tiles (tools/trace_persist.py)."""
    nat = ops.native()
    a = torch.randn(4096 * 8, 64, device=DEV).bfloat16()
    b = (torch.randn(128, 64, device=DEV) * 0.05).bfloat16()
    st = torch.zeros(256, device=DEV)
    buf = torch.zeros(3 * 16 * 8, dtype=torch.int64, device=DEV)
    nat.set_persist_trace(buf)
    try:
        ops.gemm_bf16(a, b, col_stats=st)
        torch.cuda.synchronize()
    finally:
        nat.set_persist_trace(None)
    t = buf.view(3, 16, 8).cpu()
    epi = t[2]
    assert (epi[0, :5] > 0).all()                          # tile 0: loop top .. store issued
    assert (epi[0, 1:5] >= epi[0, :4]).all()               # in order
    assert epi[1, 0] > epi[0, 4]                           # the second tile starts after the first one's store
    buf.zero_()
    ops.gemm_bf16(a, b, col_stats=st)                      # tracing off: nothing is written
    torch.cuda.synchronize()
    assert int(buf.abs().sum()) == 0


@pytest.mark.parametrize("n,h,w,cout", [(4, 224, 224, 64), (3, 65, 97, 64), (2, 32, 48, 128)])
def test_stem7_im2col_gemm_matches_the_convolution(n, h, w, cout):
    """The teacher's 7x7 / stride 2 / pad 3 stem: im2col kernel (window rows in KRSC order, zero-padded to 160 columns) +
    persistent GEMM with the folded-BN / ReLU epilogue, against the fp32 convolution; odd sizes exercise the padding."""
    torch.manual_seed(13)
    x = torch.randn(n, 3, h, w, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, 7, 7, 3, device=DEV) * 0.05).bfloat16()
    sc, sh = torch.rand(cout, device=DEV) + 0.5, torch.randn(cout, device=DEV) * 0.1
    assert ops.stem7_supported(x, wt)
    y = ops.stem7_infer(x, wt, sc, sh, relu=True)
    ref = F.conv2d(x.float(), wt.permute(0, 3, 1, 2).float(), None, 2, 3)
    ref = torch.relu(ref * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert _rel(y, ref) < 1e-2
