"""Host-side check of the 3x3 weight-gradient kernel's algorithm (csrc/conv3x3_wgrad.cu) without a GPU: the pixel-box
planner is called through the built extension, and the kernel's data movement -- 4-D boxes {64 ch, W, BH rows, NB images}
of dY and of X shifted by (s-1, r-1), out-of-image elements zero-filled, one [Cout, Cin] product per tap accumulated over
the pixel blocks, result laid out KRSC -- is replayed with torch ops and compared with autograd's weight gradient; the
same for version 2 of the kernel (one haloed X tile per filter row, taps by shifted K rows).
What this cannot cover (descriptors, swizzle, barriers) is what tests/test_round2_gpu.py is for."""
import pytest
import torch
import torch.nn.functional as F

from edl_b200 import ops

pytestmark = pytest.mark.skipif(not ops.native_available(), reason="extension not built")


def _box(t, n0, nb, h0, bh, w0, w):
    """TMA tile semantics: t[n0:n0+nb, h0:h0+bh, w0:w0+w, :] with everything outside the tensor read as zero."""
    N, H, W, C = t.shape
    out = torch.zeros(nb, bh, w, C, dtype=t.dtype)
    ns, hs, ws = max(n0, 0), max(h0, 0), max(w0, 0)
    ne, he, we = min(n0 + nb, N), min(h0 + bh, H), min(w0 + w, W)
    if ns < ne and hs < he and ws < we:
        out[ns - n0:ne - n0, hs - h0:he - h0, ws - w0:we - w0] = t[ns:ne, hs:he, ws:we]
    return out


SHAPES = [(4, 56, 56, 8, 6), (8, 28, 28, 5, 7), (8, 14, 14, 4, 4), (32, 7, 7, 3, 5),
          (3, 6, 8, 4, 4), (5, 14, 14, 2, 3), (2, 4, 28, 3, 3), (6, 10, 16, 2, 2)]


@pytest.fixture
def wgrad_version():
    C = ops.native()
    prev = C.get_wgrad3_version()
    yield C
    C.set_wgrad3_version(prev, 0)


@pytest.mark.parametrize("n,h,w,cin,cout", SHAPES)
def test_pixel_box_replay_matches_autograd(wgrad_version, n, h, w, cin, cout):
    """Version 1: three boxes of X per pixel block, each shifted by (s - 1, r - 1)."""
    C = wgrad_version
    C.set_wgrad3_version(1, 0)
    bh, nb, kb, wb = C.conv3x3_wgrad_plan(n, h, w)
    if kb == 0:
        assert not C.conv3x3_wgrad_supported(n, h, w, 64, 64)
        pytest.skip("geometry not supported by the planner")
    assert wb == w and kb == w * bh * nb and kb % 16 == 0 and kb <= 112 and bh <= h and nb <= n
    hb, ng = -(-h // bh), -(-n // nb)
    assert C.conv3x3_wgrad_kblocks(n, h, w) == hb * ng
    torch.manual_seed(0)
    x = torch.randn(n, h, w, cin, dtype=torch.float64)          # NHWC like the kernel sees it
    dy = torch.randn(n, h, w, cout, dtype=torch.float64)
    dw = torch.zeros(cout, 3, 3, cin, dtype=torch.float64)     # KRSC
    for kblk in range(hb * ng):
        h0, img0 = (kblk % hb) * bh, (kblk // hb) * nb
        a = _box(dy, img0, nb, h0, bh, 0, w).reshape(kb, cout)                  # [pixels, Cout]  (MN-major A)
        for r in range(3):
            for s in range(3):
                b = _box(x, img0, nb, h0 + r - 1, bh, s - 1, w).reshape(kb, cin)  # same box, shifted (MN-major B)
                dw[:, r, s, :] += a.t() @ b
    wt = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.permute(0, 3, 1, 2), wt, None, 1, 1).backward(dy.permute(0, 3, 1, 2))
    assert torch.allclose(dw, wt.grad.permute(0, 2, 3, 1), atol=1e-9)


@pytest.mark.parametrize("n,h,w,cin,cout", SHAPES)
def test_haloed_tile_with_shifted_rows_matches_autograd(wgrad_version, n, h, w, cin, cout):
    """Version 2 (the default): dY and X boxes start at column -1 and are Wb >= W + 1 wide (zero-filled outside the
    image); ONE X tile per (pixel block, filter row) serves the three taps -- tap s reads the tile's K rows shifted by
    s - 1, the rows just outside the tile being zero."""
    C = wgrad_version
    C.set_wgrad3_version(2, 0)
    bh, nb, kb, wb = C.conv3x3_wgrad_plan(n, h, w)
    if kb == 0:
        pytest.skip("geometry not supported by the planner")
    assert wb >= w + 1 and kb == wb * bh * nb and kb % 16 == 0 and kb <= 128 and bh <= h and nb <= n
    hb, ng = -(-h // bh), -(-n // nb)
    assert C.conv3x3_wgrad_kblocks(n, h, w) == hb * ng
    torch.manual_seed(0)
    x = torch.randn(n, h, w, cin, dtype=torch.float64)
    dy = torch.randn(n, h, w, cout, dtype=torch.float64)
    dw = torch.zeros(cout, 3, 3, cin, dtype=torch.float64)
    zero_row = torch.zeros(1, cin, dtype=torch.float64)
    for kblk in range(hb * ng):
        h0, img0 = (kblk % hb) * bh, (kblk // hb) * nb
        a = _box(dy, img0, nb, h0, bh, -1, wb).reshape(kb, cout)
        for r in range(3):
            tile = _box(x, img0, nb, h0 + r - 1, bh, -1, wb).reshape(kb, cin)
            padded = torch.cat([zero_row, tile, zero_row])                     # the zeroed rows around the tile
            for s in range(3):
                dw[:, r, s, :] += a.t() @ padded[s:s + kb]                      # K row k  <->  tile row k + s - 1
    wt = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.permute(0, 3, 1, 2), wt, None, 1, 1).backward(dy.permute(0, 3, 1, 2))
    assert torch.allclose(dw, wt.grad.permute(0, 2, 3, 1), atol=1e-9)


def test_resnet_stage_geometries():
    C = ops.native()
    prev = C.get_wgrad3_version()
    try:
        C.set_wgrad3_version(1, 0)
        for n, hw in ((32, 56), (32, 28), (32, 14), (32, 7)):
            bh, nb, kb, wb = C.conv3x3_wgrad_plan(n, hw, hw)
            assert kb == 112 and hw % bh == 0 and n % nb == 0, (hw, bh, nb, kb)      # every MMA row is a real pixel
        C.set_wgrad3_version(2, 0)
        for n, hw in ((32, 56), (32, 28), (32, 14), (32, 7)):
            bh, nb, kb, wb = C.conv3x3_wgrad_plan(n, hw, hw)
            assert kb in (112, 128) and hw * bh * nb * 8 >= kb * 7, (hw, bh, nb, kb, wb)   # >= 7/8 of the K rows are pixels
    finally:
        C.set_wgrad3_version(prev, 0)


def _strided_box(t, n0, nb, h0, bh2, w0, w2, stride):
    """TMA tile with a traversal stride: a box spanning bh2 x w2 source elements from (h0, w0) delivers every
    ``stride``-th one (ceil(span / stride) per dimension); elements outside the tensor read as zero."""
    N, H, W, C = t.shape
    hs, ws = range(h0, h0 + bh2, stride), range(w0, w0 + w2, stride)
    out = torch.zeros(nb, len(hs), len(ws), C, dtype=t.dtype)
    for bi in range(nb):
        if not 0 <= n0 + bi < N:
            continue
        for i, h in enumerate(hs):
            for j, w in enumerate(ws):
                if 0 <= h < H and 0 <= w < W:
                    out[bi, i, j] = t[n0 + bi, h, w]
    return out


@pytest.mark.parametrize("n,ho,wo,cin,cout", [(2, 4, 6, 3, 5), (3, 7, 7, 4, 4), (1, 14, 14, 2, 3)])
def test_stride2_fprop_replay_matches_conv2d(n, ho, wo, cin, cout):
    """The stride-2 mode of the persistent conv kernel (csrc/gemm_persist.cu): output tile = BH rows x Wo columns; for
    tap (r, s) the A operand is the input box starting at row 2*h0 + r - 1, column s - 1, spanning 2*BH x 2*Wo source
    elements with traversal stride 2."""
    torch.manual_seed(0)
    x = torch.randn(n, 2 * ho, 2 * wo, cin, dtype=torch.float64)            # NHWC input, twice the output size
    wt = torch.randn(cout, 3, 3, cin, dtype=torch.float64)                  # KRSC
    y = torch.zeros(n, ho, wo, cout, dtype=torch.float64)
    bh = 2 if ho % 2 == 0 else 1                                            # any row tiling works the same way
    for img in range(n):
        for h0 in range(0, ho, bh):
            rows = min(bh, ho - h0)
            acc = torch.zeros(bh * wo, cout, dtype=torch.float64)
            for r in range(3):
                for s in range(3):
                    a = _strided_box(x, img, 1, 2 * h0 + r - 1, 2 * bh, s - 1, 2 * wo, 2).reshape(bh * wo, cin)
                    acc += a @ wt[:, r, s, :].t()
            y[img, h0:h0 + rows] = acc.reshape(bh, wo, cout)[:rows]
    ref = F.conv2d(x.permute(0, 3, 1, 2), wt.permute(0, 3, 1, 2), None, 2, 1).permute(0, 2, 3, 1)
    assert torch.allclose(y, ref, atol=1e-9)
