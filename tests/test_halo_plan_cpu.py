"""CPU replay of the haloed-tile index math of the 3x3 convolutions (csrc/gemm_persist.cu, kHalo): with the REAL planner
(`conv3x3_halo_plan`) the test rebuilds what the kernel does -- one pixel box per filter row with an extra zero column on
the left, the three taps as whole-row shifts of that tile, accumulator rows of the halo column dropped while compacting --
in NumPy and compares with the convolution.  Mirrors tests/test_wgrad_plan_cpu.py for the weight-gradient plans."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from edl_b200 import ops

BLOCK_M = 128


def _replay(x, w, dgrad):
    """x [N,H,W,C] float, w [K,3,3,C] (KRSC).  fprop: y[n,h,w,k] = sum x[n,h+r-1,w+s-1,c] w[k,r,s,c];
    dgrad (x = dy [N,H,W,K]): dx[n,h,w,c] = sum dy[n,h+1-r,w+1-s,k] w[k,r,s,c]."""
    n, h, wd, cin = x.shape
    plan = ops.native().conv3x3_halo_plan(n, h, wd)
    assert plan, "shape not covered by the halo layout"
    bh, bn, tiles_h, tiles_img = plan
    wb = wd + 1
    rows_in = bn * bh * wb
    assert rows_in <= BLOCK_M
    cout = w.shape[3] if dgrad else w.shape[0]
    out = np.zeros((n, h, wd, cout), dtype=np.float64)
    written = np.zeros((n, h, wd), dtype=np.int32)
    for ti in range(tiles_img):
        for th in range(tiles_h):
            img0, h0 = ti * bn, th * bh
            acc = np.zeros((BLOCK_M, cout))
            for r in range(3):
                dh = 1 - r if dgrad else r - 1
                # A tile: 1 KB zero pad (8 rows) | 128 tile rows | pad; TMA fills the first rows_in rows, zeros outside the tensor
                tile = np.zeros((8 + BLOCK_M + 8, cin))
                for b in range(bn):
                    for hr in range(bh):
                        for col in range(wb):                       # box starts at column -1
                            ni, hi, wi = img0 + b, h0 + hr + dh, col - 1
                            if ni < n and 0 <= hi < h and 0 <= wi < wd:
                                tile[8 + (b * bh + hr) * wb + col] = x[ni, hi, wi]
                for s in range(3):
                    shift = 1 - s if dgrad else s - 1               # descriptor start shifted by whole 128-byte rows
                    a = tile[8 + shift:8 + shift + BLOCK_M]
                    wt = w[:, r, s, :]                              # [K, C]
                    acc += a @ (wt if dgrad else wt.T)
            for m in range(rows_in):                                # epilogue: drop the halo rows, compact
                ir, cw = divmod(m, wb)
                if cw == 0:
                    continue
                b, hr = divmod(ir, bh)
                ni, hi = img0 + b, h0 + hr
                if ni < n and hi < h:                               # the TMA store clips rows outside the tensor
                    out[ni, hi, cw - 1] = acc[m]
                    written[ni, hi, cw - 1] += 1
    assert (written == 1).all(), "every output pixel is produced by exactly one tile row"
    return out


@pytest.mark.parametrize("n,c,k,h,w", [(3, 4, 5, 56, 56), (5, 3, 2, 11, 20), (4, 2, 3, 14, 14), (7, 2, 2, 7, 7), (1, 3, 4, 9, 30),
                                       (2, 2, 2, 5, 127), (9, 2, 2, 6, 6), (2, 2, 3, 33, 64)])
def test_halo_tiles_reproduce_the_convolution(n, c, k, h, w):
    rng = np.random.RandomState(0)
    x = rng.randn(n, h, w, c)
    wt = rng.randn(k, 3, 3, c)
    y = _replay(x, wt, dgrad=False)
    ref = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(wt).permute(0, 3, 1, 2), None, 1, 1)
    assert np.allclose(y, ref.permute(0, 2, 3, 1).numpy(), atol=1e-9)
    dy = rng.randn(n, h, w, k)
    dx = _replay(dy, wt, dgrad=True)
    dref = torch.nn.grad.conv2d_input((n, c, h, w), torch.from_numpy(wt).permute(0, 3, 1, 2),
                                      torch.from_numpy(dy).permute(0, 3, 1, 2), 1, 1)
    assert np.allclose(dx, dref.permute(0, 2, 3, 1).numpy(), atol=1e-9)


def test_halo_plan_limits():
    nat = ops.native()
    assert nat.conv3x3_halo_plan(2, 33, 128) == []          # W + 1 columns do not fit the 128-row tile: shifted-box version
    bh, bn, th, ti = nat.conv3x3_halo_plan(32, 7, 7)
    assert bh == 7 and bn == 2 and th == 1 and ti == 16      # two 7 x 8 images per tile
    for n, h, w in [(32, 56, 56), (32, 28, 28), (32, 14, 14), (5, 11, 20)]:
        bh, bn, th, ti = nat.conv3x3_halo_plan(n, h, w)
        assert bn * bh * (w + 1) <= BLOCK_M and th * bh >= h and ti * bn >= n
