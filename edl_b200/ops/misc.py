"""RoPE, embedding-bag (CTR), uint8 image normalisation, dynamic loss scaling (csrc/misc.cu)."""
from __future__ import annotations

import torch


# --------------------------------------------------------------------------------------------- RoPE
def rope_tables(seq_len, head_dim, base=10000.0, device=None):
    inv = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.float32, device=device) / head_dim))
    ang = torch.arange(seq_len, dtype=torch.float32, device=device)[:, None] * inv[None, :]
    return torch.cos(ang).contiguous(), torch.sin(ang).contiguous()


def _rope_ref(x, cos, sin, inverse=False):
    half = x.shape[-1] // 2
    x1, x2 = x[..., :half].float(), x[..., half:].float()
    c, s = cos[:, None, :], (-sin if inverse else sin)[:, None, :]
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], -1).to(x.dtype)


class _RopeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cos, sin):
        from . import native, count_launch

        x = x.contiguous()
        y = torch.empty_like(x)
        native().rope(x, cos, sin, y, False)
        count_launch()
        ctx.save_for_backward(cos, sin)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import native, count_launch

        cos, sin = ctx.saved_tensors
        dx = torch.empty_like(dy)
        native().rope(dy.contiguous(), cos, sin, dx, True)
        count_launch()
        return dx, None, None


def rope(x, cos, sin):
    """Rotary position embedding (rotate-half).  x: [T, H, D] bf16; cos / sin: [T, D/2] fp32."""
    if x.is_cuda and x.dtype == torch.bfloat16:
        return _RopeFn.apply(x, cos, sin)
    return _rope_ref(x, cos, sin)


# --------------------------------------------------------------------------------------------- embedding bag
class _EmbBagFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, ids):
        from . import native, count_launch

        ids = ids.contiguous()
        out = torch.empty(ids.shape[0], table.shape[1], device=table.device, dtype=table.dtype)
        native().embedding_bag_fwd(table, ids, out)
        count_launch()
        ctx.save_for_backward(ids)
        ctx.shape, ctx.dtype = table.shape, table.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        from . import native, count_launch

        (ids,) = ctx.saved_tensors
        dtable = torch.zeros(ctx.shape, device=dout.device, dtype=torch.float32)
        native().embedding_bag_bwd(dout.contiguous(), ids, dtable)
        count_launch()
        return dtable.to(ctx.dtype), None


def embedding_bag_mean(table, ids):
    """out[b] = mean_l table[ids[b, l]] -- ``embedding(is_sparse) + sequence_pool(avg)`` of the CTR model.
    table [V, D] (fp32 / bf16), ids [B, L] int64.  The gradient is a dense table (all-reduced like any
    other parameter in elastic DP)."""
    if table.is_cuda:
        return _EmbBagFn.apply(table, ids)
    return table[ids].float().mean(1).to(table.dtype)


# --------------------------------------------------------------------------------------------- input pipeline
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def normalize_u8(x_u8, mean=IMAGENET_MEAN, std=IMAGENET_STD, flip=None):
    """uint8 NHWC [N, H, W, 3] -> bf16 NCHW-shaped channels_last tensor, (x/255 - mean)/std, optional
    per-image horizontal flip (uint8 [N] mask).  Shipping uint8 halves the H2D bytes per step."""
    from . import native, count_launch

    n, h, w, _ = x_u8.shape
    if x_u8.is_cuda:
        y = torch.empty((n, h, w, 3), device=x_u8.device, dtype=torch.bfloat16)
        native().normalize_u8(x_u8.contiguous(), y, list(mean), list(std), flip)
        count_launch()
        return y.permute(0, 3, 1, 2)
    xf = x_u8.float() / 255.0
    if flip is not None:
        xf = torch.where(flip.bool()[:, None, None, None], xf.flip(2), xf)
    y = (xf - torch.tensor(mean)) / torch.tensor(std)
    return y.to(torch.bfloat16).permute(0, 3, 1, 2)


# --------------------------------------------------------------------------------------------- AMP loss scaling
class DynamicLossScaler:
    """fp16-style dynamic loss scaling (reference: ``mixed_precision.decorate(init_loss_scaling,
    use_dynamic_loss_scaling)``, example/distill/resnet/train_with_fleet.py:324-327): the loss is
    multiplied by ``scale``; the fused all-reduce raises ``found_inf``; the fused optimizer divides
    gradients by ``scale`` (device scalar ``inv_scale``) and skips the step on overflow; ``update()``
    halves the scale on overflow and doubles it after ``growth_interval`` clean steps.  bf16
    training does not need it (the default path), fp16 does."""

    def __init__(self, device, init_scale=2.0 ** 15, growth_factor=2.0, backoff_factor=0.5, growth_interval=1000):
        self.scale = torch.full((1,), float(init_scale), dtype=torch.float32, device=device)
        self.inv_scale = 1.0 / self.scale
        self.found_inf = torch.zeros(1, dtype=torch.int32, device=device)
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, growth_interval
        self._good = torch.zeros(1, dtype=torch.int32, device=device)

    def scale_loss(self, loss):
        return loss * self.scale

    def attach(self, optimizer, dp=None):
        optimizer.set_grad_scale(self.inv_scale)
        optimizer.set_found_inf(self.found_inf)
        if dp is not None:
            dp.found_inf = self.found_inf

    @torch.no_grad()
    def update(self):
        """Device-only bookkeeping (graph capturable): no host sync."""
        bad = self.found_inf > 0
        self._good.copy_(torch.where(bad, torch.zeros_like(self._good), self._good + 1))
        grow = self._good >= self.growth_interval
        new = torch.where(bad, self.scale * self.backoff_factor,
                          torch.where(grow, self.scale * self.growth_factor, self.scale))
        self.scale.copy_(new.clamp_(1.0, 2.0 ** 24))
        self._good.copy_(torch.where(grow, torch.zeros_like(self._good), self._good))
        self.inv_scale.copy_(1.0 / self.scale)
        self.found_inf.zero_()
