"""FP8 (e4m3) inference path for the teacher (csrc/gemm_fp8.cu): per-tensor activation scales from a
calibration pass ("delayed scaling"), per-output-channel weight scales, dequantisation folded into the
GEMM epilogue together with the folded BatchNorm scale / shift and the ReLU.

The reference serves the teacher in fp32 through Paddle Serving (example/distill/resnet/scripts/
start_local_teacher.sh:24-30); SURVEY K12 asks for an fp8 tcgen05 path on Blackwell."""
from __future__ import annotations

from typing import Optional

import torch

E4M3_MAX = 448.0


def quantize_weight_rows(w2: torch.Tensor):
    """w2 [N, K] float/bf16 -> (uint8 e4m3 bytes [N, K], fp32 scale [N])"""
    w = w2.float()
    scale = (w.abs().amax(1).clamp_min(1e-12) / E4M3_MAX)
    q = (w / scale[:, None]).to(torch.float8_e4m3fn).view(torch.uint8).contiguous()
    return q, scale


def quantize_e4m3(x: torch.Tensor, scale: torch.Tensor, amax: Optional[torch.Tensor] = None, out=None):
    """bf16 tensor -> e4m3 bytes with a per-tensor scale (device scalar); optionally tracks max|x|."""
    from . import native, count_launch

    x = x.contiguous()
    q = out if out is not None else torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    if x.is_cuda:
        native().quantize_e4m3(x, q, scale, amax)
        count_launch()
    else:
        if amax is not None:
            amax.copy_(torch.maximum(amax, x.float().abs().max()))
        q.copy_((x.float() / scale).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).view(torch.uint8))
    return q


def gemm_fp8(a8: torch.Tensor, b8: torch.Tensor, col_scale=None, col_shift=None, relu=False, out=None):
    """out bf16 [M, N] = relu?((a8 [M,K] @ b8 [N,K]^T) * col_scale + col_shift); a8 / b8 hold e4m3 bytes."""
    from . import native, count_launch

    m, n = a8.shape[0], b8.shape[0]
    if out is None:
        out = torch.empty((m, n), device=a8.device, dtype=torch.bfloat16)
    if a8.is_cuda:
        native().gemm_fp8(a8, b8, out, col_scale, col_shift, relu)
        count_launch()
        return out
    d = a8.view(torch.float8_e4m3fn).float() @ b8.view(torch.float8_e4m3fn).float().t()
    if col_scale is not None:
        d = d * col_scale
    if col_shift is not None:
        d = d + col_shift
    out.copy_(torch.relu(d) if relu else d)
    return out
