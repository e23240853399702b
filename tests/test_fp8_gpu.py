"""e4m3 tcgen05 GEMM + quantiser (csrc/gemm_fp8.cu) and the fp8 teacher mode."""
import pytest
import torch

from edl_b200.ops import fp8

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (300, 256, 512), (6272, 2048, 1024), (32, 1000, 2048), (257, 64, 144)])
def test_gemm_fp8_matches_dequantised_reference(m, n, k):
    torch.manual_seed(0)
    a = torch.randn(m, k, device=DEV).bfloat16()
    w = torch.randn(n, k, device=DEV) * 0.05
    w8, ws = fp8.quantize_weight_rows(w)
    amax = torch.zeros(1, device=DEV)
    act_scale = (a.float().abs().max() / fp8.E4M3_MAX).view(1)
    a8 = fp8.quantize_e4m3(a, act_scale, amax)
    assert abs(float(amax) - float(a.float().abs().max())) < 1e-6
    # the quantiser agrees with torch's e4m3 cast
    ref8 = (a.float() / act_scale).to(torch.float8_e4m3fn)
    assert (a8.view(torch.float8_e4m3fn).float() - ref8.float()).abs().max() == 0
    shift = torch.randn(n, device=DEV)
    d = fp8.gemm_fp8(a8, w8, (act_scale * ws).contiguous(), shift, relu=True)
    exact = torch.relu((a8.view(torch.float8_e4m3fn).float() @ w8.view(torch.float8_e4m3fn).float().t()) * (act_scale * ws) + shift)
    assert _rel(d, exact) < 1e-2                                   # kernel vs the same fp8 operands in fp32
    full = torch.relu(a.float() @ w.t() + shift)
    assert _rel(d, full) < 8e-2                                    # fp8 quantisation error budget


def test_teacher_fp8_mode_tracks_bf16_logits():
    from edl_b200.models.resnext import ResNeXt50_32x4d, to_inference_dtype

    torch.manual_seed(0)
    m = to_inference_dtype(ResNeXt50_32x4d(class_dim=100), torch.bfloat16, DEV).eval()
    x = torch.randn(4, 3, 128, 128, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    ref = m(x).float()
    n = m.enable_fp8(x)
    assert n >= 30
    out = m(x).float()
    cos = torch.nn.functional.cosine_similarity(out.flatten(), ref.flatten(), dim=0).item()
    assert cos > 0.98, cos
    assert (out.argmax(-1) == ref.argmax(-1)).float().mean() >= 0.75
