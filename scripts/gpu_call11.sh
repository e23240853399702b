#!/bin/bash
# 1 GPU: teacher path (widened small groups, 128x256 tiles), persistent GEMM tests, teacher profile
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 600 python -m pytest tests/test_persist_gpu.py tests/test_model_gpu.py tests/test_fp8_gpu.py -q --timeout 300 -x > gpurun_out/c11_tests.log 2>&1
echo "tests: exit $? $(tail -1 gpurun_out/c11_tests.log)"
timeout 300 python tools/teacher_prof.py > gpurun_out/teacher_r2.txt 2>&1; grep -m1 "^forward" gpurun_out/teacher_r2.txt
EDL_GEMM_WIDE=0 timeout 300 python tools/teacher_prof.py > gpurun_out/teacher_r2_nowide.txt 2>&1; grep -m1 "^forward" gpurun_out/teacher_r2_nowide.txt
python - <<'P'
import torch
from edl_b200 import ops
from edl_b200.models.resnext import ResNeXt101_32x16d, to_inference_dtype
import torch.nn.functional as F
# widened small groups against the library on the real layer shapes
torch.manual_seed(0)
from edl_b200.models.resnext import FoldedConv
for c, g, hw, st in ((512, 32, 56, 1), (1024, 32, 28, 1), (1024, 32, 56, 2)):
    m = FoldedConv(c, c, 3, stride=st, groups=g).cuda().to(torch.bfloat16)
    m.scale = m.scale.float(); m.shift = m.shift.float()
    x = torch.randn(8, c, hw, hw, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    ops.reset_fallbacks()
    y = m(x)
    ref = torch.relu(F.conv2d(x.float(), m.weight.permute(0, 3, 1, 2).float(), None, st, 1, 1, g))
    err = ((y.float() - ref).norm() / ref.norm()).item()
    print("widened groups C=%d g=%d %dx%d s%d: rel err %.4f  fallbacks %s" % (c, g, hw, hw, st, err, list(ops.fallbacks())))
    assert err < 1e-2 and not ops.fallbacks()
P
