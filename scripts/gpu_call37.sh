#!/bin/bash
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 600 python -m pytest tests/test_persist_gpu.py tests/test_conv3x3_gpu.py tests/test_gemm_gpu.py -q --timeout 300 -x > gpurun_out/c37_tests.log 2>&1
rc=$?
echo "tests: exit $rc $(tail -1 gpurun_out/c37_tests.log)"; grep -E "^E  |Error" gpurun_out/c37_tests.log | head -12
if [ $rc -ne 0 ]; then exit 0; fi
b() { local tag=$1; shift
  timeout 300 env ${ENVV:-A=1} python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e "$@" > gpurun_out/b37_$tag.json 2> gpurun_out/b37_$tag.err
  python - gpurun_out/b37_$tag.json $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("bench %-14s %.0f img/s  %.3f ms/step  launches %s  fallbacks %d" % (sys.argv[2], d["value"], d["ms_per_step"], d.get("gpu_launches"), len(d.get("library_fallbacks") or {})))
except Exception as e:
    print("bench %s: no result (%s)" % (sys.argv[2], e))
P
}
b x64 --no-extras --kineto gpurun_out/kineto_r2_c37.txt
ENVV="EDL_EPI_WARPS=16" b x64epi16 --no-extras
(timeout 120 python tools/trace_persist.py fwd 100352 64 256; timeout 120 python tools/trace_persist.py dgradbn 100352 64 256 y) > gpurun_out/trace_persist_c37.txt 2>&1
grep -E "==|epilogue tile  [45]" gpurun_out/trace_persist_c37.txt | head -12
timeout 200 python tools/teacher_prof.py > gpurun_out/teacher_c37.txt 2>&1; grep -E "^forward|gemm_persist|gemm_pair" gpurun_out/teacher_c37.txt | cut -c1-100,150-260 | head -4
timeout 300 python -m pytest tests/test_round2_gpu.py tests/test_model_gpu.py -q --timeout 300 -x -k "model or bnmodel or teacher" > gpurun_out/c37_tests2.log 2>&1
echo "tests2: exit $? $(tail -1 gpurun_out/c37_tests2.log)"
