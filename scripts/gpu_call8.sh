#!/bin/bash
# 1 GPU: split-K without atomics (partials + reduce kernel), stem wgrad kernel, A/B matrix
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_round2_gpu.py tests/test_conv3x3_gpu.py tests/test_model_gpu.py -q --timeout 300 -x > gpurun_out/c8_tests.log 2>&1
echo "tests: exit $? $(tail -1 gpurun_out/c8_tests.log)"
b() { local tag=$1; shift
  timeout 300 env "$@" python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e ${EXTRA:-} > gpurun_out/b8_$tag.json 2> gpurun_out/b8_$tag.err
  python - gpurun_out/b8_$tag.json $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("bench %-22s %.0f img/s  %.3f ms/step  launches %s  fallbacks %d" % (sys.argv[2], d["value"], d["ms_per_step"], d.get("gpu_launches"), len(d.get("library_fallbacks") or {})))
except Exception as e:
    print("bench %s: no result (%s)" % (sys.argv[2], e))
P
}
b base A=1
b no_partials EDL_SPLITK_PARTIALS=0
b s2lib EDL_OWN_S2_BWD=0
EXTRA="--own-wgrad3 --kineto gpurun_out/kineto_r2_c8.txt" b ownwgrad3 A=1
EXTRA="--own-wgrad3" b ownwgrad3_s2lib EDL_OWN_S2_BWD=0
b skipwgrad EDL_DEBUG_SKIP_WGRAD=1
