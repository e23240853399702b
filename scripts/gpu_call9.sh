#!/bin/bash
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_round2_gpu.py -q --timeout 300 -x -k "wgrad or stride2 or gemm or fused_bn" > gpurun_out/c9_tests.log 2>&1
echo "tests: exit $? $(tail -1 gpurun_out/c9_tests.log)"
b() { local tag=$1; shift
  timeout 300 env "$@" python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e ${EXTRA:-} > gpurun_out/b9_$tag.json 2> gpurun_out/b9_$tag.err
  python - gpurun_out/b9_$tag.json $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("bench %-22s %.0f img/s  %.3f ms/step  launches %s  fallbacks %d" % (sys.argv[2], d["value"], d["ms_per_step"], d.get("gpu_launches"), len(d.get("library_fallbacks") or {})))
except Exception as e:
    print("bench %s: no result (%s)" % (sys.argv[2], e))
P
}
EXTRA="--kineto gpurun_out/kineto_r2_c9.txt" b base A=1
EXTRA="--own-wgrad3 --kineto gpurun_out/kineto_r2_c9_own.txt" b ownwgrad3 A=1
python tools/trace_timeline.py gpurun_out/kineto_r2_c9.txt.trace.json > gpurun_out/timeline_r2_c9.txt 2>&1
tail -25 gpurun_out/timeline_r2_c9.txt
