#!/bin/bash
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 400 python -m pytest tests/test_round2_gpu.py tests/test_model_gpu.py -q --timeout 300 -x -k "stem_conv or pixel_pair or model" > gpurun_out/c20_tests.log 2>&1
echo "tests: exit $? $(tail -1 gpurun_out/c20_tests.log)"; grep -E "Error|assert" gpurun_out/c20_tests.log | head -5
b() { local tag=$1; shift
  timeout 300 env ${ENVV:-A=1} python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e "$@" > gpurun_out/b20_$tag.json 2> gpurun_out/b20_$tag.err
  python - gpurun_out/b20_$tag.json $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("bench %-14s %.0f img/s  %.3f ms/step  launches %s  fallbacks %d" % (sys.argv[2], d["value"], d["ms_per_step"], d.get("gpu_launches"), len(d.get("library_fallbacks") or {})))
except Exception as e:
    print("bench %s: no result (%s)" % (sys.argv[2], e))
P
}
b base --kineto gpurun_out/kineto_r2_c20.txt
ENVV="EDL_OWN_STEM23=1" b stem23
b nolib --no-library
grep -E "stem_wgrad" gpurun_out/kineto_r2_c20.txt | cut -c1-60,180-260
