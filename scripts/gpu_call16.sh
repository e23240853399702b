#!/bin/bash
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 400 python -m pytest tests/test_round2_gpu.py tests/test_persist_gpu.py tests/test_conv3x3_gpu.py -q --timeout 300 -x -k "stride2 or persistent_conv or conv3x3" > gpurun_out/c16_tests.log 2>&1
echo "tests: exit $? $(tail -1 gpurun_out/c16_tests.log)"
b() { local tag=$1; shift
  timeout 300 env ${ENVV:-A=1} python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e "$@" > gpurun_out/b16_$tag.json 2> gpurun_out/b16_$tag.err
  python - gpurun_out/b16_$tag.json $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("bench %-14s %.0f img/s  %.3f ms/step  launches %s  fallbacks %d" % (sys.argv[2], d["value"], d["ms_per_step"], d.get("gpu_launches"), len(d.get("library_fallbacks") or {})))
except Exception as e:
    print("bench %s: no result (%s)" % (sys.argv[2], e))
P
}
b base
ENVV="EDL_OWN_S2_BWD=1" b s2own
ENVV="EDL_OWN_S2_BWD=1 EDL_OWN_STEM23=1" b s2own_stem23
b nolib --no-library
