#!/bin/bash
# A/B: size threshold for the fused BN-backward reduction; balanced runs check
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
b() { local tag=$1; shift
  timeout 300 env ${ENVV:-A=1} python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e "$@" > gpurun_out/b27_$tag.json 2> gpurun_out/b27_$tag.err
  python - gpurun_out/b27_$tag.json $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("bench %-14s %.0f img/s  %.3f ms/step  launches %s  fallbacks %d" % (sys.argv[2], d["value"], d["ms_per_step"], d.get("gpu_launches"), len(d.get("library_fallbacks") or {})))
except Exception as e:
    print("bench %s: no result (%s)" % (sys.argv[2], e))
P
}
b base --no-extras --kineto gpurun_out/kineto_r2_c27.txt
ENVV="EDL_FUSE_BN_BWD_MAX_MB=40" b max40 --no-extras --kineto gpurun_out/kineto_r2_c27_max40.txt
ENVV="EDL_FUSE_BN_BWD_MAX_MB=20" b max20 --no-extras
ENVV="EDL_FUSE_BN_BWD_MAX_MB=10" b max10 --no-extras
timeout 300 python -m pytest tests/test_persist_gpu.py -q --timeout 300 -x -k "halo or resident or bn_backward" > gpurun_out/c27_tests.log 2>&1
echo "tests: exit $? $(tail -1 gpurun_out/c27_tests.log)"
