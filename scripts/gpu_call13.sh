#!/bin/bash
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 400 python -m pytest tests/test_round2_gpu.py -q --timeout 300 -x -k "pixel_pair or stem_conv" > gpurun_out/c13_tests.log 2>&1
echo "tests: exit $? $(tail -1 gpurun_out/c13_tests.log)"
b() { local tag=$1; shift
  timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 "$@" > gpurun_out/b13_$tag.json 2> gpurun_out/b13_$tag.err
  python - gpurun_out/b13_$tag.json $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("bench %-14s %.0f img/s  %.3f ms/step  e2e %.0f  launches %s  fallbacks %s" % (sys.argv[2], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value") or 0, d.get("gpu_launches"), sorted(d.get("library_fallbacks") or {})[:3]))
except Exception as e:
    print("bench %s: no result (%s)" % (sys.argv[2], e))
P
}
b base
b nolib --no-library --kineto gpurun_out/kineto_r2_nolib.txt
EDL_OWN_STEM23=1 b stem23only --no-e2e
