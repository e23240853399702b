#!/bin/bash
# 2 GPUs: stride-2 backward on own kernels, wgrad upper-bound experiment, in-place hot recovery after the example fix
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 400 python -m pytest tests/test_round2_gpu.py -q --timeout 240 -k "stride2 or conv3x3_wgrad" > gpurun_out/s2bwd_tests.log 2>&1
echo "stride-2 / wgrad tests: exit $? $(tail -1 gpurun_out/s2bwd_tests.log)"
b() { local tag=$1; shift
  timeout 300 env "$@" python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e > gpurun_out/b7_$tag.json 2> gpurun_out/b7_$tag.err
  python - gpurun_out/b7_$tag.json $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("bench %-16s %.0f img/s  %.3f ms/step  launches %s  fallbacks %d" % (sys.argv[2], d["value"], d["ms_per_step"], d.get("gpu_launches"), len(d.get("library_fallbacks") or {})))
except Exception as e:
    print("bench %s: no result (%s)" % (sys.argv[2], e))
P
}
b base A=1
b s2lib EDL_OWN_S2_BWD=0
b skipwgrad EDL_DEBUG_SKIP_WGRAD=1
for mc in 0 1; do
  EDL_DISABLE_MULTICAST=$mc timeout 700 python tools/bench_elastic_launch.py --native-store --trainer resnet --gpus-per-pod 1 --leave kill \
       --out gpurun_out/elastic_launch_2gpu_kill_nomc$mc.json > gpurun_out/elastic_launch_2gpu_kill_nomc$mc.log 2>&1
  echo "elastic kill (EDL_DISABLE_MULTICAST=$mc): exit $?"; tail -n 2 gpurun_out/elastic_launch_2gpu_kill_nomc$mc.log
  rm -rf gpurun_out/elastic_fail_nomc$mc; [ -d gpurun_out/elastic_fail_inplace_kill ] && mv gpurun_out/elastic_fail_inplace_kill gpurun_out/elastic_fail_nomc$mc
done
