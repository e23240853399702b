#!/bin/bash
# Round-2 second GPU call (1 GPU): re-run the experimental tests that failed on tolerance with run-to-run noise
# yardsticks, A/B the remaining switches, 3x3-wgrad microbench, kineto of the default step.
set -u
mkdir -p gpurun_out
export EDL_TEST_EXPERIMENTAL=1
S=gpurun_out/call2_summary.txt
: > "$S"
python -c 'import torch' 2> /dev/null
t() { # name expr
  timeout 300 python -m pytest tests/test_round2_gpu.py -q --timeout 240 -k "$2" > "gpurun_out/exp2_$1.log" 2>&1
  echo "tests $1: exit $?  $(tail -1 gpurun_out/exp2_$1.log)" | tee -a "$S"
}
t pdl "programmatic_dependent_launch"
t bnmodel "fused_bn_backward_matches"
t pipeline "step_pipelined"
t jpeg "crop_resize_normalize or nvjpeg"
timeout 300 python tools/bench_loader.py --images 2048 --threads 8 > gpurun_out/loader_bench.jsonl 2> gpurun_out/loader_bench.err
sed 's/^/loader: /' gpurun_out/loader_bench.jsonl | tee -a "$S"
timeout 300 python tools/bench_wgrad3.py > gpurun_out/wgrad3.log 2>&1; sed 's/^/wgrad3: /' gpurun_out/wgrad3.log | tail -8 | tee -a "$S"
bench() { local tag=$1; shift
  timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 "$@" > "gpurun_out/ab2_$tag.json" 2> "gpurun_out/ab2_$tag.err"
  echo "bench $tag: exit $? $(python - "gpurun_out/ab2_$tag.json" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%.0f img/s  %.3f ms/step  e2e %.0f  launches %s" % (d["value"], d["ms_per_step"], e.get("value") or 0, d.get("gpu_launches")))
except Exception as ex:
    print("no result (%s)" % ex)
P
)" | tee -a "$S"
}
bench base --kineto gpurun_out/kineto_r2_base.txt
bench pdl --pdl
bench s2stem --conv3-s2 --own-stem1
bench bnbwd2 --fuse-bn-bwd 2
bench bnbwd2_pdl --fuse-bn-bwd 2 --pdl
bench all --fuse-bn-bwd 2 --pdl --conv3-s2 --own-stem1
cat "$S"
