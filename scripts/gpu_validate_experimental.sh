#!/bin/bash
# First GPU call of a round: validate what was written without a GPU, then A/B it on the flagship bench.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_validate_experimental.sh'
# Results land in gpurun_out/ (merged back by gpurun).  Nothing here changes defaults: read
# gpurun_out/experimental_summary.txt, then flip the defaults (ops/gemm.py, csrc/bn.cu, ...) of what passed AND paid.
#
# One pytest process PER FEATURE: an illegal address in one kernel poisons its CUDA context for the rest of the
# process, and a hang must only cost that feature's timeout.  A feature whose tests fail is left out of the combined
# A/B runs at the end.
set -u
mkdir -p gpurun_out
export EDL_TEST_EXPERIMENTAL=1
SUMMARY=gpurun_out/experimental_summary.txt
: > "$SUMMARY"
python -c 'import torch' 2> /dev/null            # page the image in once, outside every timeout

declare -A OK
run_group() {   # name, per-group timeout, -k expression
  local name=$1 tmo=$2 expr=$3
  timeout "$tmo" python -m pytest tests/test_experimental_gpu.py -q --timeout 240 -k "$expr" \
      > "gpurun_out/exp_$name.log" 2>&1
  local rc=$?
  OK[$name]=$rc
  echo "tests $name: exit $rc  $(tail -1 "gpurun_out/exp_$name.log")" | tee -a "$SUMMARY"
}
run_group wgrad3   400 "conv3x3_wgrad or own_wgrad"
run_group pdl      300 "programmatic_dependent_launch"
run_group conv3s2  300 "stride2_fprop"
run_group stem1    200 "stem_conv_direct"
run_group bnr2     400 "bnr_mode2 or bnr_many_tiles"
run_group bnmodel  300 "fused_bn_backward_matches"
run_group teacher  300 "teacher_residual"
run_group pipeline 200 "step_pipelined"
run_group jpeg     300 "crop_resize_normalize or nvjpeg"
if [ "${OK[jpeg]}" = 0 ]; then
  timeout 300 python tools/bench_loader.py --images 2048 --threads 8 > gpurun_out/loader_bench.jsonl 2> gpurun_out/loader_bench.err
  sed 's/^/loader: /' gpurun_out/loader_bench.jsonl | tee -a "$SUMMARY"
fi

bench() {       # tag, flags...
  local tag=$1; shift
  timeout 300 python bench.py --gpus 1 --steps 60 --warmup 5 "$@" > "gpurun_out/ab_$tag.json" 2> "gpurun_out/ab_$tag.err"
  echo "bench $tag: exit $? $(python - "gpurun_out/ab_$tag.json" <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%.0f img/s  %.3f ms/step  e2e %s  pipelined %s  launches %s" % (
        d["value"], d["ms_per_step"], e.get("value"), e.get("pipelined"), d.get("gpu_launches")))
except Exception as ex:                      # noqa: BLE001
    print("no result (%s)" % ex)
E
)" | tee -a "$SUMMARY"
}
bench base
combined=()
[ "${OK[pdl]}" = 0 ]     && { bench pdl --pdl;               combined+=(--pdl); }
[ "${OK[wgrad3]}" = 0 ]  && { bench ownwgrad3 --own-wgrad3;  combined+=(--own-wgrad3); }
[ "${OK[conv3s2]}" = 0 ] && { bench conv3s2 --conv3-s2;      combined+=(--conv3-s2); }
[ "${OK[stem1]}" = 0 ]   && { bench ownstem1 --own-stem1;    combined+=(--own-stem1); }
[ "${OK[bnr2]}" = 0 ]    && bench fusebnbwd2 --fuse-bn-bwd 2      # speed only: the model-level mismatch is "bnmodel"
[ ${#combined[@]} -gt 1 ] && bench combined "${combined[@]}"
[ ${#combined[@]} -gt 0 ] && [ "${OK[bnr2]}" = 0 ] && [ "${OK[bnmodel]}" = 0 ] && bench combined_bnbwd "${combined[@]}" --fuse-bn-bwd 2
[ "${OK[teacher]}" = 0 ] && echo "teacher residual fusion passed: A/B it with scripts/gpu_multi_metrics.sh 2 distill" | tee -a "$SUMMARY"
cat "$SUMMARY"
