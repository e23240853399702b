#!/bin/bash
# compute-sanitizer passes over the kernel tests (memcheck + racecheck + synccheck), one small subset at a time so a
# hang costs one timeout, not the call.  The reference has no sanitizer runs.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize.sh'
# Summaries land in gpurun_out/sanitize_*.log and gpurun_out/sanitizer_summary.txt (copied to profiles/sanitizer.txt).
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SUM=gpurun_out/sanitizer_summary.txt
: > "$SUM"
python -c 'import torch' 2> /dev/null
run() {  # name tool budget pytest-args...
  local name=$1 tool=$2 budget=$3; shift 3
  timeout "$budget" $SAN --tool "$tool" --print-limit 10 --error-exitcode 9 \
    python -m pytest "$@" -x -q -p no:cacheprovider --timeout $((budget - 20)) > "gpurun_out/sanitize_${name}_${tool}.log" 2>&1
  echo "$name / $tool: exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' "gpurun_out/sanitize_${name}_${tool}.log" | tail -2 | tr '\n' ' ')" | tee -a "$SUM"
}
# SUBSET=new: only the kernels added at the end of round 2 (CTA-pair GEMM, halo / resident-weight convolutions, stem im2col,
# tensor-core statistics), with small shapes
if [ "${SUBSET:-all}" = "new" ]; then
  run new_kernels memcheck 150 tests/test_persist_gpu.py -k "(cta_pair and 19000) or (haloed and (11-20 or 6-6 or 9-30)) or (stem7 and 65) or (tensor_core_bn and (shape1 or shape5)) or phase_trace"
  cat "$SUM"
  exit 0
fi
run gemm_wgrad   memcheck  200 tests/test_gemm_gpu.py -k "wgrad or split"
run persist      memcheck  200 tests/test_persist_gpu.py -k "wide_tile or epilogue or bnr"
run wgrad3       memcheck  200 tests/test_round2_gpu.py -k "conv3x3_wgrad_tcgen05 and 14-14"
run wgrad3       racecheck 200 tests/test_round2_gpu.py -k "conv3x3_wgrad_tcgen05 and 14-14"
run stem_s2      memcheck  200 tests/test_round2_gpu.py -k "stem_conv or (stride2 and 28-28)"
run bn           memcheck  200 tests/test_kernels_gpu.py -k "bn and not fused"
run optim_loss   memcheck  200 tests/test_kernels_gpu.py -k "sgd or adam or soft_ce or pool or rope"
run conv3        synccheck 200 tests/test_conv3x3_gpu.py
cat "$SUM"
