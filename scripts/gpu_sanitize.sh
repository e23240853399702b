#!/bin/bash
# compute-sanitizer passes over the kernel tests (memcheck + racecheck + synccheck), one small subset at a time so a
# hang costs one timeout, not the call.  The reference has no sanitizer runs.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_sanitize.sh'
# Summaries land in gpurun_out/sanitize_*.log; copy the ERROR SUMMARY lines into profiles/sanitizer.txt.
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool pytest-args...
  local name=$1 tool=$2; shift 2
  timeout 280 $SAN --tool "$tool" --print-limit 20 --error-exitcode 9 \
    python -m pytest "$@" -x -q -p no:cacheprovider --timeout 250 > "gpurun_out/sanitize_${name}_${tool}.log" 2>&1
  echo "$name/$tool exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' "gpurun_out/sanitize_${name}_${tool}.log" | tail -2 | tr '\n' ' ')"
}
run bn        memcheck  tests/test_kernels_gpu.py -k "bn and not fused"
run bn        racecheck tests/test_kernels_gpu.py -k "bn and not fused"
run optim     memcheck  tests/test_kernels_gpu.py -k "sgd or adam or soft_ce or pool"
run gemm      memcheck  tests/test_gemm_gpu.py -k "not ship"
run gemm      synccheck tests/test_gemm_gpu.py -k "not ship"
run conv3     memcheck  tests/test_conv3x3_gpu.py
run persist   memcheck  tests/test_persist_gpu.py
