#!/bin/bash
# ncu of the kernels that dominate after the issue-loop fix
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
NCU="ncu --set full --clock-control none --import-source on"
cap() { local name=$1 kern=$2; shift 2
  timeout 300 $NCU -k regex:$kern -s 3 -c 1 -f -o gpurun_out/prof_$name python tools/prof_one.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name: $?"
}
cap c25_conv3g gemm_persist_kernel conv3g 32 2048 2048 14 14 32
cap c25_pair gemm_pair_kernel gemmi 6272 1024 2048
cap c25_dgradbn_s2 gemm_persist_kernel dgradbn 25088 512 128
cap c25_dgradbn_s1 gemm_persist_kernel dgradbn 100352 64 256
cap c25_fwd_s1 gemm_persist_kernel fwd 100352 64 256
