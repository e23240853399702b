#!/bin/bash
# ncu --set full of the dominant tcgen05 kernels (end of round 2)
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
NCU="ncu --set full --clock-control none --import-source on"
cap() { local name=$1 kern=$2; shift 2
  timeout 300 $NCU -k regex:$kern -s 3 -c 1 -f -o gpurun_out/prof_$name python tools/prof_one.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name: $?"
}
cap final_conv3g gemm_persist_kernel conv3g 32 2048 2048 14 14 32
cap final_pair gemm_pair_kernel gemmi 6272 1024 2048
cap final_dgradbn_s2 gemm_persist_kernel dgradbn 25088 512 128
cap final_dgradbn_s1 gemm_persist_kernel dgradbn 100352 64 256
cap final_fwd_s1 gemm_persist_kernel fwd 100352 64 256
cap final_conv3_s2 gemm_persist_kernel conv3 32 128 128 28 28
cap final_wgrad gemm_tcgen05_kernel wgrad 1024 256 6272
