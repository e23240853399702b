#!/bin/bash
# 1 GPU: in-graph device times of the own 3x3 wgrad (kineto) + ncu --set full of two shapes + launch overhead check
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 --own-wgrad3 --no-e2e --kineto gpurun_out/kineto_r2_ownwgrad3.txt > gpurun_out/ab3_ownwgrad3.json 2> gpurun_out/ab3_ownwgrad3.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:conv3x3_wgrad_kernel -s 3 -c 1 -f -o gpurun_out/prof_wgrad3_c64 python tools/prof_one.py wgrad3 32 64 64 56 56 > gpurun_out/ncu_wgrad3_c64.log 2>&1
timeout 300 $NCU -k regex:conv3x3_wgrad_kernel -s 3 -c 1 -f -o gpurun_out/prof_wgrad3_c512 python tools/prof_one.py wgrad3 32 512 512 7 7 > gpurun_out/ncu_wgrad3_c512.log 2>&1
timeout 300 $NCU -k regex:conv3x3_wgrad_kernel -s 3 -c 1 -f -o gpurun_out/prof_wgrad3_c256 python tools/prof_one.py wgrad3 32 256 256 14 14 > gpurun_out/ncu_wgrad3_c256.log 2>&1
timeout 300 $NCU -k regex:sgd_momentum_kernel -s 2 -c 1 -f -o gpurun_out/prof_sgd python tools/prof_one.py sgd > gpurun_out/ncu_sgd.log 2>&1
timeout 300 $NCU -k regex:soft_ce -s 4 -c 2 -f -o gpurun_out/prof_softce python tools/prof_one.py softce 32 > gpurun_out/ncu_softce.log 2>&1
timeout 300 $NCU -k regex:rope_kernel -s 2 -c 1 -f -o gpurun_out/prof_rope python tools/prof_one.py rope 4096 16 64 > gpurun_out/ncu_rope.log 2>&1
ls -la gpurun_out/*.ncu-rep
