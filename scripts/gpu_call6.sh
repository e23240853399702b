#!/bin/bash
# 2 GPUs: wgrad3 v2 (haloed tile) descriptor variants, multi-GPU test, in-place hot recovery without the multicast alias
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
for bo in 0 1; do
  EDL_WGRAD3_BO=$bo timeout 400 python -m pytest tests/test_round2_gpu.py -q --timeout 240 -k "conv3x3_wgrad or own_wgrad" > gpurun_out/wgrad3v2_bo$bo.log 2>&1
  echo "wgrad3 v2 base-offset mode $bo: exit $? $(tail -1 gpurun_out/wgrad3v2_bo$bo.log)"
done
for bo in 0 1; do
  if tail -1 gpurun_out/wgrad3v2_bo$bo.log | grep -q " passed" && ! tail -1 gpurun_out/wgrad3v2_bo$bo.log | grep -q failed; then
    EDL_WGRAD3_BO=$bo timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --own-wgrad3 --no-e2e --kineto gpurun_out/kineto_r2_wgrad3v2.txt > gpurun_out/b6_wgrad3v2_bo$bo.json 2> gpurun_out/b6_wgrad3v2.err
    echo "bench own-wgrad3 v2 (bo $bo): $(head -c 300 gpurun_out/b6_wgrad3v2_bo$bo.json)"
    break
  fi
done
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e > gpurun_out/b6_base.json 2> gpurun_out/b6_base.err
echo "bench base: $(head -c 300 gpurun_out/b6_base.json)"
timeout 600 python -m pytest tests/test_allreduce_multigpu.py -x -q --timeout 560 -k "2" > gpurun_out/mg_test_2.log 2>&1
echo "multigpu test: exit $? $(tail -3 gpurun_out/mg_test_2.log | tr '\n' ' ')"
EDL_DISABLE_MULTICAST=1 timeout 700 python tools/bench_elastic_launch.py --native-store --trainer resnet --gpus-per-pod 1 --leave kill \
     --out gpurun_out/elastic_launch_2gpu_kill_p2p.json > gpurun_out/elastic_launch_2gpu_kill_p2p.log 2>&1
echo "elastic kill (no multicast): exit $?"; tail -n 3 gpurun_out/elastic_launch_2gpu_kill_p2p.log
