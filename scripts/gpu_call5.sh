#!/bin/bash
# 2 GPUs: full multi-GPU test + rescale recovery through the real launcher with GPU trainers (restart vs in place)
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 600 python -m pytest tests/test_allreduce_multigpu.py tests/test_device_feed_multigpu.py -x -q --timeout 560 -k "2 or feed" > gpurun_out/mg_test_2.log 2>&1
echo "multigpu tests: exit $? $(tail -3 gpurun_out/mg_test_2.log | tr '\n' ' ')"
timeout 500 python -m pytest tests/test_round2_gpu.py -q --timeout 240 -k "agreement or hierarchical" > gpurun_out/mg_round2.log 2>&1
echo "round2 multigpu tests: exit $? $(tail -2 gpurun_out/mg_round2.log | tr '\n' ' ')"
for leave in scale_in kill; do
  timeout 700 python tools/bench_elastic_launch.py --native-store --trainer resnet --gpus-per-pod 1 --leave $leave \
     --out gpurun_out/elastic_launch_2gpu_$leave.json > gpurun_out/elastic_launch_2gpu_$leave.log 2>&1
  echo "elastic $leave: exit $?"; tail -n 3 gpurun_out/elastic_launch_2gpu_$leave.log
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29631 tools/bench_rescale.py --drop 1 --out gpurun_out/rescale_2gpu.json > gpurun_out/rescale_2gpu.log 2>&1
tail -n 2 gpurun_out/rescale_2gpu.log
