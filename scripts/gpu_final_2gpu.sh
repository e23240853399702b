#!/bin/bash
# End-of-round validation on TWO GPUs: the multi-GPU tests (fused bucket kernels, engine vs single-process reference, clip,
# consolidate, in-place rescale through the launcher) and the headline bench at N = 2 with every metric term.
#   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_final_2gpu.sh'
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 900 python -m pytest tests/test_allreduce_multigpu.py tests/test_device_feed_multigpu.py -q --timeout 800 -p no:cacheprovider > gpurun_out/final_2gpu_tests.log 2>&1
echo "multi-GPU tests: exit $? $(tail -1 gpurun_out/final_2gpu_tests.log)"; grep -E "^FAILED|^ERROR|^E  " gpurun_out/final_2gpu_tests.log | head
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29652 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/final_bench_2gpu.json 2> gpurun_out/final_bench_2gpu.err
echo "bench 2: $? $(grep '^{' gpurun_out/final_bench_2gpu.json | tail -1 | cut -c1-330)"
