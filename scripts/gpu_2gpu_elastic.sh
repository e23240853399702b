#!/bin/bash
# 2 GPUs: in-place rescale through the launcher (pytest), DeepFM, all-reduce ncu (application replay), bench N=2
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 900 python -m pytest tests/test_allreduce_multigpu.py -q --timeout 800 -k "launcher_on_gpus" > gpurun_out/c21_elastic.log 2>&1
echo "elastic pytest: exit $? $(tail -1 gpurun_out/c21_elastic.log)"; grep -E "^E  " gpurun_out/c21_elastic.log | head -5
timeout 500 python tools/bench_elastic_launch.py --native-store --trainer resnet --gpus-per-pod 1 --leave kill --modes inplace \
   --out gpurun_out/elastic_launch_2gpu_kill_final.json > gpurun_out/elastic_launch_2gpu_kill_final.log 2>&1
tail -n 1 gpurun_out/elastic_launch_2gpu_kill_final.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29651 examples/ctr/train.py --model deepfm --steps 30 --vocab 1000001 --out gpurun_out/ctr_deepfm_2gpu.json > gpurun_out/ctr_deepfm_2gpu.log 2>&1
echo "deepfm: exit $? $(tail -1 gpurun_out/ctr_deepfm_2gpu.log | cut -c1-200)"
timeout 400 $TR --master-port 29652 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/b21_2gpu.json 2> gpurun_out/b21_2gpu.err
echo "bench 2: $(head -c 400 gpurun_out/b21_2gpu.json)"
for algo in multimem fused; do
timeout 400 ncu --replay-mode application --devices 0 --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis \
   --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats \
   -k regex:allreduce -s 4 -c 1 -f -o gpurun_out/prof_allreduce_2gpu_$algo python tools/prof_allreduce.py --gpus 2 --mb 16 --algo $algo \
   > gpurun_out/ncu_allreduce_2gpu_$algo.log 2>&1
echo "ncu $algo: exit $? $(tail -2 gpurun_out/ncu_allreduce_2gpu_$algo.log | tr '\n' ' ' | cut -c1-300)"
done
