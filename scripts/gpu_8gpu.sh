#!/bin/bash
# N GPUs (8): flagship bench with every metric term, 8-rank kernel/engine test, all-reduce sweep vs NCCL, CTR sweep
set -u
N=${1:-8}
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29641 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/b10_${N}gpu.json 2> gpurun_out/b10_${N}gpu.err
echo "bench $N: exit $? $(head -c 700 gpurun_out/b10_${N}gpu.json)"
tail -3 gpurun_out/b10_${N}gpu.err
timeout 420 python -m pytest tests/test_allreduce_multigpu.py -x -q --timeout 400 -k "$N" > gpurun_out/mg_test_$N.log 2>&1
echo "multigpu test ($N): exit $? $(tail -2 gpurun_out/mg_test_$N.log | tr '\n' ' ')"
timeout 300 $TR --master-port 29642 tools/bench_comm.py --sweep --max-mb 128 --iters 10 --out gpurun_out/comm_${N}gpu.json > gpurun_out/comm_${N}gpu.log 2>&1
echo "comm sweep: exit $?"; tail -n 2 gpurun_out/comm_${N}gpu.log | cut -c1-400
for algo in multimem twoshot fused; do
  timeout 120 python tools/prof_allreduce.py --gpus $N --mb 16 --algo $algo >> gpurun_out/prof_allreduce_${N}gpu.jsonl 2>> gpurun_out/prof_allreduce_${N}gpu.err
done
cat gpurun_out/prof_allreduce_${N}gpu.jsonl
timeout 240 $TR --master-port 29643 examples/ctr/train.py --sweep --out gpurun_out/ctr_sweep_${N}gpu.json > gpurun_out/ctr_sweep_${N}gpu.log 2>&1
echo "ctr sweep: exit $?"; tail -n 2 gpurun_out/ctr_sweep_${N}gpu.log | cut -c1-300
timeout 240 $TR --master-port 29644 examples/ctr/train.py --model deepfm --steps 30 --vocab 1000001 --out gpurun_out/ctr_deepfm_${N}gpu.json > gpurun_out/ctr_deepfm_${N}gpu.log 2>&1
echo "ctr deepfm: exit $?"; tail -n 1 gpurun_out/ctr_deepfm_${N}gpu.log | cut -c1-300
