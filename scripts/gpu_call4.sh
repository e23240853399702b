#!/bin/bash
# 2 GPUs: own symmetric-memory bootstrap (cuMem VMM), fused reduce-scatter+SGD+all-gather, clip, extended bench line
set -u
N=${1:-2}
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 600 python -m pytest tests/test_allreduce_multigpu.py -x -q --timeout 560 -k "2" > gpurun_out/mg_test_${N}.log 2>&1
echo "multigpu test: exit $? $(tail -3 gpurun_out/mg_test_${N}.log | tr '\n' ' ')"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 > gpurun_out/b4_1gpu.json 2> gpurun_out/b4_1gpu.err
echo "bench 1: $(head -c 400 gpurun_out/b4_1gpu.json)"
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --no-fused-opt > gpurun_out/b4_1gpu_nofuse.json 2> gpurun_out/b4_1gpu_nofuse.err
echo "bench 1 nofuse: $(head -c 300 gpurun_out/b4_1gpu_nofuse.json)"
timeout 600 $TR --master-port 29621 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/b4_${N}gpu.json 2> gpurun_out/b4_${N}gpu.err
echo "bench $N: $(head -c 600 gpurun_out/b4_${N}gpu.json)"
tail -5 gpurun_out/b4_${N}gpu.err
timeout 400 $TR --master-port 29622 bench.py --gpus $N --steps 100 --warmup 5 --no-fused-opt --no-extras > gpurun_out/b4_${N}gpu_nofuse.json 2> gpurun_out/b4_${N}gpu_nofuse.err
echo "bench $N nofuse: $(head -c 300 gpurun_out/b4_${N}gpu_nofuse.json)"
timeout 400 $TR --master-port 29623 tools/bench_comm.py --sweep --exposed --max-mb 64 --iters 10 --out gpurun_out/comm_${N}gpu.json > gpurun_out/comm_${N}gpu.log 2>&1
tail -n 3 gpurun_out/comm_${N}gpu.log
