#!/bin/bash
# Multi-GPU metrics of BASELINE.json that bench.py does not print: exposed comm ms/step, all-reduce sweep vs NCCL,
# rescale recovery time (in place vs stop-resume), CTR embedding all-reduce sweep.
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_multi_metrics.sh 8'
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 tools/bench_comm.py --sweep --exposed --out gpurun_out/comm_${N}gpu.json > gpurun_out/comm_${N}gpu.log 2>&1
timeout 600 $TR --master-port 29612 tools/bench_rescale.py --drop $(( N >= 4 ? 2 : 1 )) --out gpurun_out/rescale_${N}gpu.json > gpurun_out/rescale_${N}gpu.log 2>&1
timeout 600 $TR --master-port 29613 examples/ctr/train.py --sweep --out gpurun_out/ctr_sweep_${N}gpu.json > gpurun_out/ctr_sweep_${N}gpu.log 2>&1
timeout 300 $TR --master-port 29614 examples/ctr/train.py --model deepfm --steps 30 --vocab 1000001 --out gpurun_out/ctr_deepfm_${N}gpu.json > gpurun_out/ctr_deepfm_${N}gpu.log 2>&1
tail -n 3 gpurun_out/comm_${N}gpu.log gpurun_out/rescale_${N}gpu.log gpurun_out/ctr_sweep_${N}gpu.log gpurun_out/ctr_deepfm_${N}gpu.log
