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
# distill mode A/B: teacher residual fused into the GEMM epilogue (validate with tests/test_experimental_gpu.py first)
for flags in "" "--teacher-fuse-res" "--teacher-fuse-res --conv3-s2" "--teacher-fuse-res --conv3-s2 --pdl"; do
  tag=$(echo "distill$flags" | tr -d ' -')
  timeout 400 $TR --master-port 29615 bench.py --gpus $N --mode distill --steps 60 --warmup 8 $flags \
    > gpurun_out/ab_${tag}_${N}gpu.json 2> gpurun_out/ab_${tag}_${N}gpu.err
  echo "$tag: $(head -c 240 gpurun_out/ab_${tag}_${N}gpu.json)"
done
# launcher-level recovery time with real GPU trainers: pod A = first half of the GPUs, pod B = second half
timeout 900 python tools/bench_elastic_launch.py --native-store --trainer resnet --gpus-per-pod $(( N / 2 )) \
  --out gpurun_out/elastic_launch_${N}gpu.json > gpurun_out/elastic_launch_${N}gpu.log 2>&1
EDL_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_experimental_gpu.py -q -k "agreement or hierarchical" > gpurun_out/agree_test.log 2>&1
# hierarchical all-reduce overhead: the same box pretending to be 2 hosts (EDL_FAKE_HOST is read per rank)
EDL_FAKE_HOST_SPLIT=$(( N / 2 )) timeout 400 $TR --master-port 29616 bench.py --gpus $N --steps 60 --warmup 5 \
  > gpurun_out/bench_hier_${N}gpu.json 2> gpurun_out/bench_hier_${N}gpu.err
tail -n 3 gpurun_out/elastic_launch_${N}gpu.log gpurun_out/agree_test.log
tail -n 3 gpurun_out/comm_${N}gpu.log gpurun_out/rescale_${N}gpu.log gpurun_out/ctr_sweep_${N}gpu.log gpurun_out/ctr_deepfm_${N}gpu.log
