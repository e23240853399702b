#!/bin/bash
# Multi-GPU metrics of BASELINE.json that bench.py does not print, in selectable sections (an N-GPU call is charged N x
# its wall time: pick what the budget allows).
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_multi_metrics.sh 8 comm rescale'
# Sections (default: comm rescale tests):
#   comm     exposed comm ms/step + all-reduce sweep vs NCCL                      (~2 min)
#   rescale  in-process rebuild + sync_from vs checkpoint reload                  (~1.5 min)
#   tests    experimental multi-GPU tests (agreement kernel, hierarchical 2x2)    (~1.5 min)
#   launch   launcher-level recovery time with GPU trainers, restart vs in place  (~4 min)
#   distill  distill-mode A/B: teacher residual fusion, stride-2 convs, PDL       (~4 min)
#   hier     flagship bench with the box pretending to be 2 hosts                 (~1 min)
#   ctr      CTR embedding all-reduce sweep + DeepFM step at the reference vocab  (~3 min)
set -u
N=${1:-8}
shift || true
SECTIONS=${*:-comm rescale tests}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for sec in $SECTIONS; do
  case $sec in
    comm)
      timeout 400 $TR --master-port 29611 tools/bench_comm.py --sweep --exposed --max-mb 128 --iters 10 \
        --out gpurun_out/comm_${N}gpu.json > gpurun_out/comm_${N}gpu.log 2>&1
      tail -n 4 gpurun_out/comm_${N}gpu.log ;;
    rescale)
      timeout 400 $TR --master-port 29612 tools/bench_rescale.py --drop $(( N >= 4 ? 2 : 1 )) \
        --out gpurun_out/rescale_${N}gpu.json > gpurun_out/rescale_${N}gpu.log 2>&1
      tail -n 2 gpurun_out/rescale_${N}gpu.log ;;
    tests)
      timeout 500 python -m pytest tests/test_round2_gpu.py -q -k "agreement or hierarchical" \
        > gpurun_out/multi_tests.log 2>&1
      tail -n 3 gpurun_out/multi_tests.log ;;
    launch)
      timeout 900 python tools/bench_elastic_launch.py --native-store --trainer resnet --gpus-per-pod $(( N / 2 )) \
        --out gpurun_out/elastic_launch_${N}gpu.json > gpurun_out/elastic_launch_${N}gpu.log 2>&1
      tail -n 3 gpurun_out/elastic_launch_${N}gpu.log ;;
    distill)
      for flags in "" "--teacher-fuse-res" "--teacher-fuse-res --conv3-s2" "--teacher-fuse-res --conv3-s2 --pdl"; do
        tag=$(echo "distill$flags" | tr -d ' -')
        timeout 400 $TR --master-port 29615 bench.py --gpus $N --mode distill --steps 60 --warmup 8 $flags \
          > gpurun_out/ab_${tag}_${N}gpu.json 2> gpurun_out/ab_${tag}_${N}gpu.err
        echo "$tag: $(head -c 240 gpurun_out/ab_${tag}_${N}gpu.json)"
      done ;;
    hier)
      EDL_FAKE_HOST_SPLIT=$(( N / 2 )) timeout 400 $TR --master-port 29616 bench.py --gpus $N --steps 60 --warmup 5 \
        > gpurun_out/bench_hier_${N}gpu.json 2> gpurun_out/bench_hier_${N}gpu.err
      head -c 300 gpurun_out/bench_hier_${N}gpu.json; echo ;;
    ctr)
      timeout 400 $TR --master-port 29613 examples/ctr/train.py --sweep --out gpurun_out/ctr_sweep_${N}gpu.json \
        > gpurun_out/ctr_sweep_${N}gpu.log 2>&1
      timeout 300 $TR --master-port 29614 examples/ctr/train.py --model deepfm --steps 30 --vocab 1000001 \
        --out gpurun_out/ctr_deepfm_${N}gpu.json > gpurun_out/ctr_deepfm_${N}gpu.log 2>&1
      tail -n 2 gpurun_out/ctr_sweep_${N}gpu.log gpurun_out/ctr_deepfm_${N}gpu.log ;;
    *) echo "unknown section $sec" ;;
  esac
done
