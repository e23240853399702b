#!/bin/bash
# First GPU call of a round: validate what was written without a GPU, then A/B it on the flagship bench.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_validate_experimental.sh'
# Results land in gpurun_out/ (merged back by gpurun).  Nothing here changes defaults: read the numbers, then
# flip EDL_PDL / EDL_OWN_WGRAD3 defaults (ops/gemm.py, csrc/bn.cu) for what passed AND paid.
set -u
mkdir -p gpurun_out
export EDL_TEST_EXPERIMENTAL=1
timeout 900 python -m pytest tests/test_experimental_gpu.py -q --timeout 300 > gpurun_out/experimental_tests.log 2>&1
echo "experimental tests exit $?" | tee -a gpurun_out/experimental_tests.log
for flags in "" "--pdl" "--own-wgrad3" "--conv3-s2" "--own-stem1" "--fuse-bn-bwd 2" "--pdl --own-wgrad3 --conv3-s2 --own-stem1" "--pdl --own-wgrad3 --conv3-s2 --own-stem1 --fuse-bn-bwd 2"; do
  tag=$(echo "base$flags" | tr -d ' -')   # e.g. basepdlownwgrad3
  timeout 300 python bench.py --gpus 1 --steps 60 --warmup 5 $flags > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  echo "$tag: $(head -c 300 gpurun_out/ab_$tag.json)"
done
