#!/bin/bash
# End-of-round validation on ONE GPU: the whole `-m gpu` suite, smoke(), the headline bench (with the end-to-end and the
# extra metric terms) and the library-free variant.   gpurun --timeout 1500 -- 'bash scripts/gpu_final.sh'
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/final_gpu_tests.log 2>&1
echo "pytest -m gpu: exit $? $(tail -1 gpurun_out/final_gpu_tests.log)"; grep -E "^FAILED|^ERROR" gpurun_out/final_gpu_tests.log | head
timeout 200 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke: $? $(tail -1 gpurun_out/final_smoke.log)"
timeout 400 python bench.py --gpus 1 --steps 200 --warmup 5 > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err
echo "bench: $? $(cut -c1-400 gpurun_out/final_bench_1gpu.json | tail -1)"
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 5 --no-library --no-extras > gpurun_out/final_bench_1gpu_nolib.json 2> gpurun_out/final_bench_1gpu_nolib.err
echo "bench nolib: $? $(cut -c1-300 gpurun_out/final_bench_1gpu_nolib.json | tail -1)"
