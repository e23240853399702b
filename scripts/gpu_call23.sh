#!/bin/bash
# resident weights for the 64-channel 3x3 layers: numerics, A/B in the student step and the teacher forward, ncu of the
# teacher's two dominant kernels
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
timeout 500 python -m pytest tests/test_persist_gpu.py tests/test_conv3x3_gpu.py -q --timeout 300 -x > gpurun_out/c23_tests.log 2>&1
rc=$?
echo "tests: exit $rc $(tail -1 gpurun_out/c23_tests.log)"; grep -E "^E  |Error" gpurun_out/c23_tests.log | head -12
if [ $rc -ne 0 ]; then exit 0; fi
b() { local tag=$1; shift
  timeout 300 env ${ENVV:-A=1} python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e "$@" > gpurun_out/b23_$tag.json 2> gpurun_out/b23_$tag.err
  python - gpurun_out/b23_$tag.json $tag <<'P'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("bench %-14s %.0f img/s  %.3f ms/step  launches %s  fallbacks %d" % (sys.argv[2], d["value"], d["ms_per_step"], d.get("gpu_launches"), len(d.get("library_fallbacks") or {})))
except Exception as e:
    print("bench %s: no result (%s)" % (sys.argv[2], e))
P
}
b res --no-extras --kineto gpurun_out/kineto_r2_c23.txt
ENVV="EDL_CONV_BRES=0" b nores --no-extras
timeout 200 python tools/teacher_prof.py > gpurun_out/teacher_c23_res.txt 2>&1; grep -E "^forward|gemm_persist" gpurun_out/teacher_c23_res.txt | cut -c1-100,150-260
EDL_CONV_BRES=0 timeout 200 python tools/teacher_prof.py > gpurun_out/teacher_c23_nores.txt 2>&1; grep -E "^forward" gpurun_out/teacher_c23_nores.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm_persist_kernel -s 3 -c 1 -f -o gpurun_out/prof_teacher_conv3g python tools/prof_one.py conv3g 32 2048 2048 14 14 32 > gpurun_out/ncu_teacher_conv3g.log 2>&1
echo "ncu conv3g: $?"
timeout 300 $NCU -k regex:gemm_persist_kernel -s 3 -c 1 -f -o gpurun_out/prof_teacher_gemm python tools/prof_one.py gemmi 6272 1024 2048 > gpurun_out/ncu_teacher_gemm.log 2>&1
echo "ncu gemmi: $?"
timeout 300 python -m pytest tests/test_model_gpu.py tests/test_round2_gpu.py -q --timeout 300 -x -k "model or bnmodel or teacher" > gpurun_out/c23_tests2.log 2>&1
echo "tests2: exit $? $(tail -1 gpurun_out/c23_tests2.log)"
