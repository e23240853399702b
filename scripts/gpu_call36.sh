#!/bin/bash
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
(timeout 120 python tools/trace_persist.py fwd 100352 64 256; timeout 120 python tools/trace_persist.py fwd 25088 256 512;  timeout 120 python tools/trace_persist.py dgradbn 25088 128 512 y) > gpurun_out/trace_persist_c36.txt 2>&1
grep -E "==|epilogue tile  [45]|mma      tile  [45]" gpurun_out/trace_persist_c36.txt | head -40
