#!/bin/bash
set -u
mkdir -p gpurun_out
python -c 'import torch' 2> /dev/null
(timeout 120 python tools/trace_persist.py fwd 100352 64 256
 timeout 120 python tools/trace_persist.py dgradbn 100352 64 256 y
 timeout 120 python tools/trace_persist.py dgradbn 25088 128 512 y
 timeout 120 python tools/trace_persist.py dgradbn 25088 512 128
 timeout 120 python tools/trace_persist.py conv3 32 128 128 28 28) > gpurun_out/trace_persist_c28.txt 2>&1
echo "trace: $? $(wc -l < gpurun_out/trace_persist_c28.txt) lines"
