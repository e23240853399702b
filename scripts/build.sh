#!/bin/bash
# Build + test driver (reference: scripts/build.sh:44-75 -- protoc codegen, cmake wheel, a background etcd,
# then ctest).  Here: compile the sm_100a extension in-tree and run the CPU test-suite; no external
# daemons are needed (the tests start the in-repo KV / mini-redis servers themselves).
set -euo pipefail
cd "$(dirname "$0")/.."
python -m edl_b200.build_ext "$@"
python -m pytest tests -x -q -m "not gpu"
if python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)"; then
  python -m pytest tests -x -q -m gpu
fi
