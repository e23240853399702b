#!/bin/bash
# One ResNeXt101_32x16d teacher on GPU ${1:-1}, serving 'image' -> 'score' on port ${2:-9898}
# (reference: example/distill/resnet/scripts/start_local_teacher.sh:24-30 starts Paddle Serving).
set -eu
gpu=${1:-1}
port=${2:-9898}
CUDA_VISIBLE_DEVICES=${gpu} python -m paddle_edl.distill.teacher_server \
  --model resnext101_32x16d --port "${port}" --max_batch 16 "${@:3}"
