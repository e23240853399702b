#!/bin/bash
# Student on GPU 0 against a fixed local teacher (reference: scripts/train_student.sh).
set -eu
here=$(cd "$(dirname "$0")/.." && pwd)
export CUDA_VISIBLE_DEVICES=${CUDA_VISIBLE_DEVICES:-0}
python "${here}/train.py" \
  --model ResNet50_vd --batch_size 32 --lr 0.1 --lr_strategy cosine_warmup_decay --num_epochs 120 \
  --use_distill_service True --distill_teachers "${TEACHERS:-127.0.0.1:9898}" "$@"
