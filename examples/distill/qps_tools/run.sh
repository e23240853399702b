#!/bin/bash
# DistillReader QPS for a sweep of teacher batch sizes (reference: example/distill/qps_tools/run.sh:23-28).
#   TEACHERS=ip:port[,ip:port] bash run.sh          fixed teachers
#   bash run.sh                                     NOP teacher (the pipeline alone: no teacher needed)
set -eu
cd "$(dirname "$0")"
export PADDLE_DISTILL_SERVICE_NAME=${PADDLE_DISTILL_SERVICE_NAME:-MnistDistill}
export PADDLE_DISTILL_MAX_TEACHER=${PADDLE_DISTILL_MAX_TEACHER:-1}
out=${QPS_OUT:-qps.jsonl}
: > "$out"
for bs in ${TEACHER_BATCH_SIZES:-1 2 4 8 16 24 32}; do
  echo "-------- teacher batch size $bs ---------"
  if [ -n "${TEACHERS:-}" ]; then
    python distill_reader_qps.py --teachers "$TEACHERS" --teacher_bs "$bs" --out "$out"
  else
    python distill_reader_qps.py --nop --teacher_bs "$bs" --out "$out"
  fi
done
echo "results in $out"
