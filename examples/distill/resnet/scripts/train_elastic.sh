#!/bin/bash
# 8-GPU elastic data-parallel student with dynamically discovered teachers.
#   KV=10.0.0.2:2379 DISCOVERY=10.0.0.3:7001 bash train_elastic.sh
set -eu
here=$(cd "$(dirname "$0")/.." && pwd)
python -m paddle_edl.collective.launch --nodes_range "${NODES_RANGE:-1:4}" --nproc_per_node "${NPROC:-8}" \
  --etcd_endpoints "${KV:-127.0.0.1:2379}" --job_id "${JOB_ID:-distill_rn50vd}" \
  --hdfs_path "${CKPT:-/tmp/distill_rn50vd}" --log_dir ./log \
  "${here}/train.py" --use_distill_service True --discovery "${DISCOVERY:-127.0.0.1:7001}" \
  --service_name ResNeXt101_32x16d "$@"
