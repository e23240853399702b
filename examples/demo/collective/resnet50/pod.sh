#!/bin/bash
# One pod of the demo job: the elastic launcher around the ResNet trainer.
set -eu
source "$(dirname "$0")/../env.sh"
exec python -m paddle_edl.collective.launch --nodes_range "${PADDLE_EDLNODES_RANAGE}" \
  --nproc_per_node "${NPROC_PER_POD:-4}" --etcd_endpoints "${PADDLE_ETCD_ENDPOINTS}" --job_id "${PADDLE_JOB_ID}" \
  --hdfs_path "${PADDLE_EDL_HDFS_PATH}" --log_dir ./log \
  "${REPO}/examples/collective/resnet50/train.py" --epochs "${EPOCHS:-90}" "$@"
