#!/bin/bash
# Elastic ResNet50_vd pre-training on one node (reference: example/collective/resnet50/train_pretrain.sh).
set -e
export PADDLE_JOB_ID=${PADDLE_JOB_ID:-rn50_job}
export PADDLE_ETCD_ENDPOINTS=${PADDLE_ETCD_ENDPOINTS:-127.0.0.1:2379}
python -m paddle_edl.store.kv_server --port ${PADDLE_ETCD_ENDPOINTS##*:} &   # a real etcd is not needed
KV=$!
trap "kill $KV" EXIT
sleep 1
python -m paddle_edl.collective.launch --nodes_range 1:8 --nproc_per_node ${NPROC:-8} \
  --etcd_endpoints ${PADDLE_ETCD_ENDPOINTS} --job_id ${PADDLE_JOB_ID} --log_dir ./log \
  --hdfs_path ${CKPT:-./resnet_ckpt} "$(dirname "$0")/train.py" --epochs ${EPOCHS:-90} "$@"
