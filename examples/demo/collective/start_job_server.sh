#!/bin/bash
# Store + JobServer: the server rewrites the pod set of the job every --time_interval_to_change seconds
# (reference: example/demo/collective/start_job_server.sh).
set -eu
source "$(dirname "$0")/env.sh"
python -m paddle_edl.store.kv_server --port "${PADDLE_ETCD_ENDPOINTS##*:}" &
store=$!
trap 'kill ${store}' EXIT
python -m paddle_edl.demo.collective.job_server_demo --node_ips "${NODE_IPS:-127.0.0.1}" \
  --pod_num_of_node "${PODS_PER_NODE:-2}" --gpu_num_of_node "${GPUS_PER_NODE:-8}" \
  --time_interval_to_change "${CHANGE_EVERY:-900}" --port "${PADDLE_JOBSERVER##*:}"
