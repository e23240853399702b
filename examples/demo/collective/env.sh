#!/bin/bash
# Shared settings of the JobServer / JobClient demo (reference: example/demo/collective/env.sh).
export PADDLE_JOB_ID=${PADDLE_JOB_ID:-edl_demo_job}
export PADDLE_JOBSERVER=${PADDLE_JOBSERVER:-http://127.0.0.1:8180}
export PADDLE_ETCD_ENDPOINTS=${PADDLE_ETCD_ENDPOINTS:-127.0.0.1:2379}
export PADDLE_EDLNODES_RANAGE=${PADDLE_EDLNODES_RANAGE:-1:4}
export PADDLE_EDL_HDFS_PATH=${PADDLE_EDL_HDFS_PATH:-/tmp/edl_demo_ckpt}
export PADDLE_EDL_ONLY_FOR_CE_TEST=${PADDLE_EDL_ONLY_FOR_CE_TEST:-0}
REPO=$(cd "$(dirname "${BASH_SOURCE[0]}")/../../.." && pwd)
export REPO PYTHONPATH=${REPO}:${PYTHONPATH:-}
