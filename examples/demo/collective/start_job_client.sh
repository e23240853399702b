#!/bin/bash
# JobClient on every node: polls the JobServer, starts / stops this node's pods
# (reference: example/demo/collective/start_job_client.sh).
set -eu
source "$(dirname "$0")/env.sh"
python -m paddle_edl.demo.collective.job_client_demo --pod_path "$(dirname "$0")/resnet50/pod.sh" \
  --package_sh "$(dirname "$0")/resnet50/package.sh" --nodes_range "${PADDLE_EDLNODES_RANAGE}"
