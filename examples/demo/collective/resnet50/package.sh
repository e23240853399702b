#!/bin/bash
# Per-pod runtime directory of the JobClient demo (the role of the reference's
# example/demo/collective/resnet50/package.sh:19-52, invoked as `package.sh -pod_id <id>`): every pod gets its
# own working directory with the trainer sources it runs and links to the shared data set / checkpoint
# directory, so that pods the JobClient starts and stops on one node never share log or scratch files.
#
#   PADDLE_EDL_IMAGENET_PATH           data set root (optional: synthetic data without it)
#   PADDLE_EDL_FLEET_CHECKPOINT_PATH   shared checkpoint directory (required for stop-resume across pods)
set -eu
pod_id=""
while [ $# -gt 0 ]; do
  case "$1" in
    -pod_id) pod_id="$2"; shift 2 ;;
    "") shift ;;
    *) echo "package.sh: unsupported argument '$1' (usage: package.sh -pod_id <id>)" >&2; exit 1 ;;
  esac
done
[ -n "${pod_id}" ] || { echo "package.sh: -pod_id is required" >&2; exit 1; }
here="$(cd "$(dirname "$0")" && pwd)"
repo="$(cd "${here}/../../../.." && pwd)"
src="${repo}/examples/collective/resnet50"
dst="${PADDLE_EDL_POD_ROOT:-${here}/resnet50_pod}/${pod_id}"
mkdir -p "${dst}"
cp "${src}"/*.py "${dst}/"
cp "${here}/pod.sh" "${dst}/"
[ -z "${PADDLE_EDL_IMAGENET_PATH:-}" ] || [ -e "${dst}/ImageNet" ] || ln -s "${PADDLE_EDL_IMAGENET_PATH}" "${dst}/ImageNet"
if [ -n "${PADDLE_EDL_FLEET_CHECKPOINT_PATH:-}" ] && [ ! -e "${dst}/fleet_checkpoints" ]; then
  mkdir -p "${PADDLE_EDL_FLEET_CHECKPOINT_PATH}"
  ln -s "${PADDLE_EDL_FLEET_CHECKPOINT_PATH}" "${dst}/fleet_checkpoints"
fi
echo "packaged pod ${pod_id} into ${dst}"
