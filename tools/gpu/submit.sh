#!/bin/bash
# submit a GPU job, retrying while the pod answers "busy" (exit code 3): tools/gpu/submit.sh <timeout_s> <gpus> <command...>
timeout_s=$1; gpus=$2; shift 2
for attempt in $(seq 1 60); do
  if [ "$gpus" = "1" ]; then
    /usr/local/graft/bin/gpurun --timeout "$timeout_s" -- "$@"
  else
    /usr/local/graft/bin/gpurun --gpus "$gpus" --timeout "$timeout_s" -- "$@"
  fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[submit] busy (attempt $attempt), retrying in 45 s"
  sleep 45
done
exit 3
