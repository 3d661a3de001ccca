#!/bin/bash
# usage: [GPUS=N] tools/gpu_retry.sh <timeout_s> <logfile> <command string>   -- retries while the pod answers "busy" (exit code 3)
T=$1; LOG=$2; shift 2
G=${GPUS:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then
    /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$LOG" 2>&1
  else
    /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" > "$LOG" 2>&1
  fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
