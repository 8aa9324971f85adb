#!/bin/bash
# gpurun with retries on "no box / slot free right now" (exit code 3): tools/gpu.sh <timeout-seconds> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in 1 2 3 4 5 6 7 8; do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
