#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3; nothing is charged for those).  Usage: tools/gpu_retry.sh [--gpus N] <timeout> '<command>'
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GP --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
