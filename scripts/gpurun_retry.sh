#!/bin/bash
# scripts/gpurun_retry.sh [--gpus N] TIMEOUT 'command'   - retry while the pod answers "busy" (exit 3, nothing charged)
gpus=""
if [ "$1" == "--gpus" ]; then gpus="--gpus $2"; shift 2; fi
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $gpus --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
