#!/bin/bash
# scratch/gpu.sh <timeout-seconds> '<command>' : gpurun with retries while the pod's GPU slots are busy (exit code 3)
T=$1; shift
python -m lichtfeld_densification_plugin_b200.build >/dev/null || { echo "BUILD FAILED"; python -m lichtfeld_densification_plugin_b200.build; exit 1; }
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
