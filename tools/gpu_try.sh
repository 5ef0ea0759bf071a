#!/bin/bash
# retry wrapper around gpurun: keeps asking while the pod answers "busy" (rc 3), max ~40 min
# usage: tools/gpu_try.sh <gpurun args...>
for i in $(seq 1 16); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
