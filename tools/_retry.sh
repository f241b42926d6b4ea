#!/bin/bash
# keep asking for a GPU box until the call is accepted (exit code 3 = no box free right now, nothing charged)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
