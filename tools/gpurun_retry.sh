#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 = nothing charged).  usage: tools/gpurun_retry.sh <logfile> [--gpus N] <timeout> <command string>
LOG=$1; shift
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GP --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "EXIT $rc" >> $LOG; exit $rc; fi
  sleep 45
done
echo "EXIT 3 (gave up)" >> $LOG
