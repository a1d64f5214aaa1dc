#!/bin/bash
# usage: gpujob.sh <timeout_s> <script_file> [--gpus N]   -- retries while the pod answers "busy / transient"
T=$1; S=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$(cat "$S")" > /tmp/gpujob_last.log 2>&1
  rc=$?
  if grep -q "status=transient\|nothing was charged" /tmp/gpujob_last.log && [ $rc -ne 0 -o -n "$(grep -l 'status=transient' /tmp/gpujob_last.log)" ]; then
    echo "[gpujob] attempt $i: transient, retrying in 90 s" ; sleep 90; continue
  fi
  break
done
tail -40 /tmp/gpujob_last.log
exit $rc
