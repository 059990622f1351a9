#!/bin/bash
# usage: .gpuretry.sh <timeout> <cmd>   -- retries while the pod answers transient (rc 3)
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2" > /root/repo/gpurun_out/.retry.log 2>&1
  if ! grep -q "status=transient" /root/repo/gpurun_out/.retry.log; then break; fi
  sleep 90
done
tail -40 /root/repo/gpurun_out/.retry.log
