#!/bin/bash
# background helper: keeps trying to get an 8-GPU box for a command given as $1; log in /tmp/run8.log
cd /root/repo
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- "$1" > /tmp/run8.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && [ $rc -ne 2 ]; then echo "done rc=$rc try=$i" >> /tmp/run8.log; exit $rc; fi
  sleep 200
done
