#!/bin/bash
# background helper: keeps trying to get an 8-GPU box for the scaling bench + the one-process strong-scaling probe
cd /root/repo
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/bench_8gpu.err | tail -1 > gpurun_out/bench_8gpu.json; echo rc=$?; tail -c 300 gpurun_out/bench_8gpu.err; nproc; python tools/multi_probe.py --reads 20000000 --gpus 1,2,4,8 --sam 2>&1 | tail -4 | tee gpurun_out/multi_probe_8gpu.txt' > /tmp/run8.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && [ $rc -ne 2 ]; then echo "done rc=$rc try=$i" >> /tmp/run8.log; exit $rc; fi
  sleep 240
done
