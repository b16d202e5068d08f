#!/bin/bash
# N = 2 with the side-stream weight gradients (default) and without: data-parallel step + teardown
mkdir -p gpurun_out
for v in 1 0; do
  t0=$(date +%s)
  NSP_WGRAD_STREAM=$v timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$v bench.py --gpus 2 --steps 10 --warmup 3 --no-eager --no-cpu-baseline > gpurun_out/n2_side$v.log 2>&1
  echo "NSP_WGRAD_STREAM=$v rc=$? $(( $(date +%s) - t0 ))s $(grep '^{' gpurun_out/n2_side$v.log | tail -1 | cut -c1-200)"
  grep -i "error\|Traceback" gpurun_out/n2_side$v.log | head -5
done
