#!/bin/bash
# N = 2 sanity of the data-parallel step (PDL kernels + async NCCL buckets inside one CUDA graph) and of the process teardown.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bucketed.log 2>&1
echo "bucketed rc=$? $(( $(date +%s) - t0 ))s $(grep '^{' gpurun_out/n2_bucketed.log | tail -1 | cut -c1-160)"
