#!/bin/bash
mkdir -p gpurun_out
for v in "" _unb; do
  NSP_LIB_PATH=$PWD/neural_sp_b200/libnsp_b200$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/ab_wait$v.log 2>&1
done
timeout 300 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/suite.log 2>&1; tail -1 gpurun_out/suite.log
