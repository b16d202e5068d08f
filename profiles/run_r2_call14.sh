#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/call14_summary.txt
: > $S
stage() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/$name.log 2>&1; local rc=$?
          echo "$name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" >> $S; }
stage quick 300 python -m pytest tests/test_backward_gpu.py tests/test_modules_gpu.py -q --timeout=120 -p no:cacheprovider -x
stage suite 600 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider
stage bench 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager
cat $S
