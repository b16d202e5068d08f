#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/call9_summary.txt
: > $S
stage() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/$name.log 2>&1; local rc=$?
          echo "$name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" >> $S; }
stage ctc_trace 100 env NSP_CTC_DEBUG=8 python profiles/prof_ctc.py
stage ctc_prof 100 python profiles/prof_ctc.py
stage suite_pair 900 env NSP_GEMM_EPILOGUE=pair python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider
stage bench_pair 600 env NSP_GEMM_EPILOGUE=pair python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager
cat $S
