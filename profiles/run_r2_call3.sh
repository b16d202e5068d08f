#!/bin/bash
# round 2, call 3: where does the streaming CTC kernel lose its time (debug switches + ncu per-kernel durations), first
# hardware run of the streaming conv-module kernels, suite, bench with the cheap bounded waits.
mkdir -p gpurun_out
S=gpurun_out/call3_summary.txt
: > $S
stage() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/$name.log 2>&1; local rc=$?
          echo "$name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" >> $S; }
stage conv_stream_tests 300 python -m pytest tests/test_backward_gpu.py tests/test_modules_gpu.py -q --timeout=120 -p no:cacheprovider -k "conv_module or conformer_conv"
stage conv_prof_stream 200 python profiles/prof_conv.py
stage conv_prof_legacy 200 env NSP_CONV_PATH=legacy python profiles/prof_conv.py
stage ctc_dbg0 100 python profiles/prof_ctc.py
stage ctc_dbg1 100 env NSP_CTC_DEBUG=1 python profiles/prof_ctc.py
stage ctc_dbg2 100 env NSP_CTC_DEBUG=2 python profiles/prof_ctc.py
stage ctc_dbg3 100 env NSP_CTC_DEBUG=3 python profiles/prof_ctc.py
stage ctc_ncu 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ctc -c 40 --csv --log-file gpurun_out/ctc_ncu.csv python profiles/prof_ctc.py
stage suite_default 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider
stage bench_default 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
stage bench_pair 600 env NSP_GEMM_EPILOGUE=pair python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager
stage bench_eager 600 python bench.py --impl eager --steps 5 --warmup 3
cat $S
