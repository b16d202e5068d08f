#!/bin/bash
# round 2, call 5: streaming CTC v2 (signal warp, log2 sweep), cheap bounded waits in the tma/pair GEMMs, bench lines.
mkdir -p gpurun_out
S=gpurun_out/call5_summary.txt
: > $S
stage() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/$name.log 2>&1; local rc=$?
          echo "$name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" >> $S; }
stage ctc_tests 300 python -m pytest tests/test_ctc_gpu.py -q --timeout=120 -p no:cacheprovider
stage ctc_prof 100 python profiles/prof_ctc.py
stage ctc_prof_nolat 100 env NSP_CTC_DEBUG=1 python profiles/prof_ctc.py
stage ctc_ncu 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ctc -c 12 --csv --log-file gpurun_out/ctc_ncu.csv python profiles/prof_ctc.py
stage suite_pair 900 env NSP_GEMM_EPILOGUE=pair python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider
stage bench_pair 600 env NSP_GEMM_EPILOGUE=pair python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --shape-profile gpurun_out/shapes_pair.json
stage bench_tma 600 env NSP_GEMM_EPILOGUE=tma python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager
stage bench_m_pair 600 env NSP_GEMM_EPILOGUE=pair python bench.py --workload conformer_m_ctc --steps 10 --warmup 3 --no-cpu-baseline --no-eager
cat $S
