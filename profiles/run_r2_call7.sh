#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/call7_summary.txt
: > $S
stage() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/$name.log 2>&1; local rc=$?
          echo "$name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" >> $S; }
stage ctc_tests 300 python -m pytest tests/test_ctc_gpu.py -q --timeout=120 -p no:cacheprovider -x
stage ctc_trace 100 env NSP_CTC_DEBUG=8 python profiles/prof_ctc.py
stage ctc_prof 100 python profiles/prof_ctc.py
stage ctc_nolat 100 env NSP_CTC_DEBUG=1 python profiles/prof_ctc.py
cat $S
