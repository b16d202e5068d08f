#!/bin/bash
# round 2, call 2: streaming CTC bring-up (risky first, own timeout), whole GPU suite without the experimental gate
# (default / tma / pair GEMM epilogues), the new bench line, the reference arm, per-shape GEMM profile in pair mode.
mkdir -p gpurun_out
S=gpurun_out/call2_summary.txt
: > $S
stage() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/$name.log 2>&1; local rc=$?
          echo "$name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" >> $S; }
stage ctc_stream_tests 300 python -m pytest tests/test_ctc_gpu.py -q --timeout=120 -p no:cacheprovider -x
stage ctc_legacy_tests 300 env NSP_CTC_PATH=legacy python -m pytest tests/test_ctc_gpu.py -q --timeout=120 -p no:cacheprovider
stage ctc_prof_stream 200 python profiles/prof_ctc.py
stage ctc_prof_legacy 200 env NSP_CTC_PATH=legacy python profiles/prof_ctc.py
stage suite_default 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider
stage suite_tma 900 env NSP_GEMM_EPILOGUE=tma python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider
stage suite_pair 900 env NSP_GEMM_EPILOGUE=pair python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider
stage bench_default 900 python bench.py --steps 10 --warmup 3
stage bench_pair 600 env NSP_GEMM_EPILOGUE=pair python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --shape-profile gpurun_out/shapes_pair.json
stage bench_reference 600 python bench.py --impl reference --steps 3 --warmup 1
stage bench_m 600 env NSP_GEMM_EPILOGUE=pair python bench.py --workload conformer_m_ctc --steps 10 --warmup 3 --no-cpu-baseline
cat $S
