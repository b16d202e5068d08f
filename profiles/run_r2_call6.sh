#!/bin/bash
# round 2, call 6: programmatic dependent launch on every kernel (A/B with NSP_PDL=0), CTC sweep trace.
mkdir -p gpurun_out
S=gpurun_out/call6_summary.txt
: > $S
stage() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/$name.log 2>&1; local rc=$?
          echo "$name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" >> $S; }
stage pdl_smoke 120 env NSP_GEMM_EPILOGUE=pair python -m pytest tests/test_gemm_gpu.py tests/test_ctc_gpu.py -q --timeout=60 -p no:cacheprovider -x
stage ctc_trace 100 env NSP_CTC_DEBUG=8 python profiles/prof_ctc.py
stage suite_pdl 900 env NSP_GEMM_EPILOGUE=pair python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider
stage bench_pair_pdl 600 env NSP_GEMM_EPILOGUE=pair python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager
stage bench_pair_nopdl 600 env NSP_GEMM_EPILOGUE=pair NSP_PDL=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager
stage bench_m_pdl 600 env NSP_GEMM_EPILOGUE=pair python bench.py --workload conformer_m_ctc --steps 10 --warmup 3 --no-cpu-baseline --no-eager
cat $S
