#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_ctc_gpu.py -q --timeout=120 -p no:cacheprovider -x > gpurun_out/ctc_tests.log 2>&1; tail -1 gpurun_out/ctc_tests.log
NSP_CTC_DEBUG=8 timeout 100 python profiles/prof_ctc.py > gpurun_out/ctc_trace.log 2>&1
timeout 100 python profiles/prof_ctc.py > gpurun_out/ctc_prof.log 2>&1
NSP_CTC_DEBUG=1 timeout 100 python profiles/prof_ctc.py > gpurun_out/ctc_nolat.log 2>&1
