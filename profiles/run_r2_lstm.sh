#!/bin/bash
# round 2: first hardware run of the tensor-core LSTM recurrence (lstm_tc.cu)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_lstm_tc_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/lstm_tc_tests.log
cat gpurun_out/lstm_tc_tests.log
timeout 200 python profiles/prof_lstm.py 2>&1 | tail -8 | tee gpurun_out/lstm_tc_prof.log
timeout 400 python -m pytest tests/test_rnn_gpu.py tests/test_zz_streaming_gpu.py -q -k "lstm or rnn or blstm" 2>&1 | tail -8 | tee gpurun_out/lstm_rnn_tests.log
