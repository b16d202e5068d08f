#!/bin/bash
# round 2: tensor-core LSTM after the publish reordering + bench lines of the RNN workloads with the stock-torch eager arm
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_lstm_tc_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/lstm_tc_tests.log
timeout 200 python profiles/prof_lstm.py 2>&1 | tail -8 | tee gpurun_out/lstm_tc_prof.log
timeout 500 python bench.py --workload c4_lstm_rnnt --steps 10 --warmup 3 > gpurun_out/r02_bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 3000 gpurun_out/r02_bench_c4.json; tail -5 gpurun_out/bench_c4.err
timeout 300 python bench.py --workload c1_blstm_ctc --steps 20 --warmup 5 > gpurun_out/r02_bench_c1.json 2> gpurun_out/bench_c1.err; tail -c 2500 gpurun_out/r02_bench_c1.json; tail -5 gpurun_out/bench_c1.err
