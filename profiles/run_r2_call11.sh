#!/bin/bash
# round 2, call 11: everything on the new defaults (TMA/pair GEMMs, PDL, streaming CTC + conv module): suite, the driver's
# bench line, the other BASELINE configs, launch lists (ncu) of one train step and one forward step.
mkdir -p gpurun_out
S=gpurun_out/call11_summary.txt
: > $S
stage() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/$name.log 2>&1; local rc=$?
          echo "$name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" >> $S; }
stage suite 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider
stage bench_default 900 python bench.py --steps 20 --warmup 5 --shape-profile gpurun_out/r02_gemm_shapes_train.json
stage bench_m 600 python bench.py --workload conformer_m_ctc --steps 10 --warmup 3 --no-cpu-baseline
stage bench_c1 300 python bench.py --workload c1_blstm_ctc --steps 10 --warmup 3
stage bench_c4 600 python bench.py --workload c4_lstm_rnnt --steps 5 --warmup 3
stage bench_c5 600 python bench.py --workload c5_transformer_t3000 --steps 5 --warmup 3
stage bench_lengths 400 python bench.py --steps 10 --warmup 3 --lengths librispeech --no-cpu-baseline --no-eager
stage bench_dropout 400 python bench.py --steps 10 --warmup 3 --dropout 0.1 --no-cpu-baseline --no-eager
stage ncu_train 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train_step.csv python bench.py --ncu-step --no-cpu-baseline --no-eager
stage ncu_fwd 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_fwd_step.csv python bench.py --ncu-step --step fwd --no-cpu-baseline --no-eager
cat $S
