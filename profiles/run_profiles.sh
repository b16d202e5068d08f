#!/bin/bash
# Run on the GPU box (via gpurun) from the repo root: launch list of one steady bench step + full captures of the
# dominant GEMM and of the CTC kernels.  Outputs land in gpurun_out/ and are summarised into profiles/ afterwards.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1400 -c 760 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 5 -c 1 -f -o gpurun_out/gemm_ffn1 \
    python profiles/prof_ops.py gemm_ffn1 > gpurun_out/gemm_ffn1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctc_ -s 20 -c 4 -f -o gpurun_out/ctc \
    python profiles/prof_ops.py ctc > gpurun_out/ctc.log 2>&1
python profiles/prof_ops.py ctc > gpurun_out/ctc_times.log 2>&1
python profiles/prof_ops.py gemm_ffn1 > gpurun_out/gemm_times.log 2>&1
python profiles/prof_ops.py gemm_ffn2 >> gpurun_out/gemm_times.log 2>&1
python profiles/prof_ops.py gemm_qkv >> gpurun_out/gemm_times.log 2>&1
