#!/bin/bash
# Round-2 opening sequence, meant for ONE `gpurun --timeout 1500 -- 'bash profiles/run_round2_validation.sh'` call
# (≈ 12-15 GPU-minutes).  Every stage runs under its own `timeout`, writes its log under gpurun_out/ and never stops the
# script: read gpurun_out/round2_summary.txt first.  Order = cheapest / most informative first (NOTES.md).
mkdir -p gpurun_out
S=gpurun_out/round2_summary.txt
: > $S
stage() {            # stage <name> <seconds> <command...>
    local name=$1 secs=$2; shift 2
    local t0=$(date +%s)
    timeout $secs "$@" > gpurun_out/$name.log 2>&1
    local rc=$?
    echo "$name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/$name.log | cut -c1-200)" >> $S
}
# 1. everything added after the round-1 GPU budget was spent (simple kernels + host logic already pinned on CPU)
stage zz_tests 600 env NSP_EXPERIMENTAL=1 python -m pytest tests -m gpu -q -k "zz and not (streamed_chunks and bf16)" --timeout=120 -p no:cacheprovider
stage stream_kernels_gbs 200 python profiles/prof_ops.py stream_kernels      # algorithmic GB/s of the new HBM-bound kernels
# 2. the already-validated suite (the restructured forward paths run through it)
stage validated_suite 900 python -m pytest tests -m gpu -q -k "not zz" --timeout=300 -p no:cacheprovider
# 3. opt-in size classes / kernels, riskiest last, on the bring-up build (bounded mbarrier waits in EVERY tcgen05 kernel: a
#    protocol bug traps with a message instead of hanging the GPU)
[ -f neural_sp_b200/libnsp_b200_dbg.so ] || stage build_dbg 400 make -C neural_sp_b200/csrc debug -j 32
export NSP_LIB_PATH=$PWD/neural_sp_b200/libnsp_b200_dbg.so
stage experimental_stream_bf16 300 env NSP_EXPERIMENTAL=1 python -m pytest tests/test_zz_streaming_gpu.py -q -k "streamed_chunks and bf16" --timeout=120
stage gemm_tma_epilogue 300 env NSP_EXPERIMENTAL=1 python -m pytest tests/test_gemm_tma_epilogue_gpu.py -q --timeout=60 -k "not cta_pairs"
stage gemm_cta_pairs 300 env NSP_EXPERIMENTAL=1 python -m pytest tests/test_gemm_tma_epilogue_gpu.py -q --timeout=60 -k "cta_pairs"
unset NSP_LIB_PATH
# 4. bench lines: default, recipe dropout, the two opt-in GEMM epilogue modes
stage bench_default 600 python bench.py --steps 10 --warmup 3
stage bench_dropout 400 python bench.py --steps 10 --warmup 3 --dropout 0.1 --no-cpu-baseline
stage bench_librispeech_lengths 400 python bench.py --steps 10 --warmup 3 --lengths librispeech --no-cpu-baseline
stage bench_tma 400 env NSP_GEMM_EPILOGUE=tma python bench.py --steps 10 --warmup 3 --no-cpu-baseline
stage bench_pair 400 env NSP_GEMM_EPILOGUE=pair python bench.py --steps 10 --warmup 3 --no-cpu-baseline
cat $S
