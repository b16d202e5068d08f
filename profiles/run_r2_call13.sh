#!/bin/bash
# round 2, call 13: conv3x3_tc with the bare spin on its TMA / MMA threads; ncu --set full captures of the top kernels.
mkdir -p gpurun_out
S=gpurun_out/call13_summary.txt
: > $S
stage() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout $secs "$@" > gpurun_out/$name.log 2>&1; local rc=$?
          echo "$name rc=$rc $(( $(date +%s) - t0 ))s :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)" >> $S; }
stage suite 600 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider
stage bench 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager
NCU="ncu --set full --clock-control none --import-source on -f"
stage full_gemm_ffn1 200 $NCU -k regex:gemm_ts_kernel -s 5 -c 1 -o gpurun_out/r02_gemm_ffn1 python profiles/prof_ops.py gemm_ffn1
stage full_gemm_ffn2 200 $NCU -k regex:gemm_ts_kernel -s 5 -c 1 -o gpurun_out/r02_gemm_ffn2 python profiles/prof_ops.py gemm_ffn2
stage full_wgrad 200 $NCU -k regex:wgrad_kernel -s 5 -c 1 -o gpurun_out/r02_wgrad python profiles/prof_ops.py wgrad
stage full_ctc 200 $NCU -k regex:ctc_ -s 4 -c 2 -o gpurun_out/r02_ctc python profiles/prof_ctc.py
stage full_conv 200 $NCU -k "regex:conv_ln_kernel|dwconv_bwd_kernel" -s 6 -c 3 -o gpurun_out/r02_convmod python profiles/prof_conv.py
stage full_attn 200 $NCU -k regex:attn_tc_kernel -s 2 -c 1 -o gpurun_out/r02_attn_fwd python profiles/prof_ops.py attn_fwd
stage full_lnbwd 400 $NCU -k "regex:layernorm_bwd_kernel|act_bwd_bias_kernel" --profile-from-start off -c 4 -o gpurun_out/r02_lnbwd python bench.py --ncu-step --no-cpu-baseline --no-eager
for f in gemm_ffn1 gemm_ffn2 wgrad ctc attn_fwd; do python profiles/prof_ops.py $f 2>/dev/null | tail -3 >> gpurun_out/op_times.log; done
cat $S
