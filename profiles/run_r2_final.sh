#!/bin/bash
# round 2, final validation of the shipped build: full GPU suite, smoke(), side-stream weight-gradient A/B, bench lines, ncu lists
mkdir -p gpurun_out
stage() { local name=$1 lim=$2; shift 2; local t0=$(date +%s); timeout $lim "$@"; echo "[stage $name] rc=$? $(( $(date +%s) - t0 ))s"; }
stage tests 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/final_tests.log
stage smoke 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/final_smoke.log
for v in 0 1; do
  NSP_WGRAD_STREAM=$v stage ab$v 300 python bench.py --steps 20 --warmup 5 --no-eager --no-cpu-baseline > gpurun_out/ab_wgrad$v.json 2> gpurun_out/ab_wgrad$v.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/ab_wgrad$v.json').read().strip().splitlines()[-1]); print('NSP_WGRAD_STREAM=$v', d['ms_per_step'], d['value'], 'graph', d['config'].get('cuda_graph'), d['config'].get('cuda_graph_error'), 'loss', d.get('loss'))
except Exception as e: print('ab$v failed', e); print(open('gpurun_out/ab_wgrad$v.err').read()[-1500:])
P
done
NSP_WGRAD_STREAM=1 stage tests_side 600 python -m pytest tests/test_backward_gpu.py tests/test_benchscale_gpu.py tests/test_encoder_gpu.py -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/final_tests_side.log
stage bench 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/bench_final.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1]); print('final', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'fwd', d['fwd']['ms_per_step'], 'eager', d['eager_b200']['ms_per_step'], d.get('speedup_vs_eager_b200'), 'roofline', d['roofline']['frac'], 'ctc', d['roofline_ctc']['frac'], 'loss_check', d.get('loss_check',{}).get('rel_err'), 'cpu', d['cpu_baseline']['value'])
P
stage c4 400 python bench.py --workload c4_lstm_rnnt --steps 10 --warmup 3 > gpurun_out/r02_bench_c4.json 2>/dev/null
stage c1 300 python bench.py --workload c1_blstm_ctc --steps 20 --warmup 5 > gpurun_out/r02_bench_c1.json 2>/dev/null
python - <<P
import json
for f in ('c4','c1'):
    d=json.loads(open('gpurun_out/r02_bench_%s.json'%f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], 'eager', d['eager_b200']['ms_per_step'], d.get('speedup_vs_eager_b200'))
P
stage ncu_train 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train_step_final.csv python bench.py --ncu-step --no-cpu-baseline --no-eager > /dev/null 2>&1
stage ncu_lstm 300 ncu --set full --clock-control none --import-source on -f -k "regex:lstm_tc" -c 2 -o gpurun_out/r02_lstm_tc python profiles/prof_lstm.py > /dev/null 2>&1
ls -la gpurun_out | tail -20
