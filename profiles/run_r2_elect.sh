#!/bin/bash
# round 2: elect.sync for every single-thread role (TMA producers, MMA issuers, epilogue bulk stores): full GPU suite + benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/elect_tests.log; cat gpurun_out/elect_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-eager --no-cpu-baseline > gpurun_out/r02_bench_elect.json 2> gpurun_out/bench_elect.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_elect.json').read().strip().splitlines()[-1])
print('default', d['ms_per_step'], d['value'], 'fwd', d['fwd']['ms_per_step'], 'e2e', d['e2e']['value'], 'gemm frac', d['roofline']['frac'], 'ctc', d.get('ctc_loss_ms_per_batch'))
print({k: round(v, 3) for k, v in list(d['kernel_time_ms_per_step'].items())[:12]})
P
tail -3 gpurun_out/bench_elect.err
timeout 300 python bench.py --workload c4_lstm_rnnt --steps 10 --warmup 3 --no-eager > gpurun_out/r02_bench_c4_elect.json 2>/dev/null; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_c4_elect.json').read().strip().splitlines()[-1])
print('c4', d['ms_per_step'], d['value'], {k: round(v, 3) for k, v in list(d['kernel_time_ms_per_step'].items())[:6]})
P
