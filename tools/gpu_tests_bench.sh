#!/bin/bash
# GPU test suite (conv kernels + parity) and the MOT20 bench; usage: gpu_tests_bench.sh TAG [pytest args]   (BENCH_ARGS, SKIP_TESTS env)
TAG=${1:-r02x}
shift
O=gpurun_out/$TAG
mkdir -p $O
if [ -z "$SKIP_TESTS" ]; then
timeout 1200 python -m pytest tests -x -q -m gpu "$@" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:---sequences 2 --no-e2e --adapter-frames 0} > $O/bench.json 2> $O/bench.err
echo "bench rc=$?"; tail -3 $O/bench.err; python -c "
import json,sys
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('ms/frame', d['ms_per_frame'], 'value', d['value'], 'e2e', d['e2e']['value'], 'p50/p99', d['p50_frame_latency_ms'], d['p99_frame_latency_ms'])
r=d['roofline']; print('roofline', r['kernel'], r['frac'], 'executed', r['executed_frac'], 'step_frac', r['step_frac'], 'hbm', r['step_hbm_frac'])
print('adapter', d.get('e2e_adapter')); print('parity', d.get('parity_vs_reference_golden'), 'rows', d.get('result_rows_gathered'))
print({k:v['ms_per_frame'] for k,v in d['kernels'].items() if v['ms_per_frame']>0.4})
for k,v in sorted(d['conv_detail_ms_per_frame'].items(), key=lambda kv:-kv[1]): print('  ',k,v)
"
