#!/bin/bash
# GPU test suite (conv kernels + parity) and the MOT20 bench; usage: gpu_tests_bench.sh TAG [pytest args]
TAG=${1:-r02x}
shift
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu "$@" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench.json 2> $O/bench.err
echo "bench rc=$?"; tail -3 $O/bench.err; python -c "
import json,sys
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.4})
for k,v in sorted(d['conv_detail_ms_per_step'].items(), key=lambda kv:-kv[1]): print('  ',k,v)
"
