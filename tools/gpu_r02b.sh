#!/bin/bash
# halo kernel parity + timing, then the MOT20 bench with the experimental kernels on/off
TAG=${1:-r02b}
O=gpurun_out/$TAG
mkdir -p $O
timeout 150 python tests/probe_halo.py > $O/probe_halo.log 2>&1; echo "probe_halo rc=$?"; tail -25 $O/probe_halo.log
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg
  BUSCA_HALO=$1 BUSCA_POOL_MONO=$2 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_halo$1_mono$2.json 2> $O/bench_halo$1_mono$2.err
  echo "bench halo=$1 mono=$2 rc=$?"; python -c "
import json,sys
d=json.loads(open('$O/bench_halo$1_mono$2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.4})
"
done
