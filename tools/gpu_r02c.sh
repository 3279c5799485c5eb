#!/bin/bash
# halo v2 (resident weights / two pixel tiles per weight tile): parity + timing, then the MOT20 bench
TAG=${1:-r02c}
O=gpurun_out/$TAG
mkdir -p $O
timeout 150 python tests/probe_halo.py > $O/probe_halo.log 2>&1; echo "probe_halo rc=$?"; tail -18 $O/probe_halo.log
BUSCA_HALO_MT=1 timeout 150 python tests/probe_halo.py > $O/probe_halo_mt1.log 2>&1; echo "probe_halo mt1 rc=$?"; tail -7 $O/probe_halo_mt1.log
BUSCA_HALO=1 BUSCA_POOL_MONO=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_halo1.json 2> $O/bench_halo1.err
echo "bench halo=1 rc=$?"; tail -3 $O/bench_halo1.err; python -c "
import json,sys
d=json.loads(open('$O/bench_halo1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], {k:v['ms_per_step'] for k,v in d['kernels'].items() if v['ms_per_step']>0.4})
print({k:v for k,v in d['conv_detail_ms_per_step'].items() if '3x3' in k})
"
