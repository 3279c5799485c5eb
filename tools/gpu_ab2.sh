#!/bin/bash
# A/B visit with fewer words: GPU tests, then bench with VAR=1 and VAR=0 (one line each + per-class times).
# usage: gpurun --timeout 600 -- 'bash tools/gpu_ab2.sh TAG ENVVAR [skip-tests]'
TAG=${1:-rXX}; VAR=${2:-BUSCA_PDL}
O=gpurun_out/$TAG
mkdir -p $O
if [ "$3" != "skip-tests" ]; then
  timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
  tail -5 $O/pytest_gpu.log
fi
for v in 1 0; do
  env $VAR=$v timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_${VAR}_$v.json 2> $O/bench_${VAR}_$v.err; echo "bench $VAR=$v rc=$?"
  python tools/bench_brief.py $O/bench_${VAR}_$v.json
  tail -3 $O/bench_${VAR}_$v.err
done
