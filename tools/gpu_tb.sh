#!/bin/bash
# GPU tests + one short bench line (2 sequences).  usage: gpurun --timeout 700 -- 'bash tools/gpu_tb.sh TAG [pytest -k expr]'
TAG=${1:-rXX}
O=gpurun_out/$TAG
mkdir -p $O
if [ -n "$2" ]; then
  timeout 400 python -m pytest tests -m gpu -x -q -k "$2" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
else
  timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
fi
tail -15 $O/pytest_gpu.log
timeout 200 python bench.py --sequences 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python tools/bench_brief.py $O/bench.json -v
tail -3 $O/bench.err
