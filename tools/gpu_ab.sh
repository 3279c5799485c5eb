#!/bin/bash
# A/B visit: conv + parity tests on the default build, then bench with each value of one env toggle.
# usage: gpurun --timeout 600 -- 'bash tools/gpu_ab.sh TAG ENVVAR [skip-tests]'
TAG=${1:-rXX}; VAR=${2:-BUSCA_RESB}
O=gpurun_out/$TAG
mkdir -p $O
if [ "$3" != "skip-tests" ]; then
  timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
  tail -6 $O/pytest_gpu.log
fi
for v in 1 0; do
  env $VAR=$v timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_${VAR}_$v.json 2> $O/bench_${VAR}_$v.err; echo "bench $VAR=$v rc=$?"
  python - $O/bench_${VAR}_$v.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms_per_step",d["ms_per_step"],"value",d["value"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"],d["roofline"]["kernel"])
    print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items() if v["ms_per_step"]>0.4})
    for k,v in d["conv_detail_ms_per_step"].items(): print(f"   {k:45s} {v:7.3f}")
except Exception as e:
    print("bad json",e); print(open(sys.argv[1]).read()[:500])
PY
  tail -3 $O/bench_${VAR}_$v.err
done
