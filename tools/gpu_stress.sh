#!/bin/bash
# BASELINE.json configs[4] at full size (1000 tracks x 10 proposals, L = 30) + the GPU tests.  usage: gpurun --timeout 900 -- 'bash tools/gpu_stress.sh TAG'
TAG=${1:-rXX}
O=gpurun_out/$TAG
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
timeout 400 python bench.py --workload stress --sequences 1 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --adapter-frames 0 > $O/bench_stress.json 2> $O/bench_stress.err; echo "stress rc=$?"
python tools/bench_brief.py $O/bench_stress.json
tail -5 $O/bench_stress.err
