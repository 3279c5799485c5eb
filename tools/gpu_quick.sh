#!/bin/bash
# Short GPU-box visit (budget-bounded): parity tests, one bench line, the reference arm, ncu launch list.
# usage: gpurun --timeout 900 -- 'bash tools/gpu_quick.sh TAG'
TAG=${1:-rXX}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 420 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -14 $O/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_bf16.json 2> $O/bench_bf16.err; echo "bench bf16 rc=$?"
head -c 1500 $O/bench_bf16.json; echo; tail -3 $O/bench_bf16.err
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?"
head -c 600 $O/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
   --clock-control none -c 3000 --csv --log-file $O/launches_bf16.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_launches.log 2>&1; echo "ncu list rc=$?"
python tools/ncu_launch_summary.py $O/launches_bf16.csv --json $O/ncu_traffic.json > $O/launches_bf16_summary.md 2>&1; head -14 $O/launches_bf16_summary.md
ls -la $O
