#!/bin/bash
# One GPU-box visit: parity tests, bench (bf16 + fp32), ncu launch list, ncu --set full of the conv kernels.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_round.sh TAG [skip-tests]'
TAG=${1:-rXX}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
  tail -3 $O/pytest_gpu.log
fi
timeout 600 python bench.py --precision bf16 --steps 5 --warmup 3 > $O/bench_bf16.json 2> $O/bench_bf16.err; echo "bench bf16 rc=$?"
cat $O/bench_bf16.json | head -c 1500; echo
timeout 600 python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_fp32.json 2> $O/bench_fp32.err; echo "bench fp32 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_bf16.csv \
   python bench.py --precision bf16 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:conv_tc_kernel -c 75 -f -o /tmp/conv_full \
   python bench.py --precision bf16 --steps 1 --warmup 0 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/conv_full.ncu-rep --page raw --csv > $O/conv_full_raw.csv 2>/dev/null
ls -la /tmp/conv_full.ncu-rep
ls -la $O
