#!/bin/bash
# One GPU-box visit: parity tests, bench (bf16 default, dedup off, fp32), ncu launch list with DRAM bytes and tensor-pipe
# activity, ncu --set full of a few conv kernels.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_round.sh TAG [skip-tests] [skip-full]'
TAG=${1:-rXX}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
  tail -5 $O/pytest_gpu.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_bf16.json 2> $O/bench_bf16.err; echo "bench bf16 rc=$?"
head -c 1800 $O/bench_bf16.json; echo; tail -3 $O/bench_bf16.err
BUSCA_DEDUP=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_bf16_nodedup.json 2> $O/bench_bf16_nodedup.err; echo "bench nodedup rc=$?"
head -c 600 $O/bench_bf16_nodedup.json; echo
timeout 600 python bench.py --precision fp32 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_fp32.json 2> $O/bench_fp32.err; echo "bench fp32 rc=$?"
head -c 400 $O/bench_fp32.json; echo
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?"
cat $O/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
   --clock-control none -c 3000 --csv --log-file $O/launches_bf16.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_launches.log 2>&1; echo "ncu list rc=$?"
python tools/ncu_launch_summary.py $O/launches_bf16.csv --json $O/ncu_traffic.json > $O/launches_bf16_summary.md 2>&1; head -14 $O/launches_bf16_summary.md
if [ "$3" != "skip-full" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 60 -f -o $O/conv_full \
     python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
  ncu -i $O/conv_full.ncu-rep --page raw --csv > $O/conv_full_raw.csv 2>/dev/null
fi
ls -la $O
