#!/bin/bash
# Round record: all GPU tests, the default bench line, the reference arm, smoke(), the ncu launch list of one frame (DRAM bytes +
# tensor-pipe activity per launch) and a full ncu capture of one launch of each kernel the roofline talks about.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_record.sh TAG'
TAG=${1:-rXX}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 500 python -m pytest tests -m gpu -x -q --durations=6 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_bf16.json 2> $O/bench_bf16.err; echo "bench bf16 rc=$?"
python tools/bench_brief.py $O/bench_bf16.json; tail -3 $O/bench_bf16.err
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?"
head -c 500 $O/bench_reference.json; echo
timeout 120 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
   --clock-control none -c 1200 --csv --log-file $O/launches_bf16.csv \
   python bench.py --sequences 1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --adapter-frames 0 > $O/ncu_launches.log 2>&1; echo "ncu list rc=$?"
python tools/ncu_launch_summary.py $O/launches_bf16.csv --one-step --json $O/ncu_traffic.json > $O/launches_bf16_summary.md 2>&1; head -40 $O/launches_bf16_summary.md
