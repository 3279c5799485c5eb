#!/bin/bash
# Round-end record on a tight GPU budget: selected tests, the bench line (with the CPU sample), the ncu launch list
# (DRAM bytes + tensor-pipe activity per launch), the reference arm and smoke().
# usage: gpurun --timeout 260 -- 'bash tools/gpu_final.sh TAG "PYTEST_K_EXPR"'
TAG=${1:-rXX}; KEXPR=${2:-long_history}
O=gpurun_out/$TAG
mkdir -p $O
timeout 80 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $O/pytest_sel.log 2>&1; echo "pytest($KEXPR) rc=$?"; tail -3 $O/pytest_sel.log
timeout 90 python bench.py --steps 5 --warmup 3 > $O/bench_bf16.json 2> $O/bench_bf16.err; echo "bench rc=$?"; head -c 700 $O/bench_bf16.json; echo
timeout 100 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
   --clock-control none -c 3000 --csv --log-file $O/launches_bf16.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_launches.log 2>&1; echo "ncu list rc=$?"
python tools/ncu_launch_summary.py $O/launches_bf16.csv --json $O/ncu_traffic.json > $O/launches_bf16_summary.md 2>&1; head -12 $O/launches_bf16_summary.md
timeout 40 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"; head -c 300 $O/bench_reference.json; echo
timeout 40 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
