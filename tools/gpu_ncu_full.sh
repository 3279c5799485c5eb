#!/bin/bash
# ncu --set full of a window of conv_tc launches (memory batch), raw + source CSV export.
# usage: gpurun --timeout 600 -- 'bash tools/gpu_ncu_full.sh TAG SKIP COUNT'
TAG=${1:-rXX}; SKIP=${2:-6}; COUNT=${3:-34}
O=gpurun_out/$TAG
mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16 or dedup or reid" > $O/pytest_subset.log 2>&1; echo "pytest subset rc=$?"; tail -3 $O/pytest_subset.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip $SKIP --launch-count $COUNT -f -o $O/conv_full \
   python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 $O/ncu_full.log
ncu -i $O/conv_full.ncu-rep --page raw --csv > $O/conv_full_raw.csv 2>/dev/null
ls -la $O
