#!/bin/bash
# Light ncu capture (source counters + warp states, few passes) of a window of conv_tc launches on the small workload.
# usage: gpurun --timeout 400 -- 'bash tools/gpu_ncu_src.sh TAG SKIP COUNT [skip-tests]'
TAG=${1:-rXX}; SKIP=${2:-7}; COUNT=${3:-3}
O=gpurun_out/$TAG
mkdir -p $O
if [ "$4" != "skip-tests" ]; then
  timeout 200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
fi
timeout 170 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section MemoryWorkloadAnalysis \
   --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip $SKIP --launch-count $COUNT -f -o $O/conv_src \
   python bench.py --workload mot17 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_src.log 2>&1; echo "ncu rc=$?"
tail -4 $O/ncu_src.log
ls -la $O
