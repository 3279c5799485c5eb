#!/bin/bash
# Light ncu captures (source counters + warp states) of several windows of conv_tc launches on the small workload.
# usage: gpurun --timeout 400 -- 'bash tools/gpu_ncu_multi.sh TAG "SKIP:COUNT SKIP:COUNT ..."'
TAG=${1:-rXX}
O=gpurun_out/$TAG
mkdir -p $O
for w in $2; do
  SKIP=${w%%:*}; COUNT=${w##*:}
  timeout 100 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section MemoryWorkloadAnalysis \
     --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip $SKIP --launch-count $COUNT -f -o $O/conv_src_$SKIP \
     python bench.py --workload mot17 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_src_$SKIP.log 2>&1; echo "ncu skip=$SKIP rc=$?"
done
ls -la $O
