#!/bin/bash
# PDL A/B on the clean timed region + source-counter capture of the first conv_tc launches (stem, layer1 incl. the DUAL final pass)
TAG=${1:-r02u}
O=gpurun_out/$TAG
mkdir -p $O
for v in 1 0; do
  env BUSCA_PDL=$v timeout 200 python bench.py --sequences 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_PDL_$v.json 2> $O/bench_PDL_$v.err; echo "bench PDL=$v rc=$?"
  python tools/bench_brief.py $O/bench_PDL_$v.json
  tail -3 $O/bench_PDL_$v.err
done
timeout 400 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section MemoryWorkloadAnalysis \
   --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 0 --launch-count 12 -f -o $O/conv_src \
   python bench.py --workload mot17 --sequences 1 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_src.log 2>&1; echo "ncu rc=$?"
tail -4 $O/ncu_src.log
ls -la $O
