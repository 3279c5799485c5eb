#!/bin/bash
TAG=${1:-r02x}
O=gpurun_out/$TAG
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 200 python bench.py --sequences 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python tools/bench_brief.py $O/bench.json
tail -3 $O/bench.err
timeout 300 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section MemoryWorkloadAnalysis \
   --clock-control none --import-source on -k regex:"stem_march_kernel|stem_prepass_kernel|gram_fold_final" --launch-skip 0 --launch-count 6 -f -o $O/stem_src \
   python bench.py --workload mot17 --sequences 1 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_src.log 2>&1; echo "ncu rc=$?"
tail -2 $O/ncu_src.log | cut -c1-300
