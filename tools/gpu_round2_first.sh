#!/bin/bash
# First GPU visit of round 2: the three probes of the experimental (default-off) kernels written blind at the end of round 1.
# usage: gpurun --timeout 420 -- 'bash tools/gpu_round2_first.sh r02a'
TAG=${1:-r02a}
O=gpurun_out/$TAG
mkdir -p $O
timeout 60 python tests/probe_umma.py > $O/probe_umma.log 2>&1; echo "probe_umma rc=$?"; tail -4 $O/probe_umma.log
timeout 60 python tests/probe_pool.py > $O/probe_pool.log 2>&1; echo "probe_pool rc=$?"; tail -5 $O/probe_pool.log
# the halo kernel may hang if a barrier is wrong: short timeout, and it runs last
timeout 150 python tests/probe_halo.py > $O/probe_halo.log 2>&1; echo "probe_halo rc=$?"; tail -25 $O/probe_halo.log
# where the 18 ms between the device-timed step and the plug-in path go (cProfile of the e2e loop, stderr)
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-e2e > $O/bench_profile_e2e.json 2> $O/bench_profile_e2e.err; echo "bench --profile-e2e rc=$?"; head -45 $O/bench_profile_e2e.err
