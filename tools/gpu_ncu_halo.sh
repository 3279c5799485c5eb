#!/bin/bash
# ncu --set full + source counters of the halo kernel launches of tests/probe_halo.py (timing part: N = 512)
TAG=${1:-r02d}
SKIP=${2:-9}
COUNT=${3:-6}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo_kernel --launch-skip $SKIP --launch-count $COUNT -o $O/halo -f python tests/probe_halo.py > $O/ncu.log 2>&1
echo "ncu rc=$?"; tail -5 $O/ncu.log; ls -la $O
