#!/bin/bash
# ncu --set full of the SURVEY 8(f) kernels and the IoU / centre-distance matrices (tools/rounds_evidence.py), digested on the box.
# usage: gpurun --timeout 900 -- 'bash tools/gpu_rounds_evidence.sh TAG'
TAG=${1:-rXX}
O=gpurun_out/$TAG
mkdir -p $O
python tools/rounds_evidence.py > $O/rounds_events.json 2> $O/rounds_events.err; echo "events rc=$?"; cat $O/rounds_events.json | head -80
timeout 600 ncu --set full --clock-control none --import-source on \
   -k regex:"pair_matrix_kernel|match_cost_kernel|assignment_kernel|kalman_|duplicate_kernel|coverage_kernel|frame_ingest_kernel|ecc_" \
   --launch-count 60 -f -o $O/rounds_full python tools/rounds_evidence.py > $O/ncu_rounds.log 2>&1; echo "ncu rc=$?"
tail -3 $O/ncu_rounds.log | cut -c1-200
ncu -i $O/rounds_full.ncu-rep --page raw --csv > $O/rounds_raw.csv 2>/dev/null; rm -f $O/rounds_full.ncu-rep
python tools/ncu_full_summary.py $O/rounds_raw.csv > $O/rounds_summary.md 2>&1; cat $O/rounds_summary.md
