#!/bin/bash
# Evidence visit: compute-sanitizer memcheck of the smoke run (fp32 + bf16 paths), ncu --set full of one launch of the kernels the
# roofline discussion names (dominant conv instantiation, halo 3x3, gather, geometry), raw CSV export.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh TAG'
TAG=${1:-rXX}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > $O/memcheck_smoke.log 2>&1; echo "memcheck rc=$?"
tail -6 $O/memcheck_smoke.log
timeout 500 ncu --set full --clock-control none --import-source on \
   -k regex:"conv_tc_kernel|conv3x3_halo_kernel|crop_resize_kernel|frame_geometry_kernel|stem_march_kernel|linear_fused_kernel|gram_stats_kernel" \
   --launch-skip 330 --launch-count 140 -f -o $O/full \
   python bench.py --sequences 1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --adapter-frames 0 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 $O/ncu_full.log | cut -c1-200
ncu -i $O/full.ncu-rep --page raw --csv > $O/full_raw.csv 2>/dev/null; rm -f $O/full.ncu-rep
python tools/ncu_full_summary.py $O/full_raw.csv > $O/full_summary.md 2>&1; head -60 $O/full_summary.md
ls -la $O | head -20
