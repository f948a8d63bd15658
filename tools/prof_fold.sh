#!/bin/bash
# ncu capture of sor_stats_kernel (the exact sequential-fold statistics of SOR) on the voxelised bench frame
mkdir -p gpurun_out
python tools/profile_once.py batch 2 > gpurun_out/prof_plain.log 2>&1 || { tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"sor_stats_kernel" -c 2 -f -o gpurun_out/prof_fold python tools/profile_once.py batch 2 > gpurun_out/prof_ncu.log 2>&1
echo "ncu rc=$?"
