#!/bin/bash
mkdir -p gpurun_out
export PCR_KNN_IMPL=warp PCR_OCC_SCALE=${PCR_OCC_SCALE:-0.45}
python tools/profile_once.py batch 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sor_mean_kernel" -c 2 -f -o gpurun_out/prof_warp python tools/profile_once.py batch 2 > gpurun_out/prof_ncu.log 2>&1
echo "ncu rc=$?"
