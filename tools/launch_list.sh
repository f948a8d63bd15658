#!/bin/bash
# launch list (ncu, serialised, cold cache: shares not absolutes) of the fused pipeline on the voxelised bench frame
mkdir -p gpurun_out
python tools/profile_once.py batch 3 > gpurun_out/ll_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python tools/profile_once.py batch 3 > gpurun_out/ll_ncu.log 2>&1
echo rc=$?
python tools/parse_launches.py gpurun_out/launches_r02.csv 3
