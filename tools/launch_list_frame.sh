#!/bin/bash
# launch list of the whole device step (voxel -> fused SOR + normals) under ncu: serialised; first with ncu's default cache
# flush before every kernel (cold), then with --cache-control none (warm, as inside the step).  Shares, not absolutes.
mkdir -p gpurun_out
python tools/profile_once.py frame 4 > gpurun_out/llf_plain.log 2>&1 || { tail -5 gpurun_out/llf_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_frame.csv python tools/profile_once.py frame 4 > gpurun_out/llf_ncu.log 2>&1
echo rc=$?
python tools/parse_launches.py gpurun_out/launches_frame.csv 4 > gpurun_out/launches_frame_cold.txt
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/launches_frame_warm.csv python tools/profile_once.py frame 4 > gpurun_out/llf_ncu2.log 2>&1
echo rc=$?
python tools/parse_launches.py gpurun_out/launches_frame_warm.csv 4 > gpurun_out/launches_frame_warm.txt
paste -d'|' <(cut -c1-75 gpurun_out/launches_frame_cold.txt) <(cut -c62-75 gpurun_out/launches_frame_warm.txt)
