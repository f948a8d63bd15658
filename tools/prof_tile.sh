#!/bin/bash
# launch list + ncu capture of the level-0 cell-tile kernel (fused SOR -> normals pipeline, voxelised bench frame)
mkdir -p gpurun_out

PCR_DEBUG=1 python tools/profile_once.py batch 3 > gpurun_out/prof_plain.log 2>&1 || { tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tile.csv python tools/profile_once.py batch 3 > gpurun_out/ll_ncu.log 2>&1
python tools/parse_launches.py gpurun_out/launches_tile.csv 3
ncu --set full --clock-control none --import-source on -k regex:"knn_tile_kernel" -c 2 -f -o gpurun_out/prof_tile python tools/profile_once.py batch 2 > gpurun_out/prof_ncu.log 2>&1
echo "ncu rc=$?"
