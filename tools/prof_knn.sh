#!/bin/bash
# ncu capture of the level-0 KNN kernels of the fused SOR -> normals pipeline on the voxelised bench frame
mkdir -p gpurun_out
export PCR_OCC_SCALE=${PCR_OCC_SCALE:-0.25}
PCR_DEBUG=1 python tools/profile_once.py batch 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"knn_sel_kernel|knn_thread_kernel" -c 4 -f -o gpurun_out/prof_knn python tools/profile_once.py batch 2 > gpurun_out/prof_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/prof_plain.log
