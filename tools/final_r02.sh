#!/bin/bash
# Round-2 evidence run (under gpurun): GPU tests, the bench line of both arms, launch lists (cold / warm) + ncu capture of
# the dominant kernels.  TAG names the output files (default r02).
T=${TAG:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_ref_n1.json 2> gpurun_out/${T}_bench_ref_n1.err; echo "ref rc=$?"
bash tools/launch_list_frame.sh > gpurun_out/${T}_launches_device_step.txt 2>&1; tail -3 gpurun_out/${T}_launches_device_step.txt
python tools/profile_once.py frame 3 > gpurun_out/prof_plain.log 2>&1 || { tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"knn_tile_kernel|sor_stats_kernel|normals_from_lists" -c 8 -f -o gpurun_out/${T}_final python tools/profile_once.py frame 2 > gpurun_out/prof_ncu.log 2>&1
echo "ncu rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
