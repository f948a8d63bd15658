#!/bin/bash
# Round-2 evidence run (under gpurun): GPU tests, the bench line of both arms, launch list + ncu capture of the dominant kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err; echo "ref rc=$?"
python tools/profile_once.py batch 3 > gpurun_out/prof_plain.log 2>&1 || { tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_step.csv python tools/profile_once.py batch 3 > gpurun_out/ll_ncu.log 2>&1
python tools/parse_launches.py gpurun_out/r02_launches_step.csv 3
ncu --set full --clock-control none --import-source on -k regex:"knn_tile_kernel|sor_stats_kernel|normals_from_lists" -c 6 -f -o gpurun_out/r02_final python tools/profile_once.py batch 2 > gpurun_out/prof_ncu.log 2>&1
echo "ncu rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
