#!/bin/bash
# A/B of the level-0 KNN implementations on the bench frame (run under gpurun; writes gpurun_out/ab_*.json)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ab_tests.log
B="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-icp"
# (insert baseline: 0.684 ms/step, knn 0.203, deferred 0.101 -- gpurun_out/ab_insert.json)
for s in 0.25 0.35 0.45; do
  PCR_OCC_SCALE=$s $B > gpurun_out/ab_select_$s.json 2> gpurun_out/ab_select_$s.err; echo "select $s rc=$?"
done
for s in 0.35 0.45 0.6; do PCR_FIRST_SHELLS=1 PCR_OCC_SCALE=$s $B > gpurun_out/ab_select_fs1_$s.json 2> gpurun_out/ab_select_fs1_$s.err; echo "fs1 $s rc=$?"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'ms/step %.3f e2e %.3f knn %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], r['avg_launch_ms']), {k:round(v,3) for k,v in r['stage_ms_per_step'].items()}, 'same', d['device_and_e2e_results_identical'], 'kept', d['kept_points'])
    except Exception as e:
        print(f, 'ERR', e)
PY
