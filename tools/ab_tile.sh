#!/bin/bash
# A/B of the level-0 tile kernel against the thread-per-query selection kernel (run under gpurun; writes gpurun_out/tile_*.json)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/tile_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/tile_tests.log
B="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-icp"
for s in ${SCALES:-0.25 0.35 0.45}; do
  PCR_OCC_SCALE=$s timeout 300 $B > gpurun_out/tile_on_$s.json 2> gpurun_out/tile_on_$s.err; echo "tile $s rc=$?"
  PCR_OCC_SCALE=$s PCR_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-icp 2>&1 | grep "level 0: [0-9]* normal" | head -2
done
PCR_KNN_TILE=0 timeout 300 $B > gpurun_out/tile_off.json 2> gpurun_out/tile_off.err; echo "off rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/tile_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'ms/step %.3f e2e %.3f knn %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], r['avg_launch_ms']), {k:round(v,3) for k,v in r['stage_ms_per_step'].items()}, 'same', d['device_and_e2e_results_identical'], 'kept', d['kept_points'], 'batch dev ms', round(d['batch8m']['device']['ms'],2))
    except Exception as e:
        print(f, 'ERR', e)
PY
