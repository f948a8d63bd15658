#!/bin/bash
# second-stage sweep of the tile kernel: dense-class threshold x cell size (run under gpurun)
mkdir -p gpurun_out
B="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-icp"
for s in ${SCALES:-0.45 0.5 0.55}; do for d in ${DENSE:-160 256 384}; do
  PCR_TILE_DENSE_N=$d PCR_OCC_SCALE=$s timeout 300 $B > gpurun_out/tile2_${s}_$d.json 2> gpurun_out/tile2_${s}_$d.err || echo "rc=$? $s $d"
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/tile2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'ms/step %.3f e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {k:round(v,3) for k,v in r['stage_ms_per_step'].items()}, 'same', d['device_and_e2e_results_identical'], 'kept', d['kept_points'], 'batch dev ms', round(d['batch8m']['device']['ms'],2))
    except Exception as e:
        print(f, 'ERR', e)
PY
