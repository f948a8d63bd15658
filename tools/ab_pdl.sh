#!/bin/bash
# A/B of the programmatic-launch variants (run under gpurun): default library vs tools/_ab/libpcr_early.so
mkdir -p gpurun_out
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-icp"
PCR_LIB_OVERRIDE=$PWD/tools/_ab/libpcr_early.so timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for r in 1 2; do
  timeout 300 $B > gpurun_out/ab_base_$r.json 2> gpurun_out/ab_base_$r.err || echo "rc=$? base"
  PCR_LIB_OVERRIDE=$PWD/tools/_ab/libpcr_early.so timeout 300 $B > gpurun_out/ab_early_$r.json 2> gpurun_out/ab_early_$r.err || echo "rc=$? early"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'ms/step %.3f e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {k:round(v,3) for k,v in r['stage_ms_per_step'].items()}, 'same', d['device_and_e2e_results_identical'], 'kept', d['kept_points'], 'batch dev ms', round(d['batch8m']['device']['ms'],2))
    except Exception as e:
        print(f, 'ERR', e)
PY
