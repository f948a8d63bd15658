#!/bin/bash
# A/B of an environment hook (run under gpurun): tools/ab_env.sh "PCR_NO_EARLY_NORMALS=1" [rounds]
# The GPU tests run first with the default settings; then bench.py alternates default / hook.
mkdir -p gpurun_out
HOOK="$1"; R=${2:-2}
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-icp"
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
for r in $(seq 1 $R); do
  timeout 300 $B > gpurun_out/ab_base_$r.json 2> gpurun_out/ab_base_$r.err || echo "rc=$? base"
  env $HOOK timeout 300 $B > gpurun_out/ab_hook_$r.json 2> gpurun_out/ab_hook_$r.err || echo "rc=$? hook"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'ms/step %.3f e2e %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {k:round(v,3) for k,v in r['stage_ms_per_step'].items()}, 'same', d['device_and_e2e_results_identical'], 'kept', d['kept_points'], 'batch dev ms', round(d['batch8m']['device']['ms'],2), 'api', round(d['batch8m']['host_api']['ms'],2) if 'host_api' in d['batch8m'] else '')
    except Exception as e:
        print(f, 'ERR', e)
PY
