#!/bin/bash
# A/B of an environment hook on the ICP / config3 blocks of the bench (run under gpurun): tools/ab_icp.sh "PCR_NO_STAGED_UPLOAD=1"
mkdir -p gpurun_out
HOOK="$1"
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
for r in 1 2; do
  timeout 400 $B > gpurun_out/abi_base_$r.json 2> gpurun_out/abi_base_$r.err || echo "rc=$? base"
  env $HOOK timeout 400 $B > gpurun_out/abi_hook_$r.json 2> gpurun_out/abi_hook_$r.err || echo "rc=$? hook"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/abi_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        i=d['icp_sharded']; c=d.get('config3',{})
        print(f, 'icp e2e/iter %.3f loop %.3f setup %.2f' % (i['ms_per_iter_e2e'], i['ms_per_iter_loop_rank0'], i['ms_setup_and_host_rank0']), 'equals_unsharded', i['equals_unsharded'],
              {k:(round(v,3) if isinstance(v,float) else v) for k,v in c.items() if 'ms' in k}, 'step %.3f' % d['ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
