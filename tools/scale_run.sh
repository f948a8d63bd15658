#!/bin/bash
# bench at N GPUs under torchrun (run under gpurun --gpus N): tools/scale_run.sh N [tag]
N=$1; T=${2:-r02}
mkdir -p gpurun_out
if [ "$N" -gt 1 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py > gpurun_out/${T}_multi_gpu_check_${N}xB200.txt 2>&1; echo "check rc=$?"; tail -3 gpurun_out/${T}_multi_gpu_check_${N}xB200.txt
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench rc=$?"
else
  timeout 900 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
fi
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_n$N.json").read().strip().splitlines()[-1])
i=d["icp_sharded"]; b=d["batch8m"]
print("N", d["n_gpus"], "value %.1f M ms/step %.3f e2e %.1f M" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6))
print("batch dev %.2f ms api %.2f ms same %s" % (b["device"]["ms"], b["host_api"]["ms"], b["host_and_device_results_identical"]))
print("icp e2e/iter %.3f loop %.3f (kernels %.3f + reduce/solve %.3f) setup %.2f %s identical %s equals_unsharded %s" % (i["ms_per_iter_e2e"], i["ms_per_iter_loop_rank0"], i["ms_per_iter_step_kernels_rank0"], i["ms_per_iter_reduce_allreduce_solve_rank0"], i["ms_setup_and_host_rank0"], i["collective"][:20], i["identical_on_all_ranks"], i["equals_unsharded"]))
print("config3", json.dumps(d.get("config3"))[:700])
PY
