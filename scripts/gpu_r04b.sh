#!/bin/bash
# gpurun --gpus N: pull (default) vs push (GMS_PULL=0) exchange of the log-weights.  usage: scripts/gpu_r04b.sh <tag> <N>
tag=${1:-r04b}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_nccl.py -m gpu -q --maxfail=4 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "multi-GPU tests rc=$?"; tail -n 8 gpurun_out/${tag}_tests.log
for pull in 1 0; do
  GMS_PULL=$pull timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$pull bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/${tag}_bench_n${N}_pull$pull.json 2> gpurun_out/${tag}_bench_n${N}_pull$pull.err
  echo "bench N=$N pull=$pull rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value %.4g"%d["value"], "ms/step", round(d["ms_per_step"],4), "b2b", round(d.get("back_to_back",{}).get("ms_per_step",0),4), "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),4))
        print("   phases", {k:round(v,4) for k,v in d.get("phases_ms_per_step",{}).items()})
        pc=d.get("parity_check")
        if pc: print("   parity", pc["ranks"], pc["shared"], pc["per_particle"], pc["shared_detail"]["max_weight_rel_diff_rank0"], pc["shared_detail"]["problems_rank0"], pc["per_particle_detail"]["problems_rank0"])
    except Exception as e:
        print(f, "unparsed", e, open(f.replace('.json','.err')).read()[-400:])
PY
