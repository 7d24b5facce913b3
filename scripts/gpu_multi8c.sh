#!/bin/bash
# One gpurun --gpus 8 call: the driver's scaling invocation at N = 8 (default settings: pull exchange, extras, parity check),
# then the same workload with the push exchange (GMS_PULL=0) for the A/B.
tag=${1:-r04e}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_bench_n8.err; echo "bench N=8 rc=$?"
GMS_PULL=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29509 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu --no-extra --no-parity > gpurun_out/${tag}_bench_n8_push.json 2> gpurun_out/${tag}_bench_n8_push.err; echo "bench N=8 (push) rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value %.4g"%d["value"], "ms/step", round(d["ms_per_step"],4), "b2b", round(d.get("back_to_back",{}).get("ms_per_step",0),4), "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),4))
        print("   phases", {k:round(v,4) for k,v in d.get("phases_ms_per_step",{}).items()})
        pc=d.get("parity_check")
        if pc: print("   parity", pc["ranks"], pc["shared"], pc["per_particle"], pc["shared_detail"]["max_weight_rel_diff_rank0"], pc["shared_detail"]["problems_rank0"], pc["per_particle_detail"]["problems_rank0"])
        for k,v in (d.get("extra") or {}).items():
            print("   extra", k, "ms/step", round(v["ms_per_step"],4), "b2b", round(v["back_to_back"]["ms_per_step"],4), "e2e", round((v.get("e2e") or {}).get("ms_per_step",0),4), {kk:round(vv,4) for kk,vv in v["phases_ms_per_step"].items()}, (v.get("no_resample") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "unparsed", e, open(f.replace('.json','.err')).read()[-600:])
PY
