#!/bin/bash
# One gpurun --gpus N call: multi-GPU parity tests (optional), then bench at n = 2..N with the sharded and the replicated
# normalise / resample (GMS_SHARDED=1/0), short runs.   usage: scripts/gpu_multi2.sh <tag> <N> [notests] [noextra]
tag=${1:-r02m}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${tag}_smi.csv 2>&1
if [ "$3" != "notests" ]; then
timeout 1500 python -m pytest tests/test_multigpu_nccl.py -m gpu -q --maxfail=4 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "multi-GPU tests rc=$?"; tail -25 gpurun_out/${tag}_tests.log
fi
for n in 8 4 2; do
  [ $n -le $N ] || continue
  for sh in 1 0; do
    xtra="--no-extra --no-parity"; [ $sh = 1 ] && [ "$4" != "noextra" ] && xtra=""
    GMS_SHARDED=$sh timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n+sh*10)) bench.py --gpus $n --steps 20 --warmup 5 $xtra > gpurun_out/${tag}_bench_n${n}_s$sh.json 2> gpurun_out/${tag}_bench_n${n}_s$sh.err
    echo "bench N=$n sharded=$sh rc=$?"; tail -3 gpurun_out/${tag}_bench_n${n}_s$sh.err | cut -c1-300
  done
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
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
        print(f, "unparsed", e)
PY
