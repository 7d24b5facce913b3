#!/bin/bash
# One gpurun call (1 GPU): per-particle-map path — parity tests, K2pp / K1 bench, launch list, ncu --set full of its kernels.
tag=${1:-r02pp}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/${tag}_tests.log
for w in K2pp K1; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --workload $w > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step", round(d["ms_per_step"],4), "b2b", round(d.get("back_to_back",{}).get("ms_per_step",0),4), "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),4))
        print("   phases", {k:round(v,4) for k,v in d.get("phases_ms_per_step",{}).items()})
    except Exception as e:
        print(f, "unparsed", e)
PY
if [ "$2" != "noprof" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${tag}_launches_k2pp.csv python bench.py --steps 6 --warmup 3 --no-cpu --workload K2pp > gpurun_out/${tag}_ncu_k2pp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_map_update_red|k_score_pp|k_copy_maps" -s 9 -c 3 -o gpurun_out/${tag}_k2pp -f python bench.py --steps 3 --warmup 3 --no-cpu --workload K2pp > gpurun_out/${tag}_ncu_k2pp_full.log 2>&1
echo "ncu rc=$?"
fi
