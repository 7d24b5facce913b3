#!/bin/bash
# One gpurun call (1 GPU): full GPU suite (incl. the bit-exact trig probe / raw-sweep counts), bench lines of every
# workload, launch list of the K2pp step (k_cdf_literal before / after: profiles/r02_launches_k2pp.csv vs this one).
tag=${1:-r04a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -n 12 gpurun_out/${tag}_tests.log
for w in K4 K2 K3 K2pp K1; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra --workload $w > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step", round(d["ms_per_step"],4), "b2b", round(d.get("back_to_back",{}).get("ms_per_step",0),4), "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),4), {k:round(v,4) for k,v in d.get("phases_ms_per_step",{}).items()}, (d.get("no_resample") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "unparsed", e, open(f.replace('.json','.err')).read()[-300:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${tag}_launches_k2pp.csv python bench.py --steps 6 --warmup 3 --no-cpu --workload K2pp > gpurun_out/${tag}_ncu_k2pp.log 2>&1
grep -c k_cdf_literal gpurun_out/${tag}_launches_k2pp.csv
grep k_cdf_literal gpurun_out/${tag}_launches_k2pp.csv | tail -n 3
