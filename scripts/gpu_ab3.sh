#!/bin/bash
tag=${1:-r02ab3}
mkdir -p gpurun_out
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra $ARGS > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; }
ARGS="--workload K3"
for g in 2 4 8 16 32; do b k3_g$g GMS_SCORE_G=$g; done
ARGS="--workload K2"
for g in 8 16 32; do b k2_g$g GMS_SCORE_G=$g; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_k*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step", round(d["ms_per_step"],4), "b2b", round(d.get("back_to_back",{}).get("ms_per_step",0),4), "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),4), {k:round(v,4) for k,v in d.get("phases_ms_per_step",{}).items()})
    except Exception as e:
        print(f, "unparsed", e, open(f.replace('.json','.err')).read()[-300:])
PY
