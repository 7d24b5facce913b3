#!/bin/bash
tag=${1:-r02ab4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/${tag}_tests.log
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra $ARGS > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; }
ARGS="--workload K2pp"
b k2pp_w4 GMS_PP_WARPS=4
b k2pp_w8 GMS_PP_WARPS=8
ARGS="--workload K1"
b k1_w4 GMS_PP_WARPS=4
b k1_w8 GMS_PP_WARPS=8
ARGS="--workload K4"
b k4 GMS_PP_WARPS=4
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_k*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step", round(d["ms_per_step"],4), "b2b", round(d.get("back_to_back",{}).get("ms_per_step",0),4), "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),4), {k:round(v,4) for k,v in d.get("phases_ms_per_step",{}).items()})
    except Exception as e:
        print(f, "unparsed", e, open(f.replace('.json','.err')).read()[-300:])
PY
