#!/bin/bash
tag=${1:-r02ab5}
mkdir -p gpurun_out
for v in 8 9; do
GMS_SCORE_V=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=8 -p no:cacheprovider -k "shared or sorted or full_size or k3 or k4 or golden_slam or determinism or fast" > gpurun_out/${tag}_tests_v$v.log 2>&1
echo "tests(V=$v) rc=$?"; tail -2 gpurun_out/${tag}_tests_v$v.log
done
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra $ARGS > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; }
ARGS="--workload K4"
b k4_v7 GMS_SCORE_V=7
b k4_v8 GMS_SCORE_V=8
b k4_v9 GMS_SCORE_V=9
b k4_v8_g1 GMS_SCORE_V=8 GMS_SCORE_G=1
b k4_v8_g4 GMS_SCORE_V=8 GMS_SCORE_G=4
ARGS="--workload K4g"
b k4g_v8 GMS_SCORE_V=8
ARGS="--workload K3"
b k3_v8 GMS_SCORE_V=8
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_k*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step", round(d["ms_per_step"],4), "b2b", round(d.get("back_to_back",{}).get("ms_per_step",0),4), "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),4), {k:round(v,4) for k,v in d.get("phases_ms_per_step",{}).items()})
    except Exception as e:
        print(f, "unparsed", e, open(f.replace('.json','.err')).read()[-300:])
PY
