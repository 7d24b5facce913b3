#!/bin/bash
# One gpurun call (1 GPU): parity tests, then A/B timings of this round's switches.
tag=${1:-r02ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/${tag}_tests.log
b() { # name, env..., -- args
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra $ARGS > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
}
ARGS="--workload K2pp"
b k2pp_bulk0 GMS_COPY_BULK=0
b k2pp_bulk_c4 GMS_COPY_CHUNKS=4
b k2pp_bulk_c8 GMS_COPY_CHUNKS=8
b k2pp_bulk_c16 GMS_COPY_CHUNKS=16
b k2pp_bulk_c32 GMS_COPY_CHUNKS=32
ARGS="--workload K4"
b k4_dyn0 GMS_SCORE_DYNAMIC=0
b k4_dyn1 GMS_SCORE_DYNAMIC=1
b k4_dyn1_g1 GMS_SCORE_DYNAMIC=1 GMS_SCORE_G=1
b k4_dyn1_g4 GMS_SCORE_DYNAMIC=1 GMS_SCORE_G=4
ARGS="--workload K3"
b k3 GMS_SCORE_DYNAMIC=1
ARGS="--workload K2"
b k2 GMS_SCORE_DYNAMIC=1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_k*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step", round(d["ms_per_step"],4), "b2b", round(d.get("back_to_back",{}).get("ms_per_step",0),4), "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),4), {k:round(v,4) for k,v in d.get("phases_ms_per_step",{}).items()})
    except Exception as e:
        print(f, "unparsed", e, open(f.replace('.json','.err')).read()[-300:])
PY
