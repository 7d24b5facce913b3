#!/bin/bash
# quick 1-GPU call: one test file + score variant timings
tag=${1:-r02q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=6 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/${tag}_tests.log
for v in 2 5 6; do for g in 2 4; do
  GMS_SCORE_V=$v GMS_SCORE_G=$g timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/${tag}_sweep_v${v}_g${g}.json 2>/dev/null
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step", round(d["ms_per_step"],4), "b2b", round(d.get("back_to_back",{}).get("ms_per_step",0),4), "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),4), "score", round(d["phases_ms_per_step"]["score"],4))
    except Exception as e:
        print(f, "unparsed", e)
PY
