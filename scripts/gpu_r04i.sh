#!/bin/bash
# One gpurun call (1 GPU): k_score_sorted with CTA-chunked dynamic item drawing (GMS_SCORE_DYNAMIC=1) vs the static grid.
tag=${1:-r04i}
mkdir -p gpurun_out
for dyn in 0 1; do
  GMS_SCORE_DYNAMIC=$dyn timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra > gpurun_out/${tag}_k4_dyn$dyn.json 2> gpurun_out/${tag}_k4_dyn$dyn.err
done
GMS_SCORE_DYNAMIC=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "full_size or determinism or K4 or k4 or sorted" > gpurun_out/${tag}_tests_dyn1.log 2>&1; echo "tests(dyn=1) rc=$?"; tail -n 3 gpurun_out/${tag}_tests_dyn1.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_k4_dyn*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step", round(d["ms_per_step"],4), "b2b", round(d["back_to_back"]["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k:round(v,4) for k,v in d["phases_ms_per_step"].items()})
    except Exception as e:
        print(f, "unparsed", e, open(f.replace('.json','.err')).read()[-300:])
PY
