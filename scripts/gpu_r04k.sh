#!/bin/bash
# One gpurun call (1 GPU): k_resample_small (CDF + selection + parent list in one launch, default) vs the separate
# kernels (GMS_SMALL_FUSED=0) on the small workloads; full GPU suite with the default.
tag=${1:-r04k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -n 4 gpurun_out/${tag}_tests.log
for f in 1 0; do for w in K2 K2pp K1; do
  GMS_SMALL_FUSED=$f timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra --workload $w > gpurun_out/${tag}_${w}_fused$f.json 2> gpurun_out/${tag}_${w}_fused$f.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_K*_fused*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step", round(d["ms_per_step"],4), "b2b", round(d["back_to_back"]["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), "launches", d.get("gpu_launches"), {k:round(v,4) for k,v in d["phases_ms_per_step"].items()})
    except Exception as e:
        print(f, "unparsed", e, open(f.replace('.json','.err')).read()[-300:])
PY
