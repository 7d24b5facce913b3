#!/bin/bash
# One gpurun call (1 GPU): parity tests (default, and with the exact-sum normalise), K4 / K2 / K3 / K2pp bench lines.
tag=${1:-r02q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/${tag}_tests.log
GMS_SHARDED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/${tag}_tests_exact.log 2>&1
echo "tests(exact sums) rc=$?"; tail -4 gpurun_out/${tag}_tests_exact.log
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
