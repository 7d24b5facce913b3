#!/bin/bash
# One gpurun call (1 GPU): parity tests under the alternative kernel variants, then timing sweeps.
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=6 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests(default) rc=$?"; tail -15 gpurun_out/${tag}_tests.log
GMS_SCORE_V=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=6 -p no:cacheprovider -k "shared or full_size or golden or replay or sorted or determinism" > gpurun_out/${tag}_tests_v2.log 2>&1
echo "tests(V=2) rc=$?"; tail -5 gpurun_out/${tag}_tests_v2.log
GMS_MAP_WIN_WORDS=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=6 -p no:cacheprovider -k "per_particle or pp or golden or replay or k1 or determinism or hook" > gpurun_out/${tag}_tests_atomic.log 2>&1
echo "tests(atomic map update) rc=$?"; tail -5 gpurun_out/${tag}_tests_atomic.log
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" > /dev/null 2>&1
}
for v in 0 1 2; do for g in 1 2 4 8; do
  GMS_SCORE_V=$v GMS_SCORE_G=$g timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/${tag}_sweep_v${v}_g${g}.json 2>/dev/null
done; done
for w in 0 13000 26000 52000; do
  GMS_MAP_WIN_WORDS=$w timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --workload K2pp > gpurun_out/${tag}_k2pp_win${w}.json 2>/dev/null
done
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_k4.json 2> gpurun_out/${tag}_bench_k4.err
for w in K2 K3 K4g; do
  timeout 300 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step", round(d["ms_per_step"],4), "b2b", round(d.get("back_to_back",{}).get("ms_per_step",0),4), "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),4))
        print("   phases", {k:round(v,4) for k,v in d.get("phases_ms_per_step",{}).items()})
    except Exception as e:
        print(f, "unparsed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/${tag}_launches_k4.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/${tag}_ncu_k4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/${tag}_launches_k2pp.csv python bench.py --steps 6 --warmup 3 --no-cpu --workload K2pp > gpurun_out/${tag}_ncu_k2pp.log 2>&1
