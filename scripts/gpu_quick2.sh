#!/bin/bash
tag=${1:-r02q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=6 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/${tag}_tests.log
GMS_MAP_WIN_WORDS=13000 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=6 -p no:cacheprovider -k "per_particle or pp or golden or replay or k1 or determinism or hook or strongest" > gpurun_out/${tag}_tests_win.log 2>&1
echo "tests(windowed) rc=$?"; tail -4 gpurun_out/${tag}_tests_win.log
for w in 0 6500 13000 26000; do
  GMS_MAP_WIN_WORDS=$w timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --workload K2pp > gpurun_out/${tag}_k2pp_win${w}.json 2>/dev/null
done
for w in K2 K3 K4; do
  timeout 300 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu --no-extra > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
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
GMS_MAP_WIN_WORDS=13000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_map_update_win -s 3 -c 1 -o gpurun_out/${tag}_mapwin -f python bench.py --steps 3 --warmup 3 --no-cpu --workload K2pp > gpurun_out/${tag}_ncu_mapwin.log 2>&1
