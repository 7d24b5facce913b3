#!/bin/bash
# One gpurun call (1 GPU): tests, variant timings, ncu --set full of the scoring kernel, sanitizer logs.
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=6 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests(default) rc=$?"; tail -12 gpurun_out/${tag}_tests.log
GMS_MAP_WIN_WORDS=13000 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=6 -p no:cacheprovider -k "per_particle or pp or golden or replay or k1 or determinism or hook or strongest" > gpurun_out/${tag}_tests_win.log 2>&1
echo "tests(windowed map update) rc=$?"; tail -8 gpurun_out/${tag}_tests_win.log
for v in 0 2 3 4; do
  GMS_SCORE_V=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/${tag}_sweep_v${v}.json 2>/dev/null
done
for w in 0 3200 6500 13000; do
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
for v in 0 2; do
GMS_SCORE_V=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score_sorted -s 6 -c 1 -o gpurun_out/${tag}_score_v$v -f python bench.py --steps 4 --warmup 3 --no-cpu --no-extra > gpurun_out/${tag}_ncu_score_v$v.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score_sorted -s 6 -c 1 -o gpurun_out/${tag}_score_k4g -f python bench.py --steps 4 --warmup 3 --no-cpu --workload K4g > gpurun_out/${tag}_ncu_score_k4g.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/${tag}_launches_k4.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/${tag}_ncu_k4.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --log-file gpurun_out/${tag}_memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "golden or appendix or degenerate or out_of_bounds or strongest or hook_matches" > gpurun_out/${tag}_memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/${tag}_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --log-file gpurun_out/${tag}_racecheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_racecheck_smoke.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/${tag}_racecheck.log
