#!/bin/bash
# One gpurun call (1 GPU): everything the round's single-GPU record needs — full GPU test suite, every workload's
# bench line, the reference arm, ncu launch lists + --set full captures of the dominant kernels, sanitizer logs.
tag=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/${tag}_tests.log
timeout 400 python bench.py --steps 50 --warmup 10 > gpurun_out/${tag}_bench_k4.json 2> gpurun_out/${tag}_bench_k4.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
for w in K1 K2 K2pp K3 K4g; do
  timeout 300 python bench.py --steps 30 --warmup 5 --workload $w > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value %.4g"%d["value"], "ms/step", round(d["ms_per_step"],4), "b2b", round(d.get("back_to_back",{}).get("ms_per_step",0),4), "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),4), {k:round(v,4) for k,v in d.get("phases_ms_per_step",{}).items()}, (d.get("no_resample") or {}).get("ms_per_step"), "roofline", round((d.get("roofline") or {}).get("frac",0),3), "cpu", (d.get("cpu_baseline") or {}).get("ms_per_step_sample"))
    except Exception as e:
        print(f, "unparsed", e, open(f.replace('.json','.err')).read()[-300:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${tag}_launches_k4.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/${tag}_ncu_k4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 70 --csv --log-file gpurun_out/${tag}_launches_k2pp.csv python bench.py --steps 6 --warmup 3 --no-cpu --workload K2pp > gpurun_out/${tag}_ncu_k2pp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score_sorted -s 6 -c 1 -o gpurun_out/${tag}_score_k4 -f python bench.py --steps 4 --warmup 3 --no-cpu --no-extra > gpurun_out/${tag}_ncu_score.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_score_pp|k_copy_maps_bulk" -s 8 -c 2 -o gpurun_out/${tag}_k2pp -f python bench.py --steps 3 --warmup 3 --no-cpu --workload K2pp > gpurun_out/${tag}_ncu_k2pp_full.log 2>&1
GMS_DEFER_INTEGRATION=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_map_update_red -s 4 -c 1 -o gpurun_out/${tag}_k2pp_upd -f python bench.py --steps 3 --warmup 3 --no-cpu --workload K2pp > gpurun_out/${tag}_ncu_k2pp_upd.log 2>&1
echo "ncu done"
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/${tag}_memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "golden or appendix or degenerate or out_of_bounds or strongest or hook_matches or operator_sequences or determinism_and_profile" > gpurun_out/${tag}_memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/${tag}_memcheck.log; tail -2 gpurun_out/${tag}_memcheck_pytest.log
timeout 300 compute-sanitizer --tool racecheck --log-file gpurun_out/${tag}_racecheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_racecheck_smoke.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/${tag}_racecheck.log
