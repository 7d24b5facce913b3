#!/bin/bash
# One gpurun call: GPU parity tests, the default bench line, per-workload bench lines, the ncu launch list.
# usage: scripts/gpu_check.sh <tag> [quick]
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=6 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
tail -40 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_k4.json 2> gpurun_out/${tag}_bench_k4.err
echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench_k4.err
for w in K2 K3 K2pp; do
  timeout 300 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  echo "bench $w rc=$?"; tail -2 gpurun_out/${tag}_bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step", round(d["ms_per_step"],4), "b2b", d.get("back_to_back",{}).get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "launches", d.get("gpu_launches"))
        print("   phases", {k:round(v,4) for k,v in d.get("phases_ms_per_step",{}).items()})
    except Exception as e:
        print(f, "unparsed", e)
PY
if [ "$2" != "quick" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/${tag}_launches_k4.csv python bench.py --steps 6 --warmup 3 --no-cpu --no-extra > gpurun_out/${tag}_ncu_k4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 120 --csv --log-file gpurun_out/${tag}_launches_k2pp.csv python bench.py --steps 6 --warmup 3 --no-cpu --workload K2pp > gpurun_out/${tag}_ncu_k2pp.log 2>&1
fi
