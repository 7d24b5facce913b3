#!/bin/bash
# One gpurun call (1 GPU), last of the round: the driver's own bench invocations (default workload with cpu_baseline and
# extras; the reference arm), then compute-sanitizer over the kernels this round's last changes touched.
tag=${1:-r04g}
mkdir -p gpurun_out
timeout 200 python bench.py --steps 50 --warmup 10 > gpurun_out/${tag}_bench_k4.json 2> gpurun_out/${tag}_bench_k4.err; echo "bench rc=$?"
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "reference arm rc=$?"
python - <<PY
import json
for n in ("k4","ref"):
    try:
        d=json.loads(open("gpurun_out/${tag}_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value %.4g"%d["value"], "ms/step", round(d["ms_per_step"],4), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(n, "unparsed", e)
PY
timeout 110 compute-sanitizer --tool racecheck --log-file gpurun_out/${tag}_racecheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_racecheck_smoke.log 2>&1
echo "racecheck rc=$?"; tail -n 2 gpurun_out/${tag}_racecheck.log
timeout 170 compute-sanitizer --tool memcheck --log-file gpurun_out/${tag}_memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_trig.py tests/test_next_rows.py -m gpu -q -p no:cacheprovider -k "golden or appendix or trig or deskew or next_rows or strongest" > gpurun_out/${tag}_memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -n 2 gpurun_out/${tag}_memcheck.log; tail -n 2 gpurun_out/${tag}_memcheck_pytest.log
