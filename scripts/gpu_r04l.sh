#!/bin/bash
# Last gpurun call of the round (1 GPU): full GPU suite, smoke(), the K4 line — on the library as committed.
tag=${1:-r04l}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -n 3 gpurun_out/${tag}_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/${tag}_smoke.log
timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra > gpurun_out/${tag}_k4.json 2> gpurun_out/${tag}_k4.err
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_k4.json').read().strip().splitlines()[-1])
print('K4 ms/step', round(d['ms_per_step'],4), 'b2b', round(d['back_to_back']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))
"
