#!/bin/bash
# One gpurun call (1 GPU): the full GPU suite and the K4 line on the final code of the round.
tag=${1:-r04j}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -n 4 gpurun_out/${tag}_tests.log
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-extra > gpurun_out/${tag}_k4.json 2> gpurun_out/${tag}_k4.err
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_k4.json').read().strip().splitlines()[-1])
print('K4 ms/step', round(d['ms_per_step'],4), 'b2b', round(d['back_to_back']['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms_per_step'].items()})
"
