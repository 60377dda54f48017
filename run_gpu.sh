#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_teacher.py -m gpu -x -q -k "tall or planes or q24 or midsize or golden or host or tcgen05" > gpurun_out/pytest_teacher.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_teacher.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'])
for k in d.get('kernels',[]): print(k)
print(d.get('student'))
PY
