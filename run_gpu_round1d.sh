#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_teacher.py -m gpu -q -k "tcgen05" > gpurun_out/pytest_tc.log 2>&1; echo "tc pytest exit $?" >> gpurun_out/pytest_tc.log
tail -6 gpurun_out/pytest_tc.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
for r in d['kernels']: print(r)
print(d['student'])
PY
tail -5 gpurun_out/bench.err
