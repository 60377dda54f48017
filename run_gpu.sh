#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^E " gpurun_out/pytest_dist.log | head -20
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
tail -5 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
print($N, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
for s in d['shards'][:8]: print(s['rank'], s['rows'], s['nnz'], [(a[:34],b) for a,b in s['phases_ms']])
print(d.get('student'))
PY
