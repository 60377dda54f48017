#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_student.py tests/test_gpu_dist.py -m gpu -q > gpurun_out/pytest_student.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_student.log
tail -6 gpurun_out/pytest_student.log
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]);print($N, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
for s in d['shards'][:8]: print(s['rank'], s['rows'], s['nnz'], [(a[:14],b) for a,b in s['phases_ms']])"
tail -2 gpurun_out/bench_n$N.err
done
