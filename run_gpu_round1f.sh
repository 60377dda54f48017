#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli_and_data.py tests/test_gpu_dist.py -m gpu -q > gpurun_out/pytest_dist.log 2>&1; echo "dist pytest exit $?" >> gpurun_out/pytest_dist.log
tail -15 gpurun_out/pytest_dist.log
for N in 2 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]);print($N, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
tail -3 gpurun_out/bench_n$N.err
done
