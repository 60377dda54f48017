#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_cli_and_data.py -m gpu -q > gpurun_out/pytest_dist.log 2>&1; echo "dist pytest exit $?" >> gpurun_out/pytest_dist.log
tail -25 gpurun_out/pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
tail -c 1800 gpurun_out/bench_n2.json; tail -8 gpurun_out/bench_n2.err
