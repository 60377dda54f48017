#!/bin/bash
# experiments of round 2: sparse rows (s24) and the TMA gather A/B
set -x
mkdir -p gpurun_out
timeout 300 python tools/exp_spmm_s24.py > gpurun_out/exp_s24.jsonl 2> gpurun_out/exp_s24.err
timeout 300 python tools/exp_spmm_tma.py > gpurun_out/exp_tma.jsonl 2> gpurun_out/exp_tma.err
GLNN_S24=1 timeout 300 python bench.py --light --steps 5 > gpurun_out/bench_s24_light.json 2> gpurun_out/bench_s24_light.err
cat gpurun_out/exp_s24.jsonl gpurun_out/exp_tma.jsonl gpurun_out/bench_s24_light.json
