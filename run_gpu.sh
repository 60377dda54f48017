#!/bin/bash
# What a round-end check runs on a B200 box (from the repo root, library already built in-tree):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- ./run_gpu.sh
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc >> gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/t_all.log 2>&1
tail -5 gpurun_out/t_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 700 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# launch list of the teacher forward (cold cache, serialised: compare SHARES with the bench breakdown)
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/launches_teacher_products.csv python bench.py --steps 2 --warmup 3 --light > /dev/null 2>&1
tail -c 600 gpurun_out/bench_n1.json
