#!/bin/bash
# What a round-end check runs on a B200 box (from the repo root, library already built in-tree):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- ./run_gpu.sh
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# launch lists (cold cache, serialised): teacher forward, student steps
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/launches_teacher_products.csv python bench.py --steps 2 --warmup 3 --light > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv \
    --log-file gpurun_out/launches_student_products_warm.csv python tools/prof_student.py 3 > /dev/null 2>&1
tail -c 400 gpurun_out/bench_n1.json
