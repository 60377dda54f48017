#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_teacher.py -m gpu -q -k "tcgen05" > gpurun_out/pytest_tc.log 2>&1; echo "tc pytest exit $?" >> gpurun_out/pytest_tc.log
tail -25 gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
