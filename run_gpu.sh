#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest exit $?"
grep -v Warning gpurun_out/pytest_dist.log | tail -40
