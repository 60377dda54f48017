#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_csr -s 1 -c 1 -o gpurun_out/r1_spmm_d256 python profiles/prof_spmm.py 256 > gpurun_out/ncu_spmm.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/ncu_spmm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_csr -s 1 -c 1 -o gpurun_out/r1_spmm_d100 python profiles/prof_spmm.py 100 > gpurun_out/ncu_spmm100.log 2>&1; echo "ncu exit $?"
