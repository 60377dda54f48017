#!/bin/bash
for v in p0_b4 p1_b4 p0_b3 p1_b3 p0_b2 p1_b2; do
echo "== $v"
GLNN_B200_LIB=$PWD/gpurun_variants/lib_$v.so timeout 300 python tools/exp_spmm_formats.py 2>&1 | grep -v Warn | tail -3
done
