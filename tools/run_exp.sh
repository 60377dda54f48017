#!/bin/bash
# The single-GPU experiments of round 2 (results: profiles/r2_exp_*.jsonl, verdicts: DESIGN.md 4.1):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/run_exp.sh'
# Multi-GPU: torchrun --nproc-per-node N tools/exp_push_rate.py   (all-to-all slab push rate)
set -x
mkdir -p gpurun_out
timeout 300 python tools/exp_spmm_s24.py > gpurun_out/exp_s24.jsonl 2> gpurun_out/exp_s24.err        # sparse rows
timeout 300 python tools/exp_spmm_tma.py > gpurun_out/exp_tma.jsonl 2> gpurun_out/exp_tma.err        # TMA bulk row pull A/B
timeout 300 python tools/exp_spmm_rowstride.py > gpurun_out/exp_rowstride.jsonl 2> gpurun_out/exp_rowstride.err
cat gpurun_out/exp_s24.jsonl gpurun_out/exp_tma.jsonl gpurun_out/exp_rowstride.jsonl
