#!/bin/bash
# Radix-tile sweep, tests under the other tile sizes, and ncu evidence for the device graph builder:
#   /usr/local/graft/bin/gpurun --timeout 200 -- 'bash tools/run_exp_csr.sh'
mkdir -p gpurun_out
: > gpurun_out/exp_csr_tiles.jsonl
for it in 8 16 4; do   # GLNN_CSR_MATCH=1 selects the MATCH-instruction ranking
  GLNN_CSR_ITEMS=$it timeout 40 python tools/exp_csr_build.py --tile-sweep >> gpurun_out/exp_csr_tiles.jsonl 2>> gpurun_out/exp_csr_tiles.err
done
cat gpurun_out/exp_csr_tiles.jsonl
GLNN_CSR_ITEMS=16 timeout 40 python -m pytest tests/test_gpu_csr_build.py -q -x 2>&1 | tail -1
GLNN_CSR_ITEMS=4 timeout 40 python -m pytest tests/test_gpu_csr_build.py -q -x 2>&1 | tail -1
timeout 50 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:coo_prepare|radix|scan_|widen' -c 40 --csv \
    --log-file gpurun_out/launches_csr_build.csv python tools/exp_csr_build.py --once > /dev/null 2>&1
tail -n +1 gpurun_out/launches_csr_build.csv | cut -d, -f5,12- | tail -25
timeout 60 ncu --set full --clock-control none -k 'regex:coo_prepare|radix_hist|radix_scatter' -c 3 \
    -o gpurun_out/csr_build_full -f python tools/exp_csr_build.py --once > /dev/null 2>&1
ls -la gpurun_out/ | tail -5
