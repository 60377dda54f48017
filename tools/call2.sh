set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/t_all2.log 2>&1
tail -15 gpurun_out/t_all2.log
timeout 300 python tools/exp_spmm_rowstride.py > gpurun_out/exp_rowstride.jsonl 2> gpurun_out/exp_rowstride.err
cat gpurun_out/exp_rowstride.jsonl
