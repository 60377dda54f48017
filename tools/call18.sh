set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/t_all_final.log 2>&1
tail -4 gpurun_out/t_all_final.log
