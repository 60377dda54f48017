set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 2 -c 2 -o gpurun_out/r2_gemm_small python profiles/prof_gemm_small.py > gpurun_out/ncu_gemm_small.log 2>&1
ncu -i gpurun_out/r2_gemm_small.ncu-rep --page raw --csv > gpurun_out/r2_gemm_small.raw.csv 2>/dev/null
ncu -i gpurun_out/r2_gemm_small.ncu-rep --page source --csv > gpurun_out/r2_gemm_small.source.csv 2>/dev/null
timeout 300 python -m pytest tests/test_gpu_runners.py tests/test_gpu_student.py -m gpu -q -p no:cacheprovider -k "runner or run_transductive or adam" 2>&1 | tail -3
ls -la gpurun_out/r2_gemm_small*
