set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_tall -s 1 -c 1 -o gpurun_out/r2_gemm_tall python profiles/prof_gemm_tall.py > gpurun_out/ncu_gemm_tall.log 2>&1
ncu -i gpurun_out/r2_gemm_tall.ncu-rep --page raw --csv > gpurun_out/r2_gemm_tall.raw.csv 2>/dev/null
ncu -i gpurun_out/r2_gemm_tall.ncu-rep --page source --csv > gpurun_out/r2_gemm_tall.source.csv 2>/dev/null
ls -la gpurun_out/r2_gemm_tall*
