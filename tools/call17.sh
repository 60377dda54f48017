set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --cache-control none --import-source on -k regex:gemm_bf16x3 -s 2 -c 1 -o gpurun_out/r2_gemm_shortk python profiles/prof_gemm_shortk.py > gpurun_out/ncu_gemm_shortk.log 2>&1
ncu -i gpurun_out/r2_gemm_shortk.ncu-rep --page raw --csv > gpurun_out/r2_gemm_shortk.raw.csv 2>/dev/null
ncu -i gpurun_out/r2_gemm_shortk.ncu-rep --page source --csv > gpurun_out/r2_gemm_shortk.source.csv 2>/dev/null
ls -la gpurun_out/r2_gemm_shortk*
