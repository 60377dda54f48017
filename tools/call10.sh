set -x
mkdir -p gpurun_out
# fresh ncu capture of the dominant kernel (q24 768 B rows, d = 256): source of roofline.traffic
ncu --set full --clock-control none --import-source on -k regex:spmm_csr_kernel -s 1 -c 1 -o gpurun_out/r2_spmm_q24_d256 python profiles/prof_spmm.py 256 > gpurun_out/ncu_spmm.log 2>&1
ncu -i gpurun_out/r2_spmm_q24_d256.ncu-rep --page raw --csv > gpurun_out/r2_spmm_q24_d256.raw.csv 2>/dev/null
# launch lists: arxiv students (warm), products student
for cfg in "256 512 128 40" "1024 512 128 40" "2048 4096 100 47"; do
  set -- $cfg
  ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 120 --csv --log-file gpurun_out/r2_launches_student_h$1_bs$2.csv python tools/prof_student.py 4 $1 $2 $3 $4 > /dev/null 2>&1
done
# what one rank of the 8-way data-parallel products student computes locally: 512 rows, H = 2048
GLNN_TIME=1 python tools/prof_student.py 4 2048 512 100 47 | tail -1
GLNN_TIME=1 python tools/prof_student.py 4 2048 1024 100 47 | tail -1
GLNN_TIME=1 python tools/prof_student.py 4 2048 2048 100 47 | tail -1
GLNN_TIME=1 python tools/prof_student.py 4 2048 4096 100 47 | tail -1
GLNN_TIME=1 python tools/prof_student.py 4 256 512 128 40 | tail -1
GLNN_TIME=1 python tools/prof_student.py 4 1024 512 128 40 | tail -1
ls -la gpurun_out | tail -8
