#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench n1 exit $?"
python -c "
import json;d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]);print(1, d['value'], d['ms_per_step'], d['e2e'])"
tail -3 gpurun_out/bench.err
N=4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]);print($N, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
for s in d['shards']: print(s)"
tail -3 gpurun_out/bench_n$N.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 1 -c 1 -o gpurun_out/r1_gemm_teacher python profiles/prof_gemm.py teacher > gpurun_out/ncu_gemm_t.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 3 -c 3 -o gpurun_out/r1_gemm_student python profiles/prof_gemm.py student > gpurun_out/ncu_gemm_s.log 2>&1; echo "ncu exit $?"
