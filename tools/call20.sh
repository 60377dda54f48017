set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n2_final.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity"]["max_rel"], d["student"]["ms_per_step"], d["config"]["parallelism"][:60])
r = json.loads(open("gpurun_out/bench_ref_n2.json").read().strip().splitlines()[-1])
print(r["value"], r["config"] == d["config"])
PY
