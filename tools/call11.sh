set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n2.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity"]["max_rel"], d["parity"]["allclose_rtol1e-4_atol1e-5"], d["student"]["ms_per_step"], d["student"]["eval_sharded"])
for s in d["shards"]: print("  ", s["rank"], s["phases_ms"])
PY
tail -3 gpurun_out/r2_bench_n2.err
