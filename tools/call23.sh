set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_n4b.json 2> gpurun_out/r2_bench_n4b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n4b.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["parity"]["max_rel"], d["student"]["ms_per_step"])
PY
tail -2 gpurun_out/r2_bench_n4b.err
