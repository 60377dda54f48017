set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n8b.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["parity"]["max_rel"], d["parity"]["allclose_rtol1e-4_atol1e-5"], d["student"]["ms_per_step"], d["student"]["eval_sharded"]["max_abs_diff_vs_unsharded"])
PY
tail -2 gpurun_out/r2_bench_n8b.err
