set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
GLNN_DIST_TWO_PASS=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 4 --steps 10 --warmup 3 --light > gpurun_out/n4_onepass_final.json 2> gpurun_out/n4_onepass_final.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n4.json","gpurun_out/n4_onepass_final.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d.get("e2e",{}).get("ms_per_step"), (d.get("parity") or {}).get("max_rel"), (d.get("student") or {}).get("ms_per_step"))
    for s in d["shards"][:2]: print("  ", s["rank"], s["phases_ms"])
PY
