set -x
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 10 --warmup 3 --light > gpurun_out/n8b_$name.json 2> gpurun_out/n8b_$name.err
}
run tp_all GLNN_DIST_TWO_PASS=1
run tp_l1only GLNN_DIST_TWO_PASS=1 GLNN_DIST_TWO_PASS_Z=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n8b_*.json")) + ["gpurun_out/r2_bench_n8.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), d.get("e2e",{}).get("ms_per_step"), d.get("parity"), (d.get("student") or {}).get("ms_per_step"), (d.get("student") or {}).get("eval_sharded"))
        for s in d["shards"][:8:4]:
            print("  ", s["rank"], s["phases_ms"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-1500:])
PY
