set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider -k sharded_forward 2>&1 | tail -5
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 --light > gpurun_out/n2_$name.json 2> gpurun_out/n2_$name.err
}
run twopass GLNN_DIST_TWO_PASS=1
run onepass GLNN_DIST_TWO_PASS=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n2_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3))
        for s in d["shards"][:2]:
            print("  ", s["rank"], s["phases_ms"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-1500:])
PY
