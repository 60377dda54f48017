set -x
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 --light > gpurun_out/n4_$name.json 2> gpurun_out/n4_$name.err
}
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider -k sharded_forward 2>&1 | tail -3
run sm_c4 GLNN_PUSH_ENGINE=sm GLNN_DIST_CHUNKS=4
run sm_c8 GLNN_PUSH_ENGINE=sm GLNN_DIST_CHUNKS=8
run sm_c8_64 GLNN_PUSH_ENGINE=sm GLNN_DIST_CHUNKS=8 GLNN_PUSH_CTAS=64
run ce_c4 GLNN_PUSH_ENGINE=ce GLNN_DIST_CHUNKS=4
run sm_c4_rep GLNN_PUSH_ENGINE=sm GLNN_DIST_CHUNKS=4 GLNN_DIST_REPLICATE=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n4_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3))
        for s in d["shards"][:2]:
            print("  ", s["rank"], s["phases_ms"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-800:])
PY
