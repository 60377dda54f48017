set -x
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 --light > gpurun_out/n8_$name.json 2> gpurun_out/n8_$name.err
}
run twopass GLNN_DIST_TWO_PASS=1
run onepass GLNN_DIST_TWO_PASS=0
run onepass_c8 GLNN_DIST_TWO_PASS=0 GLNN_DIST_CHUNKS=8 GLNN_PUSH_CTAS=64
run ce GLNN_DIST_TWO_PASS=0 GLNN_PUSH_ENGINE=ce
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n8_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3))
        for s in d["shards"][:8:3]:
            print("  ", s["rank"], s["phases_ms"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-1500:])
PY
