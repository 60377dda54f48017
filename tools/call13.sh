set -x
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 --steps 10 --warmup 3 --light > gpurun_out/n8c_$name.json 2> gpurun_out/n8c_$name.err
}
run ctas64 GLNN_PUSH_CTAS=64
run ctas128 GLNN_PUSH_CTAS=128
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n8c_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3))
        for s in d["shards"][:8:4]:
            print("  ", s["rank"], s["phases_ms"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-1500:])
PY
