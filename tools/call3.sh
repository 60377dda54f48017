set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_teacher_train.py -m gpu -q -rA -p no:cacheprovider -k "dist or act_block or adam" > gpurun_out/t_dist2.log 2>&1
tail -12 gpurun_out/t_dist2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
GLNN_DIST_REPLICATE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-parity > gpurun_out/bench_n2_norep.json 2> gpurun_out/bench_n2_norep.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_n2.json", "gpurun_out/bench_n2_norep.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("parity"), d["student"].get("ms_per_step"), d["student"].get("eval_sharded"))
        for s in d["shards"]:
            print("  ", s["rank"], s["phases_ms"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/bench_n2.err
