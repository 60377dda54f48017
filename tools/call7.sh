set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_teacher.py -m gpu -q -p no:cacheprovider -k "gemm_tall or teacher_matches or midsize or full_size_products_forward or q24_projection" 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --no-parity > gpurun_out/bench_bk32.json 2> gpurun_out/bench_bk32.err
GLNN_TALL_BK=64 timeout 300 python bench.py --steps 5 --no-parity > gpurun_out/bench_bk64.json 2> gpurun_out/bench_bk64.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_bk32.json","gpurun_out/bench_bk64.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"])
        for k in d["kernels"]: print("   ", k["name"], k["ms"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
