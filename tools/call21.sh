set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_teacher.py -m gpu -q -p no:cacheprovider -k "gemm_tall or teacher_matches or midsize or full_size_products_forward or q24_projection" 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --no-parity > gpurun_out/bench_ew16.json 2> gpurun_out/bench_ew16.err
GLNN_TALL_EW=8 timeout 300 python bench.py --steps 5 --no-parity > gpurun_out/bench_ew8.json 2> gpurun_out/bench_ew8.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_ew16.json","gpurun_out/bench_ew8.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"])
        for k in d["kernels"]:
            if "gemm" in k["name"]: print("   ", k["name"], k["ms"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
