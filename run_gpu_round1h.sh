#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_teacher.py -m gpu -q -k "planes or tcgen05" > gpurun_out/pytest_tc.log 2>&1; echo "tc pytest exit $?" >> gpurun_out/pytest_tc.log
tail -25 gpurun_out/pytest_tc.log
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from glnn_b200 import ops
dev = torch.device('cuda:0')
def t(fn, it=20):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / it
x = torch.randn(4096, 2048, device=dev); w = torch.randn(2048, 2048, device=dev); dy = torch.randn(4096, 2048, device=dev)
px, pw, pdy = ops.split_planes(x), ops.split_planes(w), ops.split_planes(dy)
for name, f32, pl in [("fwd", lambda: ops.gemm(x, w, trans_b=True, impl=2), lambda: ops.gemm_planes(px, pw, trans_b=True)),
                      ("dX", lambda: ops.gemm(dy, w, impl=2), lambda: ops.gemm_planes(pdy, pw)),
                      ("dW", lambda: ops.gemm(dy, x, trans_a=True, impl=2), lambda: ops.gemm_planes(pdy, px, trans_a=True))]:
    a, b = t(f32), t(pl)
    print(f"student {name}: fp32-operand {a*1e3:.0f} us ({34.36/a:.0f} TF/s eq)  planes {b*1e3:.0f} us ({34.36/b:.0f} TF/s eq)")
a = torch.randn(2449029, 256, device=dev); w2 = torch.randn(256, 256, device=dev); out = torch.empty(2449029, 256, device=dev)
pa, pw2 = ops.split_planes(a), ops.split_planes(w2)
print("teacher 256->256: fp32-operand %.2f ms, planes %.2f ms, planes->planes %.2f ms, split %.2f ms" % (
    t(lambda: ops.gemm(a, w2, trans_b=True, out=out, impl=2), 5), t(lambda: ops.gemm_planes(pa, pw2, trans_b=True, out=out), 5),
    t(lambda: ops.gemm_planes(pa, pw2, trans_b=True, out_planes=True), 5), t(lambda: ops.split_planes(a), 5)))
PY
