"""ncu target: a few distillation steps of the products student (MLP3w8, bs 4096).
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/prof_student.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from glnn_b200 import mlp_engine
from glnn_b200.models import Model
dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
hidden = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
f = int(sys.argv[4]) if len(sys.argv) > 4 else 100
c = int(sys.argv[5]) if len(sys.argv) > 5 else 47
torch.manual_seed(0)
n = bs * 8
model = Model(dict(model_name="MLP", num_layers=3, feat_dim=f, hidden_dim=hidden, label_dim=c,
                   dropout_ratio=0.2, norm_type="batch", device=dev)).train()
opt = torch.optim.Adam(model.parameters(), lr=0.01)
x = torch.randn(n, f, device=dev)
t = torch.log_softmax(torch.randn(n, c, device=dev), 1)
idx = torch.randperm(n)[: steps * bs].view(steps, bs).to(dev)
mlp_engine.train_pass(model.encoder, opt, x, t, idx, 1.0)
torch.cuda.synchronize()
print("done")
if os.environ.get("GLNN_TIME"):
    steps_t = 40
    idx = torch.randperm(n)[: bs * 8].repeat(5).view(steps_t, bs).to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        torch.cuda.synchronize(); e0.record()
        mlp_engine.train_pass(model.encoder, opt, x, t, idx, 1.0)
        e1.record(); torch.cuda.synchronize()
    print(f"student step {e0.elapsed_time(e1) / steps_t * 1000:.1f} us (hidden {hidden}, bs {bs})")
