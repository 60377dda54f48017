"""CPU-only: the C-ABI library builds/loads and exports every symbol include/glnn_b200.h declares,
ctypes signatures cover them all, and the host-side logic (graph container, batching, config merge,
model surface) behaves like the reference's.  No kernel is launched here."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "glnn_b200.h")).read()
    return sorted(set(re.findall(r"^GLNN_API [\w\s\*]+?\b(glnn_[a-z0-9_]+)\(", hdr, flags=re.M)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    from glnn_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.glnn_version() == 2
    assert isinstance(lib.glnn_last_error(), bytes)


def test_abi_layout_queries_need_no_gpu():
    import ctypes
    from glnn_b200 import _lib
    lib = _lib.load()
    d = _lib.MlpDesc(num_layers=3, feat_dim=128, hidden_dim=256, label_dim=40, norm=1, dropout=0.2,
                     bn_eps=1e-5, bn_momentum=0.1)
    # P of SURVEY.md K10: 110,120 for the arxiv MLP (weights+biases 109,096 + 2 x 2 x 256 BN)
    assert lib.glnn_mlp_param_count(ctypes.byref(d)) == 110120
    assert lib.glnn_mlp_bn_stat_count(ctypes.byref(d)) == 4 * 256
    d8 = _lib.MlpDesc(num_layers=3, feat_dim=100, hidden_dim=2048, label_dim=47, norm=1, dropout=0.2,
                      bn_eps=1e-5, bn_momentum=0.1)
    assert lib.glnn_mlp_param_count(ctypes.byref(d8)) == 4507695
    assert lib.glnn_mlp_workspace_bytes(ctypes.byref(d8), 4096) > 0
    layers = (_lib.GnnLayer * 3)()
    for l, (a, b) in enumerate([(100, 256), (256, 256), (256, 47)]):
        layers[l].d_in, layers[l].d_out = a, b
    assert lib.glnn_gnn_forward_workspace_bytes(1000, layers, 3) >= 3 * 1000 * 256 * 4


def test_argument_errors_are_reported_without_a_gpu():
    from glnn_b200 import _lib
    lib = _lib.load()
    rc = lib.glnn_gemm_f32(None, 4, 0, None, 4, 0, None, 4, 4, 4, 4, None, None, None, None, 0, 0, None)
    assert rc == -1 and b"null" in lib.glnn_last_error()
    rc = lib.glnn_spmm_csr_f32(None, 0, None, None, 4, None, 4, -1, 4, 4, 0, 0, None, None, None, None,
                               None, 0, None)
    assert rc == -1
    with pytest.raises(ValueError):
        _lib.check(rc, "spmm")


def test_product_path_has_no_cpu_fallback():
    from glnn_b200 import _lib, ops
    with pytest.raises(_lib.GlnnError):
        ops.spmm_csr(torch.zeros(2, dtype=torch.int32), torch.zeros(0, dtype=torch.int32),
                     torch.zeros(1, 4))
    from glnn_b200.models import Model
    m = Model(dict(model_name="SAGE", num_layers=2, feat_dim=4, hidden_dim=4, label_dim=2,
                   dropout_ratio=0.0, norm_type="none", device="cpu"))
    from glnn_b200.graph import graph
    g = graph((np.array([0, 1]), np.array([1, 0])), num_nodes=2)
    with pytest.raises(_lib.GlnnError):
        m.inference(g, torch.zeros(2, 4))


def test_step_functions_raise_instead_of_falling_back(monkeypatch):
    """No silent eager-torch path: CPU tensors (or anything else outside the fused kernels) raise
    unless the caller sets GLNN_ALLOW_TORCH_FALLBACK=1."""
    from glnn_b200 import _lib, train_and_eval as TE
    from glnn_b200.models import Model
    monkeypatch.delenv("GLNN_ALLOW_TORCH_FALLBACK", raising=False)
    m = Model(dict(model_name="MLP", num_layers=2, feat_dim=6, hidden_dim=8, label_dim=3,
                   dropout_ratio=0.0, norm_type="none", device="cpu"))
    opt = torch.optim.Adam(m.parameters(), lr=0.01)
    x, y = torch.randn(20, 6), torch.randint(0, 3, (20,))
    with pytest.raises(_lib.GlnnError):
        TE.train_mini_batch(m, x, y, 10, torch.nn.NLLLoss(), opt)
    with pytest.raises(_lib.GlnnError):
        TE.evaluate_mini_batch(m, x, y, torch.nn.NLLLoss(), 10, lambda o, l: 0.0)
    with pytest.raises(_lib.GlnnError):
        m(None, x)
    m.eval()
    with torch.no_grad(), pytest.raises(_lib.GlnnError):
        m(None, x)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "graphless-neural-networks_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(import|from)\s+(glnn_oracle|dgl_shim|oracle)\b", src,
                                 flags=re.M), fn


def test_csr_graph_matches_oracle_builder():
    import glnn_oracle as O
    from glnn_b200.graph import graph
    rng = np.random.default_rng(0)
    src, dst = rng.integers(0, 50, 400), rng.integers(0, 50, 400)
    g = graph((src, dst), num_nodes=50)
    indptr, indices = O.csr_from_edges(src, dst, 50)
    assert np.array_equal(g.indptr.numpy(), indptr) and g.indptr.dtype == torch.int32
    assert np.array_equal(g.indices.numpy(), indices)
    assert np.array_equal(g.in_degrees().numpy(), np.bincount(dst, minlength=50))
    assert np.array_equal(g.out_degrees().numpy(), np.bincount(src, minlength=50))
    sub = g.subgraph(torch.tensor([3, 1, 7, 20]))
    s, d = sub.edges()
    assert sub.num_nodes() == 4 and (s.numel() == 0 or int(s.max()) < 4)


def test_model_surface_and_state_dict_keys(golden_dir):
    """Same ctor keys / dispatch / state_dict names as the reference, and identical seeded init."""
    from helpers import load
    from glnn_b200.models import Model
    from glnn_b200.utils import set_seed
    d = load("student_mlp_bn3")
    set_seed(0)
    m = Model(dict(model_name="MLP", num_layers=3, feat_dim=20, hidden_dim=32, label_dim=7,
                   dropout_ratio=0.0, norm_type="batch", device="cpu"))
    sd = m.state_dict()
    ref_keys = sorted(k[len("init."):] for k in d if k.startswith("init."))
    assert sorted(sd) == ref_keys
    for k in ref_keys:  # seeded construction consumes the RNG exactly like the reference
        assert np.array_equal(sd[k].numpy(), d["init." + k]), k
    t = load("teacher_sage_bn3")
    s = Model(dict(model_name="SAGE", num_layers=3, feat_dim=20, hidden_dim=32, label_dim=7,
                   dropout_ratio=0.5, norm_type="batch", device="cpu"))
    assert sorted(s.state_dict()) == sorted(k[3:] for k in t if k.startswith("sd."))
    c = load("teacher_gcn_cora_like")
    gm = Model(dict(model_name="GCN", num_layers=2, feat_dim=50, hidden_dim=16, label_dim=7,
                    dropout_ratio=0.5, norm_type="none", device="cpu"))
    assert sorted(gm.state_dict()) == sorted(k[3:] for k in c if k.startswith("sd."))
    assert tuple(gm.state_dict()["encoder.layers.0.weight"].shape) == (50, 16)  # [in, out]
    assert isinstance(Model(dict(model_name="GA1MLP3w4", num_layers=2, feat_dim=4, hidden_dim=4,
                                 label_dim=2, dropout_ratio=0.0, norm_type="none",
                                 device="cpu")).encoder, type(m.encoder))
    with pytest.raises(NotImplementedError):
        Model(dict(model_name="GAT", num_layers=2, feat_dim=4, hidden_dim=4, label_dim=2,
                   dropout_ratio=0.0, norm_type="none", device="cpu", attn_dropout_ratio=0.1))


def test_batch_index_rule():
    from glnn_b200.train_and_eval import _batch_index
    torch.manual_seed(0)
    assert tuple(_batch_index(1000, 64).shape) == (15, 64)   # tail of 40 rows dropped
    assert tuple(_batch_index(50, 512).shape) == (1, 50)     # n < bs -> one batch of n
    torch.manual_seed(3)
    a = _batch_index(100, 10)
    torch.manual_seed(3)
    assert torch.equal(a.view(-1), torch.randperm(100))      # same CPU randperm call as the reference


def test_cpu_generic_loop_matches_oracle_without_gpu(monkeypatch):
    """With GLNN_ALLOW_TORCH_FALLBACK=1 out-of-scope configurations run a generic autograd loop;
    check it against the oracle so the opt-in stays correct too (CPU tensors, tiny)."""
    monkeypatch.setenv("GLNN_ALLOW_TORCH_FALLBACK", "1")
    import glnn_oracle as O
    import warnings
    from glnn_b200 import train_and_eval as TE
    from glnn_b200.models import Model
    from glnn_b200.utils import set_seed
    set_seed(1)
    m = Model(dict(model_name="MLP", num_layers=2, feat_dim=6, hidden_dim=8, label_dim=3,
                   dropout_ratio=0.0, norm_type="none", device="cpu"))
    p = {k[len("encoder."):]: v.detach().clone() for k, v in m.state_dict().items()}
    st = O.init_adam_state(p)
    opt = torch.optim.Adam(m.parameters(), lr=0.01, weight_decay=1e-3)
    x, y = torch.randn(40, 6), torch.randint(0, 3, (40,))
    torch.manual_seed(9)
    perm = torch.randperm(40)[:40].view(4, 10)
    want = O.train_mini_batch(p, st, x, y, "nll", 10, perm, 0.5, 2, "none", 0.0, 0.01, 1e-3)
    torch.manual_seed(9)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = TE.train_mini_batch(m, x, y, 10, torch.nn.NLLLoss(), opt, 0.5)
    assert abs(got - want) < 1e-5
    assert torch.allclose(m.state_dict()["encoder.layers.1.weight"], p["layers.1.weight"], atol=1e-5)


def test_training_config_yaml_wins(tmp_path):
    from glnn_b200.utils import get_training_config
    y = tmp_path / "c.yaml"
    y.write_text("global:\n  num_layers: 2\n  hidden_dim: 128\nd:\n  M:\n    hidden_dim: 64\n  N:\n")
    assert get_training_config(str(y), "M", "d") == {"num_layers": 2, "hidden_dim": 64, "model_name": "M"}
    assert get_training_config(str(y), "N", "d")["hidden_dim"] == 128


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference (CPU, small workload): exactly ONE line on stdout, valid JSON, with
    the keys the driver's contract names for the reference arm."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--workload", "ogbn-arxiv", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "nodes/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    # a non-zero rank of a torchrun launch exits 0 without work and without output
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                        capture_output=True, text=True, timeout=120, cwd=ROOT,
                        env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r2.returncode == 0 and r2.stdout.strip() == ""
