"""Drop-in surface: CLI flags/defaults equal the reference's, the CPF loader keeps the graph
conventions the kernels rely on, and (GPU) train_teacher.py -> out.npz -> train_student.py runs end
to end on a synthetic CPF-format dataset with the reference's output tree."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_cpf_npz(path, n=400, f=60, classes=5, seed=0):
    rng = np.random.default_rng(seed)
    labels = np.arange(n) % classes
    rows, cols = [], []
    for i in range(1, n):  # a spanning chain keeps one big component; plus random intra-class edges
        rows.append(i)
        cols.append(i - 1)
    extra = rng.integers(0, n, (3 * n, 2))
    rows += list(extra[:, 0])
    cols += list(extra[:, 1])
    adj = sp.csr_matrix((np.ones(len(rows), dtype=np.float32), (rows, cols)), shape=(n, n))
    centers = rng.normal(size=(classes, f))
    attr = sp.csr_matrix(((centers[labels] + rng.normal(size=(n, f))) > 0.8).astype(np.float32))
    np.savez(path, adj_data=adj.data, adj_indices=adj.indices, adj_indptr=adj.indptr,
             adj_shape=adj.shape, attr_data=attr.data, attr_indices=attr.indices,
             attr_indptr=attr.indptr, attr_shape=attr.shape, labels=labels)


def test_cli_flags_match_reference():
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference tree only exists in the authoring container")
    import importlib.util
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dgl_shim
    dgl_shim.install()
    from glnn_b200 import cli
    for role, fname in (("teacher", "train_teacher.py"), ("student", "train_student.py")):
        src = open(os.path.join(ref, fname)).read()
        # evaluate only the reference's get_args() (argparse), nothing else of the script
        start = src.index("def get_args():")
        end = src.index("\ndef run(args):")
        ns = {"argparse": __import__("argparse")}
        exec(src[start:end], ns)
        old = sys.argv
        sys.argv = [fname]
        try:
            want = vars(ns["get_args"]())
        finally:
            sys.argv = old
        got = vars(cli.build_parser(role).parse_args([]))
        assert got == want, (role, {k: (got.get(k), want.get(k)) for k in set(got) | set(want)
                                    if got.get(k) != want.get(k)})


def test_cpf_loader_conventions(tmp_path, monkeypatch):
    from glnn_b200.dataloader import load_data
    (tmp_path / "data").mkdir()
    write_cpf_npz(tmp_path / "data" / "cora.npz")
    monkeypatch.chdir(tmp_path)
    g, labels, itr, iva, ite = load_data("cora", "./data", seed=0, labelrate_train=20, labelrate_val=30,
                                         split_idx=0)
    n = g.num_nodes()
    src, dst = g.edges()
    assert int((src == dst).sum()) == n                  # exactly one self-loop per node
    key = src * n + dst
    assert key.unique().numel() == key.numel()           # unweighted, de-duplicated
    rev = dst * n + src
    assert set(key.tolist()) == set(rev.tolist())        # symmetric
    assert int(g.in_degrees().min()) >= 1                # GraphConv never sees in-degree 0
    assert g.ndata["feat"].shape == (n, 60) and labels.shape == (n,)
    assert itr.numel() == 5 * 20 and iva.numel() == 5 * 30
    assert itr.numel() + iva.numel() + ite.numel() == n
    assert len(set(itr.tolist()) & set(iva.tolist())) == 0
    g2, _, itr2, _, _ = load_data("cora", "./data", seed=0, labelrate_train=20, labelrate_val=30,
                                  split_idx=0)
    assert torch.equal(itr, itr2)                        # seeded split is reproducible
    _, _, itr3, _, _ = load_data("cora", "./data", seed=1, labelrate_train=20, labelrate_val=30,
                                 split_idx=0)
    assert not torch.equal(itr, itr3)


@pytest.mark.gpu
def test_cli_teacher_then_student_end_to_end(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    (tmp_path / "data").mkdir()
    write_cpf_npz(tmp_path / "data" / "cora.npz")
    env = dict(os.environ, PYTHONPATH=ROOT)
    common = ["--dataset", "cora", "--device", "0", "--max_epoch", "30", "--patience", "30",
              "--model_config_path", os.path.join(ROOT, "train.conf.yaml"), "--save_results"]
    for teacher in ("GCN", "SAGE"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "train_teacher.py"), "--teacher",
                            teacher] + common, cwd=tmp_path, env=env, capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        score = float(r.stdout.strip().split("\t")[0])
        out_dir = tmp_path / "outputs" / "transductive" / "cora" / teacher / "seed_0"
        out = np.load(out_dir / "out.npz")["arr_0"]
        assert out.dtype == np.float32 and out.shape[1] == 5
        assert np.allclose(np.exp(out).sum(1), 1.0, atol=1e-4)   # log-probabilities
        assert (out_dir / "log").exists() and (out_dir / "model.pth").exists()
        assert (out_dir.parent / "exp_results").read_text().strip() != ""
        assert score > 0.5, (teacher, score)                     # 5 separable classes: well above chance
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train_student.py"), "--teacher", "SAGE",
                        "--student", "MLP", "--lamb", "0.5"] + common, cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    s_dir = tmp_path / "outputs" / "transductive" / "cora" / "SAGE_MLP" / "seed_0"
    s_out = np.load(s_dir / "out.npz")["arr_0"]
    assert s_out.shape == out.shape
    assert float(r.stdout.strip().split("\t")[0]) > 0.4
    # inductive setting: two scores on the printed line
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train_teacher.py"), "--teacher", "SAGE",
                        "--exp_setting", "ind"] + common, cwd=tmp_path, env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(r.stdout.strip().split("\t")) == 2


@pytest.mark.gpu
def test_teacher_to_student_hand_off_stays_on_device(tmp_path, monkeypatch):
    """SURVEY 8f row 3 (train_teacher.py:296-297 -> dataloader.py:169-170): in one process the
    student takes the teacher's log-probabilities from device memory; out.npz is still written in the
    reference's format, and a student that reads the file back scores the same."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from glnn_b200 import cli, dataloader
    (tmp_path / "data").mkdir()
    write_cpf_npz(tmp_path / "data" / "cora.npz")
    monkeypatch.chdir(tmp_path)
    common = ["--dataset", "cora", "--device", "0", "--max_epoch", "12", "--patience", "12",
              "--model_config_path", os.path.join(ROOT, "train.conf.yaml")]
    calls = []
    orig = cli.load_out_t
    monkeypatch.setattr(cli, "load_out_t", lambda d: (calls.append(d), orig(d))[1])
    t_line, s_line = cli.distill_pipeline(["--teacher", "SAGE"] + common,
                                          ["--teacher", "SAGE", "--student", "MLP", "--lamb", "0.5",
                                           "--dropout_ratio", "0"] + common)
    assert calls == []                                     # no read-back of out.npz
    t_dir = tmp_path / "outputs" / "transductive" / "cora" / "SAGE" / "seed_0"
    key = str(t_dir)
    assert key in cli._OUT_T_ON_DEVICE and cli._OUT_T_ON_DEVICE[key].is_cuda
    on_disk = np.load(t_dir / "out.npz")["arr_0"]
    assert on_disk.dtype == np.float32
    assert np.array_equal(on_disk, cli._OUT_T_ON_DEVICE[key].cpu().numpy())   # same bits both ways
    # the file path gives the same student (same seed, same inputs)
    cli._OUT_T_ON_DEVICE.clear()
    import shutil
    shutil.rmtree(tmp_path / "outputs" / "transductive" / "cora" / "SAGE_MLP")
    s_line2 = cli.main("student", ["--teacher", "SAGE", "--student", "MLP", "--lamb", "0.5",
                                   "--dropout_ratio", "0"] + common)
    assert len(calls) == 1
    assert abs(float(s_line.split("\t")[0]) - float(s_line2.split("\t")[0])) < 0.02


def test_cli_refuses_to_run_without_cuda_before_creating_outputs(tmp_path, monkeypatch):
    """--device defaults to -1 (CPU) in the reference; this implementation has no CPU path and must
    say so up front instead of failing deep inside a kernel wrapper after creating output_dir."""
    from glnn_b200 import cli
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cli.main("teacher", ["--teacher", "GCN", "--dataset", "cora"])
    assert not (tmp_path / "outputs").exists()


def test_arxiv_edge_preprocessing_known_answer():
    """dataloader.py:74-77 on a 4-node multigraph: reverse edges appended without de-duplication,
    original self-loops dropped, one self-loop per node added; in-degrees count multiplicities."""
    from glnn_b200.dataloader import arxiv_edges
    from glnn_b200.graph import graph
    src = torch.tensor([0, 1, 2, 0, 3])
    dst = torch.tensor([1, 0, 2, 1, 0])          # 0->1 twice, 1->0, a self-loop on 2, 3->0
    s, d = arxiv_edges(src, dst, 4)
    pairs = sorted(zip(s.tolist(), d.tolist()))
    want = sorted([(0, 1)] * 3 + [(1, 0)] * 3 + [(3, 0), (0, 3)] + [(i, i) for i in range(4)])
    assert pairs == want
    g = graph((s, d), num_nodes=4)
    assert g.in_degrees().tolist() == [3 + 1 + 1, 3 + 1, 1, 1 + 1]
    assert g.num_edges() == 2 * 4 + 4             # 2 x (5 - 1 self-loop) + n
