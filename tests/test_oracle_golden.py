"""CPU: pin oracle/glnn_oracle.py (the restatement that travels to the GPU box) against the fixtures
produced by the reference's own code (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

import glnn_oracle as O
from helpers import STUDENT_CASES, TEACHER_CASES, load, relerr, student_masks, sub, teacher_params

TOL = 1e-4  # north_star: 1e-4 relative fp32


@pytest.mark.parametrize("case", TEACHER_CASES)
@pytest.mark.parametrize("batched", [False, True])
def test_teacher_oracle_matches_reference(case, batched):
    d = load("teacher_" + case)
    n = int(d["n"])
    indptr, indices = O.csr_from_edges(d["src"], d["dst"], n)
    layers, norms = teacher_params(d)
    feats = torch.from_numpy(d["feats"])
    if str(d["model_name"]) == "SAGE":
        bs = int(d["batch_size"]) if batched else None
        logits = O.sage_inference(indptr, indices, feats, layers, norms, batch_size=bs)
    else:
        logits = O.gcn_forward(indptr, indices, feats, layers, norms)
    assert relerr(logits, d["logits"]) < 1e-5
    out = torch.log_softmax(logits, 1)
    idx = torch.from_numpy(d["idx_eval"])
    loss, score = O.nll_and_acc(out[idx], torch.from_numpy(d["labels"])[idx])
    assert abs(loss - float(d["loss"])) < 1e-5 * max(1.0, abs(float(d["loss"])))
    assert abs(score - float(d["score"])) < 1e-6


def test_sage_known_answers():
    """Hand-computed rule R1 cases: multi-edge multiplicity, zero in-degree, self-loop double count."""
    # edges: 0->1 twice (multi-edge), 2->1, 1->1 (explicit self loop); node 0 and 2 have in-degree 0
    src, dst = np.array([0, 0, 2, 1]), np.array([1, 1, 1, 1])
    indptr, indices = O.csr_from_edges(src, dst, 3)
    h = torch.tensor([[1.0, 2.0], [10.0, 20.0], [100.0, 200.0]])
    w, b = torch.eye(2), torch.zeros(2)
    y = O.sage_gcn_conv(indptr, indices, h, 3, w, b)
    assert torch.equal(y[0], h[0])  # in-degree 0 -> h_v / 1
    assert torch.equal(y[2], h[2])
    # node 1: (1+1+100+10 [edges] + 10 [implicit self]) / (4+1)
    assert torch.allclose(y[1], torch.tensor([(1 + 1 + 100 + 10 + 10) / 5.0, (2 + 2 + 200 + 20 + 20) / 5.0]))


def test_gcn_zero_in_degree_raises():
    indptr, indices = O.csr_from_edges(np.array([0]), np.array([1]), 2)
    with pytest.raises(ValueError):
        O.gcn_forward(indptr, indices, torch.ones(2, 3), [(torch.ones(3, 2), torch.zeros(2))])


def _run_student(d, dtype=torch.float32):
    L, norm = int(d["num_layers"]), str(d["norm"])
    p = sub(d, "init.", dtype)
    state = O.init_adam_state(p)
    feats = torch.from_numpy(d["feats"]).to(dtype)
    labels = torch.from_numpy(d["labels"])
    out_t = torch.from_numpy(d["out_t"]).to(dtype)
    n_l, bs, lamb = int(d["n_l"]), int(d["batch_size"]), float(d["lamb"])
    dropout = float(d["dropout"])
    flat_masks = student_masks(d, 0, L)
    mi = 0
    losses = []
    for ep in range(int(d["epochs"])):
        for pi, (f, t, kind, lam) in enumerate([(feats[:n_l], labels[:n_l], "nll", lamb),
                                                (feats, out_t, "kl", 1 - lamb)]):
            perm = torch.from_numpy(d[f"perm.{2 * ep + pi}"])
            nb = max(1, f.shape[0] // bs)
            idx = perm[: nb * bs].view(nb, -1) if nb > 1 else perm[: nb * bs].view(1, -1)
            masks = None
            if dropout > 0:
                masks = []
                for _ in range(nb):
                    masks.append([flat_masks[mi + l] for l in range(L - 1)])
                    mi += L - 1
            losses.append(O.train_mini_batch(p, state, f, t, kind, bs, idx, lam, L, norm, dropout,
                                             float(d["lr"]), float(d["wd"]), masks))
    out_all = O.evaluate_mini_batch(p, feats, bs, L, norm)
    return p, state, losses, out_all


def noise_driven(k, L, norm):
    """Keys whose value is driven by Adam-amplified rounding noise in the reference itself: a Linear
    bias that feeds BatchNorm has a mathematically-zero gradient (BN removes the column mean), its
    fp32 gradient is ~1e-9 summation noise, and Adam (eps=1e-8) turns that into O(lr) random moves.
    The bias cancels in train mode but leaks into running_mean; both are excluded from parity and
    the eval-mode forward is pinned separately on identical state (test_student_eval_...)."""
    if norm != "batch":
        return False
    if k.startswith("layers.") and k.endswith(".bias") and int(k.split(".")[1]) != L - 1:
        return True
    return k.endswith("running_mean")


@pytest.mark.parametrize("case", STUDENT_CASES)
def test_student_oracle_matches_reference(case):
    d = load("student_" + case)
    p, state, losses, out_all = _run_student(d)
    assert np.allclose(losses, d["losses"], rtol=1e-4, atol=1e-6)
    final = sub(d, "final.")
    L, norm = int(d["num_layers"]), str(d["norm"])
    for k, v in final.items():
        if k.endswith("num_batches_tracked"):
            assert int(p[k]) == int(v), k
        elif not noise_driven(k, L, norm):
            assert relerr(p[k], v) < 5e-4, k
    for k in state:
        if noise_driven(k, L, norm):
            continue
        assert state[k]["step"] == int(d[f"adam.encoder.{k}.step"])
        assert relerr(state[k]["exp_avg"], d[f"adam.encoder.{k}.exp_avg"]) < 2e-3, k
        assert relerr(state[k]["exp_avg_sq"], d[f"adam.encoder.{k}.exp_avg_sq"]) < 2e-3, k
    if norm != "batch" or L == 1:
        assert relerr(out_all, d["out_all"]) < TOL


@pytest.mark.parametrize("case", STUDENT_CASES)
def test_student_eval_on_reference_state(case):
    """evaluate_mini_batch (train_and_eval.py:108-136) on the reference's own final state."""
    d = load("student_" + case)
    p = sub(d, "final.")
    out_all = O.evaluate_mini_batch(p, torch.from_numpy(d["feats"]), int(d["batch_size"]),
                                    int(d["num_layers"]), str(d["norm"]))
    assert relerr(out_all, d["out_all"]) < 1e-5
    loss, score = O.nll_and_acc(out_all, torch.from_numpy(d["labels"]))
    assert abs(loss - float(d["loss_eval"])) < 1e-5
    assert abs(score - float(d["score_eval"])) < 1e-6


def test_student_lamb0_hard_pass_still_moves_parameters():
    """with lamb=0 the hard-label pass has zero gradients but Adam still moves
    parameters through stale momentum / weight decay; skipping it changes the result."""
    d = load("student_mlp_lamb0")
    assert float(d["lamb"]) == 0.0
    p, _, _, _ = _run_student(d)
    final = sub(d, "final.")
    assert relerr(p["layers.2.weight"], final["layers.2.weight"]) < 5e-4


@pytest.mark.parametrize("case", ["teacher_train_gcn2", "teacher_train_gcn3_lamb"])
def test_gcn_train_step_oracle_matches_reference(case):
    """Full-batch GCN TRAINING steps: the oracle's manual backward + Adam against
    the reference's own `train` (autograd through GCN.forward over the shim, torch.optim.Adam):
    per-step losses and every parameter after the last step."""
    d = load(case)
    indptr, indices = O.csr_from_edges(d["src"], d["dst"], int(d["n"]))
    L = int(d["num_layers"])
    p = {f"layers.{l}.{k}": torch.from_numpy(d[f"init.encoder.layers.{l}.{k}"]).double().clone()
         for l in range(L) for k in ("weight", "bias")}
    st = O.init_adam_state(p)
    feats, labels = torch.from_numpy(d["feats"]).double(), torch.from_numpy(d["labels"])
    losses = [O.gcn_train_step(indptr, indices, feats, labels, d["idx_train"], p, st, float(d["lamb"]),
                               float(d["lr"]), float(d["wd"])) for _ in range(len(d["losses"]))]
    assert np.allclose(losses, d["losses"], rtol=2e-6)
    for k, v in p.items():
        want = torch.from_numpy(d[f"final.encoder.{k}"]).double()
        assert relerr(v, want) < 2e-5, k
    # the trained parameters reproduce the reference's evaluate() output
    out = torch.log_softmax(O.gcn_forward(indptr, indices, feats,
                                          [(p[f"layers.{l}.weight"], p[f"layers.{l}.bias"]) for l in range(L)]), 1)
    assert relerr(out, torch.from_numpy(d["out"]).double()) < 2e-5


@pytest.mark.parametrize("case", ["teacher_train_sage2", "teacher_train_sage3_lamb"])
def test_sage_block_train_step_oracle_matches_reference(case):
    """Block-wise GraphSAGE TRAINING steps: the oracle's manual backward through
    full-neighbour blocks + Adam against the reference's own `train_sage` / SAGE.forward (autograd over
    the shim's SAGEConv): per-step losses and every parameter after the last step."""
    d = load(case)
    indptr, indices = O.csr_from_edges(d["src"], d["dst"], int(d["n"]))
    L = int(d["num_layers"])
    p = {f"layers.{l}.fc_neigh.{k}": torch.from_numpy(d[f"init.encoder.layers.{l}.fc_neigh.{k}"]).double().clone()
         for l in range(L) for k in ("weight", "bias")}
    st = O.init_adam_state(p)
    feats, labels = torch.from_numpy(d["feats"]).double(), torch.from_numpy(d["labels"])
    losses = [O.sage_block_train_step(indptr, indices, feats, labels, d["seeds"], p, st, float(d["lamb"]),
                                      float(d["lr"]), float(d["wd"])) for _ in range(len(d["losses"]))]
    assert np.allclose(losses, d["losses"], rtol=2e-6)
    for k, v in p.items():
        assert relerr(v, torch.from_numpy(d[f"final.encoder.{k}"]).double()) < 2e-5, k


@pytest.mark.parametrize("case", ["tran_none2", "tran_bn3", "ind_none3", "ind_bn2"])
def test_oracle_epoch_loop_matches_reference_runners(case):
    """a11 / f4: O.distill_run against the history the reference's own distill_run_transductive /
    distill_run_inductive produced (fixtures runner_*.npz): same number of epochs (early stopping),
    per-epoch evaluation losses and scores, final scores and returned log-probabilities."""
    d = load("runner_" + case)
    inductive = bool(int(d["inductive"]))
    L, norm = int(d["num_layers"]), str(d["norm"])
    p = sub(d, "init.")
    state = O.init_adam_state(p)
    feats, labels = torch.from_numpy(d["feats"]), torch.from_numpy(d["labels"])
    out_t = torch.from_numpy(d["out_t"])
    indices = tuple(torch.from_numpy(d[f"index.{i}"]) for i in range(6 if inductive else 4))
    perms = [torch.from_numpy(d[f"perm.{i}"]) for i in range(int(d["n_perms"]))]
    out, scores, hist = O.distill_run(p, state, feats, labels, out_t, indices, inductive, perms,
                                      int(d["batch_size"]), float(d["lamb"]), L, norm, float(d["lr"]),
                                      float(d["wd"]), int(d["patience"]), int(d["max_epoch"]))
    want = d["hist"]
    got = np.array(hist, dtype=np.float64)
    assert got.shape == want.shape
    nloss = (want.shape[1] - 1) // 2
    # With BatchNorm the eval-mode outputs after training carry the noise-driven Linear biases (zero
    # mathematical gradient, Adam-amplified rounding noise, seen through the lagging running_mean --
    # see noise_driven() above): the reference is not reproducible against ITSELF across summation
    # orders below ~1e-3 there, so the bound is per norm type; scores may move by one node.
    rtol = 2e-4 if norm == "none" else 2e-3
    assert np.allclose(got[:, 1:1 + nloss], want[:, 1:1 + nloss], rtol=rtol, atol=1e-6)
    sizes = [indices[0].numel(), indices[2].numel(), indices[3].numel()] + \
        ([indices[5].numel()] if inductive else [])
    for j, m in enumerate(sizes):
        tol = 1e-6 if norm == "none" else 1.0 / m + 1e-6
        assert np.all(np.abs(got[:, 1 + nloss + j] - want[:, 1 + nloss + j]) <= tol), j
    assert np.allclose(scores, d["scores"], atol=1e-6 if norm == "none" else 1.0 / min(sizes) + 1e-6)
    assert relerr(out, d["out"]) < (1e-4 if norm == "none" else 5e-3)
