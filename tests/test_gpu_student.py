"""GPU parity tests for the student path (B): the fused train_mini_batch pass and evaluate_mini_batch
against (1) the fixtures produced by the reference's own train_mini_batch / evaluate_mini_batch
(autograd + torch.optim.Adam) and (2) the CPU oracle at the real layer shapes."""
import copy

import numpy as np
import pytest
import torch

import glnn_oracle as O
from helpers import STUDENT_CASES, assert_parity, load, relerr, relerr_q, student_masks, sub
from test_oracle_golden import noise_driven

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _student_model(d, dev, prefix="init."):
    from glnn_b200.models import Model
    conf = dict(model_name=str(d["model_name"]), num_layers=int(d["num_layers"]),
                feat_dim=d["feats"].shape[1], hidden_dim=int(d["hidden"]),
                label_dim=d["out_t"].shape[1], dropout_ratio=float(d["dropout"]),
                norm_type=str(d["norm"]), device=dev)
    model = Model(conf)
    sd = {k[len(prefix):]: torch.from_numpy(np.array(v)) for k, v in d.items() if k.startswith(prefix)}
    model.load_state_dict(sd)
    return model


class _Replay:
    """Feeds the permutations the reference drew (recorded in the fixture) to torch.randperm."""

    def __init__(self, perms):
        self.perms, self.i, self._orig = perms, 0, torch.randperm

    def __enter__(self):
        def randperm(n, *a, **k):
            p = self.perms[self.i]
            self.i += 1
            assert p.numel() == n
            return p.clone()
        torch.randperm = randperm
        return self

    def __exit__(self, *a):
        torch.randperm = self._orig


def _run_fixture(d, dev, use_masks):
    from glnn_b200 import mlp_engine, train_and_eval as TE
    model = _student_model(d, dev)
    opt = torch.optim.Adam(model.parameters(), lr=float(d["lr"]), weight_decay=float(d["wd"]))
    feats = torch.from_numpy(d["feats"]).to(dev)
    labels = torch.from_numpy(d["labels"]).to(dev)
    out_t = torch.from_numpy(d["out_t"]).to(dev)
    n_l, bs, lamb = int(d["n_l"]), int(d["batch_size"]), float(d["lamb"])
    L, H = int(d["num_layers"]), int(d["hidden"])
    crit_l = torch.nn.NLLLoss()
    crit_t = torch.nn.KLDivLoss(reduction="batchmean", log_target=True)
    perms = [torch.from_numpy(d[f"perm.{i}"]) for i in range(2 * int(d["epochs"]))]
    losses = []
    if not use_masks:
        with _Replay(perms):
            for _ in range(int(d["epochs"])):
                losses.append(TE.train_mini_batch(model, feats[:n_l], labels[:n_l], bs, crit_l, opt, lamb))
                losses.append(TE.train_mini_batch(model, feats, out_t, bs, crit_t, opt, 1 - lamb))
    else:  # parity mode with the reference's recorded dropout keep-masks
        flat = student_masks(d, 0, L)
        mi = 0
        model.train()
        for ep in range(int(d["epochs"])):
            for pi, (f, t, lam) in enumerate([(feats[:n_l], labels[:n_l], lamb), (feats, out_t, 1 - lamb)]):
                perm = perms[2 * ep + pi]
                nb = max(1, f.shape[0] // bs)
                idx = perm[: nb * bs].view(nb, -1)
                rows = idx.shape[1]
                masks = torch.stack([torch.stack([flat[mi + s * (L - 1) + l] for l in range(L - 1)])
                                     for s in range(nb)]).to(dev)
                mi += nb * (L - 1)
                assert masks.shape == (nb, L - 1, rows, H)
                loss = mlp_engine.train_pass(model.encoder, opt, f, t, idx.to(dev), lam, masks)
                losses.append(loss.item() / nb)
    return model, opt, losses, feats, labels


@pytest.mark.parametrize("case", STUDENT_CASES)
def test_student_matches_reference_golden(dev, case):
    from glnn_b200 import train_and_eval as TE, utils as U
    d = load("student_" + case)
    use_masks = float(d["dropout"]) > 0
    model, opt, losses, feats, labels = _run_fixture(d, dev, use_masks)
    assert np.allclose(losses, d["losses"], rtol=TOL, atol=1e-6)
    L, norm = int(d["num_layers"]), str(d["norm"])
    sd = {k[len("encoder."):]: v.detach().cpu() for k, v in model.state_dict().items()}
    final = sub(d, "final.")
    for k, v in final.items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v), k
        elif not noise_driven(k, L, norm):
            # ~30 Adam steps from the reference's initial state: the per-pass losses above are the
            # tight check; the weights are a sanity bound only, because the training dynamics amplify
            # any rounding difference (bf16x3 GEMMs, atomic summation order) by orders of magnitude
            # -- see test_student_real_shapes_vs_oracle for the oracle's own fp32-vs-fp64 deviation
            assert relerr_q(sd[k], v, 0.99) < 5e-2, k
    for name, p in model.named_parameters():
        k = name[len("encoder."):]
        if noise_driven(k, L, norm):
            continue
        st = opt.state[p]
        assert int(st["step"]) == int(d[f"adam.{name}.step"])
        assert relerr_q(st["exp_avg"].cpu(), d[f"adam.{name}.exp_avg"], 0.9) < 5e-2, k
        assert relerr_q(st["exp_avg_sq"].cpu(), d[f"adam.{name}.exp_avg_sq"], 0.9) < 5e-2, k
    if norm != "batch" or L == 1:
        out_all, loss, score = TE.evaluate_mini_batch(model, feats, labels, torch.nn.NLLLoss(),
                                                      int(d["batch_size"]), U.get_evaluator("cora"))
        assert relerr(out_all.cpu(), d["out_all"]) < TOL


@pytest.mark.parametrize("case", STUDENT_CASES)
def test_student_eval_on_reference_state(dev, case):
    """evaluate_mini_batch on the reference's final weights: log-probs, loss and accuracy."""
    from glnn_b200 import train_and_eval as TE, utils as U
    d = load("student_" + case)
    model = _student_model(d, dev, prefix="final.")
    feats = torch.from_numpy(d["feats"]).to(dev)
    labels = torch.from_numpy(d["labels"]).to(dev)
    out_all, loss, score = TE.evaluate_mini_batch(model, feats, labels, torch.nn.NLLLoss(),
                                                  int(d["batch_size"]), U.get_evaluator("cora"))
    assert relerr(out_all.cpu(), d["out_all"]) < TOL
    assert_parity(out_all.cpu(), d["out_all"], "student log-probabilities")   # + allclose(1e-4, 1e-5)
    assert abs(loss - float(d["loss_eval"])) < 1e-4
    assert abs(score - float(d["score_eval"])) < 1e-6
    # Model.forward in eval mode returns raw logits whose log_softmax is the same thing
    with torch.no_grad():
        logits = model.eval()(None, feats)
    assert relerr(logits.log_softmax(1).cpu(), d["out_all"]) < TOL


REAL_SHAPES = [(128, 256, 40, 512, "MLP"), (100, 256, 47, 1024, "MLP3w8"), (128, 1024, 40, 512, "MLP3w4"),
               (100, 2048, 47, 4096, "MLP3w8")]


def _real_problem(shape, dev, nb):
    from glnn_b200.models import Model
    f, h, c, bs, name = shape
    n = bs * nb + 17
    gen = torch.Generator().manual_seed(11)
    feats = torch.randn(n, f, generator=gen)
    labels = torch.randint(0, c, (n,), generator=gen)
    out_t = torch.log_softmax(torch.randn(n, c, generator=gen) * 2, 1)
    torch.manual_seed(5)
    model = Model(dict(model_name=name, num_layers=3, feat_dim=f, hidden_dim=h, label_dim=c,
                       dropout_ratio=0.0, norm_type="batch", device=dev))
    idx1 = torch.randperm(n, generator=gen)[: nb * bs].view(nb, bs)
    idx2 = torch.randperm(n, generator=gen)[: nb * bs].view(nb, bs)
    return model, feats, labels, out_t, idx1, idx2


def _oracle_state(model, dtype):
    return {k[len("encoder."):]: (v.detach().cpu().clone().to(dtype) if v.is_floating_point()
                                  else v.detach().cpu().clone()) for k, v in model.state_dict().items()}


@pytest.mark.parametrize("kind", ["nll", "kl"])
@pytest.mark.parametrize("shape", REAL_SHAPES)
def test_student_single_step_gradients(dev, shape, kind):
    """The precise parity check of the fused forward + loss + backward: gradients of ONE step from
    identical state against the fp64 oracle, at the real arxiv / products layer shapes (tcgen05
    bf16x3 projections included).  Loss and the last layer's gradients (no ReLU between them and the
    loss) are held to 1e-4.  Below the top hidden layer exact agreement is impossible for ANY pair of
    finite-precision implementations: a pre-activation within the forward rounding error of zero
    flips its ReLU mask, which changes that unit's gradients by ~1/sqrt(batch) and, through
    dX = dZ W, every gradient of the layers below by ~1e-3.  Measured here: ~2 flips per step among
    524k activations with bf16x3 (forward error ~6e-6); plain fp32 vs fp64 flips ~0.2 (arxiv) to ~3
    (products) activations per step as well.  Those tensors are held to 1e-2 (99th percentile)."""
    from glnn_b200 import mlp_engine
    f, h, c, bs, _ = shape
    model, feats, labels, out_t, idx1, _ = _real_problem(shape, dev, 1)
    p = _oracle_state(model, torch.float64)
    tgt = labels if kind == "nll" else out_t.double()
    logits, cache = O.mlp_forward(feats.double()[idx1[0]], p, 3, "batch", True)
    loss, dlog = O.loss_and_dlogits(logits, tgt[idx1[0]], kind, 0.6)
    want = O.mlp_backward(dlog, cache, p, 3, "batch", 0.0)
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    model.train()
    got_loss = mlp_engine.train_pass(model.encoder, opt, feats.to(dev),
                                     (labels if kind == "nll" else out_t).to(dev), idx1.to(dev), 0.6)
    assert abs(got_loss.item() - float(loss)) < 1e-4 * abs(float(loss))
    got = mlp_engine.flat_grads(model.encoder)
    for k, w in want.items():
        if noise_driven(k, 3, "batch"):
            continue   # Linear bias in front of BatchNorm: mathematically zero, rounding noise only
        if k.startswith("layers.2."):
            assert relerr(got[k].cpu(), w) < TOL, k
        else:
            assert relerr_q(got[k].cpu(), w, 0.99) < 1e-2, k
            assert relerr(got[k].cpu(), w) < 0.3, k


@pytest.mark.parametrize("graph_mode", [True, False])
@pytest.mark.parametrize("shape", REAL_SHAPES[:3])
def test_student_real_shapes_vs_oracle(dev, shape, graph_mode):
    """6 NLL + 6 KL steps at arxiv / products layer shapes.  Per-pass losses must match the fp64
    oracle to 3e-4: from the second step on the loss inherits the trajectory noise described below
    -- the ORACLE's own fp32 run deviates from its fp64 run by up to 1.1e-4 on single steps of these
    passes (measured: MLP3w8/256 KL step 5), and split-K / bias-gradient atomics make the B200 run
    differ from itself at that level between graph replay and direct launches; the first step of a
    pass (identical state) is held to 1e-4 by test_student_single_step_gradients.  Weights after 12 Adam steps cannot be held to 1e-4 by ANY implementation: the
    dynamics (batch statistics, ReLU masks, Adam's g/sqrt(v)) amplify rounding noise ~1e4x -- the
    ORACLE's own fp32 run deviates from its fp64 run by ~0.6 % (99th percentile, printed below) --
    so the trajectory check is a loose sanity bound; exact parity lives in the single-step gradient
    test above and in the eval-on-identical-state test."""
    from glnn_b200 import mlp_engine
    f, h, c, bs, name = shape
    nb = 6
    model, feats, labels, out_t, idx1, idx2 = _real_problem(shape, dev, nb)
    runs = {}
    for dt in (torch.float32, torch.float64):
        p = _oracle_state(model, dt)
        st = O.init_adam_state(p)
        losses = [O.train_mini_batch(p, st, feats.to(dt), labels, "nll", bs, idx1, 0.3, 3, "batch", 0.0, 0.01, 0.0),
                  O.train_mini_batch(p, st, feats.to(dt), out_t.to(dt), "kl", bs, idx2, 0.7, 3, "batch", 0.0, 0.01, 0.0)]
        runs[dt] = (p, losses)
    p32, p64 = runs[torch.float32][0], runs[torch.float64][0]
    opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0.0)
    model.train()
    fd, ld, td = feats.to(dev), labels.to(dev), out_t.to(dev)
    if graph_mode:
        got = [mlp_engine.train_pass(model.encoder, opt, fd, ld, idx1.to(dev), 0.3).item() / nb,
               mlp_engine.train_pass(model.encoder, opt, fd, td, idx2, 0.7).item() / nb]
    else:  # one step per call -> direct launches (no graph replay)
        got = [0.0, 0.0]
        for i in range(nb):
            got[0] += mlp_engine.train_pass(model.encoder, opt, fd, ld, idx1[i:i + 1].to(dev), 0.3).item() / nb
        for i in range(nb):
            got[1] += mlp_engine.train_pass(model.encoder, opt, fd, td, idx2[i:i + 1].to(dev), 0.7).item() / nb
    assert np.allclose(got, runs[torch.float64][1], rtol=3e-4)
    sd = {k[len("encoder."):]: v.detach().cpu() for k, v in model.state_dict().items()}
    for k in ("layers.0.weight", "layers.1.weight", "layers.2.weight", "layers.2.bias",
              "norms.0.weight", "norms.1.bias", "norms.0.running_var", "norms.1.running_var"):
        ref_dev = relerr_q(p32[k], p64[k], 0.99)
        gpu_dev = relerr_q(sd[k], p64[k], 0.99)
        print(f"{k}: fp32-oracle vs fp64 {ref_dev:.2e}, B200 vs fp64 {gpu_dev:.2e}")
        assert gpu_dev < 0.25, (k, gpu_dev, ref_dev)
    assert int(sd["norms.0.num_batches_tracked"]) == 2 * nb
    # eval forward on identical state
    model.load_state_dict({"encoder." + k: v for k, v in p32.items()})
    got_eval = mlp_engine.eval_forward(model.encoder, fd)
    want_eval = O.evaluate_mini_batch(p32, feats, bs, 3, "batch")
    assert relerr(got_eval.cpu(), want_eval) < TOL
    assert_parity(got_eval.cpu(), want_eval, "student eval on identical state")


def test_state_dict_roundtrip_and_views(dev):
    """Flat-buffer views survive deepcopy(state_dict()) / load_state_dict (early stopping path)."""
    from glnn_b200 import mlp_engine
    from glnn_b200.models import Model
    torch.manual_seed(0)
    model = Model(dict(model_name="MLP", num_layers=3, feat_dim=16, hidden_dim=32, label_dim=5,
                       dropout_ratio=0.0, norm_type="batch", device=dev))
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    x = torch.randn(256, 16, device=dev)
    y = torch.randint(0, 5, (256,), device=dev)
    idx = torch.arange(256).view(4, 64)
    mlp_engine.train_pass(model.encoder, opt, x, y, idx, 1.0)
    snap = copy.deepcopy(model.state_dict())
    before = mlp_engine.eval_forward(model.encoder, x).clone()
    mlp_engine.train_pass(model.encoder, opt, x, y, idx, 1.0)
    assert not torch.allclose(before, mlp_engine.eval_forward(model.encoder, x))
    model.load_state_dict(snap)
    assert torch.equal(before, mlp_engine.eval_forward(model.encoder, x))
    fl = model.encoder._flat
    assert model.encoder.layers[0].weight.data_ptr() == fl.params.data_ptr()


def test_dropout_device_stream_statistics(dev):
    """Fast mode (device counter-based RNG): finite, decreasing loss on a learnable target."""
    from glnn_b200 import mlp_engine
    from glnn_b200.models import Model
    torch.manual_seed(0)
    model = Model(dict(model_name="MLP", num_layers=3, feat_dim=32, hidden_dim=256, label_dim=8,
                       dropout_ratio=0.5, norm_type="batch", device=dev))
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    x = torch.randn(4096, 32, device=dev)
    y = (x[:, :8].argmax(1)).long()
    idx = torch.randperm(4096).view(8, 512)
    l0 = mlp_engine.train_pass(model.encoder, opt, x, y, idx, 1.0).item() / 8
    for _ in range(10):
        l1 = mlp_engine.train_pass(model.encoder, opt, x, y, idx, 1.0).item() / 8
    assert np.isfinite(l1) and l1 < l0
    acts = mlp_engine.eval_forward(model.encoder, x)
    assert torch.isfinite(acts).all()


# ---------------------------------------------------------------------------------------------
# gradients with the ReLU masks under control, Adam from injected gradients  
# ---------------------------------------------------------------------------------------------
def _clear_relu_margin(p, x, num_layers, norm, margin, gen, eps=1e-5):
    """Shifts, column by column, the BatchNorm beta (or the Linear bias without norm) of every hidden
    layer until NO pre-activation of the batch `x` lies within `margin` of zero (fp64 forward over the
    fp32-rounded parameters).  Then the ReLU masks of any implementation whose forward error is below
    `margin` equal the oracle's, and gradients can be compared without mask flips.
    p: fp64 oracle state (fp32-representable values), modified in place.  Returns the number of
    shifted columns."""
    h = x
    shifted = 0
    for l in range(num_layers - 1):
        z = h @ p[f"layers.{l}.weight"].t() + p[f"layers.{l}.bias"]
        if norm == "batch":
            mu, var = z.mean(0), z.var(0, unbiased=False)
            base = (z - mu) / torch.sqrt(var + eps) * p[f"norms.{l}.weight"]
            key = f"norms.{l}.bias"
        else:
            base = z - p[f"layers.{l}.bias"]
            key = f"layers.{l}.bias"
        off = p[key].clone()
        for _ in range(200):
            y = base + off
            bad = (y.abs() < margin).any(0)
            if not bool(bad.any()):
                break
            k = int(bad.sum())
            shifted += k
            off[bad] = (off[bad] + (torch.rand(k, generator=gen, dtype=torch.float64) - 0.5) * 0.05
                        ).float().double()          # stays fp32-representable
        else:
            raise AssertionError("could not clear the ReLU margin")
        p[key] = off
        h = torch.relu(base + off)
    return shifted


@pytest.mark.parametrize("kind", ["nll", "kl"])
@pytest.mark.parametrize("shape", REAL_SHAPES)
def test_student_single_step_gradients_relu_controlled(dev, shape, kind):
    """ALL gradients of one fused step at 1e-4 (SURVEY 8d gate), at the real arxiv / products layer
    shapes, once ReLU-mask flips are excluded by construction: the hidden pre-activations of the test
    batch are given a margin of 2e-4 around zero (30x the bf16x3 forward error), so the B200 step and
    the fp64 oracle see identical masks.  What remains is rounding: every tensor -- weights and
    biases of all layers, BatchNorm gamma / beta -- must then agree to max|a-b| / max|b| <= 1e-4.
    (The Linear biases in front of BatchNorm have a mathematically zero gradient and are skipped.)"""
    from glnn_b200 import mlp_engine
    f, h, c, bs, _ = shape
    model, feats, labels, out_t, idx1, _ = _real_problem(shape, dev, 1)
    p = _oracle_state(model, torch.float64)
    x = feats.double()[idx1[0]]
    n_shift = _clear_relu_margin(p, x, 3, "batch", 2e-4, torch.Generator().manual_seed(1))
    model.load_state_dict({"encoder." + k: (v.float() if v.is_floating_point() else v)
                           for k, v in p.items()})
    p = _oracle_state(model, torch.float64)          # exactly what the device holds
    tgt = labels if kind == "nll" else out_t.double()
    logits, cache = O.mlp_forward(x, p, 3, "batch", True)
    for l in range(2):
        assert float(cache[l][3].abs().min()) >= 1.9e-4
    loss, dlog = O.loss_and_dlogits(logits, tgt[idx1[0]], kind, 0.6)
    want = O.mlp_backward(dlog, cache, p, 3, "batch", 0.0)
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    model.train()
    got_loss = mlp_engine.train_pass(model.encoder, opt, feats.to(dev),
                                     (labels if kind == "nll" else out_t).to(dev), idx1.to(dev), 0.6)
    assert abs(got_loss.item() - float(loss)) < 1e-4 * abs(float(loss))
    got = mlp_engine.flat_grads(model.encoder)
    worst = {}
    for k, w in want.items():
        if noise_driven(k, 3, "batch"):
            continue
        worst[k] = relerr(got[k].cpu(), w)
    print(f"{shape} {kind}: {n_shift} columns shifted; max-rel gradient errors "
          + ", ".join(f"{k}={v:.1e}" for k, v in worst.items()))
    for k, v in worst.items():
        assert v < TOL, (k, v)


@pytest.mark.parametrize("wd", [0.0, 5e-4])
@pytest.mark.parametrize("step", [1, 2, 37, 5000])
def test_adam_kernel_from_injected_gradients(dev, step, wd):
    """adam_kernel (csrc/mlp.cu; torch.optim.Adam, train_student.py:275-277) driven by INJECTED
    gradients through glnn_adam_step_f32 -- no forward / backward in the loop, so nothing but the
    update rule is tested: parameters and both moments against the fp64 formula at 1e-6, the applied
    update itself at 2e-5, and against torch.optim.Adam on the same device."""
    from glnn_b200 import ops
    gen = torch.Generator().manual_seed(100 + step)
    n = (1 << 20) + 3
    p0 = torch.randn(n, generator=gen)
    g = torch.randn(n, generator=gen) * torch.logspace(-9, 0, n)     # magnitudes from 1e-9 to 1
    m0 = torch.randn(n, generator=gen) * 0.1 if step > 1 else torch.zeros(n)
    v0 = torch.rand(n, generator=gen) * 0.01 if step > 1 else torch.zeros(n)
    lr, b1, b2, eps = 0.01, 0.9, 0.999, 1e-8
    pd, gd, md, vd = (t.double() for t in (p0, g, m0, v0))
    ge = gd + wd * pd
    m_want = b1 * md + (1 - b1) * ge
    v_want = b2 * vd + (1 - b2) * ge * ge
    denom = v_want.sqrt() / (1 - b2 ** step) ** 0.5 + eps
    p_want = pd - (lr / (1 - b1 ** step)) * m_want / denom
    p, m, v = p0.to(dev), m0.to(dev), v0.to(dev)
    ops.adam_step(p, g.to(dev), m, v, step, lr, (b1, b2), eps, wd)
    assert relerr(m.cpu(), m_want) < 1e-6
    assert relerr(v.cpu(), v_want) < 1e-6
    # With weight decay, the few elements whose g + wd * p cancels down to ~eps (1e-8) turn the fp32
    # rounding of that sum (3e-11) into a 1e-3 relative change of THEIR update: inherent to any fp32
    # Adam (torch's own kernel shows the same against the fp64 formula).  They are held to the size of
    # one update (lr); everything else to 1e-6 of max|p| and 5e-5 of the update.
    stable = ge.abs() > 1e-5
    assert relerr(p.cpu()[stable], p_want[stable]) < 1e-6
    assert float((p.cpu().double() - p_want).abs().max()) < lr * 5e-3
    upd_want = p_want - pd
    # the applied update, recovered from fp32 parameters: |p| eps / |update| = 3 * 6e-8 / 0.01 ~ 2e-5
    assert float((((p.cpu().double() - pd) - upd_want).abs()[stable]).max() / upd_want.abs().max()) < 5e-5
    # torch.optim.Adam from the same state on the same device
    q = torch.nn.Parameter(p0.to(dev).clone())
    opt = torch.optim.Adam([q], lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd)
    q.grad = g.to(dev).clone()
    opt.state[q] = {"step": torch.tensor(float(step - 1)), "exp_avg": m0.to(dev).clone(),
                    "exp_avg_sq": v0.to(dev).clone()}
    opt.step()
    # torch's own fp32 kernel sits up to ~3e-6 from the fp64 formula where |g| ~ eps (step 1)
    assert relerr(p.cpu()[stable], q.detach().cpu()[stable]) < 1e-5
    assert float((p.cpu() - q.detach().cpu()).abs().max()) < lr * 5e-3
    assert relerr(m.cpu(), opt.state[q]["exp_avg"].cpu()) < 1e-6
