"""The numpy oracle (oracle/rechub_oracle.py) against the golden vectors recorded from the live reference, and the
torch CPU port (oracle/ref_port.py) against the live reference itself.  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

import _golden
import _live_reference as live

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import rechub_oracle as orc  # noqa: E402


@pytest.mark.parametrize("name", _golden.NAMES)
def test_oracle_matches_reference_golden_train(name):
    rec = _golden.load(name)
    out = _golden.oracle_run(orc, name, rec, train=True)
    # the reference computes in fp32; the oracle in fp64: agreement at fp32 rounding level
    assert np.abs(out["logit"] - rec["train_logit"]).max() < 2e-6
    assert np.abs(out["prob"] - rec["train_prob"]).max() < 5e-7
    if "grads" not in out:
        return
    assert abs(out["loss"] - float(rec["loss"])) < 1e-6
    assert set(rec["grad"]) == set(out["grads"]), set(rec["grad"]) ^ set(out["grads"])
    for k, ref in rec["grad"].items():
        got = np.asarray(out["grads"][k]).reshape(ref.shape)
        scale = max(np.abs(ref).max(), 1e-3)
        if k.endswith(".bias") and k[:-4] + "weight" in rec["grad"]:  # a Linear bias in front of BatchNorm has a true gradient of 0
            scale = max(scale, np.abs(rec["grad"][k[:-4] + "weight"]).max())
        assert np.abs(got - ref).max() <= 2e-5 * scale, (k, np.abs(got - ref).max(), scale)


def test_dssm_oracle_matches_reference_golden():
    """Two towers + in-batch hard negatives + cross entropy: embeddings, logits, loss, every gradient."""
    rec = _golden.load("dssm")
    out = _golden.dssm_oracle_run(orc, rec)
    _golden.check_dssm_against_golden(rec, out, out["grads"], emb_tol=2e-6, grad_rtol=2e-5)
    rows = np.arange(out["scores"].shape[0])[:, None]
    assert np.abs(out["scores"][rows, out["neg_idx"]] - rec["scores"][rows, rec["neg_idx"]]).max() < 2e-6
    assert not np.any(out["neg_idx"] == rows)


@pytest.mark.parametrize("name", _golden.NAMES)
def test_oracle_matches_reference_golden_eval(name):
    rec = _golden.load(name)
    out = _golden.oracle_run(orc, name, rec, train=False, backward=False)
    assert np.abs(out["prob"] - rec["eval_prob"]).max() < 5e-7


def test_closed_forms():
    # FM of all-ones (B, 26, 16) = 0.5 * (26^2 - 26) * 16 = 5200; one non-zero field -> 0   
    assert np.all(orc.fm_forward(np.ones((3, 26, 16))) == 5200.0)
    e = np.zeros((2, 5, 4))
    e[:, 2, :] = 3.0
    assert np.all(orc.fm_forward(e) == 0.0)
    # CrossNetwork with w = 0: out = x + sum_l b_l
    x = np.random.RandomState(0).randn(4, 7)
    out, _ = orc.cross_forward(x, [np.zeros(7)] * 3, [np.full(7, 0.5 * (i + 1)) for i in range(3)])
    assert np.allclose(out, x + 3.0)
    # mean pooling: padding_idx=None counts id 0 as a token; padding_idx=0 masks it   
    W = np.arange(12, dtype=np.float64).reshape(4, 3)
    ids = np.array([[3, 0, 0]])
    assert np.allclose(orc.seq_pool(W, ids, "mean", None), (W[3] + 2 * W[0]) / 3)
    assert np.allclose(orc.seq_pool(W, ids, "mean", 0), W[3])
    assert np.allclose(orc.seq_pool(W, np.array([[0, 0, 0]]), "mean", 0), 0.0)
    # embedding: out of range raises like the reference, float ids truncate, padding row gets no gradient
    with pytest.raises(IndexError):
        orc.embedding_lookup(W, np.array([4]))
    assert np.all(orc.embedding_lookup(W, np.array([2.9])) == W[2])
    g = orc.embedding_grad(W.shape, np.array([1, 1, 0, 3]), np.ones((4, 3)), padding_idx=0)
    assert np.all(g[1] == 2) and np.all(g[0] == 0) and np.all(g[3] == 1) and np.all(g[2] == 0)


def test_oracle_gradients_by_finite_differences():
    """Independent of any reference: the hand-derived backward equals numerical differentiation of the forward."""
    rec = _golden.load("din")
    base = _golden.oracle_run(orc, "din", rec, train=True)
    rng = np.random.RandomState(0)
    for key in ["embedding.embed_dict.target_item_id.weight", "attention_layers.0.attention.mlp.0.weight", "attention_layers.1.attention.mlp.2.alpha", "mlp.mlp.5.weight"]:
        w = rec["sd"][key].astype(np.float64)
        d = rng.randn(*w.shape)
        d /= np.linalg.norm(d)
        eps = 1e-5
        vals = []
        for sgn in (+1, -1):
            rec["sd"][key] = w + sgn * eps * d
            vals.append(_golden.oracle_run(orc, "din", rec, train=True, backward=False)["loss"])
        rec["sd"][key] = w
        num = (vals[0] - vals[1]) / (2 * eps)
        ana = float((np.asarray(base["grads"][key]).reshape(w.shape) * d).sum())
        assert abs(num - ana) <= 1e-6 + 1e-4 * abs(ana), (key, num, ana)


@pytest.mark.skipif(not live.live_reference_available(), reason="live reference only exists in the build container")
def test_ref_port_matches_live_reference():
    import ref_port
    F = live.ref_module("basic.features")
    M = live.ref_module("models.ranking")
    for deep_sparse in (True, False):
        torch.manual_seed(3)
        dense = [F.DenseFeature("I%d" % i) for i in range(3)]
        sparse = [F.SparseFeature("C%d" % i, vocab_size=50, embed_dim=8) for i in range(4)]
        ref = M.DeepFM(dense + sparse if deep_sparse else dense, sparse, {"dims": [16, 8], "dropout": 0.0, "activation": "relu"})
        port = ref_port.PortDeepFM(3, [50] * 4, 8, mlp_dims=(16, 8), dropout=0.0, deep_includes_sparse=deep_sparse)
        sd = ref.state_dict()
        mapped = {}
        for k, v in sd.items():
            if k.startswith("embedding.embed_dict."):
                mapped["tables.%d.weight" % int(k.split(".")[2][1:])] = v
            elif k.startswith("linear.fc."):
                mapped["linear." + k.split(".")[-1]] = v
            else:
                mapped[k[len("mlp."):]] = v
        port.load_state_dict(mapped)
        g = torch.Generator().manual_seed(0)
        x = {"I%d" % i: torch.rand(32, generator=g) for i in range(3)}
        x.update({"C%d" % i: torch.randint(0, 50, (32,), generator=g) for i in range(4)})
        a = ref(x)
        b = port([x["I%d" % i] for i in range(3)], [x["C%d" % i] for i in range(4)])
        assert torch.equal(a, b)
        a.sum().backward()
        b.sum().backward()
        assert torch.equal(ref.embedding.embed_dict["C0"].weight.grad, port.tables[0].weight.grad)


@pytest.mark.parametrize("seed,B,F,D,V,n_dense", [(1, 1, 2, 4, 7, 0), (2, 33, 9, 16, 50, 5), (3, 64, 3, 8, 11, 1), (4, 17, 26, 16, 23, 13)])
def test_oracle_matches_package_cpu_route_on_fresh_inputs(seed, B, F, D, V, n_dense):
    """Beyond the golden files: random shapes, seeded inputs — the oracle against this package's CPU route (itself bit-identical
    to the live reference, tests/test_cpu_api.py) for DeepFM and DCN: logits and every parameter gradient."""
    from torch_rechub.basic.features import DenseFeature, SparseFeature
    from torch_rechub.basic.initializers import RandomNormal
    from torch_rechub.models.ranking import DCN, DeepFM
    torch.manual_seed(seed)
    init = RandomNormal(0, 0.1)
    dense = [DenseFeature("I%d" % i) for i in range(n_dense)]
    sparse = [SparseFeature("C%d" % i, V, D, initializer=init) for i in range(F)]
    g = torch.Generator().manual_seed(seed)
    x = {"I%d" % i: torch.rand(B, generator=g) for i in range(n_dense)}
    x.update({"C%d" % i: torch.randint(0, V, (B,), generator=g) for i in range(F)})
    y = torch.randint(0, 2, (B,), generator=g).float()
    xn, yn = {k: v.numpy() for k, v in x.items()}, y.numpy()
    dn, sn = [f.name for f in dense], [f.name for f in sparse]
    train = B > 1  # BatchNorm cannot take batch statistics of one row
    cases = [(DeepFM(dense + sparse, sparse, {"dims": [12, 6], "dropout": 0.0, "activation": "relu"}), lambda sd: orc.deepfm_forward_backward(sd, xn, yn, dn, sn, sn, 2, train=train)),
             (DCN(dense + sparse, n_cross_layers=2, mlp_params={"dims": [12, 6]}), lambda sd: orc.dcn_forward_backward(sd, xn, yn, dn, sn, 2, 2, train=train))]
    for model, run in cases:
        model.train(train)
        p = model(x)
        loss = torch.nn.BCELoss()(p, y)
        model.zero_grad()
        loss.backward()
        out = run({k: v.detach().numpy() for k, v in model.state_dict().items()})
        assert np.abs(out["prob"] - p.detach().numpy()).max() < 1e-6
        assert abs(out["loss"] - float(loss)) < 1e-6
        for k, prm in model.named_parameters():
            ref = prm.grad.numpy()
            scale = max(np.abs(ref).max(), 1e-3)
            if k.endswith(".bias") and k[:-4] + "weight" in out["grads"]:
                scale = max(scale, np.abs(dict(model.named_parameters())[k[:-4] + "weight"].grad.numpy()).max())
            assert np.abs(np.asarray(out["grads"][k]).reshape(ref.shape) - ref).max() <= 5e-5 * scale, (type(model).__name__, k)
